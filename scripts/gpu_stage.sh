#!/bin/bash
set -u
for st in 1 0; do
echo "== LEC_STAGE_ROWS=$st"
for w in cfg1 cfg0; do
LEC_STAGE_ROWS=$st timeout 200 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step  kernel', d['roofline']['kernel_ms']*1000)"
done
done
LEC_STAGE_ROWS=1 timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "grouped or engine or step or fused or dropin or joint" 2>&1 | tail -2
