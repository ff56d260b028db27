#!/bin/bash
set -u
mkdir -p gpurun_out
for r in 32 16 8 4 2; do
echo "== LEC_REPLICAS=$r"
LEC_REPLICAS=$r timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step  kernel', d['roofline']['kernel_ms']*1000)"
done
