#!/bin/bash
set -u
mkdir -p gpurun_out
python scripts/dbg_aux.py 2>&1 | grep step
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "engine or fused or pipelined or step or rsgd or rows or joint or grouped or pipeline" > gpurun_out/pytest_step.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_step.log
tail -5 gpurun_out/pytest_step.log
for b in 256 128 64 96; do
echo "== LEC_GROUP_BLOCK=$b"
LEC_GROUP_BLOCK=$b timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step  kernel', d['roofline']['kernel_ms']*1000)"
LEC_GROUP_BLOCK=$b timeout 200 python bench.py --workload cfg4 --steps 100 --warmup 10 --pairs 131040 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4', d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step  kernel', d['roofline']['kernel_ms']*1000)"
LEC_GROUP_BLOCK=$b timeout 200 python bench.py --workload cfg2 --steps 50 --warmup 5 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg2', d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step  kernel', d['roofline']['kernel_ms']*1000)"
done
