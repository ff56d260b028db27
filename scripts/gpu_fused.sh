#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x -k "engine or fused or pipelined or step or rsgd or rows" > gpurun_out/pytest_step.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_step.log
tail -15 gpurun_out/pytest_step.log
for f in 1 0; do
echo "== LEC_FUSED_STEP=$f"
LEC_FUSED_STEP=$f timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print(d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step  kernel', d['roofline']['kernel_ms']*1000, 'launches', d['gpu_launches'], 'e2e', d['e2e']['value']/1e9, d['e2e']['sync']['value']/1e9)"
LEC_FUSED_STEP=$f timeout 200 python bench.py --workload cfg4 --steps 100 --warmup 10 --pairs 131040 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4', d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step  kernel', d['roofline']['kernel_ms']*1000, 'launches', d['gpu_launches'])"
done
