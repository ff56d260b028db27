#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 300 python scripts/score_bench.py --dims 10,50 --iters 9 --modes topk,both_lm --engines tc 2>&1 | tee gpurun_out/score_bench.log
for mb in 2 1; do
echo "== LEC_GROUP_MINBLOCKS=$mb"
for w in cfg1 cfg0 cfg4 cfg2; do
LEC_GROUP_MINBLOCKS=$mb timeout 200 python bench.py --workload $w --steps 200 --warmup 20 --no-cpu-baseline --no-e2e $( [ $w = cfg4 ] && echo --pairs 131040 ) | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('$w', d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step  kernel', d['roofline']['kernel_ms']*1000)"
done
done
