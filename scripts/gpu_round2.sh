#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 200 python scripts/score_bench.py --dims 10,50 --iters 7 --modes topk,matrix_lm > gpurun_out/score_bench.log 2>&1
DIMS=10 MODES=topk,matrix_lm bash scripts/score_sweep.sh 4,128,4,16 2,128,1,16 > /dev/null 2>&1
cp gpurun_out/score_sweep.log gpurun_out/score_sweep_d10.log
timeout 300 python bench.py --workload cfg4 --steps 100 --warmup 10 --pairs 131040 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 30 -c 40 --csv --log-file gpurun_out/launches_cfg4.csv \
   python bench.py --workload cfg4 --steps 6 --warmup 3 --pairs 131040 --no-cpu-baseline --no-e2e --rotation 4 > gpurun_out/ncu_cfg4.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -s 60 -c 60 --csv --log-file gpurun_out/launches_cfg2.csv \
   python bench.py --workload cfg2 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --rotation 2 > gpurun_out/ncu_cfg2.log 2>&1
du -sh gpurun_out
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/score_bench.log; grep -E "==|topk|matrix" gpurun_out/score_sweep_d10.log; tail -2 gpurun_out/bench_cfg4.err; cat gpurun_out/bench_cfg4.json
