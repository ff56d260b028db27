#!/bin/bash
# v3 tensor-core scoring (FORMS template, branch-free grouped epilogue, wait backoff) + metrics kernels.
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -x > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/score_bench.py --dims 10,50 --iters 7 --modes topk,matrix_lm,both_lm --engines simt,tc > gpurun_out/score_bench.log 2>&1
cat gpurun_out/score_bench.log
timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 5 > gpurun_out/bench_cfg3_d10.json 2> gpurun_out/bench_cfg3_d10.err
timeout 300 python bench.py --workload cfg3 --dim 50 --steps 50 --warmup 5 > gpurun_out/bench_cfg3_d50.json 2> gpurun_out/bench_cfg3_d50.err
timeout 300 python bench.py --workload cfg3 --score-mode matrix --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg3_d10_matrix.json 2> gpurun_out/bench_cfg3_d10_matrix.err
timeout 300 python bench.py --workload cfg3 --score-mode topk --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg3_d10_topk.json 2> gpurun_out/bench_cfg3_d10_topk.err
for D in 10 50; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_mma_kernel -c 3 -o /tmp/prof_tc_d$D \
      python scripts/score_bench.py --images 303104 --dims $D --iters 1 --modes topk,matrix_lm,both_lm --engines tc > gpurun_out/ncu_tc_d$D.log 2>&1
  ncu -i /tmp/prof_tc_d$D.ncu-rep --page raw --csv > gpurun_out/tc_d${D}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_tc_d$D.ncu-rep --page source --csv > gpurun_out/tc_d${D}_source.csv 2>/dev/null
done
du -sh gpurun_out
for w in cfg3_d10 cfg3_d50 cfg3_d10_matrix cfg3_d10_topk; do tail -2 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json; done
