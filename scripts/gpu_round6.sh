#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -8 gpurun_out/pytest_gpu.log
timeout 100 python scripts/write_bw.py > gpurun_out/write_bw.log 2>&1; cat gpurun_out/write_bw.log
timeout 100 scripts/bin/ubench_store_pattern > gpurun_out/store_pattern.log 2>&1; cat gpurun_out/store_pattern.log
timeout 300 python scripts/score_bench.py --dims 10,50 --iters 7 --modes topk,matrix_lm,both_lm --engines tc > gpurun_out/score_bench.log 2>&1
cat gpurun_out/score_bench.log
for D in 10; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_mma_kernel -c 6 -o /tmp/prof_tc_d$D \
      python scripts/score_bench.py --images 303104 --dims $D --iters 1 --modes topk,matrix_lm --engines tc > gpurun_out/ncu_tc_d$D.log 2>&1
  ncu -i /tmp/prof_tc_d$D.ncu-rep --page raw --csv > gpurun_out/tc_d${D}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_tc_d$D.ncu-rep --page source --csv > gpurun_out/tc_d${D}_source.csv 2>/dev/null
done
