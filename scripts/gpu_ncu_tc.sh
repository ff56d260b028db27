#!/bin/bash
set -u
mkdir -p gpurun_out
for D in 10 50; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_mma_kernel -s 2 -c 2 -o /tmp/prof_tc_d$D \
      python scripts/score_bench.py --images 303104 --dims $D --iters 1 --modes topk,matrix_lm --engines tc > gpurun_out/ncu_tc_d$D.log 2>&1
  ncu -i /tmp/prof_tc_d$D.ncu-rep --page raw --csv > gpurun_out/tc_d${D}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_tc_d$D.ncu-rep --page source --csv > gpurun_out/tc_d${D}_source.csv 2>/dev/null
done
du -sh gpurun_out; tail -3 gpurun_out/ncu_tc_d50.log
