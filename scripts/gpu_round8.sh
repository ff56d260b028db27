#!/bin/bash
set -u
mkdir -p gpurun_out
for ns in 0 32 200; do
echo "== LEC_TC_SLEEP_NS=$ns"
LEC_TC_SLEEP_NS=$ns timeout 300 python scripts/score_bench.py --dims 10,50 --iters 7 --modes topk,matrix_lm --engines tc 2>&1 | tee gpurun_out/score_bench_sleep$ns.log
done
timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_mma_kernel -c 3 -o /tmp/prof_tc_d10 \
      python scripts/score_bench.py --images 303104 --dims 10 --iters 1 --modes matrix_lm --engines tc > gpurun_out/ncu_tc_d10.log 2>&1
ncu -i /tmp/prof_tc_d10.ncu-rep --page raw --csv > gpurun_out/tc_d10_raw.csv 2>/dev/null
ncu -i /tmp/prof_tc_d10.ncu-rep --page source --csv > gpurun_out/tc_d10_source.csv 2>/dev/null
