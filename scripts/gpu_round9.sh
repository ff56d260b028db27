#!/bin/bash
# alt chunk ownership on/off (matrix-only), then a full ncu capture of the top-k kernel with source counters
set -u
mkdir -p gpurun_out
for alt in 0 1 0 1; do
echo "== LEC_TC_ALT=$alt"
LEC_TC_ALT=$alt timeout 300 python scripts/score_bench.py --dims 10,50 --iters 9 --modes topk,matrix_lm,both_lm --engines tc 2>&1 | tee -a gpurun_out/score_bench_alt$alt.log
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:score_mma_kernel -c 2 -o /tmp/prof_tc_topk_d10 \
      python scripts/score_bench.py --images 303104 --dims 10 --iters 1 --modes topk --engines tc > gpurun_out/ncu_tc_topk_d10.log 2>&1
ncu -i /tmp/prof_tc_topk_d10.ncu-rep --page raw --csv > gpurun_out/tc_topk_d10_raw.csv 2>/dev/null
ncu -i /tmp/prof_tc_topk_d10.ncu-rep --page source --csv > gpurun_out/tc_topk_d10_source.csv 2>/dev/null
ls -la gpurun_out/tc_topk_d10_*
