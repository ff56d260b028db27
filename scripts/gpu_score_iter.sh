#!/bin/bash
# scoring iteration loop: parity tests of the scoring engines, the cfg3 micro-benchmark of the tc engine, and (with
# NCU=1) a full ncu capture of the top-k kernel with per-instruction counters
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "scor" > gpurun_out/pytest_score.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_score.log
tail -25 gpurun_out/pytest_score.log
timeout 300 python scripts/score_bench.py --dims 10,50 --iters 9 --modes topk,matrix_lm,both_lm --engines tc 2>&1 | tee gpurun_out/score_bench.log
if [ "${NCU:-0}" = "1" ]; then
for D in ${NCU_DIMS:-10}; do
timeout 300 ncu --set full --clock-control none --import-source on -k regex:score_mma_kernel -c 1 -o /tmp/prof_tc_topk_d$D \
      python scripts/score_bench.py --images 303104 --dims $D --iters 1 --modes ${NCU_MODE:-topk} --engines tc > gpurun_out/ncu_tc_topk_d$D.log 2>&1
ncu -i /tmp/prof_tc_topk_d$D.ncu-rep --page raw --csv > gpurun_out/tc_topk_d${D}_raw.csv 2>/dev/null
ncu -i /tmp/prof_tc_topk_d$D.ncu-rep --page source --csv > gpurun_out/tc_topk_d${D}_source.csv 2>/dev/null
done
fi
