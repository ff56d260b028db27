#!/bin/bash
# sweep of the scoring kernel's tuning knob (LEC_SCORE_CFG = images/thread, threads, min blocks, ring entries)
mkdir -p gpurun_out
: > gpurun_out/score_sweep.log
for cfg in "$@"; do
  echo "== LEC_SCORE_CFG=$cfg" >> gpurun_out/score_sweep.log
  LEC_SCORE_CFG=$cfg timeout 120 python scripts/score_bench.py --dims ${DIMS:-10} --iters 5 --modes ${MODES:-topk,matrix_lm,both_lm} >> gpurun_out/score_sweep.log 2>&1
done
cat gpurun_out/score_sweep.log
