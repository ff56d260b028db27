#!/bin/bash
set -u
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider -k "scor or classif or f1 or sampler" > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python scripts/score_bench.py --dims 10,50 --iters 7 --modes topk,matrix_lm,both_lm --engines tc > gpurun_out/score_bench.log 2>&1
cat gpurun_out/score_bench.log
