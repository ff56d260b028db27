#!/bin/bash
# GPU-box visit for the scoring kernel: parity tests, 1M x 723 timing, tuning sweep, small-N ncu captures.
set -u
mkdir -p gpurun_out
timeout 600 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "scoring or pipelined or engine" > gpurun_out/pytest_score.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_score.log
timeout 300 python scripts/score_bench.py --dims 10,50 --iters 7 --modes topk,matrix_lm,both_lm > gpurun_out/score_bench.log 2>&1
DIMS=10 MODES=topk,matrix_lm bash scripts/score_sweep.sh 2,128,1,16 6,128,1,24 8,128,1,32 4,256,1,16 4,128,1,32 > /dev/null 2>&1
cp gpurun_out/score_sweep.log gpurun_out/score_sweep_d10.log
DIMS=50 MODES=topk,matrix_lm bash scripts/score_sweep.sh 4,128,1,16 4,256,1,16 2,256,4,16 > /dev/null 2>&1
cp gpurun_out/score_sweep.log gpurun_out/score_sweep_d50.log
for D in 10 50; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_fast -s 4 -c 2 -o gpurun_out/prof_score_d$D \
      python scripts/score_bench.py --images 303104 --dims $D --iters 1 --modes topk,matrix_lm > gpurun_out/ncu_score_d$D.log 2>&1
done
ls -la gpurun_out/*.ncu-rep
for f in gpurun_out/*.ncu-rep; do s=$(stat -c %s $f); if [ $s -gt 28000000 ]; then echo "dropping $f ($s bytes)"; rm -f $f; fi; done
tail -4 gpurun_out/pytest_score.log; cat gpurun_out/score_bench.log; grep -E "==|topk|matrix" gpurun_out/score_sweep_d10.log gpurun_out/score_sweep_d50.log
