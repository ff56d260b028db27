#!/bin/bash
# Re-entry visit: tc probe, all parity tests, scoring timing (both engines), benches cfg1/cfg2/cfg4 + reference arm,
# launch list of the cfg1 bench, ncu full captures of the scoring kernels reduced to CSV on the box.
set -u
mkdir -p gpurun_out
nvidia-smi --query-gpu=name,clocks.max.sm,memory.total --format=csv > gpurun_out/gpu.txt 2>&1
timeout 900 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
timeout 300 python scripts/score_bench.py --dims 10,50 --iters 7 --modes topk,matrix_lm,both_lm --engines simt,tc > gpurun_out/score_bench.log 2>&1
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > gpurun_out/bench_cfg1_ref.json 2> gpurun_out/bench_cfg1_ref.err
timeout 300 python bench.py --workload cfg2 --steps 50 --warmup 5 > gpurun_out/bench_cfg2.json 2> gpurun_out/bench_cfg2.err
timeout 300 python bench.py --workload cfg4 --steps 100 --warmup 10 --pairs 131040 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg1.csv \
   python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_cfg1.log 2>&1
for D in 10 50; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_fast -s 2 -c 2 -o /tmp/prof_score_d$D \
      python scripts/score_bench.py --images 303104 --dims $D --iters 1 --modes topk,matrix_lm --engines simt > gpurun_out/ncu_score_d$D.log 2>&1
  ncu -i /tmp/prof_score_d$D.ncu-rep --page raw --csv > gpurun_out/score_d${D}_raw.csv 2>/dev/null
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_mma_kernel -s 2 -c 2 -o /tmp/prof_tc_d$D \
      python scripts/score_bench.py --images 303104 --dims $D --iters 1 --modes topk,matrix_lm --engines tc > gpurun_out/ncu_tc_d$D.log 2>&1
  ncu -i /tmp/prof_tc_d$D.ncu-rep --page raw --csv > gpurun_out/tc_d${D}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_tc_d$D.ncu-rep --page source --csv > gpurun_out/tc_d${D}_source.csv 2>/dev/null
done
du -sh gpurun_out
tail -3 gpurun_out/pytest_gpu.log; cat gpurun_out/score_bench.log; for w in cfg1 cfg1_ref cfg2 cfg4; do tail -2 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json; done
