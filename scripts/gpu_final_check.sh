#!/bin/bash
# Round-end state of one GPU box: smoke, every GPU test, the driver's two bench commands, the scoring micro-benchmark,
# launch lists and full ncu captures of the step kernels and the scoring kernel.  TAG names the output directory.
set -u
TAG=${TAG:-r2final}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1; echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 400 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_default_reference.json 2> $O/bench_default_reference.err
timeout 200 python scripts/score_bench.py --dims 10,50 --iters 9 --modes topk,matrix_lm,both_lm --engines simt,tc > $O/score_bench.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg1.csv \
   python bench.py --workload cfg1 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained > $O/ncu_cfg1.log 2>&1
timeout 200 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_cfg3.csv \
   python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_cfg3.log 2>&1
timeout 200 ncu --set full --clock-control none --import-source on -k regex:pairs_grouped -s 50 -c 1 -o /tmp/prof_pairs \
    python bench.py --workload cfg1 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained > $O/ncu_pairs.log 2>&1
ncu -i /tmp/prof_pairs.ncu-rep --page raw --csv > $O/pairs_cfg1_raw.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none -k regex:update_rows_kernel -s 50 -c 1 -o /tmp/prof_upd1 \
    python bench.py --workload cfg1 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e --no-sustained > $O/ncu_update_cfg1.log 2>&1
ncu -i /tmp/prof_upd1.ncu-rep --page raw --csv > $O/update_rows_cfg1_raw.csv 2>/dev/null
timeout 200 ncu --set full --clock-control none -k regex:update_rows_kernel -s 50 -c 1 -o /tmp/prof_upd4 \
    python bench.py --workload cfg4 --steps 6 --warmup 3 --pairs 131040 --rotation 4 --no-cpu-baseline --no-e2e --no-sustained > $O/ncu_update_cfg4.log 2>&1
ncu -i /tmp/prof_upd4.ncu-rep --page raw --csv > $O/update_rows_cfg4_raw.csv 2>/dev/null
for m in topk both_lm; do
  timeout 200 ncu --set full --clock-control none -k regex:score_mma_kernel -s 2 -c 1 -o /tmp/prof_tc_$m \
      python scripts/score_bench.py --images 303104 --dims 10 --iters 1 --modes $m --engines tc > $O/ncu_tc_d10_$m.log 2>&1
  ncu -i /tmp/prof_tc_$m.ncu-rep --page raw --csv > $O/tc_d10_${m}_raw.csv 2>/dev/null
done
tail -2 $O/smoke.log; tail -3 $O/pytest_gpu.log; cat $O/score_bench.log; cut -c1-400 $O/bench_default.json; echo; cut -c1-600 $O/bench_default_reference.json
