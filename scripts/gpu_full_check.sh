#!/bin/bash
# Full state check of one GPU-box visit: smoke, every GPU parity test, every bench workload (own arm + reference arm),
# ncu launch lists of cfg1 / cfg3 and full captures of the scoring kernels.  TAG names the output files.
set -u
TAG=${TAG:-r2}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > $O/pytest_gpu.log 2>&1
echo "pytest exit $?" >> $O/pytest_gpu.log
timeout 300 python bench.py > $O/bench_default.json 2> $O/bench_default.err
timeout 300 python bench.py --impl reference --steps 5 --warmup 3 > $O/bench_default_reference.json 2> $O/bench_default_reference.err
timeout 300 python bench.py --workload cfg0 --steps 200 --warmup 20 > $O/bench_cfg0.json 2> $O/bench_cfg0.err
timeout 300 python bench.py --workload cfg2 --steps 50 --warmup 5 > $O/bench_cfg2.json 2> $O/bench_cfg2.err
timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 5 > $O/bench_cfg3_d10.json 2> $O/bench_cfg3_d10.err
timeout 300 python bench.py --workload cfg3 --dim 50 --steps 50 --warmup 5 > $O/bench_cfg3_d50.json 2> $O/bench_cfg3_d50.err
timeout 300 python bench.py --workload cfg3 --score-mode matrix --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > $O/bench_cfg3_d10_matrix.json 2> $O/bench_cfg3_d10_matrix.err
timeout 300 python bench.py --workload cfg3 --score-mode topk --steps 50 --warmup 5 --no-cpu-baseline > $O/bench_cfg3_d10_topk.json 2> $O/bench_cfg3_d10_topk.err
timeout 300 python bench.py --workload cfg4 --steps 100 --warmup 10 --pairs 131040 > $O/bench_cfg4.json 2> $O/bench_cfg4.err
timeout 300 python scripts/score_bench.py --dims 10,50 --iters 9 --modes topk,matrix_lm,both_lm --engines simt,tc > $O/score_bench.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_cfg1.csv \
   python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_cfg1.log 2>&1
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 100 --csv --log-file $O/launches_cfg3.csv \
   python bench.py --workload cfg3 --steps 3 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_cfg3.log 2>&1
for w in cfg0 cfg2 cfg4; do
  case $w in cfg0) a="--workload cfg0";; cfg2) a="--workload cfg2";; cfg4) a="--workload cfg4 --pairs 131040 --rotation 4";; esac
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 200 --csv --log-file $O/launches_$w.csv \
     python bench.py $a --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_$w.log 2>&1
done
for D in ${NCU_DIMS-10 50}; do
  # score_bench runs 2 warm-up + 1 timed launch per mode: launches 2, 5, 8 of the kernel are the timed top-k / matrix / both
  for m in topk matrix_lm both_lm; do
    timeout 300 ncu --set full --clock-control none --import-source on -k regex:score_mma_kernel -s 2 -c 1 -o /tmp/prof_tc_d${D}_$m \
        python scripts/score_bench.py --images 303104 --dims $D --iters 1 --modes $m --engines tc > $O/ncu_tc_d${D}_$m.log 2>&1
    ncu -i /tmp/prof_tc_d${D}_$m.ncu-rep --page raw --csv > $O/tc_d${D}_${m}_raw.csv 2>/dev/null
  done
done
timeout 300 ncu --set full --clock-control none --import-source on -k regex:pairs_grouped -s 6 -c 1 -o /tmp/prof_pairs \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_pairs.log 2>&1
ncu -i /tmp/prof_pairs.ncu-rep --page raw --csv > $O/pairs_cfg1_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:update_rows_kernel -s 6 -c 1 -o /tmp/prof_upd4 \
    python bench.py --workload cfg4 --steps 6 --warmup 3 --pairs 131040 --no-cpu-baseline --no-e2e --rotation 4 > $O/ncu_update_cfg4.log 2>&1
ncu -i /tmp/prof_upd4.ncu-rep --page raw --csv > $O/update_rows_cfg4_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:update_rows_kernel -s 6 -c 1 -o /tmp/prof_upd1 \
    python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_update_cfg1.log 2>&1
ncu -i /tmp/prof_upd1.ncu-rep --page raw --csv > $O/update_rows_cfg1_raw.csv 2>/dev/null
timeout 300 ncu --set full --clock-control none -k regex:pairs_grouped -s 6 -c 1 -o /tmp/prof_pairs0 \
    python bench.py --workload cfg0 --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_pairs_cfg0.log 2>&1
ncu -i /tmp/prof_pairs0.ncu-rep --page raw --csv > $O/pairs_cfg0_raw.csv 2>/dev/null
for k in featnet_fwd featnet_wgrad; do
  timeout 300 ncu --set full --clock-control none -k regex:$k -s 3 -c 1 -o /tmp/prof_$k \
      python bench.py --workload cfg2 --steps 4 --warmup 3 --no-cpu-baseline --no-e2e > $O/ncu_$k.log 2>&1
  ncu -i /tmp/prof_$k.ncu-rep --page raw --csv > $O/${k}_cfg2_raw.csv 2>/dev/null
done
tail -3 $O/smoke.log; tail -8 $O/pytest_gpu.log; cat $O/score_bench.log
for w in default default_reference cfg0 cfg2 cfg3_d10 cfg3_d50 cfg3_d10_matrix cfg3_d10_topk cfg4; do tail -2 $O/bench_$w.err; cut -c1-700 $O/bench_$w.json; echo; done
