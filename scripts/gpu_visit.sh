#!/bin/bash
# One 1-GPU box visit: smoke, GPU tests, the bench workloads named in WORKLOADS, launch lists and ncu captures.
# TAG names the output directory under gpurun_out/.
set -u
TAG=${TAG:-r2a}
O=gpurun_out/$TAG
mkdir -p $O
nvidia-smi --query-gpu=name,clocks.sm,clocks.max.sm,power.draw --format=csv > $O/gpu.txt 2>&1
python -c "import __graft_entry__ as g; g.smoke()" > $O/smoke.log 2>&1; echo "smoke exit $?" >> $O/smoke.log
tail -3 $O/smoke.log
if [ "${PYTEST:-1}" = "1" ]; then
  timeout 1500 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider ${PYTEST_ARGS:-} > $O/pytest_gpu.log 2>&1
  echo "pytest exit $?" >> $O/pytest_gpu.log
  tail -25 $O/pytest_gpu.log
fi
for w in ${WORKLOADS:-default cfg0 cfg4}; do
  case $w in
    default) args="";;
    cfg0) args="--workload cfg0 --steps 200 --warmup 20";;
    cfg2) args="--workload cfg2 --steps 50 --warmup 5";;
    cfg3) args="--workload cfg3 --steps 30 --warmup 5";;
    cfg3_d50) args="--workload cfg3 --dim 50 --steps 30 --warmup 5";;
    cfg4) args="--workload cfg4 --steps 100 --warmup 10 --pairs 131040";;
  esac
  timeout 400 python bench.py $args > $O/bench_$w.json 2> $O/bench_$w.err
  tail -2 $O/bench_$w.err; cut -c1-1500 $O/bench_$w.json; echo
done
for w in ${NCU_LISTS:-cfg1 cfg4}; do
  case $w in
    cfg1) args="--steps 6 --warmup 3";;
    cfg0) args="--workload cfg0 --steps 6 --warmup 3";;
    cfg2) args="--workload cfg2 --steps 4 --warmup 3";;
    cfg4) args="--workload cfg4 --steps 6 --warmup 3 --pairs 131040 --rotation 4";;
  esac
  timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file $O/launches_$w.csv \
     python bench.py $args --no-cpu-baseline --no-e2e > $O/ncu_list_$w.log 2>&1
done
for spec in ${NCU_FULL:-update_rows_kernel:cfg4 update_rows_kernel:cfg1}; do
  k=${spec%%:*}; w=${spec##*:}
  case $w in
    cfg1) args="--steps 6 --warmup 3";;
    cfg0) args="--workload cfg0 --steps 6 --warmup 3";;
    cfg2) args="--workload cfg2 --steps 4 --warmup 3";;
    cfg4) args="--workload cfg4 --steps 6 --warmup 3 --pairs 131040 --rotation 4";;
  esac
  timeout 300 ncu --set full --clock-control none --import-source on -k regex:$k -s 5 -c 1 -o /tmp/prof_${k}_$w \
      python bench.py $args --no-cpu-baseline --no-e2e > $O/ncu_full_${k}_$w.log 2>&1
  ncu -i /tmp/prof_${k}_$w.ncu-rep --page raw --csv > $O/${k}_${w}_raw.csv 2>/dev/null
  cp /tmp/prof_${k}_$w.ncu-rep $O/ 2>/dev/null
done
# optional: rebuild the update kernel for another residency and re-run a workload (UPD_MB_SWEEP="2 3")
for mb in ${UPD_MB_SWEEP:-}; do
  rm -f learning_embeddings_b200/csrc/build/lec_update.o
  make -C learning_embeddings_b200/csrc EXTRA=-DLEC_UPD_MINBLOCKS=$mb > $O/make_mb$mb.log 2>&1
  for w in cfg4 default; do
    case $w in default) args="";; cfg4) args="--workload cfg4 --steps 100 --warmup 10 --pairs 131040";; esac
    timeout 300 python bench.py $args --no-cpu-baseline --no-e2e > $O/bench_${w}_mb$mb.json 2> $O/bench_${w}_mb$mb.err
    python -c "import json;d=json.loads(open('$O/bench_${w}_mb$mb.json').read().strip().splitlines()[-1]);print('mb=$mb $w ms_per_step',d['ms_per_step'],'kernel_ms',d['roofline']['kernel_ms'])"
  done
done
