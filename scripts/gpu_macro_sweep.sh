#!/bin/bash
# A/B of a compile-time switch on one box: rebuild the pair kernels with -D$MACRO=v for every v in $VALUES and run the
# bench workloads in $WORKLOADS (device-timed only).   MACRO=LEC_PREFETCH_INDICES VALUES="0 1" WORKLOADS="cfg0 cfg1 cfg4"
set -u
O=gpurun_out/${TAG:-r2y}; mkdir -p $O
OBJS="lec_pairs_euc32 lec_pairs_oe32 lec_pairs_hyp32 lec_pairs_hyp64 ${EXTRA_OBJS:-}"
for v in ${VALUES:-0 1}; do
  for o in $OBJS; do rm -f learning_embeddings_b200/csrc/build/$o.o; done
  make -C learning_embeddings_b200/csrc -j8 EXTRA=-D${MACRO}=$v > $O/make_$v.log 2>&1 || { tail -3 $O/make_$v.log; continue; }
  for w in ${WORKLOADS:-cfg0 cfg1 cfg4}; do
    case $w in
      cfg0) args="--workload cfg0 --steps 200 --warmup 20";;
      cfg1) args="--workload cfg1 --steps 200 --warmup 20";;
      cfg2) args="--workload cfg2 --steps 50 --warmup 5";;
      cfg4) args="--workload cfg4 --steps 100 --warmup 10 --pairs 131040";;
    esac
    timeout 300 python bench.py $args --no-cpu-baseline --no-e2e --no-sustained > $O/bench_${w}_$v.json 2> $O/bench_${w}_$v.err
    python -c "import json;d=json.loads(open('$O/bench_${w}_$v.json').read().strip().splitlines()[-1]);print('${MACRO}=$v $w ms_per_step %.4f kernel_ms %.4f' % (d['ms_per_step'],d['roofline']['kernel_ms']))" || tail -3 $O/bench_${w}_$v.err
  done
done
for o in $OBJS; do rm -f learning_embeddings_b200/csrc/build/$o.o; done
make -C learning_embeddings_b200/csrc -j8 > $O/make_restore.log 2>&1
