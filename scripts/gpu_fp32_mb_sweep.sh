#!/bin/bash
# Register-cap sweep of the fp32-core pair kernels (LEC_FP32_MINBLOCKS = resident 256-thread blocks the cap allows).
set -u
O=gpurun_out/${TAG:-r2x}; mkdir -p $O
for mbx in ${SWEEP:-2 3 4}; do
  rm -f learning_embeddings_b200/csrc/build/lec_pairs_euc32.o learning_embeddings_b200/csrc/build/lec_pairs_oe32.o learning_embeddings_b200/csrc/build/lec_pairs_hyp32.o
  make -C learning_embeddings_b200/csrc EXTRA=-DLEC_FP32_MINBLOCKS=$mbx > $O/make_mbx$mbx.log 2>&1
  for w in cfg0 cfg2; do
    case $w in cfg0) args="--workload cfg0 --steps 200 --warmup 20";; cfg2) args="--workload cfg2 --steps 50 --warmup 5";; esac
    timeout 300 python bench.py $args --no-cpu-baseline --no-e2e > $O/bench_${w}_mbx$mbx.json 2> $O/bench_${w}_mbx$mbx.err
    python -c "import json;d=json.loads(open('$O/bench_${w}_mbx$mbx.json').read().strip().splitlines()[-1]);print('mbx=$mbx $w ms_per_step %.4f kernel_ms %.4f' % (d['ms_per_step'],d['roofline']['kernel_ms']))"
  done
done
rm -f learning_embeddings_b200/csrc/build/lec_pairs_euc32.o learning_embeddings_b200/csrc/build/lec_pairs_oe32.o learning_embeddings_b200/csrc/build/lec_pairs_hyp32.o
make -C learning_embeddings_b200/csrc > $O/make_restore.log 2>&1
