#!/bin/bash
# Phase trace of the cfg1 step (debug build of the library with -DLEC_STEP_TRACE) on 1 and N GPUs.
set -u
N=${NGPU:-2}
rm -f learning_embeddings_b200/csrc/build/*.o
make -C learning_embeddings_b200/csrc -j16 EXTRA=-DLEC_STEP_TRACE > /tmp/make_trace.log 2>&1 || tail -5 /tmp/make_trace.log
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29513"
python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-sustained --step-series 10 2>&1 | grep "step phases"
$TR bench.py --gpus $N --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-sustained --step-series 10 2>&1 | grep "step phases"
rm -f learning_embeddings_b200/csrc/build/*.o
make -C learning_embeddings_b200/csrc -j16 > /tmp/make_restore.log 2>&1
