#!/bin/bash
# N-GPU visit (gpurun --gpus N): parity of the sharded step against one GPU, then the cfg1 bench per exchange mode.
set -u
N=${NGPU:-2}
mkdir -p gpurun_out
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
timeout 300 $TR scripts/multigpu_parity.py > gpurun_out/multigpu_parity_${N}gpu.log 2>&1
echo "parity exit $?" >> gpurun_out/multigpu_parity_${N}gpu.log
grep -v "^W10\|Warning\|warn" gpurun_out/multigpu_parity_${N}gpu.log | tail -30
for comm in auto nccl ${EXTRA_COMM:-}; do
  for fused in 1 ${EXTRA_FUSED:-}; do
  LEC_FUSED_STEP=$fused timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 20 --comm $comm --no-cpu-baseline > gpurun_out/bench_${N}gpu_${comm}_f$fused.json 2> gpurun_out/bench_${N}gpu_${comm}_f$fused.err
  python -c "import json,sys; d=json.loads(open('gpurun_out/bench_${N}gpu_${comm}_f$fused.json').read().strip().splitlines()[-1]); print('N=$N comm=$comm fused=$fused', d['config']['exchange'], d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step  kernel', d['roofline']['kernel_ms']*1000, 'e2e', d['e2e']['value']/1e9)" || tail -5 gpurun_out/bench_${N}gpu_${comm}_f$fused.err
  done
done
timeout 200 python bench.py --steps 200 --warmup 20 --no-cpu-baseline | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1', d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step')"
