#!/bin/bash
# N-GPU visit (gpurun --gpus N): parity of the sharded step against one GPU, then the driver-style bench at N GPUs
# (every workload) and, optionally, the cfg1 bench with the NCCL exchange for comparison.  TAG names the output files.
set -u
N=${NGPU:-2}
TAG=${TAG:-r2}
O=gpurun_out/$TAG
mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511"
if [ "${PARITY:-1}" = "1" ]; then
  timeout 300 $TR scripts/multigpu_parity.py > $O/multigpu_parity_${N}gpu.log 2>&1
  echo "parity exit $?" >> $O/multigpu_parity_${N}gpu.log
  grep -v "^W1\|Warning\|warn" $O/multigpu_parity_${N}gpu.log | tail -${PARITY_TAIL:-14}
fi
timeout 600 $TR bench.py --gpus $N --steps ${STEPS:-20} --warmup 5 ${BENCH_ARGS:-} > $O/bench_${N}gpu.json 2> $O/bench_${N}gpu.err
echo "bench exit $?"; tail -3 $O/bench_${N}gpu.err
python - <<PY
import json
try:
    d=json.loads(open('$O/bench_${N}gpu.json').read().strip().splitlines()[-1])
    print('N=$N cfg1 %.2f Gpairs/s %.1f us/step kernel %.1f us exchange=%s e2e %.2f G (%s)' % (d['value']/1e9, d['ms_per_step']*1e3, d['roofline']['kernel_ms']*1e3, d['config']['exchange'], d['e2e']['value']/1e9, d['e2e'].get('mode')))
    print(' sustained', d.get('sustained')); print(' parity', d.get('parity'))
    print(' e2e host %.2f G dev %.2f G' % (d['e2e']['host_negatives']['value']/1e9, d['e2e']['device_sampled'].get('value',0)/1e9))
    for k,v in d.get('workloads',{}).items():
        print('  %-16s %s' % (k, v.get('error') or '%.4g %s  %.4f ms/step  e2e %.4g' % (v['value'], v['unit'], v['ms_per_step'], (v.get('e2e') or {}).get('value',0))))
except Exception as e: print('no bench line', e)
PY
for comm in ${EXTRA_COMM:-}; do
  timeout 300 $TR bench.py --gpus $N --steps 200 --warmup 20 --workload cfg1 --comm $comm --no-cpu-baseline > $O/bench_${N}gpu_$comm.json 2> $O/bench_${N}gpu_$comm.err
  python -c "import json; d=json.loads(open('$O/bench_${N}gpu_$comm.json').read().strip().splitlines()[-1]); print('N=$N comm=$comm', d['config']['exchange'], d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step  e2e', d['e2e']['value']/1e9)" || tail -5 $O/bench_${N}gpu_$comm.err
done
if [ "${ONE:-1}" = "1" ]; then
  timeout 200 python bench.py --workload cfg1 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 same box', d['value']/1e9, 'Gpairs/s', d['ms_per_step']*1000, 'us/step; sustained', d['sustained']['ms_per_step']*1000)"
fi
if [ "${XCHG:-0}" = "1" ]; then
  timeout 300 $TR scripts/xchg_latency.py > $O/xchg_latency_${N}gpu.log 2>&1; grep "^world" $O/xchg_latency_${N}gpu.log
fi
