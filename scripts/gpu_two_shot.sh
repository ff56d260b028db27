set -u
O=gpurun_out/r2p; mkdir -p $O
TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29512"
timeout 300 $TR scripts/xchg_latency.py > $O/xchg_latency_2gpu.log 2>&1; grep "^world" $O/xchg_latency_2gpu.log
timeout 300 $TR bench.py --gpus 2 --workload cfg4 --steps 100 --warmup 10 --pairs 131040 --no-cpu-baseline > $O/cfg4_2gpu.json 2> $O/cfg4_2gpu.err
python -c "
import json; d=json.loads(open('$O/cfg4_2gpu.json').read().strip().splitlines()[-1]); print('cfg4 N=2', d['value']/1e9, 'G', d['ms_per_step']*1e3, 'us', d['config'].get('exchange'), d.get('parity'))"
timeout 200 python bench.py --workload cfg4 --steps 100 --warmup 10 --pairs 131040 --no-cpu-baseline --no-e2e | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('cfg4 N=1', d['value']/1e9, d['ms_per_step']*1e3)"
