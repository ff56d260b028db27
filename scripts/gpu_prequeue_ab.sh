TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
for pq in 1 0 1 0; do
  LEC_BENCH_PREQUEUE=$pq $TR bench.py --gpus 2 --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-sustained 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prequeue=$pq N=2 ms_per_step %.4f kernel %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
done
for pq in 1 0; do
  LEC_BENCH_PREQUEUE=$pq python bench.py --workload cfg1 --steps 20 --warmup 5 --no-cpu-baseline --no-e2e --no-sustained 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('prequeue=$pq N=1 ms_per_step %.4f kernel %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
done
