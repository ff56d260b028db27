TR="python -m torch.distributed.run --nnodes=1 --nproc-per-node 2 --master-addr 127.0.0.1 --master-port 29513"
for sb in 1 0; do
  if [ $sb = 1 ]; then export LEC_BENCH_SAME_BATCHES=1; else unset LEC_BENCH_SAME_BATCHES; fi
  $TR bench.py --gpus 2 --workload cfg1 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-sustained 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('same_batches=$sb N=2 ms_per_step %.4f kernel %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
done
unset LEC_BENCH_SAME_BATCHES
python bench.py --workload cfg1 --steps 200 --warmup 20 --no-cpu-baseline --no-e2e --no-sustained 2>/dev/null | python -c "import json,sys; d=json.loads(sys.stdin.read().strip().splitlines()[-1]); print('N=1 ms_per_step %.4f kernel %.4f' % (d['ms_per_step'], d['roofline']['kernel_ms']))"
