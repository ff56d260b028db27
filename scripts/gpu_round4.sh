#!/bin/bash
# v2 tensor-core scoring (three bilinear forms from the MMA), group-split cost model, native sampler: probe first,
# then tests, timing, benches and captures.
set -u
mkdir -p gpurun_out
timeout 120 python - > gpurun_out/tc_probe.log 2>&1 <<'PY'
import torch, numpy as np, sys
sys.path.insert(0, '.')
from learning_embeddings_b200 import ops, hierarchy as H
from oracle import cones
dev = torch.device('cuda:0')
for D in (10, 50, 8, 62, 100):
    h = H.ethec()
    g = torch.Generator().manual_seed(1)
    def ball(n, lo, hi):
        d = torch.randn(n, D, generator=g); return d / d.norm(dim=1, keepdim=True) * (lo + (hi - lo) * torch.rand(n, 1, generator=g))
    labels = ball(h.n, 0.1, 0.9); images = ball(300, 0.3, 0.95)
    ref = cones.score_matrix("hyp", labels.double(), images.double(), 0.1).numpy()
    for eng in ("simt", "tc"):
        idx, val, sc = ops.score_topk(labels.to(dev), images.to(dev), "hyp", 0.1, h.level_start, h.level_stop, k=5, want_scores=True, engine=eng)
        torch.cuda.synchronize()
        err = np.abs(sc.cpu().numpy() - ref)
        print("D", D, eng, "max abs err vs fp64 oracle %.3e" % np.nanmax(err), "mean %.3e" % np.nanmean(err), "nan", int(np.isnan(sc.cpu().numpy()).sum()), flush=True)
        ridx, rval = cones.topk_per_level(sc.cpu(), h.level_start, h.level_stop, 5)
        print("   topk values equal own matrix:", bool(np.array_equal(val.cpu().numpy(), rval.numpy())), flush=True)
        i2, v2, _ = ops.score_topk(labels.to(dev), images.to(dev), "hyp", 0.1, h.level_start, h.level_stop, k=5, engine=eng)
        print("   topk-only equals both-mode:", bool(torch.equal(v2, val)), bool(torch.equal(i2, idx)), flush=True)
PY
echo "probe exit $?" >> gpurun_out/tc_probe.log
cat gpurun_out/tc_probe.log
timeout 900 python -m pytest tests -m gpu -q --no-header -p no:cacheprovider > gpurun_out/pytest_gpu.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_gpu.log
tail -25 gpurun_out/pytest_gpu.log
timeout 300 python scripts/score_bench.py --dims 10,50 --iters 7 --modes topk,matrix_lm,both_lm --engines simt,tc > gpurun_out/score_bench.log 2>&1
cat gpurun_out/score_bench.log
timeout 300 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_cfg1.json 2> gpurun_out/bench_cfg1.err
timeout 300 python bench.py --workload cfg3 --steps 50 --warmup 5 > gpurun_out/bench_cfg3_d10.json 2> gpurun_out/bench_cfg3_d10.err
timeout 300 python bench.py --workload cfg3 --dim 50 --steps 50 --warmup 5 > gpurun_out/bench_cfg3_d50.json 2> gpurun_out/bench_cfg3_d50.err
timeout 300 python bench.py --workload cfg3 --score-mode matrix --steps 50 --warmup 5 --no-e2e --no-cpu-baseline > gpurun_out/bench_cfg3_d10_matrix.json 2> gpurun_out/bench_cfg3_d10_matrix.err
for s in 1 2 3 5 9 13 25; do
  LEC_GROUP_SPLIT=$s timeout 200 python bench.py --workload cfg4 --steps 100 --warmup 10 --pairs 131040 --no-cpu-baseline --no-e2e > gpurun_out/bench_cfg4_split$s.json 2> gpurun_out/bench_cfg4_split$s.err
done
timeout 300 python bench.py --workload cfg4 --steps 100 --warmup 10 --pairs 131040 > gpurun_out/bench_cfg4.json 2> gpurun_out/bench_cfg4.err
timeout 300 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches_cfg1.csv \
   python bench.py --steps 6 --warmup 3 --no-cpu-baseline --no-e2e > gpurun_out/ncu_cfg1.log 2>&1
for D in 10 50; do
  timeout 200 ncu --set full --clock-control none --import-source on -k regex:score_mma_kernel -s 3 -c 3 -o /tmp/prof_tc_d$D \
      python scripts/score_bench.py --images 303104 --dims $D --iters 1 --modes topk,matrix_lm,both_lm --engines tc > gpurun_out/ncu_tc_d$D.log 2>&1
  ncu -i /tmp/prof_tc_d$D.ncu-rep --page raw --csv > gpurun_out/tc_d${D}_raw.csv 2>/dev/null
  ncu -i /tmp/prof_tc_d$D.ncu-rep --page source --csv > gpurun_out/tc_d${D}_source.csv 2>/dev/null
done
du -sh gpurun_out
for w in cfg1 cfg3_d10 cfg3_d50 cfg3_d10_matrix cfg4; do tail -2 gpurun_out/bench_$w.err; cat gpurun_out/bench_$w.json; done
for s in 1 2 3 5 9 13 25; do python -c "import json;d=json.load(open('gpurun_out/bench_cfg4_split$s.json'));print('split',$s,d['value'],d['roofline']['kernel_ms'])"; done
