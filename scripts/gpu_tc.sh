#!/bin/bash
# GPU-box visit for the tensor-core scoring kernel: a tiny direct check first (so a bad descriptor shows up as
# numbers, not as a hang), then the parity tests, then timing.
set -u
mkdir -p gpurun_out
timeout 120 python - > gpurun_out/tc_probe.log 2>&1 <<'PY'
import torch, numpy as np, sys
sys.path.insert(0, '.')
from learning_embeddings_b200 import ops, hierarchy as H
from oracle import cones
torch.manual_seed(0)
dev = torch.device('cuda:0')
for D in (10, 50, 8):
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
timeout 600 python -m pytest tests -m gpu -q -x --no-header -p no:cacheprovider -k "scoring" > gpurun_out/pytest_score.log 2>&1
echo "pytest exit $?" >> gpurun_out/pytest_score.log
tail -15 gpurun_out/pytest_score.log
timeout 300 python scripts/score_bench.py --dims 10,50 --iters 7 --modes topk,matrix_lm,both_lm --engines simt,tc > gpurun_out/score_bench.log 2>&1
cat gpurun_out/score_bench.log
