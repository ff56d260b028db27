import sys, os
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "tests"))
import numpy as np, torch
from conftest import load_golden
from oracle import cones
from learning_embeddings_b200 import ops
def _ball(gen, n, D, lo, hi):
    d = torch.randn(n, D, generator=gen)
    return d / d.norm(dim=1, keepdim=True) * (lo + (hi - lo) * torch.rand(n, 1, generator=gen))
gen = torch.Generator().manual_seed(3)
h = load_golden("ethec_hierarchy")
L, D, n_img = 723, 10, 40000
labels = torch.zeros(L, D)
for l in range(4):
    s, e = int(h["level_start"][l]), int(h["level_stop"][l])
    labels[s:e] = _ball(gen, e - s, D, 0.10 + 0.2 * l, 0.30 + 0.2 * l)
images = _ball(gen, n_img, D, 0.30, 0.95)
idx, val, scores = ops.score_topk(labels.cuda(), images.cuda(), "hyp", 0.1, h["level_start"], h["level_stop"], k=5, want_scores=True)
ridx, rval = cones.topk_per_level(scores.cpu(), h["level_start"], h["level_stop"], 5)
print("val equal:", torch.equal(val.cpu(), rval))
ties = (rval[..., 1:] == rval[..., :-1]).any(dim=-1)
bad = (idx.cpu() != ridx.int()).any(dim=-1) & ~ties
print("bad rows:", int(bad.sum()), "of", bad.numel())
w = bad.nonzero()[:8]
for i, lv in w.tolist():
    print(i, lv, "got", idx[i, lv].tolist(), "ref", ridx[i, lv].tolist(), "val", val[i, lv].tolist())
    sc = scores[i].cpu()
    print("   E at got:", [float(sc[j]) for j in idx[i, lv].tolist()], " E at ref:", [float(sc[j]) for j in ridx[i, lv].tolist()])
