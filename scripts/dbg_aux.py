import sys, numpy as np, torch
sys.path.insert(0, '.')
from learning_embeddings_b200.engine import ConeStep, pack_index_block
from learning_embeddings_b200 import hierarchy, ops, _native as N
ethec = hierarchy.ethec()
rng = np.random.default_rng(17)
Nn, B, D = 5, 3000, 10
edges = ethec.closure_edges()
gen = torch.Generator().manual_seed(4)
w = torch.randn(ethec.n, D, generator=gen)
w = w / w.norm(dim=1, keepdim=True) * (0.05 + 0.9 * torch.rand(ethec.n, 1, generator=gen))
tab = w.cuda().clone()
e = ConeStep(tab, "hyp", Nn, B, K=0.1, alpha=0.05, lr=0.01)
for step in range(2):
    sel = rng.integers(0, len(edges), size=B)
    u, v = edges[sel, 0], edges[sel, 1]
    neg_to, neg_from = ethec.sample_negatives(u, v, Nn, rng)
    blk = pack_index_block(u, v, neg_to, neg_from)
    e.step_host(blk, B)
    rows, aux = ops.rows_forward(tab, N.ROWS_HYP_SHELL, 0.1, geom="hyp")
    d = (e.aux != aux)
    print("step", step, "rows equal", torch.equal(e.rows, rows), "aux mismatches", int(d.sum()), "nan", int(torch.isnan(aux).sum()), int(torch.isnan(e.aux).sum()))
    idx = d.nonzero()[:8]
    for i, j in idx.tolist():
        print("  row", i, "col", j, repr(float(e.aux[i, j])), repr(float(aux[i, j])), "A", repr(float(e.aux[i,0])), repr(float(aux[i,0])))
