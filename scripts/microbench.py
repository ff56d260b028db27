"""Kernel-level timing sweeps (run on the GPU box).  Prints one line per variant."""
import sys, os, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_embeddings_b200 import ops, _native as N, hierarchy as H

dev = torch.device("cuda")


def timeit(fn, iters=20, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(iters):
        fn()
    b.record()
    torch.cuda.synchronize()
    return a.elapsed_time(b) / iters * 1e3  # us


def ball(n, D, lo, hi, g):
    d = torch.randn(n, D, generator=g)
    return d / d.norm(dim=1, keepdim=True) * (lo + (hi - lo) * torch.rand(n, 1, generator=g))


def main():
    g = torch.Generator().manual_seed(0)
    h = H.ethec()
    edges = h.closure_edges()
    rng = np.random.default_rng(0)
    for D, Nn, B in ((10, 5, 190650), (50, 25, 41120), (2, 5, 190650)):
        sel = rng.integers(0, len(edges), size=B)
        u, v = edges[sel, 0], edges[sel, 1]
        nt, nf = h.sample_negatives(u, v, Nn, rng)
        P = B * (1 + 2 * Nn)
        W = ball(h.n, D, 0.1, 0.15, g).to(dev)
        rows, aux_h = ops.rows_forward(W, N.ROWS_HYP_SHELL, 0.1, 'hyp')
        _, aux_e = ops.rows_forward(W, N.ROWS_HYP_SHELL, 0.01, 'euc')
        AUX = {'hyp': aux_h, 'euc': aux_e, 'oe': None}
        ud, vd, ntd, nfd = (torch.from_numpy(np.ascontiguousarray(a, dtype=np.int32)).to(dev) for a in (u, v, nt, nf))
        grad = torch.zeros((int(os.environ.get('LEC_R', ops.default_replicas(*rows.shape))),) + tuple(rows.shape), device=dev)
        print('replicas', grad.shape[0])
        loss = torch.zeros(1, dtype=torch.float64, device=dev)
        Ep = torch.empty(B, device=dev); En = torch.empty(B, 2 * Nn, device=dev)
        for geom in ("hyp", "euc", "oe"):
            K = {"hyp": 0.1, "euc": 0.01, "oe": 0.0}[geom]
            for prec in ((0, 1) if geom == "hyp" else (0,)):
                for gr in (None, grad):
                    t = timeit(lambda: ops.pairs_grouped_raw(geom, rows, AUX[geom], D, ud, vd, ntd, nfd, Nn, K, 0.05, grad_rows=gr,
                                                             loss_out=loss, precision=prec, E_pos=Ep, E_neg=En))
                    print("grouped D=%d N=%d %s prec=%d grad=%d: %8.1f us  %6.2f Gpairs/s" % (D, Nn, geom, prec, gr is not None, t, P / t / 1e3))
        # flat kernel on the expanded list
        fi = torch.cat([ud, ud[:, None].expand(B, Nn).reshape(-1), nfd.reshape(-1)])
        ti = torch.cat([vd, ntd.reshape(-1), vd[:, None].expand(B, Nn).reshape(-1)])
        isp = torch.cat([torch.ones(B), torch.zeros(2 * Nn * B)]).to(torch.uint8).to(dev)
        E = torch.empty(P, device=dev)
        for gr in (None, grad):
            t = timeit(lambda: ops.pairs_flat_raw("hyp", rows, aux_h, D, fi, ti, 0.1, 0.05, is_pos=isp, grad_rows=gr, loss_out=loss, precision=1, E_out=E))
            print("flat    D=%d hyp prec=1 grad=%d: %8.1f us  %6.2f Gpairs/s" % (D, gr is not None, t, P / t / 1e3))
        # dense energy
        x = rows[fi.long(), :D].contiguous(); y = rows[ti.long(), :D].contiguous()
        t = timeit(lambda: ops.energy(x, y, "hyp", 0.1, 1))
        print("dense   D=%d hyp fwd: %8.1f us  %6.2f Gpairs/s  (%.0f GB/s)" % (D, t, P / t / 1e3, P * (8 * D + 4) / t / 1e3))
    # scoring
    for D in (10, 50):
        L, n_img = 723, 1 << 18
        labels = torch.zeros(L, D)
        for l in range(4):
            s, e = h.level_start[l], h.level_stop[l]
            labels[s:e] = ball(e - s, D, 0.10 + 0.2 * l, 0.30 + 0.2 * l, g)
        images = ball(n_img, D, 0.30, 0.95, g).to(dev)
        labels = labels.to(dev)
        for prec in (0, 1):
            t = timeit(lambda: ops.score_topk(labels, images, "hyp", 0.1, h.level_start, h.level_stop, 5, precision=prec), iters=5, warm=2)
            print("score D=%d prec=%d topk-only: %9.1f us  %7.2f Gscores/s" % (D, prec, t, n_img * L / t / 1e3))
        t = timeit(lambda: ops.score_topk(labels, images, "hyp", 0.1, h.level_start, h.level_stop, 5, want_scores=True), iters=5, warm=2)
        print("score D=%d prec=0 +matrix  : %9.1f us  %7.2f Gscores/s" % (D, t, n_img * L / t / 1e3))
    # rsgd
    for n, D in ((723, 10), (82115, 50)):
        W = ball(n, D, 0.1, 0.9, g).to(dev); gr = torch.randn(n, D, device=dev)
        t = timeit(lambda: ops.rsgd_update_(W, gr, 1e-3, 0.099, write_rescaled_grad=False))
        print("rsgd n=%d D=%d: %7.1f us  %.0f GB/s" % (n, D, t, n * D * 12 / t / 1e3))


main()
