"""cfg3 scoring micro-benchmark: N synthetic image embeddings x 723 ETHEC labels through lec_score_topk_ex,
outputs preallocated, CUDA-event timing.

    python scripts/score_bench.py [--images 1000000] [--dims 10,50] [--iters 10]
"""
import argparse
import ctypes
import json
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_embeddings_b200 import _native as N, hierarchy as H, ops  # noqa: E402


def ball(n, D, lo, hi, g):
    d = torch.randn(n, D, generator=g)
    return d / d.norm(dim=1, keepdim=True) * (lo + (hi - lo) * torch.rand(n, 1, generator=g))


def make_inputs(n_img, D, dev, seed=0):
    """labels with norm by level U[0.10+0.2l, 0.30+0.2l], images U[0.30, 0.95]."""
    h = H.ethec()
    g = torch.Generator().manual_seed(seed)
    labels = torch.zeros(h.n, D)
    for l in range(4):
        s, e = h.level_start[l], h.level_stop[l]
        labels[s:e] = ball(e - s, D, 0.10 + 0.2 * l, 0.30 + 0.2 * l, g)
    images = ball(n_img, D, 0.30, 0.95, g)
    return h, labels.to(dev), images.to(dev)


WS = {}


def run(lib, geom, prec, labels, images, h, k, scores, layout, idx, val, st, engine="simt"):
    L, D = labels.shape
    nl = len(h.level_start)
    ls = (ctypes.c_int32 * nl)(*h.level_start)
    le = (ctypes.c_int32 * nl)(*h.level_stop)
    if engine == "tc":
        nb = int(lib.lec_score_workspace_bytes(L, D, nl))
        if WS.get("ws") is None or WS["ws"].numel() < nb + 128:
            WS["ws"] = torch.empty(nb + 128, device=labels.device, dtype=torch.uint8)
        ws = WS["ws"][(-WS["ws"].data_ptr()) % 128:]
        N.check(lib.lec_score_topk_tc(ops.GEOM[geom], prec, N._p(labels), L, N._p(images), images.shape[0], D, 0.1,
                                      ctypes.cast(ls, ctypes.c_void_p), ctypes.cast(le, ctypes.c_void_p), nl, k,
                                      N._p(scores), N._p(idx), N._p(val), N._p(ws), nb, st), "lec_score_topk_tc")
        return
    N.check(lib.lec_score_topk_ex(ops.GEOM[geom], prec, N._p(labels), L, N._p(images), images.shape[0], D, 0.1,
                                  ctypes.cast(ls, ctypes.c_void_p), ctypes.cast(le, ctypes.c_void_p), nl, k,
                                  N._p(scores), layout, N._p(idx), N._p(val), st), "lec_score_topk_ex")


def timeit(fn, iters, warm=2):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(iters):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    ts.sort()
    return ts[len(ts) // 2], ts[0]


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--images", type=int, default=1000000)
    ap.add_argument("--dims", default="10,50")
    ap.add_argument("--iters", type=int, default=10)
    ap.add_argument("--geoms", default="hyp")
    ap.add_argument("--precs", default="0")
    ap.add_argument("--modes", default="topk,matrix_lm,both_lm,both_im")
    ap.add_argument("--engines", default="simt")
    args = ap.parse_args()
    dev = torch.device("cuda:0")
    lib = N.lib()
    st = N.stream_ptr(dev)
    peak = 6549.1
    pk = os.path.join(os.path.dirname(os.path.dirname(os.path.abspath(__file__))), "MEASURED_PEAKS.json")
    if os.path.exists(pk):
        peak = float(json.load(open(pk))["hbm_gbs"])
    out = []
    for D in [int(x) for x in args.dims.split(",")]:
        h, labels, images = make_inputs(args.images, D, dev)
        n_img, L = images.shape[0], labels.shape[0]
        idx = torch.empty((n_img, 4, 5), device=dev, dtype=torch.int32)
        val = torch.empty((n_img, 4, 5), device=dev, dtype=torch.float32)
        scores = torch.empty((L, n_img), device=dev, dtype=torch.float32)
        for geom in args.geoms.split(","):
            for prec in [int(x) for x in args.precs.split(",")]:
              for engine in args.engines.split(","):
                for mode in args.modes.split(","):
                    if engine == "tc" and (mode == "both_im" or geom != "hyp" or prec != 0):
                        continue
                    if mode == "topk":
                        fn = lambda: run(lib, geom, prec, labels, images, h, 5, None, 1, idx, val, st, engine)
                        byts = n_img * (4 * D + 160)
                    elif mode == "matrix_lm":
                        fn = lambda: run(lib, geom, prec, labels, images, h, 5, scores, 1, None, None, st, engine)
                        byts = n_img * L * 4 + n_img * 4 * D
                    elif mode == "both_lm":
                        fn = lambda: run(lib, geom, prec, labels, images, h, 5, scores, 1, idx, val, st, engine)
                        byts = n_img * L * 4 + n_img * (4 * D + 160)
                    else:
                        fn = lambda: run(lib, geom, prec, labels, images, h, 5, scores, 0, idx, val, st, engine)
                        byts = n_img * L * 4 + n_img * (4 * D + 160)
                    med, best = timeit(fn, args.iters)
                    rec = {"geom": geom, "D": D, "prec": prec, "mode": mode, "engine": engine, "images": n_img, "labels": L, "ms_median": med,
                           "ms_best": best, "Gscores_per_s": n_img * L / med / 1e6, "algorithmic_GBps": byts / med / 1e6,
                           "frac_of_measured_hbm_peak": byts / med / 1e6 / peak}
                    out.append(rec)
                    print("%s %s D=%d prec=%d %-10s: %8.3f ms (best %8.3f)  %8.1f Gscores/s  %7.1f GB/s algorithmic (%.1f%% of %.0f)"
                          % (engine, geom, D, prec, mode, med, best, rec["Gscores_per_s"], rec["algorithmic_GBps"],
                             100 * rec["frac_of_measured_hbm_peak"], peak), flush=True)
    os.makedirs("gpurun_out", exist_ok=True)
    json.dump(out, open("gpurun_out/score_bench.json", "w"), indent=1)


main()
