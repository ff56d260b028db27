#!/usr/bin/env python
"""Benchmark of the entailment-cone hot path (BASELINE.json metric: cone pairs/s fwd+bwd).

    python bench.py [--gpus N] [--steps K] [--warmup W] [--impl b200|reference] [--workload cfg1|cfg2|cfg3|cfg4]

One "step" = one pass of the hot path over one batch: row transform -> fused cone loss fwd+bwd over
B*(1+2N) pairs -> (N>1: NCCL all-reduce of the label-table gradient) -> Riemannian SGD update of the
whole table.  Workload cfg1 (BASELINE.json configs[1]): Poincare cones, label-only, ETHEC 723-node
hierarchy, D=10, RSGD, 10 negatives per edge; the 1 974 closure edges are tiled to --pairs pairs per
GPU per step and every step uses a different pre-sampled batch out of a rotation larger than L2.
Prints ONE JSON line (rank 0).
"""
import argparse
import json
import os
import sys
import threading
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

METRIC = "cone pairs/s fwd+bwd"
UNIT = "pairs/s"


def parse():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=200)
    ap.add_argument("--warmup", type=int, default=20)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="cfg1", choices=["cfg0", "cfg1", "cfg2", "cfg3", "cfg4"])
    ap.add_argument("--dim", type=int, default=10, help="cfg3: embedding dimension (10 or 50)")
    ap.add_argument("--images", type=int, default=1000000, help="cfg3: images per GPU per step")
    ap.add_argument("--score-mode", default="both", choices=["both", "matrix", "topk"],
                    help="cfg3: what a step writes: per-level top-5 + the full [L, N] energy matrix, or one of them")
    ap.add_argument("--engine", default="auto", choices=["auto", "tc", "simt"], help="cfg3: scoring engine")
    ap.add_argument("--pairs", type=int, default=1 << 21, help="pairs per GPU per step (rounded to whole groups)")
    ap.add_argument("--precision", type=int, default=None, help="0 fp32 core, 1 fp64 core (default: per workload)")
    ap.add_argument("--rotation", type=int, default=0, help="distinct batches to rotate through (0 = enough to exceed L2)")
    ap.add_argument("--comm", default="auto", choices=["auto", "p2p", "nccl"],
                    help="multi-GPU gradient exchange: fused peer-memory all-reduce+update, or NCCL")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    return ap.parse_args()


# ------------------------------------------------------------------------------------------------
# workload
# ------------------------------------------------------------------------------------------------
def workload_spec(name):
    if name == "cfg0":
        # BASELINE.json configs[0] (the reference's own CPU case: D=2, Embedder K=3, alpha=0.05, N=5, Adam,
        # order_embeddings.py:1364) with the ETHEC closure edges tiled to the cfg1 batch size, so that the Euclidean
        # pair kernel is measured at a size that fills the GPU
        return dict(name="cfg0: Euclidean cones, label-only, ETHEC 723-node hierarchy, D=2, Adam, 10 negatives/edge",
                    geom="euc", D=2, n_neg=5, K=3.0, alpha=0.05, lr=1e-3, tree="ethec")
    if name == "cfg1":
        return dict(name="cfg1: Poincare cones, label-only, ETHEC 723-node hierarchy, D=10, RSGD, 10 negatives/edge",
                    geom="hyp", D=10, n_neg=5, K=0.1, alpha=0.05, lr=1e-3, tree="ethec")
    return dict(name="cfg4: Poincare cones, 82115-node random tree, D=50, RSGD, 50 negatives/edge",
                geom="hyp", D=50, n_neg=25, K=0.1, alpha=0.05, lr=1e-3, tree="random82k")


def build_hierarchy(spec):
    from learning_embeddings_b200 import hierarchy as H
    if spec["tree"] == "ethec":
        return H.ethec()
    return H.random_tree(82115, 1.0831, seed=0)


def init_table(n, D, K, seed):
    """order_embeddings_h.py:198-203: N(0,1) direction, norm r_in + U[0, 0.05)."""
    from learning_embeddings_b200.criterion import inner_radius
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(n, D, generator=g)
    return (inner_radius(K) + 0.05 * torch.rand(n, 1, generator=g)) * w / w.norm(dim=1, keepdim=True)


def make_batches(h, spec, groups, count, seed):
    """`count` host index blocks of `groups` positives each (closure edges tiled in shuffled order)."""
    from learning_embeddings_b200.engine import pack_index_block, index_dtype_for
    rng = np.random.default_rng(seed)
    dt = index_dtype_for(h.n)
    edges = h.closure_edges()
    out = []
    for _ in range(count):
        sel = rng.integers(0, len(edges), size=groups)
        u, v = edges[sel, 0], edges[sel, 1]
        neg_to, neg_from = h.sample_negatives(u, v, spec["n_neg"], rng)
        out.append(pack_index_block(u, v, neg_to, neg_from, dtype=dt))
    return out


# ------------------------------------------------------------------------------------------------
# clocks
# ------------------------------------------------------------------------------------------------
class ClockSampler(threading.Thread):
    REASONS = {0x4: "sw_power_cap", 0x8: "hw_slowdown", 0x20: "sw_thermal_slowdown", 0x40: "hw_thermal_slowdown",
               0x80: "hw_power_brake_slowdown", 0x2: "applications_clocks_setting"}

    def __init__(self, index):
        super().__init__(daemon=True)
        self.samples, self.reasons, self.max_mhz = [], set(), None
        self._stop_evt = threading.Event()
        self.active = threading.Event()
        try:
            import pynvml
            pynvml.nvmlInit()
            self.nv = pynvml
            self.h = pynvml.nvmlDeviceGetHandleByIndex(index)
            self.max_mhz = int(pynvml.nvmlDeviceGetMaxClockInfo(self.h, pynvml.NVML_CLOCK_SM))
        except Exception:
            self.nv = None

    def run(self):
        if self.nv is None:
            return
        while not self._stop_evt.is_set():
            if self.active.is_set():
                try:
                    self.samples.append(int(self.nv.nvmlDeviceGetClockInfo(self.h, self.nv.NVML_CLOCK_SM)))
                    r = int(self.nv.nvmlDeviceGetCurrentClocksThrottleReasons(self.h))
                    for bit, name in self.REASONS.items():
                        if r & bit:
                            self.reasons.add(name)
                except Exception:
                    pass
            time.sleep(0.002)

    def stop(self):
        self._stop_evt.set()

    def summary(self):
        return {"sm_mhz": (float(np.median(self.samples)) if self.samples else None), "sm_max_mhz": self.max_mhz,
                "reasons": sorted(self.reasons), "samples": len(self.samples)}


# ------------------------------------------------------------------------------------------------
# CPU reference arm (oracle port of the reference's PyTorch-on-CPU step)
# ------------------------------------------------------------------------------------------------
def cpu_step_runner(spec, table, blk, B):
    """Returns a closure running one full reference-style step on the host cores: gather + row transform
    + E_operator + hinge + backward (autograd) + RSGD table update, all in torch fp32 on CPU."""
    from oracle import cones
    Nn = spec["n_neg"]
    b = torch.from_numpy(blk[:B * (2 + 2 * Nn)].numpy().astype(np.int64))
    u, v = b[:B], b[B:2 * B]
    neg_to = b[2 * B:2 * B + B * Nn].view(B, Nn)
    neg_from = b[2 * B + B * Nn:].view(B, Nn)
    nf = torch.cat([u[:, None].expand(B, Nn), neg_from], 1).reshape(-1)
    nt = torch.cat([neg_to, v[:, None].expand(B, Nn)], 1).reshape(-1)
    W = table.clone()
    if spec["geom"] == "euc":
        # order_embeddings.py: Embedder.soft_clip rows, EucConesLoss, torch Adam on the table
        P = torch.nn.Parameter(W, requires_grad=False)
        opt = torch.optim.Adam([P], lr=spec["lr"])

        def run_euc():
            r = cones.label_step("euc", P.data, cones.ROW_EUC_SOFTCLIP, spec["K"], spec["alpha"], u, v, nf, nt)
            P.grad = r["gW"]
            opt.step()
            return float(r["loss"])

        return run_euc
    r_in = cones.inner_radius(spec["K"])

    def run():
        nonlocal W
        r = cones.label_step(spec["geom"], W, cones.ROW_HYP_SHELL, spec["K"], spec["alpha"], u, v, nf, nt)
        _, W = cones.rsgd_step(W, r["gW"], spec["lr"], r_in)
        return float(r["loss"])

    return run


def time_cpu(spec, table, blk, groups, steps, warmup, budget_s=25.0):
    """Bounded sample: `groups` positives (x(1+2N) pairs) per CPU step."""
    torch.set_num_threads(os.cpu_count() or 1)
    run = cpu_step_runner(spec, table, blk, groups)
    for _ in range(warmup):
        run()
    times = []
    t_all = time.perf_counter()
    for _ in range(steps):
        t0 = time.perf_counter()
        run()
        times.append(time.perf_counter() - t0)
        if time.perf_counter() - t_all > budget_s:
            break
    pairs = groups * (1 + 2 * spec["n_neg"])
    return pairs / float(np.mean(times)), float(np.mean(times)), len(times), pairs


# ------------------------------------------------------------------------------------------------
# cfg2: joint image+label Euclidean cones (BASELINE.json configs[2])
# ------------------------------------------------------------------------------------------------
CFG2 = dict(name="cfg2: joint image+label Euclidean cones, 2048-d features -> FeatNet -> D=10, ETHEC labels, "
                 "23831 positives x (1+10) = 262141 pairs/step, 16384 distinct images/step, Adam",
            geom="euc", D=10, n_neg=5, K=3.0, alpha=1.0, lr=1e-3, lr_labels=0.1, B=23831, m=16384, F=2048, pool=65536)


def make_joint_batches(h, leaf_of_img, B, Nn, m, count, rng, idx_dtype):
    """`count` steps of (img_sel int64[m] rows of the feature pool, packed index block).  Positive (u, v): u a
    label; v one of the step's images w.p. 0.99 (u = its leaf label or one of that leaf's ancestors) else a
    child label (closure edge).  Negatives uniform over the step's mixed node set [labels ; images] minus the
    reference's excluded sets (oe.py:846-863: descendants of u when the child is corrupted, ancestors of v
    when the parent is corrupted).  Node ids: < n labels, >= n image slot (id - n) of the step."""
    from learning_embeddings_b200.engine import pack_index_block
    n = h.n
    edges = h.closure_edges()
    depth = h.depth
    out = []
    for _ in range(count):
        sel = rng.choice(len(leaf_of_img), size=m, replace=False)
        leaf = leaf_of_img[sel]                       # leaf label of every image slot
        is_img = rng.random(B) < 0.99
        slot = rng.integers(0, m, size=B)
        up = rng.integers(0, 4, size=B)               # how many levels above the leaf the parent sits
        u = leaf[slot].copy()
        for _k in range(3):
            step_up = (up > _k) & (h.parents[u] >= 0)
            u[step_up] = h.parents[u[step_up]]
        e = edges[rng.integers(0, len(edges), size=B)]
        u = np.where(is_img, u, e[:, 0])
        v = np.where(is_img, n + slot, e[:, 1])
        v_lab = np.where(is_img, leaf[slot], e[:, 1])  # the label whose ancestors are v's ancestors
        uu, vv, vl = (np.repeat(a[:, None], Nn, 1) for a in (u, v, v_lab))
        vimg = np.repeat(is_img[:, None], Nn, 1)

        def lab_of(x):                                 # label a node hangs under (itself for a label)
            return np.where(x >= n, leaf[np.clip(x - n, 0, m - 1)], x)

        neg_to = rng.integers(0, n + m, size=(B, Nn))
        while True:                                    # corrupted child: not u, not below u
            cl = lab_of(neg_to)
            bad = (neg_to == uu) | h.is_descendant(uu, cl) | ((neg_to >= n) & (cl == uu))
            k = int(bad.sum())
            if k == 0:
                break
            neg_to[bad] = rng.integers(0, n + m, size=k)
        neg_from = rng.integers(0, n + m, size=(B, Nn))
        while True:                                    # corrupted parent: not v, not an ancestor of v
            lab = neg_from < n
            bad = (neg_from == vv) | (lab & (h.is_descendant(np.where(lab, neg_from, 0), vl) | (vimg & (neg_from == vl))))
            k = int(bad.sum())
            if k == 0:
                break
            neg_from[bad] = rng.integers(0, n + m, size=k)
        blk = pack_index_block(u, v, neg_to, neg_from, dtype=idx_dtype)
        sel_t = torch.from_numpy(sel.astype(np.int64))
        out.append((sel_t.pin_memory() if torch.cuda.is_available() else sel_t, blk))
    return out


def cfg2_cpu_runner(c, table, fw, fb, feats, sel, blk, B):
    """The same joint step on the host cores with the oracle's torch port (oe.py forward + autograd + Adam)."""
    from oracle import cones
    Nn, n = c["n_neg"], table.shape[0]
    b = torch.from_numpy(blk[:B * (2 + 2 * Nn)].numpy().astype(np.int64))
    u, v = b[:B], b[B:2 * B]
    neg_to = b[2 * B:2 * B + B * Nn].view(B, Nn)
    neg_from = b[2 * B + B * Nn:].view(B, Nn)
    nf = torch.cat([u[:, None].expand(B, Nn), neg_from], 1).reshape(-1)
    nt = torch.cat([neg_to, v[:, None].expand(B, Nn)], 1).reshape(-1)
    W = table.clone().requires_grad_(True)
    w1 = fw.clone().requires_grad_(True)
    b1 = fb.clone().requires_grad_(True)
    opt = torch.optim.Adam([{"params": [W], "lr": c["lr_labels"]}, {"params": [w1, b1]}], lr=c["lr"])   # oe.py:1356-1357, :1714
    X = feats[sel]

    def run():
        opt.zero_grad()
        rows = torch.cat([cones.apply_rows(cones.ROW_EUC_SOFTCLIP, W, c["K"]),
                          cones.apply_rows(cones.ROW_EUC_SOFTCLIP, X @ w1.t() + b1, c["K"])], 0)
        E_pos = cones.energy(c["geom"], rows[u], rows[v], c["K"])
        E_neg = cones.energy(c["geom"], rows[nf], rows[nt], c["K"])
        loss = E_pos.sum() + (c["alpha"] - E_neg).clamp(min=0).sum()
        loss.backward()
        opt.step()
        return float(loss)

    return run


def run_cfg2(args):
    c = CFG2
    Nn, D, B, m = c["n_neg"], c["D"], c["B"], c["m"]
    pairs_per_step = B * (1 + 2 * Nn)
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    from learning_embeddings_b200 import hierarchy as H
    from learning_embeddings_b200.engine import index_dtype_for
    h = H.ethec()
    idx_dt = index_dtype_for(h.n + m)
    idx_bytes = np.dtype(idx_dt).itemsize
    cfg = {"workload": c["name"], "pairs_per_gpu_per_step": pairs_per_step, "positives_per_gpu_per_step": B,
           "images_per_gpu_per_step": m, "feature_dim": c["F"], "dim": D, "negatives_per_edge": 2 * Nn,
           "table_rows": int(h.n), "update": "adam on table (lr 0.1) + fc1 (lr 1e-3), inside the fused update kernels",
           "scalar_core": "fp32",
           "index_dtype": np.dtype(idx_dt).name, "feature_pool_images": c["pool"],
           "parallelism": "dp%d (pairs and images sharded, table + fc1 replicated)" % world}
    g = torch.Generator().manual_seed(0)
    table0 = torch.randn(h.n, D, generator=g)
    lin = torch.nn.Linear(c["F"], D)
    with torch.no_grad():
        fw0, fb0 = lin.weight.detach().clone(), lin.bias.detach().clone()
    rng = np.random.default_rng(7 + rank)
    leaves = np.arange(h.level_start[-1], h.level_stop[-1])
    leaf_of_img = leaves[rng.integers(0, len(leaves), size=c["pool"])]

    if args.impl == "reference":
        if rank != 0:
            return
        torch.set_num_threads(os.cpu_count() or 1)
        pool_cpu = 4 * m  # a bounded feature pool for the host run
        feats = torch.relu(torch.randn(pool_cpu, c["F"], generator=g))
        sel, blk = make_joint_batches(h, leaf_of_img[:pool_cpu], B, Nn, m, 1, rng, idx_dt)[0]
        run = cfg2_cpu_runner(c, table0, fw0, fb0, feats, sel, blk, B)
        for _ in range(min(2, args.warmup)):
            run()
        ts = []
        t_all = time.perf_counter()
        for _ in range(args.steps):
            t0 = time.perf_counter()
            run()
            ts.append(time.perf_counter() - t0)
            if time.perf_counter() - t_all > 150.0:
                break
        v = pairs_per_step / float(np.mean(ts))
        cfg["parallelism"] = "host cores only"
        print(json.dumps({
            "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": len(ts),
            "warmup": min(2, args.warmup), "ms_per_step": float(np.mean(ts)) * 1e3, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
                             "sample": "the full cfg2 step (%d pairs, %d images), %d timed steps" % (pairs_per_step, m, len(ts))},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD
    from learning_embeddings_b200 import _native
    from learning_embeddings_b200.engine import JointConeStep
    gd = torch.Generator(device=dev).manual_seed(1 + rank)
    feats = torch.relu(torch.randn(c["pool"], c["F"], generator=gd, device=dev))   # 537 MB, staged once
    rotation = args.rotation or 8
    batches = make_joint_batches(h, leaf_of_img, B, Nn, m, rotation, rng, idx_dt)
    dev_batches = [(s_.to(dev), b_.to(dev)) for s_, b_ in batches]
    cfg["l2_policy"] = "every step gathers %d of %d feature rows (%.0f MB read, > 126 MB L2); %d index batches rotate" % (
        m, c["pool"], m * c["F"] * 4 / 1e6, rotation)
    table, fw, fb = table0.to(dev).clone(), fw0.to(dev).clone(), fb0.to(dev).clone()
    eng = JointConeStep(table, fw, fb, feats, c["geom"], Nn, B, m, K=c["K"], alpha=c["alpha"], lr=c["lr_labels"],
                        lr_fc=c["lr"], precision=0, process_group=pg)
    cfg["exchange"] = ("packet all-reduce over peer memory inside the two update kernels (label table 35 KB, fc1 82 KB)"
                       if world > 1 else "none (1 GPU)")

    def dev_step(i):
        s_, b_ = dev_batches[i % rotation]
        eng.step_device(s_, *eng._split(b_, B))

    def sync_all():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    sampler = ClockSampler(local_rank)
    sampler.start()
    for i in range(max(3, args.warmup)):
        dev_step(i)
    sync_all()
    launches0 = _native.launch_count()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active.set()
    sync_all()
    t0.record()
    for i in range(args.steps):
        # the pair kernel is bracketed by CUDA events on every 4th step only: two event records per step sit between
        # back-to-back launches and cost more than they measure
        eng.kernel_events = kev[i] if i % 4 == 0 else None
        dev_step(i)
    t1.record()
    sync_all()
    sampler.active.clear()
    eng.kernel_events = None
    lec_launches = _native.launch_count() - launches0
    elapsed_ms = t0.elapsed_time(t1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for i, (a, b) in enumerate(kev) if i % 4 == 0]))

    def max_ranks(ms):
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            return float(tt.item())
        return ms

    elapsed_ms = max_ranks(elapsed_ms)
    value = world * pairs_per_step * args.steps / (elapsed_ms * 1e-3)
    e2e = None
    if not args.no_e2e:
        for i in range(3):
            eng.step_host(*batches[i % rotation], B)
        sync_all()
        sampler.active.set()
        w0 = time.perf_counter()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for i in range(args.steps):
            eng.step_host(*batches[i % rotation], B)
        e1.record()
        sync_all()
        sampler.active.clear()
        e_ms = max_ranks(max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3))
        e2e = {"value": world * pairs_per_step * args.steps / (e_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": B * (2 + 2 * Nn) * idx_bytes + m * 8, "d2h_bytes_per_step": 8,
               "ms_per_step": e_ms / args.steps,
               "api": "JointConeStep.step_host (features device-resident; per step: image selection + index block in, loss out)"}
    sampler.stop()
    final_loss = float(eng.loss.item())
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_per_pair = 24 + 16 * D
    achieved = pairs_per_step * bytes_per_pair / (kernel_ms * 1e-3) / 1e9
    step_ms = elapsed_ms / args.steps
    roofline = {"bound": "hbm", "kernel": "pairs_grouped_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": None, "peak_source": peak_src, "algorithmic_bytes_per_pair": bytes_per_pair,
                "kernel_ms": kernel_ms, "kernel_share_of_step": kernel_ms / step_ms,
                "step_floor_ms": 2 * m * c["F"] * 4 / (peak * 1e9) * 1e3,
                "step_hbm_frac": 2 * m * c["F"] * 4 / (peak * 1e9) * 1e3 / step_ms,
                "note": "the step is bound by the two passes over the %d x %d fp32 gathered feature rows (lec_featnet_fwd, "
                        "lec_featnet_wgrad: %.0f MB each); step_floor_ms = those bytes at the measured HBM peak, "
                        "step_hbm_frac = floor / measured step" % (m, c["F"], m * c["F"] * 4 / 1e6)}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        torch.set_num_threads(os.cpu_count() or 1)
        run = cfg2_cpu_runner(c, table0, fw0, fb0, feats[:4 * m].cpu(), torch.from_numpy(np.arange(m)), batches[0][1], B)
        run()
        ts = []
        t_all = time.perf_counter()
        while len(ts) < 20 and time.perf_counter() - t_all < 20.0:
            t0_ = time.perf_counter()
            run()
            ts.append(time.perf_counter() - t0_)
        cpu = {"value": pairs_per_step / float(np.mean(ts)), "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "the full cfg2 step (%d pairs, %d images) x %d, torch fp32 on all host threads" % (pairs_per_step, m, len(ts))}
    print(json.dumps({"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
                      "warmup": max(3, args.warmup), "ms_per_step": step_ms, "higher_is_better": True, "scaling": "weak",
                      "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": sampler.summary(),
                      "e2e": e2e, "gpu_launches": int(lec_launches), "roofline": roofline, "cpu_baseline": cpu,
                      "final_loss": final_loss}))
    if world > 1:
        torch.distributed.destroy_process_group()


# ------------------------------------------------------------------------------------------------
# cfg3: all-pairs image x label scoring (BASELINE.json metric "image x label scores/s")
# ------------------------------------------------------------------------------------------------
def cfg3_inputs(n_img, D, seed):
    """images uniform direction with |y| ~ U[0.30, 0.95]; ETHEC labels, norm by level
    U[0.10 + 0.2 l, 0.30 + 0.2 l]."""
    from learning_embeddings_b200 import hierarchy as H
    h = H.ethec()
    g = torch.Generator().manual_seed(seed)

    def ball(n, lo, hi):
        d = torch.randn(n, D, generator=g)
        return d / d.norm(dim=1, keepdim=True) * (lo + (hi - lo) * torch.rand(n, 1, generator=g))
    labels = torch.zeros(h.n, D)
    for l in range(len(h.level_start)):
        s, e = h.level_start[l], h.level_stop[l]
        labels[s:e] = ball(e - s, 0.10 + 0.2 * l, 0.30 + 0.2 * l)
    return h, labels, ball(n_img, 0.30, 0.95)


def cfg3_cpu(h, labels, images, K, steps, warmup, budget_s):
    """The reference's scoring (oe_h.py:2018-2036: E(label, image) for every pair, per-level topk(5, smallest)) as
    restated by the oracle, on all host threads, over a bounded sample of the images."""
    from oracle import cones
    torch.set_num_threads(os.cpu_count() or 1)

    def run():
        E = cones.score_matrix("hyp", labels, images, K)
        return cones.topk_per_level(E, h.level_start, h.level_stop, 5)
    for _ in range(min(1, warmup)):
        run()
    ts = []
    t_all = time.perf_counter()
    while len(ts) < steps and (not ts or time.perf_counter() - t_all < budget_s):
        t0 = time.perf_counter()
        run()
        ts.append(time.perf_counter() - t0)
    return images.shape[0] * labels.shape[0] / float(np.mean(ts)), float(np.mean(ts)), len(ts)


def run_cfg3(args):
    metric, unit = "image x label scores/s", "scores/s"
    D, n_img, K, k = args.dim, args.images, 0.1, 5
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    want_matrix, want_topk = args.score_mode in ("both", "matrix"), args.score_mode in ("both", "topk")
    cfg = {"workload": "cfg3: hyperbolic cone inference scoring, %d synthetic image embeddings x 723 ETHEC labels per GPU, D=%d, "
                       "per-level top-5%s" % (n_img, D, " + full [L, N] fp32 energy matrix" if want_matrix else ""),
           "images_per_gpu_per_step": n_img, "labels": 723, "dim": D, "levels": 4, "topk": k, "writes": args.score_mode,
           "parallelism": "dp%d (images sharded by contiguous ranges, labels replicated, no collective)" % world}

    if args.impl == "reference":
        if rank != 0:
            return
        sample = min(n_img, 32768)
        h, labels, images = cfg3_inputs(sample, D, seed=0)
        v, mean_s, n_done = cfg3_cpu(h, labels, images, K, args.steps, args.warmup, budget_s=150.0)
        cfg["images_per_gpu_per_step"] = sample
        cfg["parallelism"] = "host cores only"
        cores = os.cpu_count() or 1
        print(json.dumps({
            "metric": metric, "value": v, "unit": unit, "impl": "reference", "n_gpus": args.gpus, "steps": n_done,
            "warmup": min(1, args.warmup), "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": unit, "cores": cores, "kind": "port",
                             "sample": "%d images x 723 labels per step (energy matrix + per-level top-5), %d timed steps, torch "
                                       "fp32 on all host threads" % (sample, n_done)},
            "e2e": {"value": v, "unit": unit, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}, "gpu_launches": 0}))
        return

    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        torch.distributed.init_process_group("nccl", device_id=dev)
    from learning_embeddings_b200 import _native, ops

    h, labels, images0 = cfg3_inputs(n_img, D, seed=rank)
    L = labels.shape[0]
    labels_d = labels.to(dev)
    # rotate through image sets so that no step finds its inputs in L2; a matrix-writing step also streams 2.9 GB out
    rotation = args.rotation or max(2, int(np.ceil(160e6 / (n_img * D * 4))))
    host_sets = [images0.pin_memory()]
    for r in range(1, rotation):
        host_sets.append(images0.roll(shifts=r * 977, dims=0).contiguous().pin_memory())
    dev_sets = [x.to(dev) for x in host_sets]
    cfg["l2_policy"] = "inputs rotate through %d image sets (%.0f MB > 126 MB L2)%s" % (
        rotation, rotation * n_img * D * 4 / 1e6, "; every step also writes the %.2f GB matrix" % (L * n_img * 4 / 1e9) if want_matrix else "")
    nl = len(h.level_start)
    idx = torch.empty((n_img, nl, k), device=dev, dtype=torch.int32) if want_topk else None
    val = torch.empty((n_img, nl, k), device=dev, dtype=torch.float32) if want_topk else None
    scores = torch.empty((L, n_img), device=dev, dtype=torch.float32) if want_matrix else None

    import ctypes
    lib = _native.lib()
    ls = (ctypes.c_int32 * nl)(*h.level_start)
    le = (ctypes.c_int32 * nl)(*h.level_stop)
    tc_ok = bool(lib.lec_score_tc_supported(ops.GEOM["hyp"], 0, D, L, nl))
    if args.engine == "tc" and not tc_ok:
        raise SystemExit("tensor-core scoring does not support this case")
    # "auto" = what ops.score_topk picks: the tensor-core kernel whenever it supports the case
    use_tc = tc_ok and args.engine in ("tc", "auto")
    cfg["engine"] = "tc (tcgen05 kind::tf32 3xTF32 + fused epilogue)" if use_tc else "simt (packed FFMA2 tile kernel)"
    ws, nb = None, 0
    if use_tc:
        nb = int(lib.lec_score_workspace_bytes(L, D, nl))
        ws = ops._score_workspace(dev, nb)
    st = _native.stream_ptr(dev)

    def score(imgs, idx_o, val_o, scores_o):
        n = imgs.shape[0]
        if use_tc:
            _native.check(lib.lec_score_topk_tc(ops.GEOM["hyp"], 0, _native._p(labels_d), L, _native._p(imgs), n, D, K,
                                                ctypes.cast(ls, ctypes.c_void_p), ctypes.cast(le, ctypes.c_void_p), nl, k,
                                                _native._p(scores_o), _native._p(idx_o), _native._p(val_o), _native._p(ws), nb,
                                                _native.stream_ptr(dev)), "lec_score_topk_tc")
        else:
            _native.check(lib.lec_score_topk_ex(ops.GEOM["hyp"], 0, _native._p(labels_d), L, _native._p(imgs), n, D, K,
                                                ctypes.cast(ls, ctypes.c_void_p), ctypes.cast(le, ctypes.c_void_p), nl, k,
                                                _native._p(scores_o), 1, _native._p(idx_o), _native._p(val_o),
                                                _native.stream_ptr(dev)), "lec_score_topk_ex")

    def sync_all():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    def max_ranks(ms):
        if world > 1:
            tt = torch.tensor([ms], device=dev, dtype=torch.float64)
            torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
            return float(tt.item())
        return ms

    sampler = ClockSampler(local_rank)
    sampler.start()
    W = max(3, args.warmup)
    for i in range(W):
        score(dev_sets[i % rotation], idx, val, scores)
    sync_all()
    launches0 = _native.launch_count()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active.set()
    sync_all()
    t0.record()
    for i in range(args.steps):
        kev[i][0].record()
        score(dev_sets[i % rotation], idx, val, scores)
        kev[i][1].record()
    t1.record()
    sync_all()
    sampler.active.clear()
    lec_launches = _native.launch_count() - launches0
    elapsed_ms = max_ranks(t0.elapsed_time(t1))
    kernel_ms = float(np.mean([a.elapsed_time(b) for a, b in kev]))
    value = world * n_img * L * args.steps / (elapsed_ms * 1e-3)

    # end to end: pinned host images -> H2D -> per-level top-5 -> predictions D2H, through ops.score_topk_host, which
    # cuts the image set into slices and overlaps the copies of neighbouring slices with the kernel (what the
    # reference's caller consumes is the top-5 per level, oe_h.py:2030-2036; the matrix never leaves the device)
    e2e = None
    if not args.no_e2e:
        out_idx = torch.empty((n_img, nl, k), dtype=torch.int16).pin_memory()   # 723 label ids fit int16
        pipe = ops.ScorePipeline(labels_d, "hyp", K, h.level_start, h.level_stop, k=k, slice_images=131072,
                                 engine=("tc" if args.engine == "tc" else "auto"))
        for i in range(2):
            pipe.run(host_sets[i % rotation], out_idx)
        sync_all()
        sampler.active.set()
        w0 = time.perf_counter()
        n_e2e = max(3, min(args.steps, 50))
        for i in range(n_e2e):
            pipe.run(host_sets[i % rotation], out_idx)   # returns with the predictions in host memory
        e_ms = max_ranks((time.perf_counter() - w0) * 1e3)
        sampler.active.clear()
        e2e = {"value": world * n_img * L * n_e2e / (e_ms * 1e-3), "unit": unit, "h2d_bytes_per_step": n_img * D * 4,
               "d2h_bytes_per_step": n_img * nl * k * 2, "ms_per_step": e_ms / n_e2e, "steps": n_e2e,
               "api": "ops.ScorePipeline.run (host images in, host top-5 label ids per level out as int16; 128K-image slices, "
                      "copies overlap the kernel)"}
        # the device result of the last slice equals the host copy
        assert int(out_idx.min()) >= -1 and int(out_idx.max()) < L
    sampler.stop()
    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    # a matrix-writing step moves 4 B per score + the image rows once (+ 160 B of top-k per image);
    # a top-k-only step moves (4 D + 160) / L per score
    bytes_per_score = ((4.0 if want_matrix else 0.0) + 4.0 * D / L + (160.0 / L if want_topk else 0.0))
    achieved = n_img * L * bytes_per_score / (kernel_ms * 1e-3) / 1e9
    # measured DRAM traffic of the scoring launch: ncu dram bytes per score (profiles/traffic_cfg3.json, captured on a
    # 303 104-image launch of the same kernel) scaled to this launch's scores
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_cfg3.json")
    if use_tc and os.path.exists(tpath):
        per_score = json.load(open(tpath)).get("dram_bytes_per_score", {}).get("d%d_%s" % (D, args.score_mode))
        if per_score is not None:
            traffic = per_score * n_img * L
    roofline = {"bound": "hbm", "kernel": "score_mma_kernel" if use_tc else "score_fast_kernel", "achieved": achieved,
                "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_score": bytes_per_score, "kernel_ms": kernel_ms,
                "kernel_share_of_step": kernel_ms / (elapsed_ms / args.steps),
                "note": "kernel_ms brackets the whole library call (label repack launch + scoring kernel)"
                        + ("" if want_matrix else "; a top-k-only step is bound by FP32/MUFU issue, not HBM")}
    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        sample = min(n_img, 32768)
        v, mean_s, n_done = cfg3_cpu(h, labels, images0[:sample], K, steps=8, warmup=1, budget_s=20.0)
        cpu = {"value": v, "unit": unit, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "%d images x 723 labels per step (energy matrix + per-level top-5), %d steps, torch fp32 on all "
                         "host threads" % (sample, n_done)}
    print(json.dumps({"metric": metric, "value": value, "unit": unit, "n_gpus": world, "steps": args.steps, "warmup": W,
                      "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
                      "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": sampler.summary(), "e2e": e2e,
                      "gpu_launches": int(lec_launches), "roofline": roofline, "cpu_baseline": cpu}))
    if world > 1:
        torch.distributed.destroy_process_group()


def main():
    args = parse()
    if args.workload == "cfg2":
        return run_cfg2(args)
    if args.workload == "cfg3":
        return run_cfg3(args)
    spec = workload_spec(args.workload)
    Nn, D = spec["n_neg"], spec["D"]
    ppg = 1 + 2 * Nn
    groups = max(1, args.pairs // ppg)
    pairs_per_step = groups * ppg
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    precision = args.precision if args.precision is not None else 1
    cfg = {"workload": spec["name"], "pairs_per_gpu_per_step": pairs_per_step, "positives_per_gpu_per_step": groups,
           "dim": D, "negatives_per_edge": 2 * Nn, "table_rows": None, "update": "rsgd",
           "scalar_core": "fp64" if precision == 1 else "fp32", "index_dtype": None,
           "parallelism": "dp%d (pairs sharded, table replicated)" % world}

    h = build_hierarchy(spec)
    cfg["table_rows"] = int(h.n)
    from learning_embeddings_b200.engine import index_dtype_for
    idx_bytes = np.dtype(index_dtype_for(h.n)).itemsize
    cfg["index_dtype"] = np.dtype(index_dtype_for(h.n)).name
    table0 = init_table(h.n, D, spec["K"], seed=0)
    if spec["geom"] == "euc":
        table0 = torch.randn(h.n, D, generator=torch.Generator().manual_seed(0))   # nn.Embedding default init
        cfg["update"] = "adam (torch.optim.Adam semantics inside the fused update kernel)"
        cfg["scalar_core"] = "fp32"

    # ---------------- reference arm: CPU only, rank 0 only ----------------
    if args.impl == "reference":
        if rank != 0:
            return
        cpu_groups = min(groups, 23831)  # 262 141 pairs per CPU step at N=5: a bounded sample of the batch
        blk = make_batches(h, spec, cpu_groups, 1, seed=1)[0]
        v, mean_s, n_done, pairs = time_cpu(spec, table0, blk, cpu_groups, args.steps, args.warmup, budget_s=150.0)
        cores = os.cpu_count() or 1
        sample = "%d pairs per step (first %d positives of the batch), %d timed steps" % (pairs, cpu_groups, n_done)
        cfg["pairs_per_gpu_per_step"] = pairs
        cfg["positives_per_gpu_per_step"] = cpu_groups
        cfg["parallelism"] = "host cores only"
        print(json.dumps({
            "metric": METRIC, "value": v, "unit": UNIT, "impl": "reference", "n_gpus": args.gpus, "steps": n_done,
            "warmup": args.warmup, "ms_per_step": mean_s * 1e3, "higher_is_better": True, "scaling": "weak",
            "vs_baseline": None, "dtype": "f32", "data": "synthetic", "config": cfg,
            "cpu_baseline": {"value": v, "unit": UNIT, "cores": cores, "kind": "port", "sample": sample},
            "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
            "gpu_launches": 0}))
        return

    # ---------------- B200 arm ----------------
    if not torch.cuda.is_available():
        raise SystemExit("bench.py --impl b200 needs a CUDA device (no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    pg = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=dev)
        pg = dist.group.WORLD

    from learning_embeddings_b200 import _native, ops
    from learning_embeddings_b200.engine import ConeStep

    bytes_per_batch = groups * (2 + 2 * Nn) * idx_bytes + groups * ppg * 4  # indices in + energies out
    rotation = args.rotation or max(4, int(np.ceil(160e6 / bytes_per_batch)))
    cfg["l2_policy"] = "inputs rotate through %d distinct batches (%.0f MB > 126 MB L2)" % (
        rotation, rotation * bytes_per_batch / 1e6)
    host_batches = make_batches(h, spec, groups, rotation, seed=100 + rank)
    dev_batches = [b.to(dev) for b in host_batches]
    table = table0.to(dev).clone()
    eng = ConeStep(table, spec["geom"], Nn, groups, K=spec["K"], alpha=spec["alpha"], lr=spec["lr"],
                   precision=precision, process_group=pg, comm=args.comm,
                   update="adam" if spec["geom"] == "euc" else "auto")
    cfg["exchange"] = (eng.comm + (" " + eng.comm_note if eng.comm_note else "")) if world > 1 else "none (1 GPU)"

    def dev_step(i):
        eng.step_device(*eng._split(dev_batches[i % rotation], groups))

    sampler = ClockSampler(local_rank)
    sampler.start()

    def sync_all():
        if world > 1:
            torch.distributed.barrier()
        torch.cuda.synchronize()

    # warm-up
    for i in range(max(3, args.warmup)):
        dev_step(i)
    sync_all()

    # timed region: inputs resident in HBM
    launches0 = _native.launch_count()
    kev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(args.steps)]
    t0, t1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    sampler.active.set()
    sync_all()
    t0.record()
    for i in range(args.steps):
        # the pair kernel is bracketed by CUDA events on every 4th step only: two event records per step sit between
        # back-to-back launches and cost more than they measure
        eng.kernel_events = kev[i] if i % 4 == 0 else None
        dev_step(i)
    t1.record()
    sync_all()
    sampler.active.clear()
    eng.kernel_events = None
    lec_launches = _native.launch_count() - launches0
    elapsed_ms = t0.elapsed_time(t1)
    kernel_ms = float(np.mean([a.elapsed_time(b) for i, (a, b) in enumerate(kev) if i % 4 == 0]))
    final_loss = float(eng.global_loss().item())
    if world > 1:
        tt = torch.tensor([elapsed_ms], device=dev, dtype=torch.float64)
        torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
        elapsed_ms = float(tt.item())
    value = world * pairs_per_step * args.steps / (elapsed_ms * 1e-3)

    # end to end through the public host API: pinned host index block -> H2D -> step -> loss D2H, every step.
    # "value" is the pipelined path (ConeStep.submit_host / drain: the copy of step i+1 overlaps the kernels of
    # step i, losses are read back asynchronously, every loss is delivered); "sync" is ConeStep.step_host, which
    # returns each step's loss before the next step is issued.
    e2e = None
    if not args.no_e2e:
        def timed(fn_step, fn_end):
            for i in range(3):
                fn_step(i)
            fn_end()
            sync_all()
            sampler.active.set()
            w0 = time.perf_counter()
            e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            e0.record()
            for i in range(args.steps):
                fn_step(i)
            got = fn_end()
            e1.record()
            sync_all()
            sampler.active.clear()
            ms = max(e0.elapsed_time(e1), (time.perf_counter() - w0) * 1e3)
            if world > 1:
                tt = torch.tensor([ms], device=dev, dtype=torch.float64)
                torch.distributed.all_reduce(tt, op=torch.distributed.ReduceOp.MAX)
                ms = float(tt.item())
            return ms, got

        sync_ms, _ = timed(lambda i: eng.step_host(host_batches[i % rotation], groups), lambda: None)
        pipe_ms, losses = timed(lambda i: eng.submit_host(host_batches[i % rotation], groups), eng.drain)
        assert len(losses) == args.steps, "every step's loss must come back to the host"
        e2e = {"value": world * pairs_per_step * args.steps / (pipe_ms * 1e-3), "unit": UNIT,
               "h2d_bytes_per_step": groups * (2 + 2 * Nn) * idx_bytes, "d2h_bytes_per_step": 8,
               "ms_per_step": pipe_ms / args.steps, "api": "ConeStep.submit_host/drain (copy of step i+1 overlaps step i)",
               "sync": {"value": world * pairs_per_step * args.steps / (sync_ms * 1e-3), "ms_per_step": sync_ms / args.steps,
                        "api": "ConeStep.step_host (loss returned before the next step is issued)"}}
    sampler.stop()

    if rank != 0:
        if world > 1:
            torch.distributed.destroy_process_group()
        return

    # roofline of the dominant kernel (the fused pair kernel): algorithmic bytes model of SURVEY.md 8(d)
    peaks_path = os.path.join(ROOT, "MEASURED_PEAKS.json")
    if os.path.exists(peaks_path):
        peak, peak_src = float(json.load(open(peaks_path))["hbm_gbs"]), "measured (MEASURED_PEAKS.json hbm_gbs, burst copy)"
    else:
        peak, peak_src = 6650.0, "fallback (B200_PROFILING.md)"
    bytes_per_pair = 24 + 16 * D
    achieved = pairs_per_step * bytes_per_pair / (kernel_ms * 1e-3) / 1e9
    traffic = None
    tpath = os.path.join(ROOT, "profiles", "traffic_%s.json" % args.workload)
    if os.path.exists(tpath):
        traffic = json.load(open(tpath)).get("dram_bytes_per_launch")
    roofline = {"bound": "hbm", "kernel": "pairs_grouped_kernel", "achieved": achieved, "peak": peak, "unit": "GB/s",
                "frac": achieved / peak, "traffic": traffic, "peak_source": peak_src,
                "algorithmic_bytes_per_pair": bytes_per_pair, "kernel_ms": kernel_ms,
                "kernel_share_of_step": kernel_ms / (elapsed_ms / args.steps),
                "note": "logical-bytes model (24+16*D B per pair); the table is L2/L1-resident so DRAM traffic is far "
                        "below it -- the kernel is bound by FP32/FP64 issue and L2 vector reductions, not HBM"}

    cpu = None
    if world == 1 and not args.no_cpu_baseline:
        cpu_groups = min(groups, 23831)
        v, mean_s, n_done, pairs = time_cpu(spec, table0, host_batches[0], cpu_groups, steps=40, warmup=2, budget_s=20.0)
        cpu = {"value": v, "unit": UNIT, "cores": os.cpu_count() or 1, "kind": "port",
               "sample": "%d pairs per CPU step (first %d positives of batch 0), %d steps, torch fp32 on all host "
                         "threads, same step (gather+transform+energy+hinge+backward+update)" % (pairs, cpu_groups, n_done)}

    out = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": max(3, args.warmup),
           "ms_per_step": elapsed_ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
           "dtype": "f32", "data": "synthetic", "config": cfg, "clocks": sampler.summary(), "e2e": e2e,
           "gpu_launches": int(lec_launches), "roofline": roofline, "cpu_baseline": cpu, "final_loss": final_loss}
    print(json.dumps(out))
    if world > 1:
        torch.distributed.destroy_process_group()


if __name__ == "__main__":
    main()
