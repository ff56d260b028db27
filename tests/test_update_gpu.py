"""GPU tests of the fused table update (lec_update_rows / lec_cone_step) and of its peer exchange.

Reference for every rule: torch itself on the CPU in float64 -- torch.optim.Adam / torch.optim.SGD are what the
reference's trainers call (order_embeddings.py:563-565, oe_h.py:1520-1523) -- driven by the oracle's row transforms
(oracle/cones.py) through autograd, plus the reference's hyperbolic gradient rescale and table projection
(oe_h.py:1765-1771).  Tolerance: 2e-6 absolute + 2e-5 relative on parameters after several steps (fp32 storage).

The multi-rank exchange is run for real on ONE GPU: `world` step engines in one process, one stream each, exchange
buffers in the same device memory (sharding.LocalExchange) -- the same kernels, packets and waits as across NVLink.
"""
import numpy as np
import pytest
import torch

from oracle import cones

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from learning_embeddings_b200 import _native as N
    from learning_embeddings_b200 import ops, sharding
    from learning_embeddings_b200.engine import ConeStep, pack_index_block

DEV = "cuda"


def table_init(n, D, kind, seed):
    g = torch.Generator().manual_seed(seed)
    w = torch.randn(n, D, generator=g)
    if kind == "ball":   # hyperbolic: rows in the unit ball, a few inside the inner radius and a few at the rim
        r = 0.02 + 0.97 * torch.rand(n, 1, generator=g)
        r[::17] = 0.999995
        w = w / w.norm(dim=1, keepdim=True) * r
    return w


def reference_updates(W0, grads, rule, row_mode, K, lr, momentum=0.0, hyp_rescale=False, project_shell=False):
    """k steps of: rows = transform(W); L = <rows, G_t>; backward; [rescale]; optimizer step; [projection] in float64."""
    W = torch.nn.Parameter(W0.double().clone())
    r_in = cones.inner_radius(K) if row_mode >= cones.ROW_HYP_SHELL else 0.0
    opt = None
    if rule == "adam":
        opt = torch.optim.Adam([W], lr=lr)
    elif rule == "sgd":
        opt = torch.optim.SGD([W], lr=lr, momentum=momentum)
    seen = []
    for G in grads:
        W.grad = None
        rows = cones.apply_rows(row_mode, W, K)
        (rows * G.double()).sum().backward()
        if rule == "rsgd":
            gr, W_new = cones.rsgd_step(W.data, W.grad, lr, r_in)
            seen.append(gr.clone())
            W.data = W_new
            continue
        if hyp_rescale:
            W.grad *= ((1.0 - W.data.norm(dim=1, keepdim=True)) / 2.0) ** 2
        seen.append(W.grad.clone())
        opt.step()
        if project_shell:
            W.data = cones._shell_project(W.data, r_in)
    return W.data, seen


CASES = [
    # rule, row_mode, geom, kind, D, extra
    ("rsgd", cones.ROW_HYP_SHELL, "hyp", "ball", 10, {}),
    ("rsgd", cones.ROW_HYP_SHELL, "hyp", "ball", 50, {}),
    ("rsgd", cones.ROW_HYP_SHELL, "hyp", "ball", 2, {}),
    ("rsgd", cones.ROW_HYP_TANH, "hyp", "ball", 10, {}),                       # joint trainer with use_rsgd
    ("adam", cones.ROW_EUC_SOFTCLIP, "euc", "normal", 2, {}),                  # cfg0
    ("adam", cones.ROW_EUC_SOFTCLIP, "euc", "normal", 10, {}),                 # cfg2 label table
    ("adam", cones.ROW_EUC_SOFTCLIP, "euc", "normal", 50, {}),
    ("sgd", cones.ROW_EUC_SOFTCLIP, "euc", "normal", 10, {"momentum": 0.9}),   # order_embeddings.py:563
    ("sgd", cones.ROW_NONE, "oe", "normal", 7, {}),
    ("adam", cones.ROW_NONE, "oe", "normal", 130, {}),
    ("adam", cones.ROW_HYP_TANH, "hyp", "ball", 10, {"hyp_rescale": True, "project_shell": True}),   # oe_h.py:1765-1771
    ("adam", cones.ROW_HYP_TANH, "hyp", "ball", 50, {"hyp_rescale": True, "project_shell": True}),
]


@pytest.mark.parametrize("rule,row_mode,geom,kind,D,extra", CASES)
def test_update_rows_matches_torch_optimizers(rule, row_mode, geom, kind, D, extra):
    n, K, lr, steps, R = 300, {"euc": 3.0, "hyp": 0.1, "oe": 0.0}[geom], 0.01, 4, 3
    W0 = table_init(n, D, kind, seed=D)
    if row_mode == cones.ROW_HYP_SHELL:
        W0 = cones._shell_project(W0, cones.inner_radius(K))   # the label-only trainer keeps its table inside the shell
    g = torch.Generator().manual_seed(100 + D)
    # the gradient arrives split over R replicas, as the pair kernel leaves it; the parts are dyadic rationals so that
    # their fp32 sum is exact in any order (Adam turns a rounding-noise gradient into a full +-lr step)
    parts_all = [torch.round(torch.randn(R, n, D, generator=g) * 128.0) / (1024.0 if kind == "ball" else 512.0)
                 for _ in range(steps)]
    for P in parts_all:
        P[:, ::5] = 0.0   # rows without a gradient this step (RSGD still moves them by its 1e-6)
    grads = [P.sum(0) for P in parts_all]
    W_ref, g_seen = reference_updates(W0, grads, rule, row_mode, K, lr, **extra)
    ld = ops.padded_dim(D)
    table = W0.to(DEV).clone()
    m = torch.zeros(n, ld, device=DEV)
    v = torch.zeros(n, ld, device=DEV)
    rows = torch.empty(n, ld, device=DEV)
    aux = torch.empty(n, 4, device=DEV, dtype=torch.float64)
    gout = torch.empty(n, D, device=DEV)
    r_in = cones.inner_radius(K) if geom == "hyp" else 0.0
    for k, G in enumerate(grads):
        gr = torch.zeros(R, n, ld, device=DEV)
        gr[:, :, :D] = parts_all[k].to(DEV)
        ops.update_rows_(table, gr, rule, row_mode, geom, K, lr, r_in=r_in, state_m=m, state_v=v, opt_step=k + 1,
                         rows_out=rows, aux_out=aux, grad_out=gout, **extra)
        assert float(gr.abs().max()) == 0.0          # replicas cleared for the next pair kernel
        scale = float(g_seen[k].abs().max())
        np.testing.assert_allclose(gout.cpu().numpy(), g_seen[k].numpy(), rtol=2e-4, atol=2e-5 * scale,
                                   err_msg="gradient as the optimizer saw it, step %d" % k)
    np.testing.assert_allclose(table.cpu().numpy(), W_ref.numpy(), rtol=2e-5, atol=2e-6)
    # the rows / aperture terms left behind are those of the updated table
    rows_ref, aux_ref = ops.rows_forward(table, row_mode, K, geom)
    np.testing.assert_allclose(rows.cpu().numpy(), rows_ref.cpu().numpy(), rtol=2e-6, atol=1e-30)
    if geom != "oe":
        a, b = aux.cpu().numpy(), aux_ref.cpu().numpy()
        ok = np.isfinite(b)
        assert (np.isfinite(a) == ok).all()
        np.testing.assert_allclose(a[ok], b[ok], rtol=1e-5, atol=1e-12)
    if D % 4:
        assert float(rows[:, D:].abs().max()) == 0.0 and float(m[:, D:].abs().max()) == 0.0


def ethec_batches(n_batches, B, Nn, seed):
    from learning_embeddings_b200 import hierarchy
    h = hierarchy.ethec()
    rng = np.random.default_rng(seed)
    edges = h.closure_edges()
    out = []
    for _ in range(n_batches):
        sel = rng.integers(0, len(edges), size=B)
        u, v = edges[sel, 0], edges[sel, 1]
        neg_to, neg_from = h.sample_negatives(u, v, Nn, rng)
        out.append((u, v, neg_to, neg_from))
    return h, out


def engine_kwargs(geom):
    if geom == "hyp":
        return dict(K=0.1, alpha=0.05, lr=0.01, update="rsgd")
    return dict(K=3.0, alpha=0.05, lr=0.01, update="adam")


@pytest.mark.parametrize("geom,D,world,mode", [("hyp", 10, 2, 0), ("hyp", 10, 4, 0), ("euc", 2, 2, 0), ("hyp", 50, 3, 0),
                                               ("hyp", 10, 2, 1), ("hyp", 50, 4, 1), ("euc", 10, 3, 1)])
def test_multi_rank_exchange_in_process_equals_single_rank_step(geom, D, world, mode):
    """`world` engines on one GPU, one stream each, each taking its slice of every batch and exchanging gradients
    through lec_update_rows -- mode 0: one-shot packets inside the update kernel; mode 1: two-shot (scatter to the row's
    owner, owner update, all-gather of the updated rows) -- the replicas of the table stay BIT-identical, and equal the
    table of one engine that ran the whole batch (up to the fp32 summation order of the gradient)."""
    Nn, B, steps = 5, 4096, 5
    h, batches = ethec_batches(steps, B, Nn, seed=3)
    W0 = table_init(h.n, D, "ball" if geom == "hyp" else "normal", seed=1)
    if geom == "hyp":
        W0 = cones._shell_project(W0 * 0.6, cones.inner_radius(0.1))
    kw = engine_kwargs(geom)
    single = ConeStep(W0.to(DEV).clone(), geom, Nn, B, **kw)
    xs = sharding.LocalExchange.make(world, h.n, ops.padded_dim(D), torch.device(DEV), mode=mode)
    streams = [torch.cuda.Stream() for _ in range(world)]
    engs = []
    for r in range(world):
        with torch.cuda.stream(streams[r]):
            engs.append(ConeStep(W0.to(DEV).clone(), geom, Nn, B, exchange=xs[r], **kw))
    torch.cuda.synchronize()
    for (u, v, nt, nf) in batches:
        blk = pack_index_block(u, v, nt, nf, n_rows=h.n).to(DEV)
        single.step_device(*single._split(blk, B))
        parts = []
        for r in range(world):
            lo, hi = sharding.shard_bounds(B, r, world)
            parts.append((pack_index_block(u[lo:hi], v[lo:hi], nt[lo:hi], nf[lo:hi]).to(DEV), hi - lo))
        if mode == 0:
            # every rank's whole step on its own stream: the update kernels of the ranks run side by side and wait for one
            # another's packets
            for r in range(world):
                with torch.cuda.stream(streams[r]):
                    parts[r][0].record_stream(streams[r])
                    engs[r].step_device(*engs[r]._split(*parts[r]))
        else:
            # the three launches of the two-shot exchange, rank by rank on ONE stream (lec_exchange_t.phases): each phase
            # only waits for the phase before it, so this order needs no concurrency between the ranks' kernels
            torch.cuda.synchronize()
            for r in range(world):
                engs[r].forward_backward(*engs[r]._split(*parts[r]))
            for phase in (1, 2, 4):
                for r in range(world):
                    engs[r].reduce_and_update(phases=phase)
        torch.cuda.synchronize()
        for r in range(world):
            engs[r].check_exchange()
        total = sum(float(e.loss.item()) for e in engs)
        assert abs(total - float(single.loss.item())) <= 1e-6 * abs(total)   # tables agree to ~1e-7 after the first step
        for e in engs:
            assert abs(float(e.loss_global.item()) - total) <= 1e-12 * abs(total)
    for r in range(1, world):
        assert torch.equal(engs[r].table, engs[0].table), "table replicas diverged"
        assert torch.equal(engs[r].rows, engs[0].rows) and torch.equal(engs[r].aux, engs[0].aux)
    a, b = engs[0].table.cpu().numpy(), single.table.cpu().numpy()
    if kw["update"] == "adam":
        # Adam moves an element by ~lr * sign(g) whatever |g| is, so the rare element whose gradient cancels to rounding
        # noise can land elsewhere when the summation order changes: bound the bulk tightly and the outliers by count
        far = np.abs(a - b) > 2e-6 + 2e-5 * np.abs(b)
        assert far.mean() < 0.005 and np.abs(a - b).max() < 2.5 * kw["lr"] * len(batches)
    else:
        # fp32 L2 reductions are order-dependent: gradient sums of ~1e3 with heavy cancellation, times lr / lambda^2 -- two
        # runs of the SAME single-engine step differ by ~1e-5 on a few elements; that is the resolution of this check
        np.testing.assert_allclose(a, b, rtol=2e-4, atol=5e-5)
        assert (np.abs(a - b) > 2e-6 + 2e-5 * np.abs(b)).mean() < 0.01


def test_exchange_timeout_is_reported_and_leaves_the_table_alone():
    """A rank that never shows up: the waiting rank gives up after timeout_ms, raises on the host, and neither this
    step nor any later one touches its table."""
    Nn, B, D = 5, 512, 10
    h, batches = ethec_batches(1, B, Nn, seed=9)
    W0 = cones._shell_project(table_init(h.n, D, "ball", seed=2) * 0.5, cones.inner_radius(0.1))
    xs = sharding.LocalExchange.make(2, h.n, ops.padded_dim(D), torch.device(DEV), timeout_ms=100)
    eng = ConeStep(W0.to(DEV).clone(), "hyp", Nn, B, exchange=xs[0], **engine_kwargs("hyp"))
    blk = pack_index_block(*batches[0])
    with pytest.raises(N.LecError, match="peer exchange timed out"):
        eng.step_host(blk, B)
    assert torch.equal(eng.table.cpu(), W0)
    eng.step_device(*eng._split(blk.to(DEV), B))     # poisoned: returns without updating
    torch.cuda.synchronize()
    assert torch.equal(eng.table.cpu(), W0)
    with pytest.raises(N.LecError):
        eng.global_loss()


def test_learning_rate_changes_between_steps_take_effect():
    """The reference decays lr every epoch (order_embeddings_h.py:620); the engine re-reads lr on every step."""
    Nn, B, D = 5, 2048, 10
    h, batches = ethec_batches(2, B, Nn, seed=4)
    W0 = cones._shell_project(table_init(h.n, D, "ball", seed=5) * 0.5, cones.inner_radius(0.1))
    tabs = []
    for lr2 in (0.01, 0.0):
        eng = ConeStep(W0.to(DEV).clone(), "hyp", Nn, B, K=0.1, alpha=0.05, lr=0.01)
        eng.step_host(pack_index_block(*batches[0]), B)
        after_first = eng.table.clone()
        eng.set_lr(lr2)
        eng.step_host(pack_index_block(*batches[1]), B)
        tabs.append((after_first, eng.table.clone()))
    # lr = 0 leaves only RSGD's constant 1e-6 drift; lr = 0.01 moves rows by far more
    moved_lr = float((tabs[0][1] - tabs[0][0]).abs().max())
    moved_0 = float((tabs[1][1] - tabs[1][0]).abs().max())
    assert moved_0 < 5e-6 < moved_lr


def test_out_of_range_ids_are_counted_not_dereferenced():
    """nn.Embedding raises IndexError on a bad id in the reference; the kernels skip the pair (energy NaN, no gradient,
    nothing read or written out of bounds) and count it (lec_index_errors)."""
    n, D, B, Nn = 50, 10, 64, 3
    W = table_init(n, D, "normal", seed=0).to(DEV)
    rows, aux = ops.rows_forward(W, N.ROWS_EUC_SOFTCLIP, 3.0, "euc")
    g = torch.Generator().manual_seed(1)
    u = torch.randint(0, n, (B,), generator=g)
    v = (u + 1 + torch.randint(0, n - 1, (B,), generator=g)) % n
    nt = torch.randint(0, n, (B, Nn), generator=g)
    nf = torch.randint(0, n, (B, Nn), generator=g)
    N.index_errors(torch.device(DEV))
    good = torch.zeros(1, n, rows.shape[1], device=DEV)
    ops.pairs_grouped_raw("euc", rows, aux, D, u.to(DEV), v.to(DEV), nt.to(DEV), nf.to(DEV), Nn, 3.0, 1.0, grad_rows=good)
    assert N.index_errors(torch.device(DEV)) == 0
    u2, nt2 = u.clone(), nt.clone()
    u2[7] = n + 5            # a bad positive endpoint: the whole group is dropped
    nt2[9, 1] = 10 ** 6      # a bad negative: that pair only
    grad = torch.zeros(1, n, rows.shape[1], device=DEV)
    _, E_pos, E_neg = ops.pairs_grouped_raw("euc", rows, aux, D, u2.to(DEV), v.to(DEV), nt2.to(DEV), nf.to(DEV), Nn, 3.0, 1.0,
                                            grad_rows=grad)
    torch.cuda.synchronize()
    assert torch.isnan(E_pos[7]) and torch.isnan(E_neg[7]).all() and torch.isnan(E_neg[9, 1])
    assert int(torch.isnan(E_pos).sum()) == 1 and int(torch.isnan(E_neg).sum()) == 2 * Nn + 1
    assert torch.isfinite(grad).all()
    with pytest.raises(IndexError):
        N.raise_on_index_errors(torch.device(DEV))
    assert N.index_errors(torch.device(DEV)) == 0      # reading resets the counter
    fi = torch.tensor([0, 1, n, 3], device=DEV)
    ti = torch.tensor([1, 2, 3, -1], device=DEV)
    _, E = ops.pairs_flat_raw("euc", rows, aux, D, fi, ti, 3.0, 1.0)
    torch.cuda.synchronize()
    assert torch.isnan(E[2]) and torch.isnan(E[3]) and torch.isfinite(E[:2]).all()
    assert N.index_errors(torch.device(DEV)) == 2
