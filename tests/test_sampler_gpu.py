"""Device fast-mode sampler (lec_sample_negatives_philox): same candidate sets as the reference's
np.where(negative_G[row]) lists, draws identical to the host restatement of the Philox stream."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_embeddings_b200 import _native as N, sampler as S, hierarchy as H  # noqa: E402
from oracle import sampler as osampler  # noqa: E402

pytestmark = pytest.mark.gpu


def _expected(A, u, v, Nn, seed, step, level_start, level_stop, ppl, n_labels=0):
    lib = N.lib()
    nl = len(level_start)
    nt = np.empty((len(u), Nn), dtype=np.int64)
    nf = np.empty((len(u), Nn), dtype=np.int64)
    for i in range(len(u)):
        for p in range(Nn):
            for side in (0, 1):
                c = osampler.candidates_dense(A, u_ix=int(u[i])) if side == 0 else osampler.candidates_dense(A, v_ix=int(v[i]))
                if n_labels:
                    fixed = int(u[i]) if side == 0 else int(v[i])
                    c = osampler.filter_level_joint(c, p, nl, level_start, level_stop, ppl, fixed >= n_labels)
                else:
                    c = osampler.filter_level(c, p, nl, level_start, level_stop, ppl)
                r = lib.lec_philox_below(seed, step, (i * Nn + p) * 2 + side, len(c))
                (nt if side == 0 else nf)[i, p] = c[r]
    return nt, nf


@pytest.mark.parametrize("ppl", [False, True])
@pytest.mark.parametrize("dtype", [torch.uint16, torch.int32, torch.int64])
def test_philox_draws_match_host_restatement_on_ethec(ppl, dtype):
    h = H.ethec()
    A = h.negative_adjacency()
    graph = S.SamplerGraph.from_hierarchy(h, pick_per_level=ppl)
    e = h.closure_edges()[::11]
    u, v = e[:, 0], e[:, 1]
    dev = torch.device("cuda:0")
    np_dt = {torch.uint16: np.uint16, torch.int32: np.int32, torch.int64: np.int64}[dtype]
    ud, vd = torch.from_numpy(u.astype(np_dt)).to(dev), torch.from_numpy(v.astype(np_dt)).to(dev)
    nt, nf = graph.draw_philox(ud, vd, 5, seed=1234, step=7)
    want_t, want_f = _expected(A, u, v, 5, 1234, 7, h.level_start, h.level_stop, ppl)
    assert np.array_equal(nt.cpu().numpy().astype(np.int64), want_t)
    assert np.array_equal(nf.cpu().numpy().astype(np.int64), want_f)
    # every draw is a candidate of the reference's list; another step gives another stream
    assert A[u[:, None], want_t].all() and A[want_f, v[:, None]].all()
    nt2, _ = graph.draw_philox(ud, vd, 5, seed=1234, step=8)
    assert not torch.equal(nt, nt2)


def test_philox_joint_image_level_and_uniformity():
    h = H.random_tree(300, 1.0831, seed=2, roots=6)
    A = h.negative_adjacency()
    e = h.closure_edges()
    ls, le, n_lab = [0, 6, 40], [6, 40, 120], 120
    graph = S.SamplerGraph.from_closure(h.n, e[:, 0], e[:, 1], level_start=ls, level_stop=le, pick_per_level=True,
                                        n_labels=n_lab, level_mod=4)
    sel = np.random.default_rng(1).integers(0, len(e), 200)
    u, v = e[sel, 0], e[sel, 1]
    dev = torch.device("cuda:0")
    nt, nf = graph.draw_philox(torch.from_numpy(u).to(dev), torch.from_numpy(v).to(dev), 8, seed=5, step=0)
    want_t, want_f = _expected(A, u, v, 8, 5, 0, ls, le, True, n_labels=n_lab)
    assert np.array_equal(nt.cpu().numpy(), want_t) and np.array_equal(nf.cpu().numpy(), want_f)
    # uniformity over one node's candidate list
    g2 = S.SamplerGraph.from_closure(h.n, e[:, 0], e[:, 1])
    uu = torch.full((20000,), int(u[0]), dtype=torch.int64, device=dev)
    vv = torch.full((20000,), int(v[0]), dtype=torch.int64, device=dev)
    t, _ = g2.draw_philox(uu, vv, 1, seed=9, step=1)
    cands = osampler.candidates_dense(A, u_ix=int(u[0]))
    counts = np.bincount(t.cpu().numpy().reshape(-1), minlength=h.n)
    assert counts[cands].min() > 0 and counts.sum() == counts[cands].sum()
    exp = 20000 / len(cands)
    assert abs(counts[cands].mean() - exp) < 1e-9 and counts[cands].std() < 3 * np.sqrt(exp)


def test_philox_reports_empty_candidate_list():
    graph = S.SamplerGraph.from_closure(3, [0, 0, 1], [1, 2, 2])
    dev = torch.device("cuda:0")
    with pytest.raises(IndexError):
        graph.draw_philox(torch.tensor([0], device=dev), torch.tensor([1], device=dev), 1, seed=0, step=0)


def test_engine_step_with_device_sampler_equals_step_on_the_same_negatives():
    """ConeStep.step_sampled = draw_philox + step_device: same loss and table as feeding those negatives explicitly."""
    from learning_embeddings_b200.engine import ConeStep
    from learning_embeddings_b200.criterion import inner_radius
    h = H.ethec()
    graph = S.SamplerGraph.from_hierarchy(h)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(0)
    w = torch.randn(h.n, 10, generator=g)
    table0 = (inner_radius(0.1) + 0.05 * torch.rand(h.n, 1, generator=g)) * w / w.norm(dim=1, keepdim=True)
    e = h.closure_edges()
    u = torch.from_numpy(e[:, 0].astype(np.int32)).to(dev)
    v = torch.from_numpy(e[:, 1].astype(np.int32)).to(dev)
    a = ConeStep(table0.to(dev).clone(), "hyp", 5, len(e), K=0.1, alpha=0.05, lr=1e-3)
    b = ConeStep(table0.to(dev).clone(), "hyp", 5, len(e), K=0.1, alpha=0.05, lr=1e-3)
    for step in range(3):
        la = float(a.step_sampled(graph, u, v, seed=42, step=step).item())
        nt, nf = graph.draw_philox(u, v, 5, seed=42, step=step)
        lb = float(b.step_device(u, v, nt.view(-1), nf.view(-1)).item())
        assert abs(la - lb) <= 1e-6 * abs(lb)   # the fp32 gradient reductions are not order-deterministic
    assert torch.allclose(a.table, b.table, rtol=0, atol=1e-5)
    status = graph.device_struct(dev)[2]
    assert int(status.item()) == 0


def test_host_pipe_sampled_submissions_equal_step_sampled():
    """ConeStep.submit_host_sampled (one lec_host_pipe_submit per step: copy of the positives, Philox draw, step, loss
    read-back, slots rotating) = the same steps issued one by one through step_sampled on device-resident positives."""
    from learning_embeddings_b200.engine import ConeStep
    from learning_embeddings_b200.criterion import inner_radius
    h = H.ethec()
    graph = S.SamplerGraph.from_hierarchy(h)
    dev = torch.device("cuda:0")
    g = torch.Generator().manual_seed(1)
    w = torch.randn(h.n, 10, generator=g)
    table0 = (inner_radius(0.1) + 0.05 * torch.rand(h.n, 1, generator=g)) * w / w.norm(dim=1, keepdim=True)
    e = h.closure_edges()
    B = len(e)
    rng = np.random.default_rng(0)
    blocks = []
    for _ in range(5):   # more steps than staging slots: every slot is reused
        perm = rng.permutation(B)
        blk = torch.from_numpy(np.concatenate([e[perm, 0], e[perm, 1]]).astype(np.uint16)).pin_memory()
        blocks.append(blk)
    a = ConeStep(table0.to(dev).clone(), "hyp", 5, B, K=0.1, alpha=0.05, lr=1e-3)
    b = ConeStep(table0.to(dev).clone(), "hyp", 5, B, K=0.1, alpha=0.05, lr=1e-3)
    for blk in blocks:
        a.submit_host_sampled(graph, blk, B, 42)
    la = a.drain()
    lb = []
    for step, blk in enumerate(blocks):
        d = blk.to(dev)
        lb.append(float(b.step_sampled(graph, d[:B], d[B:], seed=42, step=step).item()))
    assert len(la) == 5 and a.drain() == []
    np.testing.assert_allclose(la, lb, rtol=1e-6)
    assert torch.allclose(a.table, b.table, rtol=0, atol=1e-5)
    assert int(graph.device_struct(dev)[2].item()) == 0
    a.close()
    a.close()
