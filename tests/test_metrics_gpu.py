"""Evaluation bookkeeping on the device (lec_f1_sweep, lec_classify_counts, metrics.py) against the reference's own
outputs (tests/golden/metrics_sweep.npz, classify_hyp_*.npz: the unmodified EmbeddingMetrics and
JointEmbeddings.calculate_classification_metrics) and the oracle restatements."""
import os
import sys

import numpy as np
import pytest
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from conftest import load_golden  # noqa: E402
from oracle import cones  # noqa: E402

pytestmark = pytest.mark.gpu

if torch.cuda.is_available():
    from learning_embeddings_b200 import metrics, ops

DEV = "cuda"


def test_f1_sweep_reproduces_reference_pool_bit_for_bit():
    g = load_golden("metrics_sweep")
    Ep, En = torch.from_numpy(g["E_pos"]), torch.from_numpy(g["E_neg"])
    row = metrics.EmbeddingMetrics(Ep, En, 0.0, "val").calculate_metrics()
    np.testing.assert_array_equal(np.asarray(row, dtype=np.float64), g["val_row"])   # the reference's floats, exactly
    fixed = metrics.EmbeddingMetrics(Ep, En, float(g["fixed_threshold"]), "test").calculate_metrics()
    np.testing.assert_array_equal(np.asarray(fixed, dtype=np.float64), g["fixed_row"])


@pytest.mark.parametrize("n_pos,n_neg,seed", [(1, 1, 0), (7, 5, 1), (1974, 52000, 2), (300000, 700000, 3)])
def test_f1_sweep_matches_oracle_with_ties_and_nans(n_pos, n_neg, seed):
    gen = torch.Generator().manual_seed(seed)
    # coarse grid -> many equal energies, across the two classes too; zeros as the hinge produces them
    Ep = (torch.rand(n_pos, generator=gen) * 40).floor() / 64
    En = (torch.rand(n_neg, generator=gen) * 90).floor() / 64 + 0.25
    if n_pos > 5:
        Ep[3] = float("nan")   # the reference's off-by-one leaves NaN energies
        En[2] = float("nan")
    want = np.asarray(cones.best_f1_sweep(Ep, En), dtype=np.float64)
    got = metrics.best_f1_sweep(Ep.to(DEV), En.to(DEV))
    np.testing.assert_array_equal(got, want)


def test_classification_counts_match_literal_restatement():
    gen = torch.Generator().manual_seed(5)
    ls, le, L, n_img = [0, 6, 27, 162], [6, 27, 162, 723], 723, 4000
    top = torch.stack([torch.stack([torch.randperm(e - s, generator=gen)[:5] + s for s, e in zip(ls, le)]) for _ in range(n_img)])
    truth = torch.stack([torch.randint(s, e, (n_img,), generator=gen) for s, e in zip(ls, le)], dim=1)
    half = n_img // 2   # make half the top-1 predictions correct
    top[:half, :, 0] = truth[:half]
    got = metrics.classification_counts(top.to(torch.int32).to(DEV), truth, L, ls, le, (1, 3, 5))
    want = cones.classification_counts(top.numpy(), truth.numpy(), L, ls, le, (1, 3, 5))
    for a, b in zip(got, want):
        np.testing.assert_array_equal(a, b)


class _LM:
    def __init__(self, g):
        self.level_start, self.level_stop = g["level_start"].tolist(), g["level_stop"].tolist()
        self.levels = [e - s for s, e in zip(self.level_start, self.level_stop)]


@pytest.mark.parametrize("name", ["classify_hyp_D10", "classify_hyp_D50"])
@pytest.mark.parametrize("engine", ["simt", "tc"])
def test_classification_metrics_equal_reference_dict(name, engine):
    g = load_golden(name)
    lab, img = torch.from_numpy(g["labels"]).clone(), torch.from_numpy(g["images"]).clone()
    # the reference's slicing leaves the last label row and the last image row zero (oe_h.py:1995-2002, :2011-2014)
    lab[-1] = 0.0
    img[-1] = 0.0
    m = metrics.classification_metrics(lab.to(DEV), img.to(DEV), g["truth"], _LM(g), "hyp", float(g["K"]), engine=engine)
    flat = {k: v for k, v in m.items() if k != "level_metrics"}
    for lvl, d in m["level_metrics"].items():
        for k, v in d.items():
            flat["level%d_%s" % (lvl, k)] = v
    for key, v in flat.items():
        np.testing.assert_allclose(float(v), float(g[key]), rtol=1e-6 if key.startswith("median") else 0, atol=0, err_msg=key)
    assert set(flat) == {k for k in g if g[k].shape == () and k != "K"}


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_caption_ranking_hinge_matches_reference_and_oracle(tag):
    """lec_caption_hinge (ops.caption_ranking_hinge) against the reference's get_image_label_loss
    (order_embeddings_images.py:533-542, golden) -- values within fp32 summation order, gradients exactly -- and chained
    behind the CUDA energy op, against autograd through the oracle."""
    g = load_golden("caption_hinge_" + tag)
    alpha = float(g["alpha"])
    Ep = torch.from_numpy(g["E_pos"]).to(DEV).requires_grad_(True)
    En = torch.from_numpy(g["E_neg"]).to(DEV).requires_grad_(True)
    S = ops.caption_ranking_hinge(Ep, En, alpha)
    np.testing.assert_allclose(S.detach().cpu().numpy(), g["S"], rtol=1e-6, atol=1e-7)
    (S * torch.from_numpy(g["gS"]).to(DEV)).sum().backward()
    np.testing.assert_allclose(Ep.grad.cpu().numpy(), g["gE_pos"], rtol=2e-6, atol=0)   # count * gS vs repeated adds
    np.testing.assert_array_equal(En.grad.cpu().numpy(), g["gE_neg"])
    # NaN energies propagate like torch.clamp and carry no gradient
    bad = torch.from_numpy(g["E_neg"]).to(DEV)
    bad[0, 0] = float("nan")
    S2 = ops.caption_ranking_hinge(torch.from_numpy(g["E_pos"]).to(DEV), bad, alpha)
    assert torch.isnan(S2[0]) and not torch.isnan(S2[1:]).any()
    # end to end: Euclidean cone energies of embedded pairs -> ranking hinge, gradients w.r.t. the embeddings
    gen = torch.Generator().manual_seed(5)
    B, M, D = 33, 7, 10
    x = (torch.randn(B, D, generator=gen) * 2 + 4).to(DEV).requires_grad_(True)
    y = (torch.randn(B, D, generator=gen) * 2 + 6).to(DEV).requires_grad_(True)
    yn = (torch.randn(B, M, D, generator=gen) * 2 + 5).to(DEV).requires_grad_(True)
    E_pos = ops.energy(x, y, "euc", 3.0)
    E_neg = ops.energy(x[:, None, :].expand(B, M, D).reshape(-1, D), yn.reshape(-1, D), "euc", 3.0).view(B, M)
    loss = ops.caption_ranking_hinge(E_pos, E_neg, 0.3).sum()
    loss.backward()
    xc, yc, ync = (t.detach().cpu().double().requires_grad_(True) for t in (x, y, yn))
    Ep_o = cones.energy("euc", xc, yc, 3.0)
    En_o = cones.energy("euc", xc[:, None, :].expand(B, M, D).reshape(-1, D), ync.reshape(-1, D), 3.0).view(B, M)
    So, _, _ = cones.caption_ranking_hinge(Ep_o, En_o, 0.3)
    So.sum().backward()
    assert abs(float(loss) - float(So.sum())) <= 1e-5 * abs(float(So.sum()))
    for got, ref in ((x.grad, xc.grad), (y.grad, yc.grad), (yn.grad, ync.grad)):
        np.testing.assert_allclose(got.cpu().numpy(), ref.numpy(), rtol=2e-4, atol=2e-5)
