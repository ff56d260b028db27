"""Pins oracle/ against the golden vectors produced by the unmodified reference
(tests/golden/make_golden.py).  CPU only."""
import numpy as np
import pytest
import torch

from conftest import load_golden
from oracle import cones, sampler

T = torch.from_numpy
DIMS = (2, 10, 50)


def t(a, dtype=None):
    x = torch.from_numpy(np.ascontiguousarray(a))
    return x.to(dtype) if dtype is not None else x


def pair_files():
    out = []
    for D in DIMS:
        for a in ("1p0", "0p05"):
            out.append(("euc", "pairs_euc_D%d_a%s" % (D, a)))
            out.append(("hyp", "pairs_hyp_D%d_a%s" % (D, a)))
        out.append(("oe", "pairs_oe_D%d" % D))
    return out


@pytest.mark.parametrize("geom,name", pair_files())
def test_pair_energy_and_autograd_match_reference(geom, name):
    g = load_golden(name)
    K = float(g.get("K", 0.0))
    alpha = float(g["alpha"])
    for dt, sfx, tol in ((torch.float32, "32", 2e-6), (torch.float64, "64", 1e-12)):
        x = t(g["x"], dt).requires_grad_(True)
        y = t(g["y"], dt).requires_grad_(True)
        E = cones.energy(geom, x, y, K)
        loss = cones.hinge_loss(E, t(g["is_pos"]), t(g["w"], dt), alpha)
        loss.backward()
        np.testing.assert_allclose(E.detach().numpy(), g["E" + sfx], rtol=tol, atol=tol)
        np.testing.assert_allclose(float(loss), float(g["L" + sfx]), rtol=max(tol, 1e-6) if dt == torch.float32 else tol)
        scale = np.abs(g["gx" + sfx]).max()
        np.testing.assert_allclose(x.grad.numpy(), g["gx" + sfx], rtol=tol * 10, atol=tol * scale)
        np.testing.assert_allclose(y.grad.numpy(), g["gy" + sfx], rtol=tol * 10, atol=tol * scale)


@pytest.mark.parametrize("geom,name", pair_files())
def test_closed_form_gradients_match_reference_fp64(geom, name):
    g = load_golden(name)
    K = float(g.get("K", 0.0))
    x, y = t(g["x"], torch.float64), t(g["y"], torch.float64)
    E, loss, gx, gy = cones.pair_grads(geom, x, y, t(g["is_pos"]), t(g["w"], torch.float64), float(g["alpha"]), K)
    np.testing.assert_allclose(E.numpy(), g["E64"], rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(float(loss), float(g["L64"]), rtol=1e-12)
    scale = np.abs(g["gx64"]).max()
    np.testing.assert_allclose(gx.numpy(), g["gx64"], rtol=1e-8, atol=1e-12 * scale)
    np.testing.assert_allclose(gy.numpy(), g["gy64"], rtol=1e-8, atol=1e-12 * scale)


ROW_CASES = [("rows_euc", cones.ROW_EUC_SOFTCLIP), ("rows_hyp_shell", cones.ROW_HYP_SHELL),
             ("rows_hyp_tanh", cones.ROW_HYP_TANH)]


@pytest.mark.parametrize("D", DIMS)
@pytest.mark.parametrize("prefix,mode", ROW_CASES)
def test_row_transforms_match_reference(prefix, mode, D):
    g = load_golden("%s_D%d" % (prefix, D))
    K = float(g["K"])
    W = t(g["W"]).requires_grad_(True)
    idx = t(g["idx"])
    out = cones.apply_rows(mode, W[idx], K)
    out.backward(t(g["G_up"]))
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=2e-6, atol=1e-7)
    scale = np.abs(g["gW"]).max()
    np.testing.assert_allclose(W.grad.numpy(), g["gW"], rtol=1e-4, atol=2e-6 * scale)
    # closed-form VJP (fp64 of the same inputs agrees with fp32 golden to fp32 accuracy)
    W64 = t(g["W"], torch.float64)
    gr = cones.rows_bwd(mode, W64[idx], t(g["G_up"], torch.float64), K)
    gW = torch.zeros_like(W64).index_add_(0, idx, gr)
    np.testing.assert_allclose(gW.numpy(), g["gW"], rtol=1e-3, atol=1e-5 * scale)


@pytest.mark.parametrize("D", DIMS)
@pytest.mark.parametrize("tag,mode", [("euc", cones.ROW_EUC_SOFTCLIP), ("hyp", cones.ROW_HYP_TANH_FEAT)])
def test_featnet_tail_matches_reference(tag, mode, D):
    g = load_golden("feat_%s_D%d" % (tag, D))
    K = float(g["K"])
    Z = t(g["Z"]).requires_grad_(True)
    out = cones.apply_rows(mode, Z, K)
    out.backward(t(g["G_up"]))
    np.testing.assert_allclose(out.detach().numpy(), g["out"], rtol=2e-6, atol=1e-7)
    scale = np.abs(g["gZ"]).max()
    np.testing.assert_allclose(Z.grad.numpy(), g["gZ"], rtol=1e-4, atol=2e-6 * scale)
    gz = cones.rows_bwd(mode, t(g["Z"], torch.float64), t(g["G_up"], torch.float64), K)
    np.testing.assert_allclose(gz.numpy(), g["gZ"], rtol=1e-3, atol=1e-5 * scale)


@pytest.mark.parametrize("D", DIMS)
@pytest.mark.parametrize("lr", ("0p001", "0p1"))
def test_rsgd_step_matches_reference(D, lr):
    g = load_golden("rsgd_D%d_lr%s" % (D, lr))
    rg, Wn = cones.rsgd_step(t(g["W"]), t(g["grad"]), float(g["lr"]), float(g["r_in"]))
    np.testing.assert_allclose(rg.numpy(), g["rescaled_grad"], rtol=2e-6, atol=0)
    np.testing.assert_allclose(Wn.numpy(), g["W_new"], rtol=2e-6, atol=1e-8)
    _, Wn64 = cones.rsgd_step(t(g["W"], torch.float64), t(g["grad"], torch.float64), float(g["lr"]), float(g["r_in"]))
    np.testing.assert_allclose(Wn64.numpy(), g["W_new64"], rtol=1e-11, atol=1e-13)


def test_mt19937_reproduces_cpython_choice_streams():
    g = load_golden("mt_choice_streams")
    ns = [int(v) for v in g["ns"]]
    for key in g:
        if not key.startswith("seed_"):
            continue
        rng = sampler.MT19937(int(key[5:]))
        got = [rng.randbelow(m) for _ in range(40) for m in ns]
        assert got == g[key].tolist(), key


def _neg_adj(h):
    n = len(h["parents"])
    A = np.ones((n, n), dtype=bool)
    A[h["tc_edges"][:, 0], h["tc_edges"][:, 1]] = False
    np.fill_diagonal(A, False)
    return A


STEP_CASES = [
    ("step_euc_D2", "euc", cones.ROW_EUC_SOFTCLIP), ("step_euc_D2_a0p05", "euc", cones.ROW_EUC_SOFTCLIP),
    ("step_euc_D10_ppl", "euc", cones.ROW_EUC_SOFTCLIP), ("step_hyp_D10", "hyp", cones.ROW_HYP_SHELL),
    ("step_hyp_D10_a0p05", "hyp", cones.ROW_HYP_SHELL), ("step_hyp_D50_ppl", "hyp", cones.ROW_HYP_SHELL),
    ("step_oe_D10", "oe", cones.ROW_NONE),
]


@pytest.mark.parametrize("name,geom,mode", STEP_CASES)
def test_label_only_step_matches_reference(name, geom, mode, ethec):
    g = load_golden(name)
    N, K, alpha = int(g["N"]), float(g["K"]), float(g["alpha"])
    rng = sampler.MT19937(0)
    nf, nt, drawn = sampler.draw_step_negatives(rng, _neg_adj(ethec), g["u"], g["v"], N, ethec["level_start"],
                                                ethec["level_stop"], bool(g["pick_per_level"]))
    assert drawn.tolist() == g["drawn"].tolist()  # bit-exact negative indices
    r = cones.label_step(geom, t(g["W0"]), mode, K, alpha, t(g["u"]), t(g["v"]), t(nf), t(nt))
    np.testing.assert_allclose(r["E_pos"].numpy(), g["E_pos"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(r["E_neg"].numpy(), g["E_neg"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(float(r["loss"]), float(g["loss"]), rtol=2e-6)
    np.testing.assert_allclose(r["from_emb"].numpy(), g["from_emb"], rtol=2e-6, atol=1e-7)
    scale = np.abs(g["gW"]).max()
    np.testing.assert_allclose(r["gW"].numpy(), g["gW"], rtol=1e-4, atol=1e-5 * scale)
    ev = cones.eval_step(geom, t(g["W0"]), mode, K, alpha, t(g["ev_from"]), t(g["ev_to"]), t(g["ev_status"]))
    np.testing.assert_allclose(ev["E_pos"].numpy(), g["ev_E_pos"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(ev["E_neg"].numpy(), g["ev_E_neg"], rtol=2e-6, atol=2e-6)
    np.testing.assert_allclose(float(ev["loss"]), float(g["ev_loss"]), rtol=2e-6)


def test_weighted_order_embedding_step_matches_reference(ethec):
    """order_embeddings.py:868-915: level weights on positives, n_nodes/N and 1/deg_tc weights on negatives."""
    g = load_golden("step_oe_D10_weighted")
    N, alpha = int(g["N"]), float(g["alpha"])
    n = len(ethec["parents"])
    rng = sampler.MT19937(0)
    nf, nt, drawn = sampler.draw_step_negatives(rng, _neg_adj(ethec), g["u"], g["v"], N, ethec["level_start"],
                                                ethec["level_stop"], True)
    assert drawn.tolist() == g["drawn"].tolist()
    lw = g["level_weights"]
    level_of = np.searchsorted(ethec["level_stop"], np.arange(n), side="right")
    w_pos = lw[level_of[g["v"]]]
    in_deg = np.bincount(ethec["tc_edges"][:, 1], minlength=n)
    out_deg = np.bincount(ethec["tc_edges"][:, 0], minlength=n)
    B = len(g["u"])
    w_neg = np.full(2 * N * B, n / N, dtype=np.float32)
    for i in range(B):
        for p in range(N):
            j = 2 * N * i + p
            if in_deg[nt[j]]:
                w_neg[j] *= np.float32(1.0 / in_deg[nt[j]])
            w_neg[j] *= w_pos[i]
            j = 2 * N * i + N + p
            if out_deg[nf[j]]:
                w_neg[j] *= np.float32(1.0 / out_deg[nf[j]])
            w_neg[j] *= w_pos[i]
    r = cones.label_step("oe", t(g["W0"]), cones.ROW_NONE, 0.0, alpha, t(g["u"]), t(g["v"]), t(nf), t(nt),
                         w_pos=t(w_pos.astype(np.float32)), w_neg=t(w_neg))
    np.testing.assert_allclose(float(r["loss"]), float(g["loss"]), rtol=1e-5)
    scale = np.abs(g["gW"]).max()
    np.testing.assert_allclose(r["gW"].numpy(), g["gW"], rtol=1e-4, atol=1e-5 * scale)


JOINT = [("joint_euc", "euc", cones.ROW_EUC_SOFTCLIP, cones.ROW_EUC_SOFTCLIP),
         ("joint_euc_ppl", "euc", cones.ROW_EUC_SOFTCLIP, cones.ROW_EUC_SOFTCLIP),
         ("joint_oe", "oe", cones.ROW_NONE, cones.ROW_NONE),
         ("joint_hyp", "hyp", cones.ROW_HYP_TANH, cones.ROW_HYP_TANH_FEAT)]


@pytest.mark.parametrize("name,geom,lab_mode,img_mode", JOINT)
def test_joint_image_label_step_matches_reference(name, geom, lab_mode, img_mode):
    g = load_golden(name)
    N, K, alpha, n_lab, nn = int(g["N"]), float(g["K"]), float(g["alpha"]), int(g["n_lab"]), int(g["n_nodes"])
    A = np.unpackbits(g["neg_adj"], axis=1)[:, :nn].astype(bool)
    rng = sampler.MT19937(0)
    nf, nt, drawn = sampler.draw_step_negatives_joint(rng, A, g["b_from"], g["b_to"], N, g["level_start"],
                                                      g["level_stop"], n_lab, bool(g["pick_per_level"]))
    assert drawn.tolist() == g["drawn"].tolist()
    W = t(g["W0"]).requires_grad_(True)
    fw = t(g["fc_w"]).requires_grad_(True)
    fb = t(g["fc_b"]).requires_grad_(True)
    rows = torch.cat([cones.apply_rows(lab_mode, W, K),
                      cones.apply_rows(img_mode, t(g["feat"]) @ fw.t() + fb, K)], dim=0)
    E_pos = cones.energy(geom, rows[t(g["b_from"])], rows[t(g["b_to"])], K)
    E_neg = cones.energy(geom, rows[t(nf)], rows[t(nt)], K)
    loss = E_pos.sum() + (alpha - E_neg).clamp(min=0).sum()
    loss.backward()
    np.testing.assert_allclose(E_pos.detach().numpy(), g["E_pos"].reshape(-1), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(E_neg.detach().numpy(), g["E_neg"].reshape(-1), rtol=1e-5, atol=2e-6)
    np.testing.assert_allclose(float(loss), float(g["loss"]), rtol=1e-5)
    for got, key in ((W.grad, "gW"), (fw.grad, "g_fc_w"), (fb.grad, "g_fc_b")):
        scale = np.abs(g[key]).max()
        np.testing.assert_allclose(got.numpy(), g[key], rtol=1e-3, atol=2e-5 * scale)


@pytest.mark.parametrize("name,geom", [("scoring_hyp_D10", "hyp"), ("scoring_hyp_D50", "hyp"),
                                       ("scoring_euc_D10", "euc"), ("scoring_oe_D10", "oe")])
def test_scoring_matches_reference_loop(name, geom):
    g = load_golden(name)
    K = float(g["K"])
    E = cones.score_matrix(geom, t(g["labels"]), t(g["images"]), K)
    np.testing.assert_allclose(E.numpy(), g["E"], rtol=2e-6, atol=2e-6, equal_nan=True)
    assert np.isnan(E[:, -1].numpy()).all() == np.isnan(g["E"][:, -1]).all()
    idx, val = cones.topk_per_level(E, g["level_start"], g["level_stop"], 5)
    np.testing.assert_allclose(val.numpy(), g["top_val"], rtol=2e-6, atol=2e-6)
    # predicted labels identical wherever the energies are not tied
    tv = g["top_val"]
    distinct = np.ones_like(tv, dtype=bool)
    distinct[..., 1:] &= (tv[..., 1:] - tv[..., :-1]) > 1e-5
    distinct[..., :-1] &= (tv[..., 1:] - tv[..., :-1]) > 1e-5
    assert (idx.numpy()[distinct] == g["top_idx"][distinct]).all()
    E64 = cones.score_matrix(geom, t(g["labels"], torch.float64), t(g["images"], torch.float64), K)
    np.testing.assert_allclose(E64.numpy(), g["E64"], rtol=1e-12, atol=1e-12, equal_nan=True)


def test_best_f1_sweep_matches_reference_pool():
    g = load_golden("metrics_sweep")
    row = cones.best_f1_sweep(t(g["E_pos"]), t(g["E_neg"]))
    np.testing.assert_allclose(np.array(row), g["val_row"], rtol=1e-12)
    fixed = cones.metrics_at_threshold(t(g["E_pos"]), t(g["E_neg"]), float(g["fixed_threshold"]))
    np.testing.assert_allclose(np.array(fixed), g["fixed_row"], rtol=1e-12)


def test_frozen_facts(ethec):
    assert len(ethec["parents"]) == 723 and len(ethec["edges"]) == 717 and len(ethec["tc_edges"]) == 1974
    assert ethec["levels"].tolist() == [6, 21, 135, 561] and ethec["level_start"].tolist() == [0, 6, 27, 162]
    assert abs(cones.inner_radius(0.1) - 0.09901951359278482) < 1e-15
    assert float(load_golden("step_hyp_D10_a0p05")["loss"]) == pytest.approx(2712.4351, abs=1e-3)


@pytest.mark.parametrize("name", ["classify_hyp_D10", "classify_hyp_D50"])
def test_classification_counts_reproduce_reference_metric_dict(name):
    """oracle scoring + bookkeeping against the dict the unmodified calculate_classification_metrics returned."""
    g = load_golden(name)
    lab, img = t(g["labels"]).clone(), t(g["images"]).clone()
    lab[-1] = 0.0   # the reference's slicing leaves the last label / image row zero
    img[-1] = 0.0
    ls, le = g["level_start"].tolist(), g["level_stop"].tolist()
    E = cones.score_matrix("hyp", lab, img, float(g["K"]))
    idx, _ = cones.topk_per_level(E, ls, le, 5)
    hit, tp, fp, tn, fn = cones.classification_counts(idx.numpy(), g["truth"], lab.shape[0], ls, le, (1, 3, 5))
    n_img = img.shape[0]
    for j, kv in enumerate((1, 3, 5)):
        assert int(hit[j].sum()) / (4 * n_img) == float(g["hit@%d" % kv])
        for lvl, (s, e) in enumerate(zip(ls, le)):
            assert int(hit[j][s:e].sum()) / n_img == float(g["level%d_hit@%d" % (lvl, kv)])
    TP, FP, TN, FN = int(tp.sum()), int(fp.sum()), int(tn.sum()), int(fn.sum())
    assert (TP + TN) / (TP + TN + FP + FN) == float(g["accuracy"])
    assert TP / (TP + FP) == float(g["m-precision"]) and TP / (TP + FN) == float(g["m-recall"])
    for lvl, (s, e) in enumerate(zip(ls, le)):
        a = (int(tp[s:e].sum()) + int(tn[s:e].sum())) / int((tp + tn + fp + fn)[s:e].sum())
        assert a == float(g["level%d_accuracy" % lvl])


@pytest.mark.parametrize("tag", ["a", "b", "c"])
def test_caption_ranking_hinge_matches_reference(tag):
    """oracle.cones.caption_ranking_hinge against OrderEmbeddingWithImagesLossvCaption.get_image_label_loss run on the
    same energies (order_embeddings_images.py:533-542; tests/golden/caption_hinge_*.npz), values and autograd gradients."""
    g = load_golden("caption_hinge_" + tag)
    Ep, En = torch.from_numpy(g["E_pos"]), torch.from_numpy(g["E_neg"])
    S, dP, dN = cones.caption_ranking_hinge(Ep, En, float(g["alpha"]))
    np.testing.assert_allclose(S.numpy(), g["S"], rtol=1e-6, atol=1e-7)
    gS = torch.from_numpy(g["gS"])
    # (the reference's autograd adds gS_i once per active negative; count * gS_i differs by that summation's rounding)
    np.testing.assert_allclose((dP * gS).numpy(), g["gE_pos"], rtol=2e-6, atol=0)
    np.testing.assert_array_equal((dN * gS[:, None]).numpy(), g["gE_neg"])


def test_trainer_side_rsgd_helpers_compose_to_the_reference_update():
    """order_embeddings_h.soft_clip / mob_add / lambda_x / exp_map_x (API-compatibility helpers, plain tensor
    expressions) composed as order_embeddings_h.py:764-775 composes them give the oracle's -- i.e. the reference's,
    tests/golden/rsgd_*.npz -- update; the training path itself is the fused CUDA kernel behind rsgd_step."""
    from learning_embeddings_b200 import order_embeddings_h as oeh
    gen = torch.Generator().manual_seed(8)
    n, D, lr = 200, 10, 0.05
    r_in = cones.inner_radius(0.1)
    W = torch.randn(n, D, generator=gen, dtype=torch.float64)
    W = W / W.norm(dim=1, keepdim=True) * (0.05 + 0.94 * torch.rand(n, 1, generator=gen, dtype=torch.float64))
    grad = torch.randn(n, D, generator=gen, dtype=torch.float64) * 3
    g_ref, W_ref = cones.rsgd_step(W, grad, lr, r_in)
    g = grad * (1.0 / oeh.lambda_x(W)) ** 2
    W_new = oeh.exp_map_x(W, -lr * g, r_in)
    np.testing.assert_allclose(g.numpy(), g_ref.numpy(), rtol=1e-13, atol=0)
    np.testing.assert_allclose(W_new.numpy(), W_ref.numpy(), rtol=1e-12, atol=1e-15)
    norms = W_new.norm(dim=1)
    assert float(norms.min()) >= r_in * (1 - 1e-12) and float(norms.max()) <= 1 - 1e-5 + 1e-12
