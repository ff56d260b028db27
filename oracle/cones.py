"""Oracle (TEST INFRASTRUCTURE): the reference's cone / order-embedding hot path restated on the CPU.

All functions are dtype-generic torch code: pass float32 tensors for a like-for-like
run against the reference's FP32, float64 for ground truth (SURVEY.md 8c tolerance
contract).  Every block cites the reference lines it follows.  The closed-form
gradient functions (`*_pair_grads`, `*_rows_bwd`) do not use autograd; they are the
formulas the CUDA kernels implement and are themselves pinned against the reference's
autograd in tests/test_oracle_golden.py.

Nothing in the product package imports this file.
"""
import math

import torch

EPS_CLAMP = 1e-5      # acos/asin argument clamp, order_embeddings_h.py:1113-1114
NORM_EPS = 1e-12      # F.normalize eps used by the reference's soft_clip / E_operator
GEOMETRIES = ("euc", "hyp", "oe")


def inner_radius(K):
    """order_embeddings_h.py:1089 -- 2K / (1 + sqrt(1 + 4K^2))."""
    return 2.0 * K / (1.0 + math.sqrt(1.0 + 4.0 * K * K))


# ----------------------------------------------------------------------------------------------
# Energies (differentiable through autograd)
# ----------------------------------------------------------------------------------------------
def energy_euc(x, y, K=3.0):
    """order_embeddings.py:954-969 (same body oe.py:721-739).  Cos-space Euclidean cone energy.

    x is the cone apex (parent), y the child.  NaN when |x| < K, exactly like the reference."""
    shp = x.shape
    x = x.reshape(-1, shp[-1])
    y = y.reshape(-1, shp[-1])
    a = x.norm(p=2, dim=1)
    d = y - x
    xh = x / a.clamp_min(NORM_EPS).unsqueeze(1)
    dh = d / d.norm(p=2, dim=1).clamp_min(NORM_EPS).unsqueeze(1)
    minus_cos = -(xh * dh).sum(dim=1)
    half_aperture = torch.sqrt(1 - (K * K / a ** 2))
    return (minus_cos + half_aperture).clamp(min=0.0).reshape(shp[:-1])


def energy_hyp(x, y, K=0.1, eps=EPS_CLAMP):
    """order_embeddings_h.py:1097-1120 (same body oe_h.py:811-833).  Poincare-ball cone energy."""
    shp = x.shape
    x = x.reshape(-1, shp[-1])
    y = y.reshape(-1, shp[-1])
    a = x.norm(p=2, dim=1)
    b = y.norm(p=2, dim=1)
    s = (x - y).norm(p=2, dim=1)
    p = (x * y).sum(dim=1)
    num = p * (1 + a ** 2) - (a ** 2) * (1 + b ** 2)
    den = a * s * torch.sqrt(1 + (a * b) ** 2 - 2 * p)
    theta = torch.acos((num / den).clamp(min=-1 + eps, max=1 - eps))
    psi = torch.asin((K * (1 - a ** 2) / a).clamp(min=-1 + eps, max=1 - eps))
    return (theta - psi).clamp(min=0.0).reshape(shp[:-1])


def energy_oe(x, y):
    """order_embeddings.py:818-824.  Order-violation energy sum_d max(0, x_d - y_d)^2."""
    shp = x.shape
    x = x.reshape(-1, shp[-1])
    y = y.reshape(-1, shp[-1])
    return ((x - y).clamp(min=0.0) ** 2).sum(dim=1).reshape(shp[:-1])


def energy(geom, x, y, K):
    if geom == "euc":
        return energy_euc(x, y, K)
    if geom == "hyp":
        return energy_hyp(x, y, K)
    if geom == "oe":
        return energy_oe(x, y)
    raise ValueError(geom)


def hinge_loss(E, is_pos, w, alpha):
    """order_embeddings.py:971-975 + :1056-1102.  sum_pos w*E + sum_neg w*max(0, alpha-E)."""
    is_pos = is_pos.bool()
    pos = (w * E)[is_pos].sum()
    neg = (w * (alpha - E).clamp(min=0.0))[~is_pos].sum()
    return pos + neg


# ----------------------------------------------------------------------------------------------
# Closed-form pair gradients (what the kernels implement; SURVEY.md 8a "analytic backward")
# ----------------------------------------------------------------------------------------------
def _coef(E_raw_ge0, E, is_pos, w, alpha):
    """d loss / d z for each pair: +w on active positives, -w on active negatives."""
    is_pos = is_pos.bool()
    act = E_raw_ge0.to(E.dtype)
    pos_c = w * act
    neg_c = -w * act * ((alpha - E) >= 0).to(E.dtype)
    return torch.where(is_pos, pos_c, neg_c)


def euc_pair_grads(x, y, is_pos, w, alpha, K=3.0):
    a = x.norm(dim=1, keepdim=True)
    d = y - x
    b = d.norm(dim=1, keepdim=True)
    xh = x / a.clamp_min(NORM_EPS)
    dh = d / b.clamp_min(NORM_EPS)
    c = (xh * dh).sum(dim=1, keepdim=True)
    root = torch.sqrt(1 - K * K / a ** 2)
    z = (-c + root).squeeze(1)
    E = z.clamp(min=0.0)
    z_d = -(xh - c * dh) / b.clamp_min(NORM_EPS)
    z_x_fixed_d = -(dh - c * xh) / a.clamp_min(NORM_EPS) + (K * K) * x / (a ** 4 * root)
    coef = _coef(z >= 0, E, is_pos, w, alpha).unsqueeze(1)
    gy = coef * z_d
    gx = coef * (z_x_fixed_d - z_d)
    return E, hinge_loss(E, is_pos, w, alpha), gx, gy


def hyp_pair_grads(x, y, is_pos, w, alpha, K=0.1, eps=EPS_CLAMP):
    A = (x * x).sum(dim=1, keepdim=True)
    B = (y * y).sum(dim=1, keepdim=True)
    p = (x * y).sum(dim=1, keepdim=True)
    diff = x - y
    s = diff.norm(dim=1, keepdim=True)
    a = torch.sqrt(A)
    w2 = 1 + A * B - 2 * p
    den = a * s * torch.sqrt(w2)
    g = (p * (1 + A) - A * (1 + B)) / den
    h = K * (1 - A) / a
    gc = g.clamp(-1 + eps, 1 - eps)
    hc = h.clamp(-1 + eps, 1 - eps)
    z = (torch.acos(gc) - torch.asin(hc)).squeeze(1)
    E = z.clamp(min=0.0)
    g_in = ((g >= -1 + eps) & (g <= 1 - eps)).to(x.dtype)
    h_in = ((h >= -1 + eps) & (h <= 1 - eps)).to(x.dtype)
    th = -g_in / torch.sqrt(1 - gc * gc)
    ps = h_in / torch.sqrt(1 - hc * hc)
    g_p = (1 + A) / den + g / w2
    g_A = (p - 1 - B) / den - g * (1 / (2 * A) + B / (2 * w2))
    g_B = -A / den - g * A / (2 * w2)
    g_s = -g / s
    h_A = -K * (1 + A) / (2 * a * A)
    zx = th * (g_p * y + 2 * g_A * x + g_s * diff / s) - ps * h_A * 2 * x
    zy = th * (g_p * x + 2 * g_B * y - g_s * diff / s)
    coef = _coef(z >= 0, E, is_pos, w, alpha).unsqueeze(1)
    return E, hinge_loss(E, is_pos, w, alpha), coef * zx, coef * zy


def oe_pair_grads(x, y, is_pos, w, alpha):
    r = (x - y).clamp(min=0.0)
    E = (r * r).sum(dim=1)
    coef = _coef(torch.ones_like(E, dtype=torch.bool), E, is_pos, w, alpha).unsqueeze(1)
    return E, hinge_loss(E, is_pos, w, alpha), coef * 2 * r, -coef * 2 * r


def pair_grads(geom, x, y, is_pos, w, alpha, K):
    if geom == "euc":
        return euc_pair_grads(x, y, is_pos, w, alpha, K)
    if geom == "hyp":
        return hyp_pair_grads(x, y, is_pos, w, alpha, K)
    if geom == "oe":
        return oe_pair_grads(x, y, is_pos, w, alpha)
    raise ValueError(geom)


# ----------------------------------------------------------------------------------------------
# Row transforms (Embedder.forward / FeatNet tail) -- forward is autograd-differentiable
# ----------------------------------------------------------------------------------------------
ROW_NONE, ROW_EUC_SOFTCLIP, ROW_HYP_SHELL, ROW_HYP_TANH, ROW_HYP_TANH_FEAT = 0, 1, 2, 3, 4


def rows_euc_softclip(e, K):
    """order_embeddings.py:195-200 (= oe.py:75-80, FeatNet oe.py:133-138): e/|e| * (|e| + K)."""
    r = e.norm(dim=1, keepdim=True)
    return e / r.clamp_min(NORM_EPS) * (r + K)


def _shell_project(x, r_in, eps=EPS_CLAMP, feat_variant=False):
    """order_embeddings_h.py:217-228 (no_grad, in place): pull rows into [r_in, 1-eps]."""
    r = x.norm(dim=1, keepdim=True)
    if feat_variant:  # oe_h.py:222
        inner = (1e-6 + x) / (1e-6 + r) * r_in
    else:
        inner = x / r * r_in
    out = torch.where(r <= r_in, inner, x)
    out = torch.where(r >= 1.0, x / r * (1.0 - eps), out)
    return out


def rows_hyp_shell(e, r_in):
    """order_embeddings_h.py:205-228: +1e-15 then shell projection, straight-through gradient."""
    e = e + 1e-15
    return e + (_shell_project(e.detach(), r_in) - e.detach())


def atanh_clamped(v):
    """oe_h.py:106-110."""
    v = min(max(v, -1 + 1e-5), 1 - 1e-5)
    return 0.5 * (math.log(1 + v) - math.log(1 - v))


def rows_hyp_tanh(e, r_in, feat_variant=False):
    """oe_h.py:77-104 (Embedder) / oe_h.py:168-224 (FeatNet tail, feat_variant=True).

    exp-map style re-parametrisation tanh(clamp(atanh(r_in) + |e|, +-15)) * e/|e| followed by the
    straight-through shell projection."""
    c0 = torch.tensor(atanh_clamped(r_in), dtype=torch.float32).to(e.dtype)  # reference keeps it as an fp32 tensor
    e = e + 1e-15
    r = e.norm(dim=1, keepdim=True)
    out = torch.tanh((c0 + r).clamp(min=-15.0, max=15.0)) * (e / r.clamp_min(NORM_EPS))
    return out + (_shell_project(out.detach(), r_in, feat_variant=feat_variant) - out.detach())


def apply_rows(mode, e, K):
    if mode == ROW_NONE:
        return e
    if mode == ROW_EUC_SOFTCLIP:
        return rows_euc_softclip(e, K)
    if mode == ROW_HYP_SHELL:
        return rows_hyp_shell(e, inner_radius(K))
    if mode == ROW_HYP_TANH:
        return rows_hyp_tanh(e, inner_radius(K))
    if mode == ROW_HYP_TANH_FEAT:
        return rows_hyp_tanh(e, inner_radius(K), feat_variant=True)
    raise ValueError(mode)


def rows_bwd(mode, e, g, K):
    """Closed-form vector-Jacobian product of apply_rows (per row; g is d loss / d out)."""
    if mode in (ROW_NONE, ROW_HYP_SHELL):
        return g.clone()
    if mode == ROW_EUC_SOFTCLIP:
        r = e.norm(dim=1, keepdim=True)
        return (1 + K / r) * g - K * e * (e * g).sum(dim=1, keepdim=True) / r ** 3
    if mode in (ROW_HYP_TANH, ROW_HYP_TANH_FEAT):
        c0 = torch.tensor(atanh_clamped(inner_radius(K)), dtype=torch.float32).to(e.dtype)
        e = e + 1e-15
        r = e.norm(dim=1, keepdim=True)
        eh = e / r
        arg = c0 + r
        t = torch.tanh(arg.clamp(-15.0, 15.0))
        tp = torch.where((arg >= -15.0) & (arg <= 15.0), 1 - t * t, torch.zeros_like(t))
        eg = (eh * g).sum(dim=1, keepdim=True)
        return tp * eg * eh + (t / r) * (g - eh * eg)
    raise ValueError(mode)


# ----------------------------------------------------------------------------------------------
# RSGD step on the whole table
# ----------------------------------------------------------------------------------------------
def rsgd_step(W, grad, lr, r_in, textbook_lambda=False):
    """order_embeddings_h.py:764-775 -> lambda_x :662, exp_map_x :668, mob_add :649, soft_clip :634.

    Returns (rescaled_grad, W_new).  The conformal factor uses the NORM (not its square) exactly as
    the reference does unless textbook_lambda is set."""
    wn = W.norm(dim=1, keepdim=True)
    lam = 2.0 / (1 - (wn * wn if textbook_lambda else wn))
    g = grad * (1.0 / lam) ** 2
    v = -lr * g + 1e-15
    vn = v.norm(dim=1, keepdim=True)
    t = torch.tanh((lam * vn / 2).clamp(-15.0, 15.0)) * v / vn
    t = t + 1e-6
    uv2 = 2.0 * (W * t).sum(dim=1, keepdim=True)
    uu = (W * W).sum(dim=1, keepdim=True)
    tt = (t * t).sum(dim=1, keepdim=True)
    den = 1.0 + uv2 + tt * uu
    res = (1.0 + uv2 + tt) / den * W + (1.0 - uu) / den * t
    return g, _shell_project(res, r_in)


# ----------------------------------------------------------------------------------------------
# Loss assembly for one training / eval step of the label-only trainers
# ----------------------------------------------------------------------------------------------
def label_step(geom, W, row_mode, K, alpha, u, v, neg_from, neg_to, w_pos=None, w_neg=None):
    """order_embeddings.py:1018-1105 'train' branch with the negatives already drawn.

    u, v: LongTensor[B]; neg_from, neg_to: LongTensor[2N*B] in the reference's [2N*i + p] layout.
    Returns dict(loss, E_pos, E_neg, from_emb, to_emb, gW) with gW from autograd."""
    W = W.clone().requires_grad_(True)
    fe, te = apply_rows(row_mode, W[u], K), apply_rows(row_mode, W[v], K)
    nfe, nte = apply_rows(row_mode, W[neg_from], K), apply_rows(row_mode, W[neg_to], K)
    E_pos = energy(geom, fe, te, K)
    E_neg = energy(geom, nfe, nte, K)
    w_pos = torch.ones_like(E_pos) if w_pos is None else w_pos.to(E_pos.dtype)
    w_neg = torch.ones_like(E_neg) if w_neg is None else w_neg.to(E_neg.dtype)
    loss = (w_pos * E_pos).sum() + (w_neg * (alpha - E_neg).clamp(min=0.0)).sum()
    loss.backward()
    return dict(loss=loss.detach(), E_pos=E_pos.detach(), E_neg=E_neg.detach(), from_emb=fe.detach(),
                to_emb=te.detach(), gW=W.grad)


def eval_step(geom, W, row_mode, K, alpha, frm, to, status):
    """order_embeddings.py:1029-1042: split by status, unit weights, no gradient."""
    with torch.no_grad():
        fe, te = apply_rows(row_mode, W[frm], K), apply_rows(row_mode, W[to], K)
        pos = status == 1
        E_pos = energy(geom, fe[pos], te[pos], K)
        E_neg = energy(geom, fe[~pos], te[~pos], K)
        loss = E_pos.sum() + (alpha - E_neg).clamp(min=0.0).sum()
    return dict(loss=loss, E_pos=E_pos, E_neg=E_neg)


# ----------------------------------------------------------------------------------------------
# All-pairs image x label scoring + per-level top-k
# ----------------------------------------------------------------------------------------------
def score_matrix(geom, labels, images, K, chunk=4096):
    """oe.py:1764-1773 / oe_h.py:2018-2028: e[i, l] = E(x = label_l, y = image_i)."""
    out = torch.empty(images.shape[0], labels.shape[0], dtype=labels.dtype)
    L, D = labels.shape
    for s in range(0, images.shape[0], chunk):
        img = images[s:s + chunk]
        x = labels.unsqueeze(0).expand(img.shape[0], L, D)
        y = img.unsqueeze(1).expand(img.shape[0], L, D)
        out[s:s + chunk] = energy(geom, x, y, K)
    return out


def topk_per_level(E, level_start, level_stop, k=5):
    """oe.py:1775-1779: torch.topk(e[start:stop], k, largest=False) per level, indices offset back."""
    idx, val = [], []
    for s, e in zip(level_start, level_stop):
        v, i = torch.topk(E[:, int(s):int(e)], k=k, dim=1, largest=False)
        idx.append(i + int(s))
        val.append(v)
    return torch.stack(idx, dim=1), torch.stack(val, dim=1)


# ----------------------------------------------------------------------------------------------
# F1 threshold sweep (EmbeddingMetrics)
# ----------------------------------------------------------------------------------------------
def metrics_at_threshold(E_pos, E_neg, t):
    """order_embeddings.py:258-270 / :289-306.  Row = (f1, t, acc, precision, recall, cp, cn)."""
    cp = int((E_pos <= t).sum())
    cn = int((E_neg > t).sum())
    n_pos, n_neg = E_pos.numel(), E_neg.numel()
    acc = (cp + cn) / (n_pos + n_neg)
    denom = cp + (n_neg - cn)
    prec = cp / denom if denom else 0.0
    rec = cp / n_pos
    f1 = 0.0 if prec + rec == 0 else 2 * prec * rec / (prec + rec)
    return (f1, float(t), acc, prec, rec, cp, cn)


def best_f1_sweep(E_pos, E_neg):
    """order_embeddings.py:272-287: every unique energy is a candidate threshold; first arg-max of F1.

    Restated as sort + searchsorted (O(n log n)) instead of the reference's O(T*n) process pool."""
    E_pos = E_pos.reshape(-1)
    E_neg = E_neg.reshape(-1)
    ts = torch.unique(torch.cat([E_pos, E_neg]))  # ascending, like np.unique
    # NaN energies (the reference's zero label row) satisfy neither E <= t nor E > t: they are counted in
    # the sizes but never in cp / cn, and a NaN threshold scores F1 = 0, so only finite thresholds can win
    ts = ts[~torch.isnan(ts)]
    sp, _ = torch.sort(E_pos[~torch.isnan(E_pos)])
    sn, _ = torch.sort(E_neg[~torch.isnan(E_neg)])
    cp = torch.searchsorted(sp, ts, right=True).double()
    cn = (sn.numel() - torch.searchsorted(sn, ts, right=True)).double()
    n_pos, n_neg = float(E_pos.numel()), float(E_neg.numel())
    prec = cp / (cp + (n_neg - cn))
    rec = cp / n_pos
    f1 = torch.where(prec + rec == 0, torch.zeros_like(prec), 2 * prec * rec / (prec + rec))
    best = int(torch.argmax(f1))  # first maximum, like np.argmax
    acc = (cp + cn) / (n_pos + n_neg)
    return (float(f1[best]), float(ts[best]), float(acc[best]), float(prec[best]), float(rec[best]),
            float(cp[best]), float(cn[best]))


# ----------------------------------------------------------------------------------------------
# Classification bookkeeping (JointEmbeddings.calculate_classification_metrics)
# ----------------------------------------------------------------------------------------------
def classification_counts(top_idx, truth, n_labels, level_start, level_stop, k_vals=(1, 3, 5)):
    """oe.py:1775-1796 / oe_h.py:2030-2051, literally: per image and level, hit@k at the true label, tp / tn on a
    correct top-1 (tn for every OTHER label of the level), else fp at the prediction and fn at the true label.
    Pinned through tests/golden/classify_hyp_*.npz (the reference's own metric dict)."""
    import numpy as np
    top_idx = np.asarray(top_idx)
    truth = np.asarray(truth)
    hit = np.zeros((len(k_vals), n_labels), dtype=np.int64)
    tp, fp, tn, fn = (np.zeros(n_labels, dtype=np.int64) for _ in range(4))
    for i in range(top_idx.shape[0]):
        for lvl in range(top_idx.shape[1]):
            want = int(truth[i, lvl])
            indices = top_idx[i, lvl]
            for j, kv in enumerate(k_vals):
                if want in indices[:kv]:
                    hit[j, want] += 1
            if want == int(indices[0]):
                tp[want] += 1
                for other in range(int(level_start[lvl]), int(level_stop[lvl])):
                    if other != want:
                        tn[other] += 1
            else:
                fp[int(indices[0])] += 1
                fn[want] += 1
    return hit, tp, fp, tn, fn


def caption_ranking_hinge(E_pos, E_neg, alpha):
    """order_embeddings_images.py:533-542 (OrderEmbeddingWithImagesLossvCaption.get_image_label_loss): per positive i
    with its M negatives, S_i = sum_j max(0, alpha + E+_i - E-_ij).  Returns (S [B], dS/dE+ [B], dS/dE- [B, M]);
    clamp(min=0) passes the gradient at equality, like torch."""
    margin = alpha + E_pos[:, None] - E_neg
    act = (margin >= 0).to(E_pos.dtype)
    return margin.clamp(min=0.0).sum(dim=1), act.sum(dim=1), -act

