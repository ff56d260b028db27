#!/usr/bin/env python
"""Generate the committed golden vectors by running the UNMODIFIED reference here.

Run once in the build container (where /root/reference exists):

    python tests/golden/make_golden.py

Every array written below comes out of the reference's own classes
(network/order_embeddings.py, order_embeddings_h.py, oe.py, oe_h.py) executed on
CPU by the installed torch (2.11); file:line of the reference entry point is
noted beside each block.  The reference ships no tests or fixtures of its own
(SURVEY.md F11), so these files are what pins oracle/ and, through it, the CUDA
path.  The GPU box never runs this script: it only reads the .npz files.
"""
import math
import os
import random
import sys
import types

import numpy as np
import torch

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import ref_shim  # noqa: E402

ref_shim.install()
import networkx as nx  # noqa: E402

db = ref_shim.load("data.db")
ref_e = ref_shim.load("network.order_embeddings")
ref_h = ref_shim.load("network.order_embeddings_h")
ref_oe = ref_shim.load("network.oe")
ref_oeh = ref_shim.load("network.oe_h")
ref_toy = ref_shim.load("network.embed_toy")

torch.set_num_threads(1)  # fixed reduction order for the fp32 goldens


def save(name, **arrays):
    out = {}
    for k, v in arrays.items():
        if isinstance(v, torch.Tensor):
            v = v.detach().cpu().numpy()
        out[k] = np.asarray(v)
    path = os.path.join(HERE, name + ".npz")
    np.savez_compressed(path, **out)
    print("wrote %-28s %7.1f KB  %s" % (name + ".npz", os.path.getsize(path) / 1024, sorted(out)))


# ---------------------------------------------------------------------------------------
# A. ETHEC hierarchy (data/db.py:3470-3510 child_of_*_ix; order_embeddings.py:363-371)
# ---------------------------------------------------------------------------------------
def build_ethec():
    lm = db.ETHECLabelMapMerged()
    G = nx.DiGraph()
    ls = lm.level_start
    for lvl, table in enumerate([lm.child_of_family_ix, lm.child_of_subfamily_ix, lm.child_of_genus_ix]):
        for p in sorted(table):
            for c in table[p]:
                G.add_edge(p + ls[lvl], c + ls[lvl + 1])
    G_tc = nx.transitive_closure(G)
    return lm, G, G_tc


def neg_adjacency(G_tc, n_nodes):
    # order_embeddings.py:417-423
    A = np.ones((n_nodes, n_nodes), dtype=bool)
    for u, v in G_tc.edges():
        A[u, v] = 0
    np.fill_diagonal(A, 0)
    return A


lm, G, G_tc = build_ethec()
n = lm.n_classes
parents = -np.ones(n, dtype=np.int64)
for u, v in G.edges():
    assert parents[v] == -1
    parents[v] = u
edges = np.array(sorted(G.edges()), dtype=np.int64)
tc_edges = np.array(sorted(G_tc.edges()), dtype=np.int64)
print("ETHEC: nodes", n, "edges", len(edges), "closure", len(tc_edges), "levels", lm.levels)
save("ethec_hierarchy", parents=parents, levels=np.array(lm.levels), level_start=np.array(lm.level_start),
     level_stop=np.array(lm.level_stop), edges=edges, tc_edges=tc_edges)


# ---------------------------------------------------------------------------------------
# B. Pairwise energies + autograd gradients, fp32 and fp64
#    Euclidean  order_embeddings.py:954-975 ; hyperbolic order_embeddings_h.py:1097-1126 ;
#    order-embedding order_embeddings.py:818-830
# ---------------------------------------------------------------------------------------
def pair_case(crit, x, y, is_pos, w, alpha, dtype):
    """loss = sum_pos w*E + sum_neg w*max(0, alpha-E)  (a4/a5 assembly)."""
    crit.alpha = alpha
    x = x.to(dtype).clone().requires_grad_(True)
    y = y.to(dtype).clone().requires_grad_(True)
    w = w.to(dtype)
    pos = is_pos.bool()
    E_pos = crit.positive_pair(x[pos], y[pos])
    neg_term, E_neg = crit.negative_pair(x[~pos], y[~pos])
    loss = torch.sum(w[pos] * E_pos) + torch.sum(w[~pos] * neg_term)
    loss.backward()
    E = torch.zeros(x.shape[0], dtype=dtype)
    E[pos] = E_pos.detach()
    E[~pos] = E_neg.detach()
    return E, loss.detach(), x.grad, y.grad


def ball_points(g, P, D, lo, hi):
    d = torch.randn(P, D, generator=g)
    d = d / d.norm(dim=1, keepdim=True)
    r = lo + (hi - lo) * torch.rand(P, 1, generator=g)
    return (d * r).float()


def gen_pairs():
    g = torch.Generator().manual_seed(1234)
    lab_e = ref_e.EucConesLoss(lm, 5, alpha=1.0)
    lab_h = ref_h.EucConesLoss(lm, 5, alpha=1.0)
    lab_o = ref_e.OrderEmbeddingLoss(lm, 5, alpha=1.0)
    emb_e = ref_e.Embedder(2, lm, K=lab_e.K)  # only for soft_clip
    for D in (2, 10, 50):
        P = 768 if D < 50 else 256
        is_pos = (torch.rand(P, generator=g) < 0.3)
        w = torch.where(torch.rand(P, generator=g) < 0.5, torch.ones(P), 0.25 + torch.rand(P, generator=g))
        # --- Euclidean: points pushed outside radius K by the reference's own soft_clip
        x = emb_e.soft_clip(torch.randn(P, D, generator=g)).detach()
        y = emb_e.soft_clip(torch.randn(P, D, generator=g)).detach()
        # a block of children placed inside / near the parent's cone (hinge boundary cases)
        k = P // 4
        y[:k] = x[:k] * (1.0 + torch.rand(k, 1, generator=g)) + 0.3 * torch.randn(k, D, generator=g)
        for alpha in (1.0, 0.05):
            E32, L32, gx32, gy32 = pair_case(lab_e, x, y, is_pos, w, alpha, torch.float32)
            E64, L64, gx64, gy64 = pair_case(lab_e, x, y, is_pos, w, alpha, torch.float64)
            save("pairs_euc_D%d_a%s" % (D, str(alpha).replace(".", "p")), x=x, y=y, is_pos=is_pos, w=w,
                 alpha=alpha, K=lab_e.K, E32=E32, L32=L32, gx32=gx32, gy32=gy32, E64=E64, L64=L64,
                 gx64=gx64, gy64=gy64)
        # --- hyperbolic: points in the shell [r_in, 1)
        r_in = lab_h.inner_radius
        x = ball_points(g, P, D, r_in, 0.9)
        y = ball_points(g, P, D, r_in, 0.999)
        y[:k] = x[:k] * (1.0 + 2.0 * torch.rand(k, 1, generator=g)) + 0.02 * torch.randn(k, D, generator=g)
        ny = y.norm(dim=1, keepdim=True)
        y = torch.where(ny >= 1.0, y / ny * (1 - 1e-5), y)
        # exact-on-axis children (acos argument hits the clamp)
        y[k:k + 8] = x[k:k + 8] * 1.5
        for alpha in (1.0, 0.05):
            E32, L32, gx32, gy32 = pair_case(lab_h, x, y, is_pos, w, alpha, torch.float32)
            E64, L64, gx64, gy64 = pair_case(lab_h, x, y, is_pos, w, alpha, torch.float64)
            save("pairs_hyp_D%d_a%s" % (D, str(alpha).replace(".", "p")), x=x, y=y, is_pos=is_pos, w=w,
                 alpha=alpha, K=lab_h.K, E32=E32, L32=L32, gx32=gx32, gy32=gy32, E64=E64, L64=L64,
                 gx64=gx64, gy64=gy64)
        # --- order embeddings
        x = torch.randn(P, D, generator=g)
        y = torch.randn(P, D, generator=g)
        y[:k] = x[:k] + torch.rand(k, D, generator=g)  # y >= x: zero energy
        E32, L32, gx32, gy32 = pair_case(lab_o, x, y, is_pos, w, 1.0, torch.float32)
        E64, L64, gx64, gy64 = pair_case(lab_o, x, y, is_pos, w, 1.0, torch.float64)
        save("pairs_oe_D%d" % D, x=x, y=y, is_pos=is_pos, w=w, alpha=1.0, E32=E32, L32=L32, gx32=gx32,
             gy32=gy32, E64=E64, L64=L64, gx64=gx64, gy64=gy64)


# ---------------------------------------------------------------------------------------
# C. Row transforms (Embedder / FeatNet forward) + backward
#    order_embeddings.py:188-200 ; order_embeddings_h.py:205-228 ; oe_h.py:77-104 ; oe_h.py:168-224
# ---------------------------------------------------------------------------------------
def gen_transforms():
    g = torch.Generator().manual_seed(99)
    for D in (2, 10, 50):
        n_rows, P = 64, 200
        idx = torch.randint(0, n_rows, (P,), generator=g)
        G_up = torch.randn(P, D, generator=g)

        def run(model, W):
            with torch.no_grad():
                model.embeddings.weight.copy_(W)
            model.zero_grad()
            out = model(idx)
            out.backward(G_up)
            return out.detach().clone(), model.embeddings.weight.grad.detach().clone()

        # Euclidean soft_clip, K=3
        m = ref_e.Embedder(D, types.SimpleNamespace(n_classes=n_rows), K=3.0)
        W = torch.randn(n_rows, D, generator=g)
        out, gW = run(m, W)
        save("rows_euc_D%d" % D, W=W, idx=idx, G_up=G_up, out=out, gW=gW, K=3.0)

        # hyperbolic label-only shell projection (straight-through), K=0.1
        m = ref_h.Embedder(D, types.SimpleNamespace(n_classes=n_rows), K=0.1)
        W = ball_points(g, n_rows, D, 0.02, 1.2)  # some rows inside r_in, some outside the ball
        out, gW = run(m, W)
        save("rows_hyp_shell_D%d" % D, W=W, idx=idx, G_up=G_up, out=out, gW=gW, K=0.1, r_in=m.inner_radius)

        # hyperbolic joint: tanh re-parametrisation + projection, K=0.1
        m = ref_oeh.Embedder(D, types.SimpleNamespace(n_classes=n_rows), normalize=None, K=0.1)
        W = torch.randn(n_rows, D, generator=g) * torch.logspace(-3, 1.3, n_rows).unsqueeze(1)  # up to |e|~20: clamp binds
        out, gW = run(m, W)
        save("rows_hyp_tanh_D%d" % D, W=W, idx=idx, G_up=G_up, out=out, gW=gW, K=0.1, r_in=m.inner_radius,
             r_in_h=float(m.inner_radius_h))

        # FeatNet tails (fc1 output -> transform); feed the pre-activation directly through the
        # module's own code by making fc1 the identity.
        for tag, mod, K in (("euc", ref_oe, 3.0), ("hyp", ref_oeh, 0.1)):
            fn = mod.FeatNet(normalize=None, input_dim=D, output_dim=D, K=K)
            with torch.no_grad():
                fn.fc1.weight.copy_(torch.eye(D))
                fn.fc1.bias.zero_()
            Z = (torch.randn(P, D, generator=g) * (torch.logspace(-3, 1.3, P).unsqueeze(1) if tag == "hyp" else 1.0))
            Zr = Z.clone().requires_grad_(True)
            out = fn(Zr)
            out.backward(G_up)
            save("feat_%s_D%d" % (tag, D), Z=Z, G_up=G_up, out=out, gZ=Zr.grad, K=K)


# ---------------------------------------------------------------------------------------
# D. RSGD step  order_embeddings_h.py:764-775 (lambda_x 662, exp_map_x 668, mob_add 649, soft_clip 634)
# ---------------------------------------------------------------------------------------
def gen_rsgd():
    g = torch.Generator().manual_seed(7)
    T = ref_h.OrderEmbedding
    for D in (2, 10, 50):
        n_rows = 300 if D < 50 else 120
        r_in = 2 * 0.1 / (1 + np.sqrt(1 + 4 * 0.1 * 0.1))
        fake = types.SimpleNamespace(embedding_dim=D, criterion=types.SimpleNamespace(inner_radius=r_in))
        for name in ("soft_clip", "mob_add", "lambda_x", "exp_map_x"):
            setattr(fake, name, types.MethodType(getattr(T, name), fake))
        W = ball_points(g, n_rows, D, r_in, 0.98)
        grad = torch.randn(n_rows, D, generator=g) * torch.logspace(-4, 3, n_rows).unsqueeze(1)
        grad[::7] = 0.0  # rows untouched by the batch still drift
        for lr in (0.001, 0.1):
            Wd, gd = W.clone(), grad.clone()
            gd *= (1.0 / fake.lambda_x(Wd)) ** 2
            W_new = fake.exp_map_x(Wd, -lr * gd)
            W64, g64 = W.double().clone(), grad.double().clone()
            g64 *= (1.0 / fake.lambda_x(W64)) ** 2
            W_new64 = fake.exp_map_x(W64, -lr * g64)
            save("rsgd_D%d_lr%s" % (D, str(lr).replace(".", "p")), W=W, grad=grad, lr=lr, r_in=r_in,
                 rescaled_grad=gd, W_new=W_new, W_new64=W_new64)


# ---------------------------------------------------------------------------------------
# E. Full label-only training / eval steps incl. the Python sampler
#    order_embeddings.py:1018-1105 ; order_embeddings_h.py:1169-1243 ; order_embeddings.py:840-923
# ---------------------------------------------------------------------------------------
class Recorder:
    def __init__(self, crit):
        self.drawn = []
        orig = crit.sample_negative_edge

        def wrapped(u=None, v=None, level_id=None):
            ix = orig(u=u, v=v, level_id=level_id)
            self.drawn.append(ix)
            return ix

        crit.sample_negative_edge = wrapped


def gen_steps():
    A = neg_adjacency(G_tc, n)
    ident = {i: i for i in range(n)}
    u_all = [int(e[0]) for e in tc_edges]
    v_all = [int(e[1]) for e in tc_edges]
    cases = [
        ("step_euc_D2", ref_e, ref_e.EucConesLoss, 2, dict(alpha=1.0), True),
        ("step_euc_D2_a0p05", ref_e, ref_e.EucConesLoss, 2, dict(alpha=0.05), True),
        ("step_euc_D10_ppl", ref_e, ref_e.EucConesLoss, 10, dict(alpha=0.05, pick_per_level=True), True),
        ("step_hyp_D10", ref_h, ref_h.EucConesLoss, 10, dict(alpha=1.0), True),
        ("step_hyp_D10_a0p05", ref_h, ref_h.EucConesLoss, 10, dict(alpha=0.05), True),
        ("step_hyp_D50_ppl", ref_h, ref_h.EucConesLoss, 50, dict(alpha=0.05, pick_per_level=True), True),
        ("step_oe_D10", ref_e, ref_e.OrderEmbeddingLoss, 10, dict(alpha=1.0, pick_per_level=False), False),
        ("step_oe_D10_weighted", ref_e, ref_e.OrderEmbeddingLoss, 10,
         dict(alpha=1.0, pick_per_level=True, weigh_neg_term=True, level_weights=torch.tensor([4.0, 3.0, 2.0, 1.0])), False),
    ]
    N = 5
    for name, mod, cls, D, kw, is_cone in cases:
        torch.manual_seed(0)
        crit = cls(lm, N, **kw)
        crit.device = torch.device("cpu")
        model = mod.Embedder(D, lm, K=crit.K) if is_cone else mod.Embedder(D, lm)
        crit.set_negative_graph(A, ident, ident)
        crit.set_graph_tc(G_tc)
        rec = Recorder(crit)
        W0 = model.embeddings.weight.detach().clone()
        random.seed(0)
        status = torch.ones(len(u_all), dtype=torch.int64)
        from_emb, to_emb, loss, Ep, En = crit(model, u_all, v_all, status, "train", N)
        loss.backward()
        gW = model.embeddings.weight.grad.detach().clone()
        print(name, "loss", float(loss), "first draws", rec.drawn[:10])
        # eval-phase call on a mixed-status batch built from the first 400 positives + the drawn negatives
        m = 400
        neg_to = rec.drawn[0:2 * N * m:2][:m]
        ev_from = u_all[:m] + u_all[:m]
        ev_to = v_all[:m] + [int(t) for t in neg_to]
        ev_status = torch.tensor([1] * m + [0] * m, dtype=torch.int64)
        with torch.no_grad():
            _, _, ev_loss, ev_Ep, ev_En = crit(model, ev_from, ev_to, ev_status, "val", N)
        save(name, W0=W0, u=np.array(u_all), v=np.array(v_all), N=N, alpha=crit.alpha,
             K=(crit.K if is_cone else 0.0), drawn=np.array(rec.drawn), loss=loss.detach(), E_pos=Ep, E_neg=En,
             from_emb=from_emb, to_emb=to_emb, gW=gW,
             ev_from=np.array(ev_from), ev_to=np.array(ev_to), ev_status=ev_status, ev_loss=ev_loss,
             ev_E_pos=ev_Ep, ev_E_neg=ev_En,
             pick_per_level=int(bool(kw.get("pick_per_level", False))),
             weigh_neg_term=int(bool(kw.get("weigh_neg_term", False))),
             level_weights=(kw["level_weights"] if "level_weights" in kw else torch.ones(4)))

    # toy tree (embed_toy.py:29-62): 4-ary depth 3, the reference's only self-contained workload
    toy = ref_toy.ToyGraph(4, 3)
    Gt = nx.DiGraph()
    Gt.add_edges_from(sorted(toy.edges))
    print("toy:", toy.levels, "edges", Gt.size())
    save("toy_hierarchy", levels=np.array(toy.levels), level_start=np.array(toy.level_start),
         edges=np.array(sorted(Gt.edges())), tc_edges=np.array(sorted(nx.transitive_closure(Gt).edges())))


# ---------------------------------------------------------------------------------------
# F. Joint image+label losses  oe.py:810-873 (Euclidean cones), oe.py:1074-1221 (OE), oe_h.py:904-1058
# ---------------------------------------------------------------------------------------
def gen_joint():
    g = torch.Generator().manual_seed(2024)
    feat_dim, D, N = 96, 10, 3
    # small label hierarchy: the reference's 32-label debug map (data/db.py:3661)
    slm = db.ETHECLabelMapMergedSmall()
    Gs = nx.DiGraph()
    for lvl, table in enumerate([slm.child_of_family_ix, slm.child_of_subfamily_ix, slm.child_of_genus_ix]):
        for p in sorted(table):
            for c in table[p]:
                Gs.add_edge(p + slm.level_start[lvl], c + slm.level_start[lvl + 1])
    n_lab = slm.n_classes
    leaves = list(range(slm.level_start[3], slm.level_stop[3]))
    par = {v: u for u, v in Gs.edges()}
    n_img = 48
    files = ["img_%03d.jpg" % i for i in range(n_img)]
    feats = {f: torch.relu(torch.randn(feat_dim, generator=g)).tolist() for f in files}
    G_tc = Gs.copy()
    img_leaf = {}
    for i, f in enumerate(files):
        leaf = leaves[i % len(leaves)]
        img_leaf[f] = leaf
        node = leaf
        while True:  # oe.py:432-444: every level's label -> image
            G_tc.add_edge(node, f)
            if node not in par:
                break
            node = par[node]
    G_tc = nx.transitive_closure(G_tc)
    # oe.py:452-474
    mapping_ix_to_node, img_label = {}, n_lab
    for node in list(G_tc.nodes()):
        if type(node) == int:
            mapping_ix_to_node[node] = node
        else:
            mapping_ix_to_node[img_label] = node
            img_label += 1
    mapping_node_to_ix = {mapping_ix_to_node[k]: k for k in mapping_ix_to_node}
    nn_ = len(G_tc.nodes())
    A = np.ones((nn_, nn_), dtype=bool)
    for u, v in G_tc.edges():
        A[mapping_node_to_ix[u], mapping_node_to_ix[v]] = 0
    np.fill_diagonal(A, 0)
    # batch: label->label and label->image edges
    all_edges = list(G_tc.edges())
    rnd = random.Random(5)
    batch = rnd.sample(all_edges, 64)
    b_from = [u for u, v in batch]
    b_to = [v for u, v in batch]
    img_order = [mapping_ix_to_node[i] for i in range(n_lab, nn_)]
    feat_mat = np.array([feats[f] for f in img_order], dtype=np.float32)

    def enc(lst):
        return np.array([mapping_node_to_ix[e] for e in lst], dtype=np.int64)

    for name, mod, cls, kw in (
        ("joint_euc", ref_oe, ref_oe.EuclideanConesWithImagesHypernymLoss, dict(K=3.0)),
        ("joint_euc_ppl", ref_oe, ref_oe.EuclideanConesWithImagesHypernymLoss, dict(K=3.0, pick_per_level=True)),
        ("joint_oe", ref_oe, ref_oe.OrderEmbeddingWithImagesHypernymLoss, dict()),
        ("joint_hyp", ref_oeh, ref_oeh.EuclideanConesWithImagesHypernymLoss, dict(K=0.1)),
    ):
        torch.manual_seed(0)
        crit = cls(slm, N, feats, 1.0, **kw)
        crit.device = torch.device("cpu")
        crit.set_negative_graph(A, mapping_node_to_ix, mapping_ix_to_node)
        K = kw.get("K", None)
        model = mod.Embedder(D, slm, normalize=None, K=K)
        fnet = mod.FeatNet(normalize=None, input_dim=feat_dim, output_dim=D, K=K)
        rec = Recorder(crit)
        random.seed(0)
        status = torch.ones(len(b_from), dtype=torch.int64)
        loss, Ep, En = crit(model, fnet, b_from, b_to, b_from, b_to, status, "train")
        loss.backward()
        print(name, "loss", float(loss), "draws", rec.drawn[:8])
        save(name, W0=model.embeddings.weight.detach(), fc_w=fnet.fc1.weight.detach(), fc_b=fnet.fc1.bias.detach(),
             feat=feat_mat, n_lab=n_lab, level_start=np.array(slm.level_start), level_stop=np.array(slm.level_stop),
             neg_adj=np.packbits(A, axis=1), n_nodes=nn_, b_from=enc(b_from), b_to=enc(b_to), N=N, alpha=1.0,
             K=(K or 0.0), drawn=np.array(rec.drawn), loss=loss.detach(), E_pos=Ep, E_neg=En,
             gW=model.embeddings.weight.grad, g_fc_w=fnet.fc1.weight.grad, g_fc_b=fnet.fc1.bias.grad,
             pick_per_level=int(bool(kw.get("pick_per_level", False))))


# ---------------------------------------------------------------------------------------
# F2. Joint trainers, whole training iterations INCLUDING the parameter update:
#     oe.py:1505-1524 (Adam over two parameter groups, oe.py:1356-1357, :1714) and
#     oe_h.py:1755-1771 (default: grad *= (1/lambda_x)^2, Adam over labels + FeatNet, soft_clip on the table;
#     use_rsgd: exp_map_x on the table, Adam on FeatNet).  The loss / criterion objects are the reference's; the four
#     update lines are executed verbatim with JointEmbeddings' own lambda_x / exp_map_x / mob_add / soft_clip bound to
#     a stand-in trainer object (JointEmbeddings.__init__ needs the image dataset on disk).
# ---------------------------------------------------------------------------------------
def gen_joint_update():
    g = torch.Generator().manual_seed(77)
    feat_dim, D, N, steps, B = 96, 10, 3, 3, 48
    slm = db.ETHECLabelMapMergedSmall()
    Gs = nx.DiGraph()
    for lvl, table in enumerate([slm.child_of_family_ix, slm.child_of_subfamily_ix, slm.child_of_genus_ix]):
        for p in sorted(table):
            for c in table[p]:
                Gs.add_edge(p + slm.level_start[lvl], c + slm.level_start[lvl + 1])
    n_lab = slm.n_classes
    leaves = list(range(slm.level_start[3], slm.level_stop[3]))
    par = {v: u for u, v in Gs.edges()}
    n_img = 40
    files = ["img_%03d.jpg" % i for i in range(n_img)]
    feats = {f: torch.relu(torch.randn(feat_dim, generator=g)).tolist() for f in files}
    G_tc = Gs.copy()
    for i, f in enumerate(files):
        node = leaves[(3 * i) % len(leaves)]
        while True:
            G_tc.add_edge(node, f)
            if node not in par:
                break
            node = par[node]
    G_tc = nx.transitive_closure(G_tc)
    mapping_ix_to_node, img_label = {}, n_lab
    for node in list(G_tc.nodes()):
        if type(node) == int:
            mapping_ix_to_node[node] = node
        else:
            mapping_ix_to_node[img_label] = node
            img_label += 1
    mapping_node_to_ix = {mapping_ix_to_node[k]: k for k in mapping_ix_to_node}
    nn_ = len(G_tc.nodes())
    A = np.ones((nn_, nn_), dtype=bool)
    for u, v in G_tc.edges():
        A[mapping_node_to_ix[u], mapping_node_to_ix[v]] = 0
    np.fill_diagonal(A, 0)
    all_edges = list(G_tc.edges())
    rnd = random.Random(11)
    batches = [rnd.sample(all_edges, B) for _ in range(steps)]
    img_order = [mapping_ix_to_node[i] for i in range(n_lab, nn_)]
    feat_mat = np.array([feats[f] for f in img_order], dtype=np.float32)

    def enc(lst):
        return np.array([mapping_node_to_ix[e] for e in lst], dtype=np.int64)

    for name, mod, cls, K, variant in (
        ("joint_upd_euc", ref_oe, ref_oe.EuclideanConesWithImagesHypernymLoss, 3.0, "euc"),
        ("joint_upd_hyp", ref_oeh, ref_oeh.EuclideanConesWithImagesHypernymLoss, 0.1, "hyp_adam"),
        ("joint_upd_hyp_rsgd", ref_oeh, ref_oeh.EuclideanConesWithImagesHypernymLoss, 0.1, "hyp_rsgd"),
    ):
        torch.manual_seed(1)
        crit = cls(slm, N, feats, 1.0, K=K)
        crit.device = torch.device("cpu")
        crit.set_negative_graph(A, mapping_node_to_ix, mapping_ix_to_node)
        model = mod.Embedder(D, slm, normalize=None, K=K)
        fnet = mod.FeatNet(normalize=None, input_dim=feat_dim, output_dim=D, K=K)
        W0 = model.embeddings.weight.detach().clone()
        fw0, fb0 = fnet.fc1.weight.detach().clone(), fnet.fc1.bias.detach().clone()
        lr_labels, lr_images, lr = 0.01, 1e-3, 1e-3
        trainer = None
        if variant == "euc":
            # oe.py:1356-1357 + :1714
            opt = torch.optim.Adam([{'params': model.parameters(), 'lr': 0.1},
                                    {'params': fnet.parameters(), 'weight_decay': 0.0}], lr=lr)
        else:
            T = ref_oeh.JointEmbeddings
            trainer = types.SimpleNamespace(embedding_dim=D, criterion=crit)
            for mname in ("soft_clip", "mob_add", "lambda_x", "exp_map_x"):
                setattr(trainer, mname, types.MethodType(getattr(T, mname), trainer))
            if variant == "hyp_adam":   # oe_h.py:1523
                opt = torch.optim.Adam([{'params': list(model.parameters()) + list(fnet.parameters())}], lr=lr_labels)
            else:                       # oe_h.py:1514-1515
                opt = torch.optim.Adam([{'params': fnet.parameters()}], lr=lr_images)
        rec = Recorder(crit)
        random.seed(0)
        out = dict(W0=W0, fc_w0=fw0, fc_b0=fb0, feat=feat_mat, n_lab=n_lab, N=N, alpha=1.0, K=K, steps=steps,
                   lr_labels=(0.1 if variant == "euc" else lr_labels), lr_fc=(lr if variant == "euc" else
                                                                              (lr_labels if variant == "hyp_adam" else lr_images)))
        for t_ in range(steps):
            b_from = [u for u, v in batches[t_]]
            b_to = [v for u, v in batches[t_]]
            n0 = len(rec.drawn)
            opt.zero_grad()
            model.zero_grad()
            status = torch.ones(B, dtype=torch.int64)
            loss, Ep, En = crit(model, fnet, b_from, b_to, b_from, b_to, status, "train")
            loss.backward()
            Wp = model.embeddings.weight
            if variant == "euc":
                opt.step()                                                       # oe.py:1524
            elif variant == "hyp_adam":
                Wp.grad.data *= (1.0 / trainer.lambda_x(Wp.data)) ** 2           # oe_h.py:1766
                opt.step()                                                       # oe_h.py:1767
                Wp.data = trainer.soft_clip(Wp.data)                             # oe_h.py:1771
            else:
                Wp.grad.data *= (1.0 / trainer.lambda_x(Wp.data)) ** 2           # oe_h.py:1761
                Wp.data = trainer.exp_map_x(Wp.data, -lr_labels * Wp.grad.data)  # oe_h.py:1762
                opt.step()                                                       # oe_h.py:1764
            out["b_from%d" % t_], out["b_to%d" % t_] = enc(b_from), enc(b_to)
            out["drawn%d" % t_] = np.array(rec.drawn[n0:])
            out["loss%d" % t_] = loss.detach()
            out["W%d" % (t_ + 1)] = Wp.detach().clone()
            out["fc_w%d" % (t_ + 1)] = fnet.fc1.weight.detach().clone()
            out["fc_b%d" % (t_ + 1)] = fnet.fc1.bias.detach().clone()
        print(name, "losses", [float(out["loss%d" % t_]) for t_ in range(steps)])
        save(name, **out)


# ---------------------------------------------------------------------------------------
# G. All-pairs image x label scoring, reference-literal loop  oe.py:1764-1779 / oe_h.py:2018-2036
# ---------------------------------------------------------------------------------------
def gen_scoring():
    g = torch.Generator().manual_seed(31)
    L = n
    for geom, crit, D in (("hyp", ref_oeh.EuclideanConesWithImagesHypernymLoss(lm, 5, {}, 1.0, K=0.1), 10),
                          ("hyp", ref_oeh.EuclideanConesWithImagesHypernymLoss(lm, 5, {}, 1.0, K=0.1), 50),
                          ("euc", ref_oe.EuclideanConesWithImagesHypernymLoss(lm, 5, {}, 1.0, K=3.0), 10),
                          ("oe", ref_oe.OrderEmbeddingWithImagesHypernymLoss(lm, 5, {}, 1.0), 10)):
        n_img = 40
        if geom == "hyp":
            lab = torch.zeros(L, D)
            for l in range(4):  # label norm by level
                s, e = lm.level_start[l], lm.level_stop[l]
                lab[s:e] = ball_points(g, e - s, D, 0.10 + 0.2 * l, 0.30 + 0.2 * l)
            img = ball_points(g, n_img, D, 0.30, 0.95)
        elif geom == "euc":
            sc = ref_e.Embedder(D, lm, K=3.0).soft_clip
            lab = sc(torch.randn(L, D, generator=g)).detach()
            img = sc(torch.randn(n_img, D, generator=g) * 3).detach()
        else:
            lab = torch.randn(L, D, generator=g).abs()
            img = torch.randn(n_img, D, generator=g).abs() + 0.5
        lab[L - 1] = 0.0  # the reference's off-by-one leaves the last label row zero
        label_rep = lab.unsqueeze(0)
        img_rep = img.unsqueeze(0)
        E = torch.zeros(n_img, L)
        top_idx = torch.zeros(n_img, 4, 5, dtype=torch.int64)
        top_val = torch.zeros(n_img, 4, 5)
        for i in range(n_img):
            img_emb = img_rep[:, i, :]
            img_emb = img_emb.repeat(1, label_rep.shape[1]).view(-1, label_rep.shape[1], img_emb.shape[1])
            e = crit.E_operator(label_rep, img_emb)
            E[i] = e[0]
            for level_id in range(4):
                values, indices = torch.topk(e[0, lm.level_start[level_id]:lm.level_stop[level_id]], k=5, largest=False)
                top_idx[i, level_id] = indices + lm.level_start[level_id]
                top_val[i, level_id] = values
        E64 = crit.E_operator(label_rep.double().repeat(n_img, 1, 1),
                              img.double().unsqueeze(1).repeat(1, L, 1))
        save("scoring_%s_D%d" % (geom, D), labels=lab, images=img, E=E, E64=E64, top_idx=top_idx, top_val=top_val,
             level_start=np.array(lm.level_start), level_stop=np.array(lm.level_stop), K=getattr(crit, "K", 0.0))


# ---------------------------------------------------------------------------------------
# H. Best-F1 threshold sweep  order_embeddings.py:250-306 (EmbeddingMetrics, 'val' and fixed-threshold phases)
# ---------------------------------------------------------------------------------------
def gen_metrics():
    g = torch.Generator().manual_seed(11)
    Ep = torch.rand(300, generator=g) * 0.6
    En = torch.rand(900, generator=g) * 1.5 + 0.2
    Ep[:40] = 0.0
    En[:25] = 0.0
    val = ref_e.EmbeddingMetrics(Ep, En, 0.0, "val", n_proc=2).calculate_metrics()
    fixed = ref_e.EmbeddingMetrics(Ep, En, 0.37, "test", n_proc=2).calculate_metrics()
    save("metrics_sweep", E_pos=Ep, E_neg=En, val_row=np.array(val, dtype=np.float64),
         fixed_row=np.array(fixed, dtype=np.float64), fixed_threshold=0.37)


# ---------------------------------------------------------------------------------------
# I. CPython Mersenne-Twister stream facts used by the sampler restatement (random.choice)
# ---------------------------------------------------------------------------------------
def gen_mt():
    out = {}
    for seed in (0, 1, 12345, 2 ** 40 + 7):
        random.seed(seed)
        ns = [1, 2, 3, 7, 100, 561, 717, 722, 1000, 65536, 82114]
        draws = []
        for _ in range(40):
            for m in ns:
                draws.append(random.choice(range(m)))
        out["seed_%d" % seed] = np.array(draws, dtype=np.int64)
    out["ns"] = np.array(ns)
    save("mt_choice_streams", **out)


# ---------------------------------------------------------------------------------------
# J. Classification metrics of the joint trainers: the UNMODIFIED JointEmbeddings.calculate_classification_metrics
#    (oe.py:1721-1921, oe_h.py:1971-2178) called unbound on a stand-in `self` that carries only what the method
#    reads: graph_dict, labelmap, embedding_dim, use_CNN, device, model, img_feat_net, criterion.
# ---------------------------------------------------------------------------------------
def gen_classify():
    # hyperbolic only: random Euclidean cones give many exact E = 0 ties (torch.topk's tie order is unspecified) and
    # levels without a single hit, where the reference itself divides by zero (oe.py:1809)
    for tag, mod, crit, D in (("hyp", ref_oeh, ref_oeh.EuclideanConesWithImagesHypernymLoss(lm, 5, {}, 1.0, K=0.1), 10),
                              ("hyp", ref_oeh, ref_oeh.EuclideanConesWithImagesHypernymLoss(lm, 5, {}, 1.0, K=0.1), 50)):
        g = torch.Generator().manual_seed(77)
        n_img = 57
        if tag == "hyp":
            lab = torch.zeros(n, D)
            for l in range(4):
                s, e = lm.level_start[l], lm.level_stop[l]
                lab[s:e] = ball_points(g, e - s, D, 0.10 + 0.2 * l, 0.30 + 0.2 * l)
        else:
            lab = ref_e.Embedder(D, lm, K=3.0).soft_clip(torch.randn(n, D, generator=g)).detach()
        leaves = torch.randint(lm.level_start[3], lm.level_stop[3], (n_img,), generator=g)
        truth = np.zeros((n_img, 4), dtype=np.int64)
        img = torch.zeros(n_img, D)
        for i in range(n_img):
            node = int(leaves[i])
            chain = [node]
            while parents[chain[-1]] >= 0:
                chain.append(int(parents[chain[-1]]))
            truth[i] = sorted(chain)
            # an image sits a little further out than its leaf label, with noise: some predictions hit, some miss
            v = lab[node] * (1.25 if tag == "hyp" else 1.6) + (0.05 if tag == "hyp" else 0.01) * torch.randn(D, generator=g) * lab[node].norm()
            if tag == "hyp" and v.norm() >= 0.97:
                v = v / v.norm() * 0.97
            img[i] = v
        names = ["img_%03d.jpg" % i for i in range(n_img)]
        Gv = nx.DiGraph()
        Gv.add_nodes_from(range(n))
        for i, name in enumerate(names):
            for lbl in truth[i]:
                Gv.add_edge(int(lbl), name)
        crit.feature_dict = {name: img[i].tolist() for i, name in enumerate(names)}
        stub = types.SimpleNamespace(graph_dict={"G_val": Gv}, labelmap=lm, embedding_dim=D, use_CNN=False,
                                     device=torch.device("cpu"), criterion=crit, model=lambda ix: lab[ix],
                                     img_feat_net=lambda x: x.squeeze(0))
        m = mod.JointEmbeddings.calculate_classification_metrics(stub, "val")
        out = {k: np.float64(v) for k, v in m.items() if k != "level_metrics"}
        for lvl, d in m["level_metrics"].items():
            for k, v in d.items():
                out["level%d_%s" % (lvl, k)] = np.float64(v)
        save("classify_%s_D%d" % (tag, D), labels=lab, images=img, truth=truth, level_start=np.array(lm.level_start),
             level_stop=np.array(lm.level_stop), K=crit.K, **out)


# ---------------------------------------------------------------------------------------
# J. caption-style ranking hinge, order_embeddings_images.py:533-542
#    (OrderEmbeddingWithImagesLossvCaption.get_image_label_loss): S_i = sum_j max(0, alpha + E+_i - E-_ij)
# ---------------------------------------------------------------------------------------
def gen_caption():
    ref_img = ref_shim.load("network.order_embeddings_images")
    fn = ref_img.OrderEmbeddingWithImagesLossvCaption.get_image_label_loss
    g = torch.Generator().manual_seed(91)
    for tag, B, M, alpha in (("a", 37, 10, 1.0), ("b", 5, 1, 0.05), ("c", 64, 50, 0.3)):
        E_pos = (torch.rand(B, generator=g) * 1.5).requires_grad_(True)
        E_neg = (torch.rand(B, M, generator=g) * 2.0).requires_grad_(True)
        with torch.no_grad():   # exact ties at the hinge and inactive rows
            E_neg[0, 0] = alpha + E_pos[0]
            E_neg[1] = 10.0
        S = fn(types.SimpleNamespace(alpha=alpha), E_pos, E_neg)
        gS = torch.randn(B, generator=g)
        (S * gS).sum().backward()
        save("caption_hinge_%s" % tag, E_pos=E_pos, E_neg=E_neg, alpha=alpha, S=S, gS=gS, gE_pos=E_pos.grad, gE_neg=E_neg.grad)


if __name__ == "__main__":
    which = sys.argv[1:] or ["pairs", "transforms", "rsgd", "steps", "joint", "joint_update", "scoring", "metrics", "mt", "classify",
                             "caption"]
    fns = dict(pairs=gen_pairs, transforms=gen_transforms, rsgd=gen_rsgd, steps=gen_steps, joint=gen_joint, joint_update=gen_joint_update,
               scoring=gen_scoring, metrics=gen_metrics, mt=gen_mt, classify=gen_classify, caption=gen_caption)
    for w in which:
        print("==", w)
        fns[w]()
