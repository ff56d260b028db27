"""Host-side mirror of the reference's criterion / Embedder protocol for the label-only trainers.

The classes keep the reference's constructor arguments, attributes, method names and return tuples
 so that the reference's trainers can use them unchanged; the arithmetic is delegated to
the CUDA kernels through `ops`.  The negative sampler stays on the host and consumes Python's `random`
stream in exactly the reference's order, so negative indices are bit-identical.
"""
import random

import numpy as np
import torch
from torch import nn

from . import _native as N
from .sampler import SamplerGraph
from . import ops


def inner_radius(K):
    """order_embeddings_h.py:1089."""
    return 2 * K / (1 + np.sqrt(1 + 4 * K * K))


def unwrap(model):
    """The reference wraps its Embedder in nn.DataParallel (order_embeddings.py:360)."""
    return model.module if hasattr(model, "module") else model


class _EmbedderBase(nn.Module):
    """nn.Embedding table + a per-row transform evaluated by lec_rows_fwd / lec_rows_bwd."""

    row_mode = N.ROWS_NONE

    def table(self):
        return self.embeddings.weight

    def rows(self, geom=None):
        """(rows [n, ld], aux): transformed table (differentiable w.r.t. the embedding weight) and the
        per-row aperture terms of energy `geom`."""
        return ops.transform_rows(self.embeddings.weight, self.row_mode, self.K, geom)

    def forward(self, inputs):
        D = self.embedding_dim
        shape = tuple(inputs.shape)
        rows, _ = self.rows()
        out = rows.index_select(0, inputs.reshape(-1).to(self.embeddings.weight.device))[:, :D]
        return out.reshape(shape + (D,))

    def soft_clip(self, x):
        shp = x.shape
        x2 = x.reshape(-1, shp[-1])
        return ops.transform_rows(x2, self.row_mode, self.K)[0][:, :shp[-1]].reshape(shp)


class EuclideanEmbedder(_EmbedderBase):
    """order_embeddings.py:179-200 (and oe.py:51-80 with normalize=None)."""

    def __init__(self, embedding_dim, labelmap, K=None, normalize=None):
        super().__init__()
        if normalize is not None:
            raise N.LecError("normalize=%r is outside the cone hot path (only normalize=None is built)" % (normalize,))
        self.labelmap = labelmap
        self.embedding_dim = embedding_dim
        self.normalize = normalize
        self.K = K
        self.embeddings = nn.Embedding(self.labelmap.n_classes, self.embedding_dim)
        self.row_mode = N.ROWS_EUC_SOFTCLIP if K else N.ROWS_NONE


class HyperbolicEmbedder(_EmbedderBase):
    """order_embeddings_h.py:181-228: rows initialised at norm r_in + U[0, 0.05), shell projection forward."""

    row_mode = N.ROWS_HYP_SHELL

    def __init__(self, embedding_dim, labelmap, K=None):
        super().__init__()
        self.labelmap = labelmap
        self.embedding_dim = embedding_dim
        self.normalize = None
        self.K = K
        self.inner_radius = inner_radius(K)
        self.epsilon = 1e-5
        self.embeddings = nn.Embedding(self.labelmap.n_classes, self.embedding_dim)
        with torch.no_grad():  # order_embeddings_h.py:198-203
            w = self.embeddings.weight.data
            norm = torch.norm(w, dim=1, keepdim=True)
            new_norm = self.inner_radius + torch.rand((w.shape[0],)) * 0.05
            self.embeddings.weight.data = new_norm.unsqueeze(1).to(w.dtype) * w / norm


class HyperbolicTanhEmbedder(HyperbolicEmbedder):
    """oe_h.py:51-110: exp-map style tanh re-parametrisation followed by the shell projection."""

    row_mode = N.ROWS_HYP_TANH

    def __init__(self, embedding_dim, labelmap, normalize=None, K=None):
        if normalize is not None:
            raise N.LecError("normalize=%r is outside the cone hot path (only normalize=None is built)" % (normalize,))
        super().__init__(embedding_dim, labelmap, K=K)


class CandidateCache:
    """np.where(negative_G[row]) / [:, col] (order_embeddings.py:993-996) memoised per node and level.

    The adjacency is fixed after set_negative_graph, so the candidate list of a node never changes; the
    reference recomputes it (and `.tolist()`s it) on every draw, which is ~75% of its step."""

    def __init__(self, neg_adj):
        self.A = neg_adj
        self.rows = {}
        self.cols = {}

    def get(self, u_ix, v_ix, filt_key, filt):
        store, key = (self.rows, (u_ix, filt_key)) if u_ix is not None else (self.cols, (v_ix, filt_key))
        hit = store.get(key)
        if hit is None:
            c = np.where(self.A[u_ix, :] == 1)[0] if u_ix is not None else np.where(self.A[:, v_ix] == 1)[0]
            if filt is not None:
                c = filt(c)
            hit = c.tolist()
            store[key] = hit
        return hit


class _PairCriterion(nn.Module):
    """Shared machinery of OrderEmbeddingLoss / EucConesLoss (Euclidean and hyperbolic)."""

    geom = None
    precision = ops.PREC_F64CORE  # strict parity by default; PREC_F32 is the fast mode

    def _common_init(self, labelmap, neg_to_pos_ratio, alpha, pick_per_level, weigh_neg_term, level_weights,
                     weigh_pos_term):
        nn.Module.__init__(self)
        self.labelmap = labelmap
        self.neg_to_pos_ratio = neg_to_pos_ratio
        self.alpha = alpha
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.pick_per_level = pick_per_level
        self.weigh_neg_term = weigh_neg_term
        self.weigh_pos_term = weigh_pos_term
        self.level_weights = level_weights
        if self.level_weights is None:
            self.level_weights = torch.ones((len(self.labelmap.levels)))
        self.G_tc = None
        self.nodes_in_G_tc = None
        self.n_nodes_G_tc = None
        self.negative_G = None
        self._cands = None
        self.last_negatives = None  # (negative_from, negative_to) of the latest train-phase forward

    # ---- graph plumbing (order_embeddings.py:949-952, :977-987) ----
    def set_graph_tc(self, graph_tc):
        self.G_tc = graph_tc
        self.nodes_in_G_tc = set(list(self.G_tc))
        self.n_nodes_G_tc = len(set(list(self.G_tc)))

    def set_negative_graph(self, n_G, mapping_from_node_to_ix, mapping_from_ix_to_node):
        self.negative_G = n_G
        self.mapping_from_node_to_ix = mapping_from_node_to_ix
        self.mapping_from_ix_to_node = mapping_from_ix_to_node
        self._cands = CandidateCache(n_G)
        self._graph = None  # native sampler graph, built on the first batch draw

    # ---- sampler (order_embeddings.py:989-1008) ----
    def sample_negative_edge(self, u=None, v=None, level_id=None):
        if level_id is not None:
            level_id = level_id % len(self.labelmap.level_names)
        if (u is None) == (v is None):
            raise ValueError("sample_negative_edge: exactly one of u, v must be given")
        filt, key = None, None
        if self.pick_per_level and level_id < len(self.labelmap.levels):
            lo, hi = self.labelmap.level_start[level_id], self.labelmap.level_stop[level_id]
            key = level_id
            filt = lambda c: c[np.where(np.logical_and(c >= lo, c < hi))]  # noqa: E731
        u_ix = self.mapping_from_node_to_ix[u] if u is not None else None
        v_ix = self.mapping_from_node_to_ix[v] if v is not None else None
        choose_from = self._cands.get(u_ix, v_ix, key, filt)
        return random.choice(choose_from)

    # ---- energies ----
    def E_operator(self, x, y):
        """Drop-in for the reference's E_operator; CPU tensors are moved to the GPU and back (the
        reference's reconstruction check calls it with CPU tensors, order_embeddings.py:548-551)."""
        on_cpu = not x.is_cuda
        if on_cpu:
            x, y = x.cuda(), y.cuda()
        e = ops.energy(x, y, self.geom, getattr(self, "K", None), self.precision)
        return e.cpu() if on_cpu else e

    def positive_pair(self, x, y):
        return self.E_operator(x, y)

    def negative_pair(self, x, y):
        e = self.E_operator(x, y)
        return torch.clamp(self.alpha - e, min=0.0), e

    def get_level_weight_for_edge(self, to):
        """order_embeddings.py:832-838, vectorised."""
        to_arr = np.asarray(to)
        retval = torch.ones((len(to)))
        for level_ix, (lo, hi) in enumerate(zip(self.labelmap.level_start, self.labelmap.level_stop)):
            m = torch.from_numpy((to_arr >= lo) & (to_arr < hi))
            retval[m] = self.level_weights[level_ix]
        return retval

    # ---- loss assembly ----
    def _weights(self, inputs_to, negative_from, negative_to):
        """Pair weights; None means 1 (cones: order_embeddings.py:1048-1052)."""
        return None, None

    def draw_negatives(self, inputs_from, inputs_to):
        """order_embeddings.py:1070-1091: N corrupt-v then corrupt-u draws per positive, in order.  One native
        call (sampler.SamplerGraph.draw_exact) that consumes Python's global `random` stream exactly as the
        reference's B x N x 2 sample_negative_edge calls would, so the indices are bit-identical."""
        if getattr(self, "_graph", None) is None:
            n_names = len(self.labelmap.level_names)
            self._graph = SamplerGraph.from_negative_adjacency(
                self.negative_G, level_start=list(self.labelmap.level_start)[:len(self.labelmap.levels)],
                level_stop=list(self.labelmap.level_stop)[:len(self.labelmap.levels)],
                pick_per_level=self.pick_per_level, level_mod=n_names)
            n = self._graph.n
            ix2node = [self.mapping_from_ix_to_node[i] for i in range(n)]
            self._ix2node_arr = None if ix2node == list(range(n)) else np.asarray(ix2node, dtype=np.int64)
        node2ix = self.mapping_from_node_to_ix
        u_ix = [node2ix[u] for u in inputs_from]
        v_ix = [node2ix[v] for v in inputs_to]
        neg_to, neg_from = self._graph.draw_exact(u_ix, v_ix, self.neg_to_pos_ratio)
        if self._ix2node_arr is not None:
            neg_to, neg_from = self._ix2node_arr[neg_to], self._ix2node_arr[neg_from]
        return neg_to, neg_from

    def forward(self, model, inputs_from, inputs_to, status, phase, neg_to_pos_ratio):
        m = unwrap(model)
        if not torch.cuda.is_available():
            raise N.LecError("no CUDA device: the cone losses have no CPU path")
        dev = m.embeddings.weight.device
        if dev.type != "cuda":
            raise N.LecError("model must live on a CUDA device (got %s)" % dev)
        D = m.embedding_dim
        K = getattr(self, "K", None)
        if hasattr(m, "rows"):
            rows, aux = m.rows(self.geom)
        else:
            rows, aux = ops.transform_rows(m.embeddings.weight, self.default_row_mode, K, self.geom)
        frm = torch.as_tensor(np.asarray(inputs_from, dtype=np.int64))
        to = torch.as_tensor(np.asarray(inputs_to, dtype=np.int64))
        frm_d, to_d = frm.to(dev, non_blocking=True), to.to(dev, non_blocking=True)
        from_emb = rows.index_select(0, frm_d)[:, :D]
        to_emb = rows.index_select(0, to_d)[:, :D]

        if phase != "train":  # order_embeddings.py:1029-1042
            loss, E = ops.flat_pair_loss(rows, aux, D, frm_d, to_d, self.geom, K, self.alpha,
                                         is_pos=(status == 1), precision=self.precision)
            st = status.to(dev)
            return from_emb, to_emb, loss, E[st == 1], E[st == 0]

        if not bool((status == 1).all()):
            raise N.LecError("train phase expects status == 1 for every edge (negatives are sampled here)")
        Nn = self.neg_to_pos_ratio
        neg_to, neg_from = self.draw_negatives(inputs_from, inputs_to)
        # the reference's flat layout, kept for inspection / parity tests
        nf = np.concatenate([np.repeat(np.asarray(inputs_from, dtype=np.int64)[:, None], Nn, 1), neg_from], axis=1)
        nt = np.concatenate([neg_to, np.repeat(np.asarray(inputs_to, dtype=np.int64)[:, None], Nn, 1)], axis=1)
        self.last_negatives = (nf.reshape(-1), nt.reshape(-1))
        w_pos, w_neg = self._weights(inputs_to, nf.reshape(-1), nt.reshape(-1))
        loss, E_pos, E_neg = ops.grouped_pair_loss(
            rows, aux, D, frm_d, to_d, torch.from_numpy(neg_to).to(dev, non_blocking=True),
            torch.from_numpy(neg_from).to(dev, non_blocking=True), Nn, self.geom, K, self.alpha,
            w_pos=w_pos, w_neg=w_neg, precision=self.precision)
        return from_emb, to_emb, loss, E_pos, E_neg.reshape(-1)


class OrderEmbeddingLoss(_PairCriterion):
    """order_embeddings.py:760-923."""

    geom = "oe"
    default_row_mode = N.ROWS_NONE

    def __init__(self, labelmap, neg_to_pos_ratio, alpha=1.0, pick_per_level=True, weigh_neg_term=False,
                 level_weights=None, weigh_pos_term=False):
        self._common_init(labelmap, neg_to_pos_ratio, alpha, pick_per_level, weigh_neg_term, level_weights,
                          weigh_pos_term)

    def _weights(self, inputs_to, negative_from, negative_to):
        """order_embeddings.py:868-915: level weight per positive; negatives n_nodes/N, 1/deg_tc, level weight."""
        Nn = self.neg_to_pos_ratio
        lw = self.get_level_weight_for_edge(inputs_to).float()
        B = len(inputs_to)
        if self.weigh_neg_term:
            w_neg = torch.ones(2 * Nn * B) * self.n_nodes_G_tc / Nn
            in_deg = dict(self.G_tc.in_degree())
            out_deg = dict(self.G_tc.out_degree())
            wn = w_neg.view(B, 2 * Nn)
            deg_u = torch.tensor([in_deg[int(t)] for t in negative_to], dtype=torch.float32).view(B, 2 * Nn)[:, :Nn]
            deg_v = torch.tensor([out_deg[int(f)] for f in negative_from], dtype=torch.float32).view(B, 2 * Nn)[:, Nn:]
            wn[:, :Nn] *= torch.where(deg_u != 0, 1.0 / deg_u, torch.ones_like(deg_u))
            wn[:, Nn:] *= torch.where(deg_v != 0, 1.0 / deg_v, torch.ones_like(deg_v))
        else:
            w_neg = torch.ones(2 * Nn * B)
        if not self.weigh_pos_term:
            w_neg = (w_neg.view(B, 2 * Nn) * lw.unsqueeze(1)).reshape(-1)
        return lw, w_neg


class EucConesLoss(_PairCriterion):
    """Euclidean entailment cones, order_embeddings.py:926-1105 (cos-space energy, K = 3)."""

    geom = "euc"
    default_row_mode = N.ROWS_EUC_SOFTCLIP

    def __init__(self, labelmap, neg_to_pos_ratio, alpha=1.0, pick_per_level=False, weigh_neg_term=False,
                 level_weights=None, weigh_pos_term=False):
        self._common_init(labelmap, neg_to_pos_ratio, alpha, pick_per_level, weigh_neg_term, level_weights,
                          weigh_pos_term)
        self.epsilon = 1e-5
        self.K = 3.0


class HypConesLoss(_PairCriterion):
    """Poincare-ball entailment cones, order_embeddings_h.py:1072-1243 (class is named EucConesLoss there)."""

    geom = "hyp"
    default_row_mode = N.ROWS_HYP_SHELL

    def __init__(self, labelmap, neg_to_pos_ratio, alpha=1.0, pick_per_level=False):
        self._common_init(labelmap, neg_to_pos_ratio, alpha, pick_per_level, False, None, False)
        self.level_weights = torch.ones((len(self.labelmap.levels)))
        self.K = 0.1
        self.inner_radius = inner_radius(self.K)
        self.epsilon = 1e-5


# --------------------------------------------------------------------------------------------------
# Riemannian SGD helpers the hyperbolic trainers define on themselves
# (order_embeddings_h.py:634-674, applied at :764-775; joint copy oe_h.py:1604-1644, :1757-1771)
# --------------------------------------------------------------------------------------------------
def rsgd_step(model, lr, r_in, textbook_lambda=False):
    """weight.grad *= (1/lambda)^2 ; weight = soft_clip(mob_add(weight, exp-map(-lr * grad))) on the whole table."""
    w = unwrap(model).embeddings.weight
    if w.grad is None:
        raise N.LecError("rsgd_step: embeddings.weight.grad is None (call loss.backward() first)")
    ops.rsgd_update_(w.data, w.grad.data, lr, r_in, textbook_lambda=textbook_lambda)
    return w
