"""Host-side mirror of the reference's JOINT image+label criteria (network/oe.py, network/oe_h.py).

Every endpoint of an edge is either a label id (int) or an image (filename str).  Labels come from the
embedding table, images from `FeatNet` = one linear layer over a 2048-d feature vector followed by the same
kind of row transform.  One training step here is

    label rows, aux  = lec_rows_fwd(table)                                   (all labels, once)
    image rows, aux  = lec_rows_fwd(fc1(features of the step's distinct images))   (cuBLAS GEMM + our kernel)
    loss, dL/drows   = lec_pairs_grouped over the concatenated [labels ; images] row table

with autograd carrying dL/drows back through `lec_rows_bwd` into the table and through cuBLAS into fc1.
The reference instead looks features up filename by filename in a Python dict (oe.py:680-707), runs FeatNet
twice per step and scatters into zero tensors (oe.py:875-964).
"""
import random

import numpy as np
import torch
from torch import nn

from . import _native as N
from . import ops
from .sampler import SamplerGraph
from .criterion import CandidateCache, HyperbolicTanhEmbedder, EuclideanEmbedder, inner_radius, unwrap


class _FeatNetBase(nn.Module):
    """oe.py:83-138 / oe_h.py:113-224 with normalize=None: fc1 then the row transform."""

    row_mode = N.ROWS_NONE

    def __init__(self, normalize, input_dim=2048, output_dim=10, K=None):
        super().__init__()
        if normalize is not None:
            raise N.LecError("normalize=%r is outside the cone hot path (only normalize=None is built)" % (normalize,))
        self.fc1 = nn.Linear(input_dim, output_dim)
        self.normalize = normalize
        self.output_dim = output_dim
        self.K = K
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")

    def rows(self, x, geom=None):
        """(rows [m, ld], aux) for a feature batch x [m, input_dim]; differentiable w.r.t. fc1 and x."""
        return ops.transform_rows(self.fc1(x), self.row_mode, self.K, geom)

    def forward(self, x):
        shp = x.shape
        rows, _ = self.rows(x.reshape(-1, shp[-1]))
        return rows[:, :self.output_dim].reshape(shp[:-1] + (self.output_dim,))


class EuclideanFeatNet(_FeatNetBase):
    def __init__(self, normalize, input_dim=2048, output_dim=10, K=None):
        super().__init__(normalize, input_dim, output_dim, K)
        self.row_mode = N.ROWS_EUC_SOFTCLIP if K else N.ROWS_NONE


class HyperbolicFeatNet(_FeatNetBase):
    row_mode = N.ROWS_HYP_TANH_FEAT

    def __init__(self, normalize, input_dim=2048, output_dim=10, K=None):
        super().__init__(normalize, input_dim, output_dim, K)
        self.inner_radius = inner_radius(K)
        self.epsilon = 1e-5


class _JointCriterion(nn.Module):
    geom = None
    precision = ops.PREC_F64CORE
    lab_row_mode = N.ROWS_NONE
    img_row_mode = N.ROWS_NONE

    def _init(self, labelmap, neg_to_pos_ratio, feature_dict, alpha, pick_per_level, use_CNN):
        nn.Module.__init__(self)
        if use_CNN:
            raise N.LecError("use_CNN=True feeds raw images through a CNN backbone; that path is out of scope "
                             "(the backbone stays stock PyTorch) -- precompute fc7 features as the reference does")
        self.labelmap = labelmap
        self.neg_to_pos_ratio = neg_to_pos_ratio
        self.alpha = alpha
        self.device = torch.device("cuda" if torch.cuda.is_available() else "cpu")
        self.mapping_from_node_to_ix = None
        self.mapping_from_ix_to_node = None
        self.negative_G = None
        self.feature_dict = feature_dict
        self.pick_per_level = pick_per_level
        self.use_CNN = use_CNN
        self.dataloader = None
        self.levels_to_hide = []
        self._cands = None
        self._feat_rows = None   # filename -> row of the device feature matrix
        self._feat_dev = None
        self.last_negatives = None

    # ---- reference plumbing ----
    def set_levels_to_hide(self, list_of_levels):
        self.levels_to_hide = list_of_levels

    def set_dataloader(self, dataloader):
        self.dataloader = dataloader

    def set_negative_graph(self, n_G, mapping_from_node_to_ix, mapping_from_ix_to_node):
        self.negative_G = n_G
        self.mapping_from_node_to_ix = mapping_from_node_to_ix
        self.mapping_from_ix_to_node = mapping_from_ix_to_node
        self._cands = CandidateCache(n_G)
        self._graph = None  # native sampler graph, built on the first batch draw

    def get_img_features(self, x):
        """oe.py:680-707 for a flat list of filenames: [1, len(x), F] host tensor."""
        return torch.tensor(np.asarray([self.feature_dict[f] for f in x], dtype=np.float32)).unsqueeze(0)

    def feature_matrix(self, device):
        """Device-resident [n_images, F] matrix + filename->row map, built once from feature_dict (replaces
        the per-filename dict lookups + torch.tensor(list) of oe.py:687-700)."""
        if self._feat_dev is None or self._feat_dev.device != device:
            names = list(self.feature_dict.keys())
            self._feat_rows = {f: i for i, f in enumerate(names)}
            mat = np.asarray([self.feature_dict[f] for f in names], dtype=np.float32)
            self._feat_dev = torch.from_numpy(mat).to(device)
        return self._feat_dev, self._feat_rows

    # ---- sampler (oe.py:755-808) ----
    def sample_negative_edge(self, u=None, v=None, level_id=None):
        n_names = len(self.labelmap.level_names)
        if level_id is not None:
            if len(self.levels_to_hide) > 0:
                level_id = level_id % (n_names - len(self.levels_to_hide) + 1)
                level_id = list(set(list(range(n_names + 1))) - set(self.levels_to_hide))[level_id]
            else:
                level_id = level_id % (n_names + 1)
        if (u is None) == (v is None):
            raise ValueError("sample_negative_edge: exactly one of u, v must be given")
        given = u if u is not None else v
        filt, key = None, None
        if self.pick_per_level:
            if level_id < len(self.labelmap.levels):
                lo, hi = self.labelmap.level_start[level_id], self.labelmap.level_stop[level_id]
                key = level_id
                filt = lambda c: c[np.where(np.logical_and(c >= lo, c < hi))]  # noqa: E731
            else:
                cut = self.labelmap.level_stop[-1]
                if type(given) == str:
                    key, filt = "labels", (lambda c: c[np.where(c < cut)[0]])
                else:
                    key, filt = "images", (lambda c: c[np.where(c >= cut)[0]])
        u_ix = self.mapping_from_node_to_ix[u] if u is not None else None
        v_ix = self.mapping_from_node_to_ix[v] if v is not None else None
        return random.choice(self._cands.get(u_ix, v_ix, key, filt))

    # ---- energies ----
    def E_operator(self, x, y):
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

    def get_image_label_loss(self, e_for_u_v_positive, e_for_u_v_negative, weights=None):
        """oe.py:747-753."""
        if weights is None:
            return torch.sum(e_for_u_v_positive) + torch.sum(torch.clamp(self.alpha - e_for_u_v_negative, min=0.0))
        weights = torch.as_tensor(weights, device=e_for_u_v_positive.device)
        return torch.sum(weights * e_for_u_v_positive) + torch.sum(
            weights * torch.sum(torch.clamp(self.alpha - e_for_u_v_negative, min=0.0), dim=1))

    # ---- embeddings of a mixed endpoint list ----
    def _row_table(self, model, img_feat_net, endpoints_lists):
        """Rows/aux of [all labels ; the distinct images appearing in endpoints_lists] and a function mapping
        an endpoint list to row numbers."""
        m, f = unwrap(model), unwrap(img_feat_net)
        dev = m.embeddings.weight.device
        if dev.type != "cuda":
            raise N.LecError("model must live on a CUDA device (got %s); there is no CPU path" % dev)
        K = getattr(self, "K", None)
        lab_rows, lab_aux = m.rows(self.geom) if hasattr(m, "rows") else ops.transform_rows(
            m.embeddings.weight, self.lab_row_mode, K, self.geom)
        n_lab = lab_rows.shape[0]
        slot = {}
        for lst in endpoints_lists:
            for e in lst:
                if type(e) == str and e not in slot:
                    slot[e] = n_lab + len(slot)
        rows, aux = lab_rows, lab_aux
        if slot:
            feat_dev, feat_rows = self.feature_matrix(dev)
            sel = torch.as_tensor([feat_rows[e] for e in slot], device=dev)
            x = feat_dev.index_select(0, sel)
            if hasattr(f, "rows"):
                img_rows, img_aux = f.rows(x, self.geom)
            else:
                img_rows, img_aux = ops.transform_rows(f.fc1(x), self.img_row_mode, K, self.geom)
            rows = torch.cat([lab_rows, img_rows], dim=0)
            aux = torch.cat([lab_aux, img_aux], dim=0) if lab_aux.numel() else lab_aux

        def to_rows(lst):
            return np.fromiter((e if type(e) != str else slot[e] for e in lst), dtype=np.int64, count=len(lst))

        return rows, aux, to_rows, m.embedding_dim

    def calculate_from_and_to_emb(self, model, img_feat_net, from_elem, to_elem):
        """oe.py:875-964 for flat lists: embeddings of the two endpoint lists."""
        rows, _, to_rows, D = self._row_table(model, img_feat_net, (from_elem, to_elem))
        dev = rows.device
        fe = rows.index_select(0, torch.from_numpy(to_rows(from_elem)).to(dev))[:, :D]
        te = rows.index_select(0, torch.from_numpy(to_rows(to_elem)).to(dev))[:, :D]
        return fe, te

    def draw_negatives(self, original_from, original_to):
        """oe.py:846-863.  One native call that consumes Python's `random` stream exactly as the reference's loop
        (sampler.SamplerGraph.draw_exact); with hidden levels (oe.py:756-761 remaps level ids through a set) the
        per-draw path below is used, which is the same stream drawn one choice at a time."""
        Nn = self.neg_to_pos_ratio
        ix2node = self.mapping_from_ix_to_node
        if len(self.levels_to_hide) == 0:
            if getattr(self, "_graph", None) is None:
                nl = len(self.labelmap.levels)
                self._graph = SamplerGraph.from_negative_adjacency(
                    self.negative_G, level_start=list(self.labelmap.level_start)[:nl],
                    level_stop=list(self.labelmap.level_stop)[:nl], pick_per_level=self.pick_per_level,
                    n_labels=int(self.labelmap.level_stop[-1]), level_mod=len(self.labelmap.level_names) + 1)
            node2ix = self.mapping_from_node_to_ix
            nt, nf = self._graph.draw_exact([node2ix[u] for u in original_from], [node2ix[v] for v in original_to], Nn)
            return ([[ix2node[i] for i in row] for row in nt.tolist()],
                    [[ix2node[i] for i in row] for row in nf.tolist()])
        neg_to, neg_from = [], []
        for u, v in zip(original_from, original_to):
            a, b = [None] * Nn, [None] * Nn
            for p in range(Nn):
                a[p] = ix2node[self.sample_negative_edge(u=u, v=None, level_id=p)]
                b[p] = ix2node[self.sample_negative_edge(u=None, v=v, level_id=p)]
            neg_to.append(a)
            neg_from.append(b)
        return neg_to, neg_from

    def forward(self, model, img_feat_net, inputs_from, inputs_to, original_from, original_to, status, phase):
        if phase != "train":
            raise N.LecError("the joint criteria are only called in the 'train' phase by the reference's trainers "
                             "(oe.py:1489-1560: val/test go through calculate_classification_metrics -> "
                             "ops.score_topk); the nested-list eval format is not built")
        Nn = self.neg_to_pos_ratio
        B = len(original_from)
        neg_to, neg_from = self.draw_negatives(original_from, original_to)
        flat_nt = [e for row in neg_to for e in row]
        flat_nf = [e for row in neg_from for e in row]
        # the reference's flat [2N*i + p] layout, kept for inspection / parity tests
        nf = [None] * (2 * Nn * B)
        nt = [None] * (2 * Nn * B)
        for i in range(B):
            nf[2 * Nn * i:2 * Nn * i + Nn] = [original_from[i]] * Nn
            nt[2 * Nn * i:2 * Nn * i + Nn] = neg_to[i]
            nf[2 * Nn * i + Nn:2 * Nn * (i + 1)] = neg_from[i]
            nt[2 * Nn * i + Nn:2 * Nn * (i + 1)] = [original_to[i]] * Nn
        self.last_negatives = (nf, nt)
        rows, aux, to_rows, D = self._row_table(model, img_feat_net,
                                                (original_from, original_to, flat_nt, flat_nf))
        dev = rows.device
        idx = [torch.from_numpy(to_rows(l)).to(dev, non_blocking=True)
               for l in (original_from, original_to, flat_nt, flat_nf)]
        loss, E_pos, E_neg = ops.grouped_pair_loss(rows, aux, D, idx[0], idx[1], idx[2].view(B, Nn), idx[3].view(B, Nn),
                                                   Nn, self.geom, getattr(self, "K", None), self.alpha,
                                                   precision=self.precision)
        return loss, E_pos, E_neg.view(B, 2 * Nn, 1)


class EuclideanConesWithImagesHypernymLoss(_JointCriterion):
    """oe.py:650-964."""
    geom = "euc"
    lab_row_mode = N.ROWS_EUC_SOFTCLIP
    img_row_mode = N.ROWS_EUC_SOFTCLIP

    def __init__(self, labelmap, neg_to_pos_ratio, feature_dict, alpha, pick_per_level=False, K=3.0, use_CNN=False):
        self._init(labelmap, neg_to_pos_ratio, feature_dict, alpha, pick_per_level, use_CNN)
        self.K = K
        self.epsilon = 1e-5


class OrderEmbeddingWithImagesHypernymLoss(_JointCriterion):
    """oe.py:967-1221."""
    geom = "oe"

    def __init__(self, labelmap, neg_to_pos_ratio, feature_dict, alpha, pick_per_level=False, use_CNN=False):
        self._init(labelmap, neg_to_pos_ratio, feature_dict, alpha, pick_per_level, use_CNN)


class HyperbolicConesWithImagesHypernymLoss(_JointCriterion):
    """oe_h.py:739-1058 (named EuclideanConesWithImagesHypernymLoss there)."""
    geom = "hyp"
    lab_row_mode = N.ROWS_HYP_TANH
    img_row_mode = N.ROWS_HYP_TANH_FEAT

    def __init__(self, labelmap, neg_to_pos_ratio, feature_dict, alpha, pick_per_level=False, K=0.1, use_CNN=False):
        self._init(labelmap, neg_to_pos_ratio, feature_dict, alpha, pick_per_level, use_CNN)
        self.K = K
        self.inner_radius = inner_radius(K)
        self.epsilon = 1e-5


__all__ = ["EuclideanFeatNet", "HyperbolicFeatNet", "EuclideanConesWithImagesHypernymLoss",
           "OrderEmbeddingWithImagesHypernymLoss", "HyperbolicConesWithImagesHypernymLoss",
           "EuclideanEmbedder", "HyperbolicTanhEmbedder"]
