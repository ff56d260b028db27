"""Negative-edge sampler of the cone losses on the native library (include/lec_b200.h, lec_sampler.cu).

Replaces the reference's per-draw `np.where(negative_G[row]) -> .tolist() -> random.choice` and the Python
B x N x 2 loop around it (order_embeddings.py:989-1008, :1070-1091; joint oe.py:755-808, :846-863 -- about 75 %
of a reference training step).

* `SamplerGraph` holds the two CSR lists of EXCLUDED nodes (a node plus its closure descendants / ancestors)
  that stand in for the dense n x n `negative_G`; it is built either from closure edges or from a dense
  adjacency as the reference's `set_negative_graph` receives it.
* `draw_exact` consumes CPython's `random` stream exactly like the reference's random.choice calls (the
  global state is loaded into the native Mersenne Twister and stored back), so indices are bit-exact and
  single `sample_negative_edge` calls can be interleaved with it freely.
* `draw_philox` is the device fast mode: same candidate sets, same uniform law, counter-based Philox stream.
"""
import ctypes
import random

import numpy as np
import torch

from . import _native as N


class LecSamplerGraph(ctypes.Structure):
    """lec_sampler_graph of include/lec_b200.h."""
    _fields_ = [
        ("n_nodes", ctypes.c_int64),
        ("row_excl_ptr", ctypes.c_void_p), ("row_excl", ctypes.c_void_p),
        ("col_excl_ptr", ctypes.c_void_p), ("col_excl", ctypes.c_void_p),
        ("pick_per_level", ctypes.c_int), ("n_levels", ctypes.c_int), ("level_mod", ctypes.c_int),
        ("level_start", ctypes.c_int32 * 8), ("level_stop", ctypes.c_int32 * 8),
        ("n_labels", ctypes.c_int64),
    ]


class LecMT19937(ctypes.Structure):
    """lec_mt19937: the 624 state words + position of CPython's generator."""
    _fields_ = [("mt", ctypes.c_uint32 * 624), ("index", ctypes.c_int32)]


def mt_from_python(state=None):
    """random.getstate() -> native state."""
    st = random.getstate() if state is None else state
    if st[0] != 3 or len(st[1]) != 625:
        raise N.LecError("unexpected random.getstate() layout (version %r)" % (st[0],))
    s = LecMT19937()
    s.mt[:] = st[1][:624]
    s.index = st[1][624]
    return s, st[2]


def mt_to_python(s, gauss_next=None):
    """native state -> tuple accepted by random.setstate."""
    return (3, tuple(int(w) for w in s.mt) + (int(s.index),), gauss_next)


def mt_seeded(a):
    """Native generator in the state random.seed(a) (a an int) would leave."""
    a = abs(int(a))
    key = []
    while a:
        key.append(a & 0xFFFFFFFF)
        a >>= 32
    key = key or [0]
    s = LecMT19937()
    N.check(N.lib().lec_mt_seed(ctypes.byref(s), (ctypes.c_uint32 * len(key))(*key), len(key)), "lec_mt_seed")
    return s


def _csr(owner, member, n):
    """CSR of `member` grouped by `owner` (both int arrays), members ascending within a group."""
    o = np.lexsort((member, owner))
    ptr = np.zeros(n + 1, dtype=np.int64)
    np.cumsum(np.bincount(owner, minlength=n), out=ptr[1:])
    return ptr, np.ascontiguousarray(member[o].astype(np.int32))


class SamplerGraph:
    """Candidate sets of the negative sampler: [0, n) minus a sorted excluded list per node."""

    def __init__(self, n_nodes, row_csr, col_csr, level_start=None, level_stop=None, pick_per_level=False, n_labels=0,
                 level_mod=None):
        """row_csr / col_csr = (ptr int64[n+1], excluded int32[...]) with each node's excluded ids ascending."""
        self.n = int(n_nodes)
        self.row_ptr, self.row_excl = row_csr
        self.col_ptr, self.col_excl = col_csr
        self.level_start = [int(v) for v in (level_start if level_start is not None else [])]
        self.level_stop = [int(v) for v in (level_stop if level_stop is not None else [])]
        self.pick_per_level = bool(pick_per_level)
        self.n_labels = int(n_labels)
        self.level_mod = int(level_mod if level_mod is not None else max(1, len(self.level_start)))
        if len(self.level_start) > 8:
            raise N.LecError("at most 8 label levels")
        self._host = self._struct(self.row_ptr.ctypes.data, self.row_excl.ctypes.data, self.col_ptr.ctypes.data,
                                  self.col_excl.ctypes.data)
        self._dev = {}

    @classmethod
    def from_closure(cls, n_nodes, anc, desc, **kw):
        """anc[i] -> desc[i] are the edges of the transitive closure (node indices in [0, n_nodes)): row u
        excludes u and its descendants, column v excludes v and its ancestors."""
        n = int(n_nodes)
        anc = np.asarray(anc, dtype=np.int64).reshape(-1)
        desc = np.asarray(desc, dtype=np.int64).reshape(-1)
        me = np.arange(n, dtype=np.int64)
        return cls(n, _csr(np.concatenate([anc, me]), np.concatenate([desc, me]), n),
                   _csr(np.concatenate([desc, me]), np.concatenate([anc, me]), n), **kw)

    @classmethod
    def from_negative_adjacency(cls, n_G, **kw):
        """From the dense matrix the reference passes to set_negative_graph (order_embeddings.py:417-423,
        oe.py:465-474): the candidates of row u are np.where(n_G[u, :] == 1), so its excluded list is every
        other column; likewise per column."""
        A = np.asarray(n_G)
        n = A.shape[0]
        r, c = np.nonzero(A != 1)
        return cls(n, _csr(r, c, n), _csr(c, r, n), **kw)

    @classmethod
    def from_hierarchy(cls, h, **kw):
        e = h.closure_edges()
        kw.setdefault("level_start", getattr(h, "level_start", None))
        kw.setdefault("level_stop", getattr(h, "level_stop", None))
        return cls.from_closure(h.n, e[:, 0], e[:, 1], **kw)

    def _struct(self, rp, re, cp, ce):
        g = LecSamplerGraph()
        g.n_nodes = self.n
        g.row_excl_ptr, g.row_excl, g.col_excl_ptr, g.col_excl = rp, re, cp, ce
        g.pick_per_level = int(self.pick_per_level)
        g.n_levels = len(self.level_start)
        g.level_mod = self.level_mod
        for i, (a, b) in enumerate(zip(self.level_start, self.level_stop)):
            g.level_start[i], g.level_stop[i] = a, b
        g.n_labels = self.n_labels
        return g

    # ---- exact mode ------------------------------------------------------------------------------
    def draw_exact(self, u_ix, v_ix, n_neg, rng=None):
        """(neg_to [B, n_neg], neg_from [B, n_neg]) int64 node indices, consuming the Mersenne-Twister stream
        bit-exactly like the reference's loop.  rng=None uses (and advances) Python's global `random`."""
        u = np.ascontiguousarray(u_ix, dtype=np.int64)
        v = np.ascontiguousarray(v_ix, dtype=np.int64)
        B = len(u)
        neg_to = np.empty((B, n_neg), dtype=np.int64)
        neg_from = np.empty((B, n_neg), dtype=np.int64)
        use_global = rng is None
        gauss = None
        if use_global:
            rng, gauss = mt_from_python()
        code = N.lib().lec_sample_negatives(ctypes.byref(rng), ctypes.byref(self._host), u.ctypes.data, v.ctypes.data,
                                            B, int(n_neg), neg_to.ctypes.data, neg_from.ctypes.data)
        if use_global:
            random.setstate(mt_to_python(rng, gauss))
        if code == -9:  # LEC_E_EMPTY
            raise IndexError("Cannot choose from an empty sequence")  # what random.choice([]) raises
        N.check(code, "lec_sample_negatives")
        return neg_to, neg_from

    # ---- fast mode -------------------------------------------------------------------------------
    def device_struct(self, dev):
        key = str(dev)
        hit = self._dev.get(key)
        if hit is None:
            t = [torch.from_numpy(a).to(dev) for a in (self.row_ptr, self.row_excl, self.col_ptr, self.col_excl)]
            status = torch.zeros(1, dtype=torch.int32, device=dev)
            hit = (self._struct(*[x.data_ptr() for x in t]), t, status)
            self._dev[key] = hit
        return hit

    def draw_philox(self, u_dev, v_dev, n_neg, seed, step, out=None, check=True):
        """Device draw: u_dev, v_dev are CUDA index tensors (uint16 / int32 / int64); returns (neg_to, neg_from)
        of the same dtype, shape [B, n_neg].  Stream = (seed, step); a given (seed, step, i, p, side) always
        yields the same draw regardless of batch size or launch geometry."""
        N.require_cuda(u_dev, v_dev)
        if u_dev.dtype != v_dev.dtype or u_dev.dtype not in (torch.uint16, torch.int32, torch.int64):
            raise N.LecError("index tensors must both be uint16, int32 or int64")
        g, _keep, status = self.device_struct(u_dev.device)
        B = u_dev.numel()
        if out is None:
            out = (torch.empty((B, n_neg), dtype=u_dev.dtype, device=u_dev.device),
                   torch.empty((B, n_neg), dtype=u_dev.dtype, device=u_dev.device))
        N.check(N.lib().lec_sample_negatives_philox(ctypes.byref(g), N._p(u_dev), N._p(v_dev), u_dev.element_size(), B,
                                                    int(n_neg), int(seed), int(step), N._p(out[0]), N._p(out[1]),
                                                    N._p(status), N.stream_ptr(u_dev.device)),
                "lec_sample_negatives_philox")
        if check:
            code = int(status.item())
            if code:
                status.zero_()
                if code == -9:
                    raise IndexError("Cannot choose from an empty sequence")
                N.check(code, "lec_sample_negatives_philox")
        return out
