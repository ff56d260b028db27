"""Oracle (TEST INFRASTRUCTURE): the reference's negative-edge sampler restated.

The reference draws negatives with CPython's `random.choice` over
`np.where(negative_G[row, :] == 1)[0]` (order_embeddings.py:989-1008; joint variant
oe.py:755-808).  "Bit-exact negative indices" therefore means reproducing CPython's
Mersenne Twister stream.  That algorithm lives in CPython (Modules/_randommodule.c,
Lib/random.py, version 3.12 here), not in /root/reference; it is restated below from its
published definition (Matsumoto & Nishimura MT19937, `init_by_array`, and
`Random._randbelow_with_getrandbits`) and pinned against `random` itself through
tests/golden/mt_choice_streams.npz.
"""
import numpy as np

_N, _M = 624, 397
_MASK32 = 0xFFFFFFFF


class MT19937:
    """CPython-compatible generator: seed(int) == random.seed(int); choice_index == random.choice index."""

    def __init__(self, seed=0):
        self.mt = [0] * _N
        self.idx = _N
        self.seed(seed)

    def _init_genrand(self, s):
        mt = self.mt
        mt[0] = s & _MASK32
        for i in range(1, _N):
            mt[i] = (1812433253 * (mt[i - 1] ^ (mt[i - 1] >> 30)) + i) & _MASK32
        self.idx = _N

    def seed(self, a):
        """random.seed(int): key = little-endian 32-bit words of |a| ([0] for 0), then init_by_array."""
        a = abs(int(a))
        key = []
        while a:
            key.append(a & _MASK32)
            a >>= 32
        if not key:
            key = [0]
        self._init_genrand(19650218)
        mt = self.mt
        i, j = 1, 0
        for _ in range(max(_N, len(key))):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1664525)) + key[j] + j) & _MASK32
            i += 1
            j += 1
            if i >= _N:
                mt[0] = mt[_N - 1]
                i = 1
            if j >= len(key):
                j = 0
        for _ in range(_N - 1):
            mt[i] = ((mt[i] ^ ((mt[i - 1] ^ (mt[i - 1] >> 30)) * 1566083941)) - i) & _MASK32
            i += 1
            if i >= _N:
                mt[0] = mt[_N - 1]
                i = 1
        mt[0] = 0x80000000
        self.idx = _N

    def _twist(self):
        mt = self.mt
        for k in range(_N):
            y = (mt[k] & 0x80000000) | (mt[(k + 1) % _N] & 0x7FFFFFFF)
            mt[k] = mt[(k + _M) % _N] ^ (y >> 1) ^ (0x9908B0DF if y & 1 else 0)
        self.idx = 0

    def uint32(self):
        if self.idx >= _N:
            self._twist()
        y = self.mt[self.idx]
        self.idx += 1
        y ^= y >> 11
        y ^= (y << 7) & 0x9D2C5680
        y ^= (y << 15) & 0xEFC60000
        y ^= y >> 18
        return y & _MASK32

    def randbelow(self, n):
        """Random._randbelow_with_getrandbits for 0 < n < 2**32: rejection on the top k bits."""
        if n <= 0:
            raise IndexError("Cannot choose from an empty sequence")
        k = int(n).bit_length()
        r = self.uint32() >> (32 - k)
        while r >= n:
            r = self.uint32() >> (32 - k)
        return r


def candidates_dense(neg_adj, u_ix=None, v_ix=None):
    """order_embeddings.py:993-996: ascending node indices with a 1 in row u (or column v)."""
    if u_ix is not None:
        return np.where(neg_adj[u_ix, :] == 1)[0]
    return np.where(neg_adj[:, v_ix] == 1)[0]


def filter_level(cands, level_id, n_level_names, level_start, level_stop, pick_per_level):
    """order_embeddings.py:990-1006 (label-only): level_id % n_levels, keep candidates inside the level."""
    if not pick_per_level:
        return cands
    level_id = level_id % n_level_names
    lo, hi = level_start[level_id], level_stop[level_id]
    return cands[(cands >= lo) & (cands < hi)]


def draw_step_negatives(rng, neg_adj, u_list, v_list, N, level_start, level_stop, pick_per_level=False):
    """order_embeddings.py:1070-1091: for each positive i, for p in range(N): corrupt v, then corrupt u.

    Returns (neg_from, neg_to, drawn) with the reference's [2N*i + p] / [2N*i + N + p] layout."""
    B = len(u_list)
    neg_from = np.empty(2 * N * B, dtype=np.int64)
    neg_to = np.empty(2 * N * B, dtype=np.int64)
    drawn = []
    nl = len(level_start)
    for i in range(B):
        u, v = int(u_list[i]), int(v_list[i])
        for p in range(N):
            c = filter_level(candidates_dense(neg_adj, u_ix=u), p, nl, level_start, level_stop, pick_per_level)
            ix = int(c[rng.randbelow(len(c))])
            drawn.append(ix)
            neg_from[2 * N * i + p] = u
            neg_to[2 * N * i + p] = ix
            c = filter_level(candidates_dense(neg_adj, v_ix=v), p, nl, level_start, level_stop, pick_per_level)
            ix = int(c[rng.randbelow(len(c))])
            drawn.append(ix)
            neg_from[2 * N * i + N + p] = ix
            neg_to[2 * N * i + N + p] = v
    return neg_from, neg_to, np.array(drawn, dtype=np.int64)


def filter_level_joint(cands, level_id, n_level_names, level_start, level_stop, pick_per_level, endpoint_is_image):
    """oe.py:786-804 (levels_to_hide empty): level_id % (n_levels+1); the extra slot is the image level.

    For the extra slot the reference keeps label candidates (< level_stop[-1]) when the fixed endpoint
    is an image, otherwise image candidates (>= level_stop[-1])."""
    if not pick_per_level:
        return cands
    level_id = level_id % (n_level_names + 1)
    if level_id < len(level_start):
        lo, hi = level_start[level_id], level_stop[level_id]
        return cands[(cands >= lo) & (cands < hi)]
    cut = level_stop[-1]
    return cands[cands < cut] if endpoint_is_image else cands[cands >= cut]


def draw_step_negatives_joint(rng, neg_adj, u_list, v_list, N, level_start, level_stop, n_labels,
                              pick_per_level=False):
    """oe.py:846-863 with node indices already mapped (labels < n_labels <= images)."""
    B = len(u_list)
    neg_from = np.empty(2 * N * B, dtype=np.int64)
    neg_to = np.empty(2 * N * B, dtype=np.int64)
    drawn = []
    nl = len(level_start)
    for i in range(B):
        u, v = int(u_list[i]), int(v_list[i])
        for p in range(N):
            c = filter_level_joint(candidates_dense(neg_adj, u_ix=u), p, nl, level_start, level_stop,
                                   pick_per_level, u >= n_labels)
            ix = int(c[rng.randbelow(len(c))])
            drawn.append(ix)
            neg_from[2 * N * i + p] = u
            neg_to[2 * N * i + p] = ix
            c = filter_level_joint(candidates_dense(neg_adj, v_ix=v), p, nl, level_start, level_stop,
                                   pick_per_level, v >= n_labels)
            ix = int(c[rng.randbelow(len(c))])
            drawn.append(ix)
            neg_from[2 * N * i + N + p] = ix
            neg_to[2 * N * i + N + p] = v
    return neg_from, neg_to, np.array(drawn, dtype=np.int64)
