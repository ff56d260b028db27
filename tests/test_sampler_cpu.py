"""Native negative sampler (lec_sampler.cu, host exact mode) against CPython's `random`, the oracle restatement
and the `drawn` streams recorded from the unmodified reference (tests/golden/step_*.npz, joint_*.npz)."""
import ctypes
import os
import random
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from learning_embeddings_b200 import _native as N, sampler as S, hierarchy as H  # noqa: E402
from oracle import sampler as osampler  # noqa: E402

GOLD = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")


def gold(name):
    with np.load(os.path.join(GOLD, name + ".npz"), allow_pickle=False) as z:
        return {k: z[k] for k in z.files}


@pytest.mark.parametrize("seed", [0, 1, 12345, 2 ** 40 + 7, 2 ** 70 + 3])
def test_mt_seed_and_stream_match_cpython(seed):
    s = S.mt_seeded(seed)
    r = random.Random(seed)
    lib = N.lib()
    assert [lib.lec_mt_uint32(ctypes.byref(s)) for _ in range(1500)] == [r.getrandbits(32) for _ in range(1500)]
    for n in (1, 2, 3, 7, 561, 716, 722, 1 << 20, (1 << 31) + 5):
        assert lib.lec_mt_randbelow(ctypes.byref(s), n) == r._randbelow(n)
    assert lib.lec_mt_randbelow(ctypes.byref(s), 0) == -9


def test_golden_choice_streams():
    g = gold("mt_choice_streams")
    ns = [int(v) for v in g["ns"]]
    lib = N.lib()
    for key in g:
        if key.startswith("seed_"):
            s = S.mt_seeded(int(key[5:]))
            got = [lib.lec_mt_randbelow(ctypes.byref(s), m) for _ in range(40) for m in ns]
            assert got == g[key].tolist(), key


def test_state_round_trip_with_python_random():
    random.seed(99)
    random.random()
    s, gauss = S.mt_from_python()
    a = [N.lib().lec_mt_uint32(ctypes.byref(s)) for _ in range(700)]
    b = [random.getrandbits(32) for _ in range(700)]
    assert a == b
    random.setstate(S.mt_to_python(s, gauss))
    assert N.lib().lec_mt_uint32(ctypes.byref(s)) == random.getrandbits(32)


def _flat(u, v, neg_to, neg_from):
    """the reference's draw order: per positive, per p: corrupted child, corrupted parent."""
    return np.stack([neg_to, neg_from], axis=2).reshape(-1)


STEPS = ["step_euc_D2", "step_euc_D10_ppl", "step_hyp_D10", "step_hyp_D50_ppl", "step_oe_D10", "step_oe_D10_weighted"]


@pytest.mark.parametrize("name", STEPS)
def test_label_only_draws_equal_reference_stream(name):
    g = gold(name)
    h = H.ethec()
    ppl = bool(g["pick_per_level"])
    graph = S.SamplerGraph.from_hierarchy(h, pick_per_level=ppl)
    random.seed(0)
    neg_to, neg_from = graph.draw_exact(g["u"], g["v"], int(g["N"]))
    assert _flat(g["u"], g["v"], neg_to, neg_from).tolist() == g["drawn"].tolist()
    # the global stream was advanced exactly as the reference's loop advances it
    rng = osampler.MT19937(0)
    osampler.draw_step_negatives(rng, h.negative_adjacency(), g["u"], g["v"], int(g["N"]), h.level_start, h.level_stop, ppl)
    assert random.getrandbits(32) == rng.uint32()
    # same graph from the dense adjacency the reference hands to set_negative_graph
    g2 = S.SamplerGraph.from_negative_adjacency(h.negative_adjacency(), level_start=h.level_start,
                                                level_stop=h.level_stop, pick_per_level=ppl)
    random.seed(0)
    nt2, nf2 = g2.draw_exact(g["u"], g["v"], int(g["N"]))
    assert np.array_equal(nt2, neg_to) and np.array_equal(nf2, neg_from)


@pytest.mark.parametrize("name", ["joint_euc", "joint_euc_ppl", "joint_oe", "joint_hyp"])
def test_joint_draws_equal_reference_stream(name):
    g = gold(name)
    nn, n_lab = int(g["n_nodes"]), int(g["n_lab"])
    A = np.unpackbits(g["neg_adj"], axis=1)[:, :nn].astype(bool)
    ls, le = [int(v) for v in g["level_start"]], [int(v) for v in g["level_stop"]]
    graph = S.SamplerGraph.from_negative_adjacency(A, level_start=ls, level_stop=le, pick_per_level=bool(g["pick_per_level"]),
                                                   n_labels=n_lab, level_mod=len(ls) + 1)
    rng = S.mt_seeded(0)
    neg_to, neg_from = graph.draw_exact(g["b_from"], g["b_to"], int(g["N"]), rng=rng)
    assert _flat(g["b_from"], g["b_to"], neg_to, neg_from).tolist() == g["drawn"].tolist()


def test_random_forest_against_oracle_all_windows():
    h = H.random_tree(400, 1.0831, seed=3, roots=5)
    A = h.negative_adjacency()
    e = h.closure_edges()
    level_start, level_stop = [0, 5, 60, 200], [5, 60, 200, 400]
    for ppl in (False, True):
        graph = S.SamplerGraph.from_closure(h.n, e[:, 0], e[:, 1], level_start=level_start, level_stop=level_stop,
                                            pick_per_level=ppl)
        sel = np.random.default_rng(0).integers(0, len(e), 300)
        u, v = e[sel, 0], e[sel, 1]
        rng = S.mt_seeded(7)
        try:
            nt, nf = graph.draw_exact(u, v, 6, rng=rng)
            got = _flat(u, v, nt, nf).tolist()
        except IndexError:
            got = "IndexError"
        orng = osampler.MT19937(7)
        try:
            _, _, drawn = osampler.draw_step_negatives(orng, A, u, v, 6, level_start, level_stop, ppl)
            want = drawn.tolist()
        except IndexError:
            want = "IndexError"
        assert got == want


def test_empty_candidate_list_raises_like_random_choice():
    # a single chain 0 -> 1 -> 2: row 0 excludes everything
    graph = S.SamplerGraph.from_closure(3, [0, 0, 1], [1, 2, 2])
    with pytest.raises(IndexError):
        graph.draw_exact([0], [1], 1, rng=S.mt_seeded(0))
    with pytest.raises(N.LecError):
        graph.draw_exact([5], [1], 1, rng=S.mt_seeded(0))


def test_philox_host_draw_is_uniform_and_in_range():
    lib = N.lib()
    n = 7
    c = np.bincount([lib.lec_philox_below(11, 3, d, n) for d in range(7000)], minlength=n)
    assert c.min() > 850 and c.max() < 1150
    assert lib.lec_philox_below(11, 3, 5, n) == lib.lec_philox_below(11, 3, 5, n)
    assert lib.lec_philox_below(1, 2, 3, 0) == -9


class _LabelMap:
    def __init__(self, h):
        self.level_start, self.level_stop, self.levels = h.level_start, h.level_stop, h.levels
        self.level_names = ["family", "subfamily", "genus", "genus_specific_epithet"]
        self.n_classes = h.n


@pytest.mark.parametrize("ppl", [False, True])
def test_criterion_batch_draw_equals_per_draw_calls_and_interleaves(ppl):
    """draw_negatives (one native call) == the reference-style loop of sample_negative_edge (random.choice), and the
    global `random` stream stays in step so both can be mixed."""
    from learning_embeddings_b200.criterion import HypConesLoss
    h = H.ethec()
    crit = HypConesLoss(_LabelMap(h), 3, alpha=0.05, pick_per_level=ppl)
    ident = {i: i for i in range(h.n)}
    crit.set_negative_graph(h.negative_adjacency(), ident, ident)
    e = h.closure_edges()[::37]
    u, v = e[:, 0].tolist(), e[:, 1].tolist()
    random.seed(5)
    nt, nf = crit.draw_negatives(u, v)
    single = crit.sample_negative_edge(u=u[0], v=None, level_id=0)
    random.seed(5)
    want_t = np.empty_like(nt)
    want_f = np.empty_like(nf)
    for i in range(len(u)):
        for p in range(3):
            want_t[i, p] = crit.sample_negative_edge(u=u[i], v=None, level_id=p)
            want_f[i, p] = crit.sample_negative_edge(u=None, v=v[i], level_id=p)
    assert np.array_equal(nt, want_t) and np.array_equal(nf, want_f)
    assert single == crit.sample_negative_edge(u=u[0], v=None, level_id=0)
