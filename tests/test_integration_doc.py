"""The ctypes example printed in INTEGRATION.md section 3 is executed here, so the document cannot drift from the ABI
again (round-1 finding: it bound lec_rows_fwd with one argument too few).  Without a GPU the binding part runs and its
argtypes / struct layout are compared with learning_embeddings_b200/_native.py; on a GPU box the example's training
iteration also runs and is compared with engine.ConeStep."""
import ctypes
import os
import re

import numpy as np
import pytest
import torch

from conftest import ROOT
from learning_embeddings_b200 import _native


def doc_namespace():
    text = open(os.path.join(ROOT, "INTEGRATION.md")).read()
    m = re.search(r"<!-- doc-test: begin -->\s*```python\n(.*?)```\s*<!-- doc-test: end -->", text, re.S)
    assert m, "INTEGRATION.md lost its doc-test block"
    src = m.group(1).replace('"learning_embeddings_b200/_lib/liblec_b200.so"', repr(_native.LIB_PATH))
    ns = {}
    exec(compile(src, "INTEGRATION.md", "exec"), ns)
    return ns


def test_doc_binding_matches_native_binding():
    ns = doc_namespace()
    native = _native.lib()
    for fn in ("lec_rows_fwd", "lec_pairs_grouped"):
        assert list(getattr(ns["lib"], fn).argtypes) == list(getattr(native, fn).argtypes), fn
    doc_fields = [(n, t) for n, t in ns["lec_update_t"]._fields_]
    assert doc_fields == [(n, t) for n, t in _native.LecUpdate._fields_]
    assert ctypes.sizeof(ns["lec_update_t"]) == ctypes.sizeof(_native.LecUpdate)
    # the header declares the same number of parameters as the example binds
    hdr = open(os.path.join(ROOT, "include", "lec_b200.h")).read()
    for fn in ("lec_rows_fwd", "lec_pairs_grouped"):
        decl = re.search(r"int %s\((.*?)\);" % fn, hdr, re.S).group(1)
        assert len(decl.split(",")) == len(getattr(ns["lib"], fn).argtypes), fn


@pytest.mark.gpu
def test_doc_training_iteration_equals_engine():
    from learning_embeddings_b200.engine import ConeStep, pack_index_block
    from learning_embeddings_b200 import hierarchy
    from oracle import cones
    ns = doc_namespace()
    h = hierarchy.ethec()
    rng = np.random.default_rng(0)
    B, Nn, D = 2048, 5, 10
    edges = h.closure_edges()
    g = torch.Generator().manual_seed(0)
    w = torch.randn(h.n, D, generator=g)
    r_in = cones.inner_radius(0.1)
    W0 = (r_in + 0.05 * torch.rand(h.n, 1, generator=g)) * w / w.norm(dim=1, keepdim=True)
    Wa, Wb = W0.cuda().clone(), W0.cuda().clone()
    eng = ConeStep(Wb, "hyp", Nn, B, K=0.1, alpha=0.05, lr=0.01)
    state = {}
    for _ in range(3):
        sel = rng.integers(0, len(edges), size=B)
        u, v = edges[sel, 0], edges[sel, 1]
        nt, nf = h.sample_negatives(u, v, Nn, rng)
        dev = [torch.from_numpy(np.ascontiguousarray(x, dtype=np.int32)).cuda() for x in (u, v, nt, nf)]
        loss, E_pos, E_neg = ns["poincare_step"](Wa, *dev, Nn, 0.05, 0.01, r_in, state)
        blk = pack_index_block(u, v, nt, nf).cuda()
        eng.step_device(*eng._split(blk, B))
        torch.cuda.synchronize()
        assert abs(float(loss) - float(eng.loss)) <= 1e-6 * abs(float(eng.loss))
        # (the two tables drift apart by fp32 rounding after the first step: the gradient scatter is not order-deterministic)
        np.testing.assert_allclose(E_neg.cpu().numpy(), eng.E_neg[:B].cpu().numpy(), rtol=1e-4, atol=1e-5)
        np.testing.assert_allclose(Wa.cpu().numpy(), Wb.cpu().numpy(), rtol=1e-4, atol=2e-5)
