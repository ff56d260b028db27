"""FeatNet.fc1 kernels (lec_featnet_fwd / lec_featnet_wgrad: gather fused into the projection and its weight gradient,
oe.py:97,113,680-707) through the C ABI against a plain torch reference of the same op (fp64 accumulation), over the
shapes the launcher dispatches on (output pairs NP = 1..8, one or two column chunks per thread, ragged row counts).
Tolerance: fp32 sums of F = 2048 products in a different order -> 2e-5 relative to the row / column scale."""
import numpy as np
import pytest
import torch

pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs CUDA")]

if torch.cuda.is_available():
    from learning_embeddings_b200 import _native as N

DEV = "cuda"


def _run(n_pool, F, D, m, sel_dtype, replicas, bad_rows=(), seed=0):
    g = torch.Generator().manual_seed(seed)
    X = torch.randn(n_pool, F, generator=g)
    W = torch.randn(D, F, generator=g) / F ** 0.5
    b = torch.randn(D, generator=g)
    gY = torch.randn(m, D, generator=g)
    if sel_dtype is None:
        sel = None
        idx = torch.arange(m)
    else:
        idx = torch.randint(0, n_pool, (m,), generator=g)
        sel = idx.to(sel_dtype).clone()
        for r in bad_rows:
            sel[r] = n_pool + 5          # outside the pool: counted, contributes zeros
    Xd, Wd, bd, gYd = X.to(DEV), W.to(DEV), b.to(DEV), gY.to(DEV)
    seld = sel.to(DEV) if sel is not None else None
    Y = torch.empty(m, D, device=DEV)
    stride = (D * F + D + 3) // 4 * 4
    grad = torch.zeros(replicas, stride, device=DEV)
    lib, st = N.lib(), N.stream_ptr(torch.device(DEV))
    N.index_errors(torch.device(DEV))
    sb = 0 if sel is None else sel.element_size()
    N.check(lib.lec_featnet_fwd(N._p(Xd), n_pool, F, N._p(seld), sb, m, N._p(Wd), N._p(bd), D, N._p(Y), st), "lec_featnet_fwd")
    N.check(lib.lec_featnet_wgrad(N._p(Xd), n_pool, F, N._p(seld), sb, m, N._p(gYd), D, N._p(grad), replicas, stride, st),
            "lec_featnet_wgrad")
    torch.cuda.synchronize()
    Xs = X[idx].clone()
    for r in bad_rows:
        Xs[r] = 0.0
    Y_ref = Xs.double() @ W.double().T + b.double()
    dW_ref = gY.double().T @ Xs.double()
    db_ref = gY.double().sum(0)
    got = grad.sum(0).cpu().double()
    np.testing.assert_allclose(Y.cpu().double().numpy(), Y_ref.numpy(), rtol=0, atol=2e-5 * float(Y_ref.abs().max()))
    np.testing.assert_allclose(got[:D * F].view(D, F).numpy(), dW_ref.numpy(), rtol=0, atol=2e-5 * float(dW_ref.abs().max()))
    np.testing.assert_allclose(got[D * F:D * F + D].numpy(), db_ref.numpy(), rtol=0, atol=2e-5 * float(db_ref.abs().max()))
    # each out-of-range id is counted once per pass
    assert N.index_errors(torch.device(DEV)) == 2 * len(bad_rows)


@pytest.mark.parametrize("F,D,m,sel_dtype,replicas", [
    (2048, 10, 16384 // 8, torch.int64, 8),     # cfg2 shape (an eighth of its rows), NP = 5
    (2048, 10, 1237, torch.int32, 1),           # ragged row count: CTAs with different ranges, partial last item
    (2048, 7, 600, torch.int64, 4),             # odd D: NP = 4, one padded output
    (516, 3, 300, torch.int32, 2),              # 129 chunks per row
    (1024, 1, 257, torch.int64, 1),             # D = 1
    (2048, 10, 700, None, 2),                   # identity selection
    (2048, 10, 40, torch.int64, 2),             # a handful of rows
    (4096, 10, 300, torch.int64, 2),            # two column chunks per thread in the weight-gradient kernel
    (2048, 16, 300, torch.int64, 2),            # NP = 8
])
def test_featnet_kernels_match_torch(F, D, m, sel_dtype, replicas):
    _run(n_pool=max(m, 900), F=F, D=D, m=m, sel_dtype=sel_dtype, replicas=replicas)


def test_featnet_out_of_range_ids_are_counted_and_read_nothing():
    _run(n_pool=900, F=2048, D=10, m=500, sel_dtype=torch.int64, replicas=2, bad_rows=(0, 77, 499))
    _run(n_pool=900, F=2048, D=10, m=30, sel_dtype=torch.int64, replicas=2, bad_rows=(3,))
