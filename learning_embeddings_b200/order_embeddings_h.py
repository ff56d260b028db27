"""Drop-in names of network/order_embeddings_h.py (label-only Poincare cones + RSGD).

The reference keeps the class name `EucConesLoss` for the hyperbolic loss (order_embeddings_h.py:1072).

The trainer-side helpers `soft_clip`, `mob_add`, `lambda_x`, `exp_map_x` (order_embeddings_h.py:634-674) are kept
as plain tensor expressions for callers that use them one by one (visualisation, notebooks); they are NOT the training
path -- `rsgd_step` is, and it runs the whole update (rescale, exp-map, Moebius add, shell projection) as one CUDA
kernel (lec_rsgd_update / lec_rsgd_update_rows).  `exp_map_x(W, -lr * grad * (1 / lambda_x(W)) ** 2)` is what that
kernel computes; tests/test_oracle_golden.py holds the helpers to the oracle's restatement of the reference update."""
import torch

from .criterion import HyperbolicEmbedder as Embedder  # order_embeddings_h.py:181
from .criterion import HypConesLoss as EucConesLoss  # order_embeddings_h.py:1072
from .criterion import OrderEmbeddingLoss, inner_radius, rsgd_step  # :906, :1089, :764-775

__all__ = ["Embedder", "EucConesLoss", "OrderEmbeddingLoss", "inner_radius", "rsgd_step", "soft_clip", "mob_add",
           "lambda_x", "exp_map_x"]


def soft_clip(x, r_in):
    """order_embeddings_h.py:634-647: rows with |x| <= r_in are scaled onto the inner shell, rows with |x| >= 1 onto
    radius 1 - 1e-5 (no gradient through the projection).  Returns a new tensor (the reference writes in place)."""
    shape = x.shape
    x = x.reshape(-1, shape[-1]).clone()
    with torch.no_grad():
        norm = x.norm(dim=1, keepdim=True)
        inner = (norm <= r_in).squeeze(1)
        x[inner] = x[inner] / norm[inner] * r_in
        outer = (norm >= 1.0).squeeze(1)
        x[outer] = x[outer] / norm[outer] * (1.0 - 1e-5)
    return x.view(shape)


def mob_add(u, v, r_in):
    """order_embeddings_h.py:649-660: Moebius addition u (+) (v + 1e-6), then soft_clip."""
    v = v + 1e-6
    uv2 = 2.0 * (u * v).sum(dim=1, keepdim=True)
    uu = (u * u).sum(dim=1, keepdim=True)
    vv = (v * v).sum(dim=1, keepdim=True)
    den = 1.0 + uv2 + vv * uu
    return soft_clip((1.0 + uv2 + vv) / den * u + (1.0 - uu) / den * v, r_in)


def lambda_x(x):
    """order_embeddings_h.py:662-666: the reference's conformal factor 2 / (1 - |x|) -- the norm, not its square
     -- broadcast over the row."""
    return (2.0 / (1.0 - x.norm(p=2, dim=1, keepdim=True))).expand_as(x)


def exp_map_x(x, v, r_in):
    """order_embeddings_h.py:668-674: exponential map at x of the tangent vector v + 1e-15."""
    v = v + 1e-15
    nv = v.norm(p=2, dim=1, keepdim=True)
    second = torch.tanh((lambda_x(x) * nv / 2).clamp(min=-15.0, max=15.0)) * v / nv
    return mob_add(x, second, r_in)
