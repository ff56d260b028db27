"""Drop-in names of network/order_embeddings_h.py (label-only Poincare cones + RSGD).

The reference keeps the class name `EucConesLoss` for the hyperbolic loss (order_embeddings_h.py:1072)."""
from .criterion import HyperbolicEmbedder as Embedder  # order_embeddings_h.py:181
from .criterion import HypConesLoss as EucConesLoss  # order_embeddings_h.py:1072
from .criterion import OrderEmbeddingLoss, inner_radius, rsgd_step  # :906, :1089, :764-775

__all__ = ["Embedder", "EucConesLoss", "OrderEmbeddingLoss", "inner_radius", "rsgd_step"]
