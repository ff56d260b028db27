"""Drop-in names of network/order_embeddings.py (label-only order embeddings + Euclidean cones)."""
from .criterion import EuclideanEmbedder as Embedder  # order_embeddings.py:179
from .criterion import EucConesLoss, OrderEmbeddingLoss  # order_embeddings.py:926, :760

__all__ = ["Embedder", "EucConesLoss", "OrderEmbeddingLoss"]
