"""Drop-in names of network/oe.py (joint image+label Euclidean cones / order embeddings)."""
from .criterion import EuclideanEmbedder
from .joint import EuclideanFeatNet as FeatNet  # oe.py:83
from .joint import EuclideanConesWithImagesHypernymLoss, OrderEmbeddingWithImagesHypernymLoss  # oe.py:650, :967


class Embedder(EuclideanEmbedder):
    """oe.py:51-80 (positional order: embedding_dim, labelmap, normalize, K)."""

    def __init__(self, embedding_dim, labelmap, normalize=None, K=None):
        super().__init__(embedding_dim, labelmap, K=K, normalize=normalize)


__all__ = ["Embedder", "FeatNet", "EuclideanConesWithImagesHypernymLoss", "OrderEmbeddingWithImagesHypernymLoss"]
