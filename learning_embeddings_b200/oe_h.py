"""Drop-in names of network/oe_h.py (joint image+label hyperbolic cones).

The reference keeps the class name `EuclideanConesWithImagesHypernymLoss` for the hyperbolic loss
(oe_h.py:739)."""
from .criterion import HyperbolicTanhEmbedder as Embedder  # oe_h.py:51
from .criterion import inner_radius, rsgd_step  # oe_h.py:1604-1644, :1757-1771
from .joint import HyperbolicFeatNet as FeatNet  # oe_h.py:113
from .joint import HyperbolicConesWithImagesHypernymLoss as EuclideanConesWithImagesHypernymLoss  # oe_h.py:739
from .joint import OrderEmbeddingWithImagesHypernymLoss  # oe_h.py:1060

__all__ = ["Embedder", "FeatNet", "EuclideanConesWithImagesHypernymLoss", "OrderEmbeddingWithImagesHypernymLoss",
           "inner_radius", "rsgd_step"]
