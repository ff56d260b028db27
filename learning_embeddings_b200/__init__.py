"""B200-native (sm_100a) entailment-cone hot path of ankitdhall/learning_embeddings.

Drop-in criterion / Embedder classes live in the modules named after the reference files
(`order_embeddings`, `order_embeddings_h`, `oe`, `oe_h`); `ops` holds the torch-facing operators and
`_native` the ctypes binding of the C ABI in include/lec_b200.h.  CUDA only -- no CPU fallback.
"""
from . import _native, ops  # noqa: F401

__all__ = ["ops", "_native"]
