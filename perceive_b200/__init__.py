"""perceive_b200 — B200-native (sm_100a) exact top-k vector search behind the
`perceive_core::search::Searcher` API of dimfeld/perceive.

The product is `libperceive_cuda.so` (C ABI: include/perceive_cuda.h); this
package is the Python host-side mirror of the reference's Rust interface and the
build recipe.  Importing it never falls back to a CPU implementation.
"""
from ._ffi import (PCV_BF16, PCV_DIST_SCALED, PCV_DIST_UNIT_SPHERE, PCV_F32, PCV_F32_SPLIT, PCV_FLAG_NO_TIMING, PCV_FLAG_PRENORMALISE, PCV_MAX_K,
                   PCV_METRIC_COSINE, PCV_METRIC_DOT_REF, PcvError, PcvStats)
from .searcher import (Index, SearchItem, Searcher, comm_unique_id, deserialize_embedding, serialize_embedding)

__all__ = [
    "Index", "Searcher", "SearchItem", "serialize_embedding", "deserialize_embedding", "comm_unique_id",
    "PcvError", "PcvStats", "PCV_F32", "PCV_BF16", "PCV_F32_SPLIT", "PCV_METRIC_DOT_REF", "PCV_METRIC_COSINE",
    "PCV_FLAG_PRENORMALISE", "PCV_FLAG_NO_TIMING", "PCV_DIST_UNIT_SPHERE", "PCV_DIST_SCALED", "PCV_MAX_K",
]
