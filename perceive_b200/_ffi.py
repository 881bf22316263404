"""ctypes binding of libperceive_cuda.so — the C ABI in include/perceive_cuda.h.

This is the Python stand-in for the Rust `extern "C"` block shown in
INTEGRATION.md.  Nothing here computes: a missing library, a missing symbol or
a missing CUDA device is a loud error, never a fallback.
"""
from __future__ import annotations

import ctypes as C
from pathlib import Path

PKG = Path(__file__).resolve().parent
import os as _os

LIB_PATH = Path(_os.environ["PCV_LIB"]) if _os.environ.get("PCV_LIB") else PKG / "libperceive_cuda.so"  # PCV_LIB: A/B experiments

PCV_OK = 0
PCV_ERR_INVALID, PCV_ERR_CUDA, PCV_ERR_NONFINITE, PCV_ERR_OOM = 1, 2, 3, 4
PCV_ERR_UNSUPPORTED, PCV_ERR_NCCL, PCV_ERR_STATE, PCV_ERR_ZERO_NORM = 5, 6, 7, 8
PCV_F32, PCV_BF16, PCV_F32_SPLIT = 0, 1, 2
PCV_METRIC_DOT_REF, PCV_METRIC_COSINE = 0, 1
PCV_FLAG_PRENORMALISE = 1
PCV_FLAG_NO_TIMING = 2
PCV_DIST_UNIT_SPHERE, PCV_DIST_SCALED = 0, 1
PCV_MAX_K = 1024
PCV_MAX_DIM = 4096

# every symbol include/perceive_cuda.h declares (tests check the .so exports all of them)
SYMBOLS = [
    "pcv_index_create", "pcv_index_create_multi", "pcv_index_destroy", "pcv_index_set_rows", "pcv_index_replace_source",
    "pcv_index_generate_synthetic", "pcv_synthetic_rows_host", "pcv_index_get_rows", "pcv_index_find_id",
    "pcv_index_set_hidden", "pcv_rowset_from_sqlite", "pcv_rowset_view", "pcv_rowset_destroy", "pcv_search",
    "pcv_search_device", "pcv_index_best_chunks", "pcv_index_set_stream", "pcv_index_synchronize", "pcv_index_stats",
    "pcv_comm_unique_id", "pcv_index_attach_comm", "pcv_index_p2p_export", "pcv_index_p2p_attach", "pcv_index_p2p_detach",
    "pcv_merge_candidates_device", "pcv_decode_embedding", "pcv_decode_embeddings_bulk",
    "pcv_encode_embedding", "pcv_distance_from_dot", "pcv_last_error", "pcv_abi_version", "pcv_device_count",
]


class PcvStats(C.Structure):
    _fields_ = [
        ("n_rows", C.c_uint64), ("n_rows_global", C.c_uint64), ("dim", C.c_uint32), ("dim_padded", C.c_uint32),
        ("n_sources", C.c_uint32), ("dtype", C.c_uint32), ("matrix_bytes", C.c_uint64),
        ("last_scan_bytes", C.c_uint64), ("last_search_ms", C.c_float), ("last_launches", C.c_uint32),
        ("sm_count", C.c_uint32), ("world", C.c_uint32), ("rank", C.c_uint32), ("last_kernel", C.c_uint32),
        ("last_fallback_queries", C.c_uint32),
    ]


class PcvError(RuntimeError):
    def __init__(self, code: int, message: str):
        super().__init__(f"libperceive_cuda error {code}: {message}")
        self.code = code
        self.message = message


_lib = None


def load() -> C.CDLL:
    """Load the in-tree library; raises if it has not been built (no fallback)."""
    global _lib
    if _lib is not None:
        return _lib
    if not LIB_PATH.exists():
        raise RuntimeError(
            f"{LIB_PATH} is missing: build it with `python -m perceive_b200._build` "
            "(there is no CPU fallback for the search path)")
    L = C.CDLL(str(LIB_PATH))
    vp, f32p, i64p, u32p, u8p = C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p
    i32, u32, u64 = C.c_int32, C.c_uint32, C.c_uint64
    sig = {
        "pcv_index_create": ([i32, u32, C.c_int, C.c_int, u32, C.POINTER(vp)], i32),
        "pcv_index_create_multi": ([C.c_void_p, i32, u32, C.c_int, C.c_int, u32, C.POINTER(vp)], i32),
        "pcv_index_destroy": ([vp], i32),
        "pcv_index_set_rows": ([vp, f32p, i64p, i64p, u64], i32),
        "pcv_index_replace_source": ([vp, C.c_int64, f32p, i64p, u64], i32),
        "pcv_index_generate_synthetic": ([vp, u64, u64, C.c_int, u64], i32),
        "pcv_synthetic_rows_host": ([u64, C.c_int, u64, u64, u32, f32p], i32),
        "pcv_index_get_rows": ([vp, u64, u64, f32p, i64p, i64p], i32),
        "pcv_index_find_id": ([vp, C.c_int64, C.POINTER(u64)], i32),
        "pcv_index_set_hidden": ([vp, i64p, u64], i32),
        "pcv_rowset_from_sqlite": ([C.c_char_p, u32, u32, i64p, u32, C.POINTER(vp)], i32),
        "pcv_rowset_view": ([vp, C.POINTER(u64), C.POINTER(u32), C.POINTER(vp), C.POINTER(vp), C.POINTER(vp)], i32),
        "pcv_rowset_destroy": ([vp], i32),
        "pcv_search": ([vp, f32p, u32, u32, i64p, u32, i64p, f32p, f32p, u32p], i32),
        "pcv_search_device": ([vp, f32p, u32, u32, i64p, u32, i64p, f32p, f32p, u32p], i32),
        "pcv_index_best_chunks": ([vp, f32p, f32p, u32, u32p, u32, vp, f32p, f32p], i32),
        "pcv_index_set_stream": ([vp, vp], i32),
        "pcv_index_synchronize": ([vp], i32),
        "pcv_index_stats": ([vp, C.POINTER(PcvStats)], i32),
        "pcv_comm_unique_id": ([u8p], i32),
        "pcv_index_attach_comm": ([vp, u8p, i32, i32], i32),
        "pcv_index_p2p_export": ([vp, i32, u32, u8p], i32),
        "pcv_index_p2p_attach": ([vp, u8p, i32, i32], i32),
        "pcv_index_p2p_detach": ([vp], i32),
        "pcv_merge_candidates_device": ([vp, f32p, i64p, u32, u32, u32, i64p, f32p, f32p, u32p], i32),
        "pcv_decode_embedding": ([u8p, C.c_size_t, f32p, C.c_size_t, C.POINTER(C.c_size_t)], i32),
        "pcv_decode_embeddings_bulk": ([u8p, C.c_void_p, C.c_size_t, C.c_size_t, f32p], i32),
        "pcv_encode_embedding": ([f32p, C.c_size_t, u8p, C.c_size_t], i32),
        "pcv_distance_from_dot": ([C.c_float, u32], C.c_float),
        "pcv_last_error": ([], C.c_char_p),
        "pcv_abi_version": ([], u32),
        "pcv_device_count": ([C.POINTER(i32)], i32),
    }
    for name in SYMBOLS:
        fn = getattr(L, name)  # AttributeError if the .so does not export it
        fn.argtypes, fn.restype = sig[name]
    _lib = L
    return L


def check(rc: int) -> None:
    if rc != PCV_OK:
        raise PcvError(rc, load().pcv_last_error().decode("utf-8", "replace"))
