"""oracle.py — TEST INFRASTRUCTURE, NOT PRODUCT CODE.

Python face of the CPU oracle: a ctypes loader for `oracle/_build/liboracle.so`
(built from oracle.c / baseline.c by `make -C oracle`) and an independent numpy
float64 twin of the same definitions.  Only tests/, `__graft_entry__.smoke()`
and bench.py's cpu_baseline / `--impl reference` legs may import this module.

PARITY UNPINNED by the reference (it has no search tests or golden vectors and
cannot be built here, see oracle.c's header); pinned by hand-derived
known-answer vectors in tests/golden/ and by the C-vs-numpy cross-check.

Definitions restated (all paths under /root/reference/crates/perceive-core/):
  distance      search.rs:266-279    max(0, 1 - dot/len), fp32
  search shape  search.rs:157-182    per-source top-k, concat, sort asc, truncate
  BLOB codec    search.rs:281-294    little-endian f32, no header
  cosine        lib.rs:63-77         rows / ||row||_2 (no epsilon), then dot
  normalise     model/worker.rs:95-103   x / max(||x||_2, 1e-12)
"""
from __future__ import annotations

import ctypes as C
import subprocess
from pathlib import Path

import numpy as np

HERE = Path(__file__).resolve().parent
LIB_PATH = HERE / "_build" / "liboracle.so"

MODE_F64, MODE_F32_SEQ, MODE_F32_V1 = 0, 1, 2
METRIC_DOT_REF, METRIC_COSINE = 0, 1
DIST_UNIT_SPHERE, DIST_SCALED = 0, 1


def _stale() -> bool:
    return not LIB_PATH.exists() or any(
        (HERE / f).stat().st_mtime > LIB_PATH.stat().st_mtime for f in ("oracle.c", "baseline.c", "Makefile"))


def build(force: bool = False) -> Path:
    if force or _stale():
        import fcntl
        (HERE / "_build").mkdir(exist_ok=True)
        with open(HERE / "_build" / ".lock", "w") as lock:  # one builder at a time under torchrun
            fcntl.flock(lock, fcntl.LOCK_EX)
            try:
                if force or _stale():
                    r = subprocess.run(["make", "-C", str(HERE)], capture_output=True, text=True)
                    if r.returncode != 0:
                        raise RuntimeError(f"oracle build failed:\n{r.stdout}\n{r.stderr}")
            finally:
                fcntl.flock(lock, fcntl.LOCK_UN)
    return LIB_PATH


_lib = None


def lib() -> C.CDLL:
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(str(LIB_PATH))
        f32p, i64p, u8p, f64p = (C.POINTER(C.c_float), C.POINTER(C.c_int64), C.POINTER(C.c_uint8), C.POINTER(C.c_double))
        L.orc_decode_embedding.argtypes = [u8p, C.c_size_t, f32p]
        L.orc_decode_embedding.restype = C.c_int
        L.orc_encode_embedding.argtypes = [f32p, C.c_size_t, u8p]
        L.orc_encode_embedding.restype = None
        L.orc_distance_from_dot.argtypes = [C.c_float, C.c_uint32]
        L.orc_distance_from_dot.restype = C.c_float
        L.orc_dot.argtypes = [f32p, f32p, C.c_uint32, C.c_int, C.c_uint32]
        L.orc_dot.restype = C.c_float
        L.orc_dot_f64.argtypes = [f32p, f32p, C.c_uint32]
        L.orc_dot_f64.restype = C.c_double
        L.orc_normalise_rows.argtypes = [f32p, C.c_uint64, C.c_uint32]
        L.orc_normalise_rows.restype = None
        L.orc_round_bf16.argtypes = [f32p, C.c_uint64]
        L.orc_round_bf16.restype = None
        L.orc_synth_rows.argtypes = [C.c_uint64, C.c_int, C.c_uint64, C.c_uint64, C.c_uint32, f32p]
        L.orc_synth_rows.restype = None
        L.orc_search.argtypes = [f32p, C.c_uint64, C.c_uint32, i64p, i64p, i64p, C.c_uint32, f32p, C.c_uint32,
                                 C.c_int, C.c_int, C.c_uint32, i64p, f32p, f64p]
        L.orc_search.restype = C.c_uint32
        L.orc_search_fast.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_int64, f32p, C.c_uint32, C.c_int, i64p, f32p, f32p]
        L.orc_search_fast.restype = C.c_uint32
        L.orc_search_batch_fast.argtypes = [f32p, C.c_uint64, C.c_uint32, C.c_int64, f32p, C.c_uint32, C.c_uint32, C.c_int,
                                            i64p, f32p, f32p, C.POINTER(C.c_uint32)]
        L.orc_search_batch_fast.restype = None
        L.orc_max_threads.argtypes = []
        L.orc_max_threads.restype = C.c_int
        _lib = L
    return _lib


def _f32(a):
    return np.ascontiguousarray(a, dtype=np.float32)


def _p(a, t):
    return a.ctypes.data_as(C.POINTER(t))


# --------------------------------------------------------------------------- C oracle
def decode_embedding(blob: bytes) -> np.ndarray:
    """search.rs:281-286.  Raises ValueError where the reference would panic."""
    buf = np.frombuffer(blob, dtype=np.uint8)
    out = np.empty(len(blob) // 4, dtype=np.float32)
    if lib().orc_decode_embedding(_p(buf, C.c_uint8) if len(blob) else None, len(blob), _p(out, C.c_float)) != 0:
        raise ValueError("embedding blob length is not a multiple of 4")
    return out


def encode_embedding(v) -> bytes:
    """search.rs:288-294."""
    v = _f32(v)
    out = np.empty(v.size * 4, dtype=np.uint8)
    lib().orc_encode_embedding(_p(v, C.c_float), v.size, _p(out, C.c_uint8))
    return out.tobytes()


def distance_from_dot(dot: float, dim: int) -> float:
    """search.rs:274-277 given the dot product."""
    return float(lib().orc_distance_from_dot(float(dot), int(dim)))


def dot(a, b, mode: int = MODE_F32_V1, epc: int = 4) -> float:
    a, b = _f32(a), _f32(b)
    return float(lib().orc_dot(_p(a, C.c_float), _p(b, C.c_float), a.size, mode, epc))


def normalise_rows(rows) -> np.ndarray:
    """model/worker.rs:95-103 in the device's summation order."""
    rows = _f32(rows).copy()
    n, d = rows.shape
    lib().orc_normalise_rows(_p(rows, C.c_float), n, d)
    return rows


def round_bf16(v) -> np.ndarray:
    v = _f32(v).copy()
    lib().orc_round_bf16(_p(v, C.c_float), v.size)
    return v


def synth_rows(seed: int, dist: int, first_row: int, n: int, dim: int) -> np.ndarray:
    out = np.empty((n, dim), dtype=np.float32)
    lib().orc_synth_rows(seed, dist, first_row, n, dim, _p(out, C.c_float))
    return out


def search(rows, ids, query, k: int, source_ids=None, sources=None, metric: int = METRIC_DOT_REF,
           mode: int = MODE_F32_V1, epc: int = 4):
    """Exact restatement of search.rs:157-182 over every row.
    Returns (ids[cnt], scores[cnt] float32, sims[cnt] float64)."""
    rows = _f32(rows)
    n, d = rows.shape if rows.ndim == 2 else (0, int(np.asarray(query).size))
    ids = np.ascontiguousarray(ids, dtype=np.int64)
    q = _f32(query)
    src = None if source_ids is None else np.ascontiguousarray(source_ids, dtype=np.int64)
    flt = None if sources is None else np.ascontiguousarray(sources, dtype=np.int64)
    n_flt = 0 if flt is None else int(flt.size)
    if flt is not None and flt.size == 0:
        flt = np.zeros(1, dtype=np.int64)  # non-null pointer, zero live entries: nothing selected
    o_ids = np.empty(k, dtype=np.int64)
    o_scores = np.empty(k, dtype=np.float32)
    o_sims = np.empty(k, dtype=np.float64)
    cnt = lib().orc_search(
        _p(rows, C.c_float), n, d, _p(ids, C.c_int64), None if src is None else _p(src, C.c_int64),
        None if flt is None else _p(flt, C.c_int64), n_flt, _p(q, C.c_float), k, metric, mode, epc,
        _p(o_ids, C.c_int64), _p(o_scores, C.c_float), _p(o_sims, C.c_double))
    return o_ids[:cnt].copy(), o_scores[:cnt].copy(), o_sims[:cnt].copy()


def search_fast(rows, query, k: int, id_base: int = 1, threads: int = 0):
    """Timed CPU baseline (baseline.c): OpenMP + SIMD full scan, dense ids."""
    rows = _f32(rows)
    n, d = rows.shape
    q = _f32(query)
    o_ids = np.empty(k, dtype=np.int64)
    o_scores = np.empty(k, dtype=np.float32)
    o_sims = np.empty(k, dtype=np.float32)
    cnt = lib().orc_search_fast(_p(rows, C.c_float), n, d, id_base, _p(q, C.c_float), k, threads,
                                _p(o_ids, C.c_int64), _p(o_scores, C.c_float), _p(o_sims, C.c_float))
    return o_ids[:cnt].copy(), o_scores[:cnt].copy(), o_sims[:cnt].copy()


def search_batch_fast(rows, queries, k: int, id_base: int = 1, threads: int = 0):
    """Timed CPU baseline for a BATCH (baseline.c): one blocked sgemm-style pass over the rows for
    all queries, per-thread top-k.  Returns (ids[B,k], scores[B,k], sims[B,k], counts[B])."""
    rows = _f32(rows)
    n, d = rows.shape
    q = _f32(queries).reshape(-1, d)
    b = q.shape[0]
    o_ids = np.empty((b, k), dtype=np.int64)
    o_scores = np.empty((b, k), dtype=np.float32)
    o_sims = np.empty((b, k), dtype=np.float32)
    o_cnt = np.empty(b, dtype=np.uint32)
    lib().orc_search_batch_fast(_p(rows, C.c_float), n, d, id_base, _p(q, C.c_float), b, k, threads,
                                _p(o_ids, C.c_int64), _p(o_scores, C.c_float), _p(o_sims, C.c_float),
                                _p(o_cnt, C.c_uint32))
    return o_ids, o_scores, o_sims, o_cnt


def max_threads() -> int:
    return int(lib().orc_max_threads())


# --------------------------------------------------------------------------- numpy twin
def np_distance(dot, dim):
    """search.rs:274-277 in float32 numpy."""
    r = np.float32(1.0) - (np.asarray(dot, dtype=np.float32) / np.float32(dim))
    return np.maximum(r, np.float32(0.0)).astype(np.float32)


def np_search(rows, ids, query, k: int, source_ids=None, sources=None, metric: int = METRIC_DOT_REF):
    """float64 brute force, independent of the C code.  Same shape as
    search.rs:157-182: per-source top-k, concat, sort, truncate."""
    rows64 = np.asarray(rows, dtype=np.float64)
    q64 = np.asarray(query, dtype=np.float64)
    ids = np.asarray(ids, dtype=np.int64)
    n = rows64.shape[0]
    src = np.zeros(n, dtype=np.int64) if source_ids is None else np.asarray(source_ids, dtype=np.int64)
    sims = rows64 @ q64 if n else np.zeros(0)
    if metric == METRIC_COSINE and n:
        sims = sims / (np.linalg.norm(rows64, axis=1) * np.linalg.norm(q64))
    picked = []
    for s in np.unique(src):
        if sources is not None and s not in set(np.asarray(sources).tolist()):
            continue
        rows_s = np.nonzero(src == s)[0]
        order = np.lexsort((ids[rows_s], -sims[rows_s]))[:k]
        picked.append(rows_s[order])
    if not picked:
        return np.zeros(0, np.int64), np.zeros(0, np.float32), np.zeros(0, np.float64)
    cat = np.concatenate(picked)
    order = np.lexsort((ids[cat], -sims[cat]))[:k]
    sel = cat[order]
    s = sims[sel]
    scores = s.astype(np.float32) if metric == METRIC_COSINE else np_distance(s.astype(np.float32), rows64.shape[1])
    return ids[sel], scores, s


def np_best_chunks(query, chunks, doc_chunk_end):
    """float64 restatement of the scoring half of Highlighter::highlight
    (crates/perceive-core/model/highlight.rs:103-127): scores = query . chunk for every
    chunk (lib.rs:63-65 `dot_product`), then per document the position of the maximum
    inside the document's slice of `scores` — itertools `position_max_by` (line 124)
    keeps the LAST of several equal maxima; an empty slice gives None (-1 here).
    Returns (best[n_docs] int32, scores[n_chunks] float64)."""
    q64 = np.asarray(query, dtype=np.float64)
    c64 = np.asarray(chunks, dtype=np.float64).reshape(-1, q64.shape[0])
    scores = c64 @ q64 if c64.shape[0] else np.zeros(0)
    best = np.full(len(doc_chunk_end), -1, dtype=np.int32)
    start = 0
    for d, end in enumerate(doc_chunk_end):
        sl = scores[start:end]
        if sl.size:
            best[d] = sl.size - 1 - int(np.argmax(sl[::-1]))  # last maximum
        start = end
    return best, scores
