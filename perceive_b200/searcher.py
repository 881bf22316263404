"""Host-side mirror of `perceive_core::search` over the C ABI.

Same names, argument meaning and error behaviour as the reference's Rust API
(crates/perceive-core/search.rs) so tests read like tests of the reference:

    Searcher.build(db, model_id, model_version)            search.rs:38-56
    Searcher.rebuild_source(db, source_id, model_id, ver)  search.rs:58-79
    Searcher.search_vector(sources, num_results, vector)   search.rs:157-182
    Searcher.hidden                                        search.rs:34
    SearchItem(id, score)                                  search.rs:18-22
    serialize_embedding / deserialize_embedding            search.rs:281-294

`Index` is the thin handle wrapper (one device, one shard); `Searcher` adds the
reference's source bookkeeping and SQLite loading.  Everything computes in
libperceive_cuda.so; this module holds no arithmetic of its own.
"""
from __future__ import annotations

import ctypes as C
import sqlite3
from dataclasses import dataclass
from typing import Iterable, Optional, Sequence

import numpy as np

from . import _ffi
from ._ffi import (PCV_BF16, PCV_DIST_SCALED, PCV_DIST_UNIT_SPHERE, PCV_F32, PCV_F32_SPLIT, PCV_FLAG_NO_TIMING, PCV_FLAG_PRENORMALISE,
                   PCV_METRIC_COSINE, PCV_METRIC_DOT_REF, PcvError, PcvStats, check)


@dataclass(frozen=True)
class SearchItem:
    """search.rs:18-22 — `score` is the reference distance: lower is better."""
    id: int
    score: float


def _ptr(a: Optional[np.ndarray]):
    return None if a is None else a.ctypes.data


def serialize_embedding(embedding) -> bytes:
    """search.rs:288-294."""
    v = np.ascontiguousarray(embedding, dtype=np.float32)
    out = np.empty(v.size * 4, dtype=np.uint8)
    check(_ffi.load().pcv_encode_embedding(_ptr(v), v.size, _ptr(out), out.size))
    return out.tobytes()


def deserialize_embedding(value: bytes) -> np.ndarray:
    """search.rs:281-286.  A length that is not a multiple of 4 raises (the
    reference panics on the trailing partial chunk)."""
    buf = np.frombuffer(bytes(value), dtype=np.uint8)
    out = np.empty(len(buf) // 4, dtype=np.float32)
    dim = C.c_size_t(0)
    check(_ffi.load().pcv_decode_embedding(_ptr(buf) if buf.size else None, buf.size, _ptr(out) if out.size else None,
                                           out.size, C.byref(dim)))
    return out


class Index:
    """The device-resident document matrix (opaque `pcv_index`): one shard on one GPU, or — `devices=[...]`
    — one handle over several GPUs of this process (pcv_index_create_multi: one row-range shard per device,
    exchange over NVLink peer access).  Every method is the same on both."""

    def __init__(self, dim: int, device: int = 0, store: int = PCV_F32, metric: int = PCV_METRIC_DOT_REF,
                 flags: int = 0, devices: Optional[Sequence[int]] = None):
        self._lib = _ffi.load()
        self._h = C.c_void_p()
        self.dim = int(dim)
        self.store = store
        self.metric = metric
        if devices is not None and len(devices) > 1:
            devs = np.ascontiguousarray(list(devices), dtype=np.int32)
            check(self._lib.pcv_index_create_multi(_ptr(devs), devs.size, dim, store, metric, flags, C.byref(self._h)))
        else:
            check(self._lib.pcv_index_create(device if not devices else int(devices[0]), dim, store, metric, flags,
                                             C.byref(self._h)))

    def close(self) -> None:
        if getattr(self, "_h", None) is not None and self._h.value:
            self._lib.pcv_index_destroy(self._h)
            self._h = C.c_void_p()

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()

    # -- rows ----------------------------------------------------------------
    def set_rows(self, rows, ids, source_ids=None) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        n = ids.size
        if n and rows.shape != (n, self.dim):
            raise ValueError(f"rows shape {rows.shape} != ({n}, {self.dim})")
        src = None if source_ids is None else np.ascontiguousarray(source_ids, dtype=np.int64)
        if src is not None and src.size != n:
            raise ValueError("source_ids length mismatch")
        check(self._lib.pcv_index_set_rows(self._h, _ptr(rows) if n else None, _ptr(ids) if n else None,
                                           _ptr(src) if (src is not None and n) else None, n))

    def replace_source(self, source_id: int, rows, ids) -> None:
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        n = ids.size
        if n and rows.shape != (n, self.dim):
            raise ValueError(f"rows shape {rows.shape} != ({n}, {self.dim})")
        check(self._lib.pcv_index_replace_source(self._h, source_id, _ptr(rows) if n else None,
                                                 _ptr(ids) if n else None, n))

    def generate_synthetic(self, n: int, seed: int, dist: int = PCV_DIST_UNIT_SPHERE, first_row: int = 0) -> None:
        check(self._lib.pcv_index_generate_synthetic(self._h, n, seed, dist, first_row))

    def get_rows(self, first_row: int, n: int):
        rows = np.empty((n, self.dim), dtype=np.float32)
        ids = np.empty(n, dtype=np.int64)
        src = np.empty(n, dtype=np.int64)
        check(self._lib.pcv_index_get_rows(self._h, first_row, n, _ptr(rows), _ptr(ids), _ptr(src)))
        return rows, ids, src

    def find_id(self, item_id: int) -> Optional[int]:
        """Row holding items.id `item_id` on this shard, or None."""
        row = C.c_uint64(0)
        check(self._lib.pcv_index_find_id(self._h, int(item_id), C.byref(row)))
        return None if row.value == 2 ** 64 - 1 else int(row.value)

    def embedding_of(self, item_id: int):
        """Stored embedding of one item: the `--like ID` query (perceive-cli/cmd/search.rs:64-85)
        read back from the device matrix instead of SQLite.  None when the id is not resident."""
        row = self.find_id(item_id)
        return None if row is None else self.get_rows(row, 1)[0][0]

    def set_hidden(self, ids: Sequence[int]) -> None:
        """Opt-in: rows with these items.id are cut out of every later search (see
        pcv_index_set_hidden; the reference ignores `hidden` while searching, search.rs:34)."""
        a = np.ascontiguousarray(sorted(int(i) for i in ids), dtype=np.int64)
        check(self._lib.pcv_index_set_hidden(self._h, _ptr(a) if a.size else None, a.size))

    def best_chunks(self, query, chunks, doc_chunk_end):
        """Highlighter scoring (model/highlight.rs:103-127): `chunks` [n_chunks, dim] are the chunk
        encodings of several documents laid end to end, `doc_chunk_end` the cumulative chunk count
        after each document.  Returns (best[n_docs] int32 — position inside the document, last of
        equal maxima, -1 without chunks —, best_score[n_docs], scores[n_chunks])."""
        q = np.ascontiguousarray(query, dtype=np.float32).reshape(-1)
        ch = np.ascontiguousarray(chunks, dtype=np.float32).reshape(-1, self.dim)
        ends = np.ascontiguousarray(doc_chunk_end, dtype=np.uint32)
        if q.size != self.dim:
            raise ValueError(f"query dimension {q.size} != {self.dim}")
        best = np.empty(ends.size, dtype=np.int32)
        best_score = np.empty(ends.size, dtype=np.float32)
        scores = np.empty(ch.shape[0], dtype=np.float32)
        check(self._lib.pcv_index_best_chunks(self._h, _ptr(q), _ptr(ch) if ch.size else None, ch.shape[0],
                                              _ptr(ends) if ends.size else None, ends.size,
                                              _ptr(best) if ends.size else None, _ptr(best_score) if ends.size else None,
                                              _ptr(scores) if scores.size else None))
        return best, best_score, scores

    # -- search ----------------------------------------------------------------
    def search(self, queries, k: int, sources: Optional[Sequence[int]] = None):
        """Batched search_vector.  Returns (ids[B,k], scores[B,k], sims[B,k], counts[B])."""
        q = np.ascontiguousarray(queries, dtype=np.float32)
        if q.ndim == 1:
            q = q[None, :]
        if q.shape[1] != self.dim:
            raise ValueError(f"query dimension {q.shape[1]} != {self.dim}")
        b = q.shape[0]
        ids = np.empty((b, k), dtype=np.int64)
        scores = np.empty((b, k), dtype=np.float32)
        sims = np.empty((b, k), dtype=np.float32)
        counts = np.empty(b, dtype=np.uint32)
        if sources is None:
            sp, ns = None, 0
        else:
            sa = np.ascontiguousarray(list(sources) or [0], dtype=np.int64)
            sp, ns = _ptr(sa), len(list(sources))
        check(self._lib.pcv_search(self._h, _ptr(q), b, k, sp, ns, _ptr(ids), _ptr(scores), _ptr(sims), _ptr(counts)))
        return ids, scores, sims, counts

    def search_host(self, queries: int, n_queries: int, k: int, out_ids: int, out_scores: int, out_sims: int,
                    out_counts: int, sources: Optional[Sequence[int]] = None) -> None:
        """`pcv_search` on raw HOST addresses (caller-owned buffers, e.g. `ndarray.ctypes.data`): the
        call a Rust or C integrator makes, without this module's per-call numpy allocations."""
        if sources is None:
            sp, ns = None, 0
        else:
            sa = np.ascontiguousarray(list(sources) or [0], dtype=np.int64)
            sp, ns = _ptr(sa), len(list(sources))
        check(self._lib.pcv_search(self._h, queries, n_queries, k, sp, ns, out_ids, out_scores, out_sims or None,
                                   out_counts or None))

    def search_device(self, d_queries: int, n_queries: int, k: int, d_ids: int, d_scores: int, d_sims: int,
                      d_counts: int, sources: Optional[Sequence[int]] = None) -> None:
        """Raw device-pointer entry (buffers owned by the caller, e.g. torch tensors)."""
        if sources is None:
            sp, ns = None, 0
        else:
            sa = np.ascontiguousarray(list(sources) or [0], dtype=np.int64)
            sp, ns = _ptr(sa), len(list(sources))
        check(self._lib.pcv_search_device(self._h, d_queries, n_queries, k, sp, ns, d_ids, d_scores, d_sims or None,
                                          d_counts or None))

    def merge_candidates_device(self, d_sims: int, d_ids: int, n_lists: int, n_queries: int, k: int, d_out_ids: int,
                                d_out_scores: int, d_out_sims: int, d_out_counts: int) -> None:
        check(self._lib.pcv_merge_candidates_device(self._h, d_sims, d_ids, n_lists, n_queries, k, d_out_ids,
                                                    d_out_scores or None, d_out_sims or None, d_out_counts or None))

    def set_stream(self, cuda_stream: Optional[int]) -> None:
        check(self._lib.pcv_index_set_stream(self._h, cuda_stream))

    def synchronize(self) -> None:
        check(self._lib.pcv_index_synchronize(self._h))

    def stats(self) -> PcvStats:
        st = PcvStats()
        check(self._lib.pcv_index_stats(self._h, C.byref(st)))
        return st

    def attach_comm(self, unique_id: bytes, rank: int, world: int) -> None:
        buf = (C.c_uint8 * 128).from_buffer_copy(unique_id)
        check(self._lib.pcv_index_attach_comm(self._h, buf, rank, world))


    def p2p_export(self, world: int, max_records: int) -> bytes:
        """Allocate this shard's peer receive buffer; returns its 64-byte CUDA IPC handle."""
        buf = (C.c_uint8 * 64)()
        check(self._lib.pcv_index_p2p_export(self._h, world, max_records, buf))
        return bytes(buf)

    def p2p_attach(self, handles: bytes, rank: int, world: int) -> None:
        """Map every rank's receive buffer (handles: world x 64 bytes, in rank order)."""
        if len(handles) != 64 * world:
            raise ValueError("handles must be world x 64 bytes")
        buf = (C.c_uint8 * len(handles)).from_buffer_copy(handles)
        check(self._lib.pcv_index_p2p_attach(self._h, buf, rank, world))


    def p2p_detach(self) -> None:
        check(self._lib.pcv_index_p2p_detach(self._h))


def comm_unique_id() -> bytes:
    buf = (C.c_uint8 * 128)()
    check(_ffi.load().pcv_comm_unique_id(buf))
    return bytes(buf)


_LOAD_SQL = """SELECT items.id, source_id, embedding
        FROM items
        JOIN item_embeddings ie ON model_id=? AND model_version=? AND ie.item_id=items.id
        WHERE skipped IS NULL AND hidden_at IS NULL"""  # search.rs:87-92, verbatim semantics


def _open(database) -> sqlite3.Connection:
    if isinstance(database, sqlite3.Connection):
        return database
    return sqlite3.connect(f"file:{database}?mode=ro", uri=True)


def _load_rows_native(path, model_id: int, model_version: int, sources: Optional[Iterable[int]]):
    """Same result as `_load_rows`, read by the library itself (pcv_rowset_from_sqlite): SQLite's C
    API through dlopen, every BLOB decoded into one matrix.  `path` is a database FILE (an
    in-memory connection cannot be shared with C); sources=None keeps every source."""
    lib = _ffi.load()
    h = C.c_void_p()
    if sources is None:
        sp, ns = None, 0
    else:
        sa = np.ascontiguousarray([int(s) for s in sources] or [0], dtype=np.int64)
        sp, ns = _ptr(sa), len([int(s) for s in sources])
    check(lib.pcv_rowset_from_sqlite(str(path).encode(), model_id, model_version, sp, ns, C.byref(h)))
    try:
        n, dim = C.c_uint64(), C.c_uint32()
        pr, pi, ps = C.c_void_p(), C.c_void_p(), C.c_void_p()
        check(lib.pcv_rowset_view(h, C.byref(n), C.byref(dim), C.byref(pr), C.byref(pi), C.byref(ps)))
        if n.value == 0:
            return np.zeros((0, 0), dtype=np.float32), np.zeros(0, np.int64), np.zeros(0, np.int64), None
        rows = np.ctypeslib.as_array(C.cast(pr, C.POINTER(C.c_float)), shape=(n.value, dim.value)).copy()
        ids = np.ctypeslib.as_array(C.cast(pi, C.POINTER(C.c_int64)), shape=(n.value,)).copy()
        srcs = np.ctypeslib.as_array(C.cast(ps, C.POINTER(C.c_int64)), shape=(n.value,)).copy()
        return rows, ids, srcs, int(dim.value)
    finally:
        lib.pcv_rowset_destroy(h)


def _load(database, model_id: int, model_version: int, sources: Iterable[int]):
    """A database file goes through the native reader, an open connection through Python's sqlite3."""
    if isinstance(database, sqlite3.Connection):
        return _load_rows(database, model_id, model_version, sources)
    return _load_rows_native(database, model_id, model_version, sources)


def _load_rows(conn: sqlite3.Connection, model_id: int, model_version: int, sources: Iterable[int]):
    """The decode half of Searcher::build_sources (search.rs:87-113): rows of the
    listed sources, embeddings decoded from their BLOBs."""
    wanted = set(int(s) for s in sources)
    ids, srcs, blobs = [], [], []
    for item_id, source_id, blob in conn.execute(_LOAD_SQL, (model_id, model_version)):
        if source_id not in wanted:  # search.rs:107-112 drops rows of unlisted sources
            continue
        ids.append(item_id)
        srcs.append(source_id)
        blobs.append(bytes(blob))
    if not blobs:
        return np.zeros((0, 0), dtype=np.float32), np.zeros(0, np.int64), np.zeros(0, np.int64), None
    if len(blobs[0]) % 4:
        raise ValueError(f"item {ids[0]}: embedding of {len(blobs[0])} bytes is not a whole number of f32")
    dim = len(blobs[0]) // 4
    # one bulk decode (pcv_decode_embeddings_bulk) instead of one call and one allocation per row
    lens = np.asarray([len(b) for b in blobs], dtype=np.uintp)
    raw = np.frombuffer(b"".join(blobs), dtype=np.uint8)
    rows = np.empty((len(blobs), dim), dtype=np.float32)
    try:
        check(_ffi.load().pcv_decode_embeddings_bulk(_ptr(raw), _ptr(lens), len(blobs), dim, _ptr(rows)))
    except PcvError as e:
        raise ValueError(f"inconsistent embedding sizes: {e.message}") from e
    return rows, np.asarray(ids, dtype=np.int64), np.asarray(srcs, dtype=np.int64), dim


def default_store(dim: int, metric: int = PCV_METRIC_DOT_REF) -> int:
    """What a drop-in Searcher stores: the fp32 values kept exactly as two 16-bit planes (PCV_F32_SPLIT) — one
    query at a time is the exact scan, a batch (`search_vectors`) goes through the tensor-core filter, and both
    return the bits a plain fp32 index returns.  Plain fp32 rows outside what that layout supports."""
    return PCV_F32_SPLIT if (metric == PCV_METRIC_DOT_REF and 64 <= dim <= 768) else PCV_F32


class Searcher:
    """Drop-in for `perceive_core::search::Searcher` (search.rs:29-260): one
    exact device-resident index instead of one HNSW graph per source."""

    def __init__(self, index: Optional[Index], sources: Sequence[int]):
        self._index = index
        self._sources = [int(s) for s in sources]
        #: search.rs:31-34 — ids hidden after the build.  Kept for API parity; like
        #: the reference, search_vector does not consult it (rows hidden later are
        #: dropped by the hydrate query, search.rs:210-212) ...
        self.hidden: set[int] = set()
        #: ... unless this is switched on: then the scan itself skips
        #: `hidden`, so a search still returns num_results visible items.
        self.filter_hidden = False
        self._hidden_on_device: frozenset = frozenset()

    def _sync_hidden(self) -> None:
        want = frozenset(self.hidden) if self.filter_hidden else frozenset()
        if want != self._hidden_on_device and self._index is not None:
            self._index.set_hidden(want)
            self._hidden_on_device = want

    # -- construction ------------------------------------------------------------
    @classmethod
    def build(cls, database, model_id: int, model_version: int, *, device: int = 0, store: Optional[int] = None,
              metric: int = PCV_METRIC_DOT_REF, flags: int = PCV_FLAG_NO_TIMING,
              devices: Optional[Sequence[int]] = None) -> "Searcher":
        """search.rs:38-56: every source in `sources`, rows from `item_embeddings`."""
        conn = _open(database)
        sources = [r[0] for r in conn.execute("SELECT id FROM sources")]  # search.rs:45-48
        rows, ids, srcs, dim = _load(database, model_id, model_version, sources)
        index = None
        if dim:
            index = Index(dim, device=device, store=default_store(dim, metric) if store is None else store, metric=metric,
                          flags=flags, devices=devices)
            index.set_rows(rows, ids, srcs)
        s = cls(index, sources)
        s._cfg = dict(device=device, store=store, metric=metric, flags=flags, devices=devices)
        return s

    @classmethod
    def from_rows(cls, rows, ids, source_ids=None, *, device: int = 0, store: Optional[int] = None,
                  metric: int = PCV_METRIC_DOT_REF, flags: int = PCV_FLAG_NO_TIMING,
              devices: Optional[Sequence[int]] = None) -> "Searcher":
        """Same index from in-memory rows (what build() does after the SQL load)."""
        rows = np.ascontiguousarray(rows, dtype=np.float32)
        ids = np.ascontiguousarray(ids, dtype=np.int64)
        src = np.zeros(ids.size, dtype=np.int64) if source_ids is None else np.asarray(source_ids, dtype=np.int64)
        index = Index(rows.shape[1], device=device, store=default_store(rows.shape[1], metric) if store is None else store,
                      metric=metric, flags=flags, devices=devices)
        index.set_rows(rows, ids, src)
        s = cls(index, sorted(set(src.tolist())))
        s._cfg = dict(device=device, store=store, metric=metric, flags=flags, devices=devices)
        return s

    def rebuild_source(self, database, source_id: int, model_id: int, model_version: int) -> None:
        """search.rs:58-79: replace (or add) one source's rows, keep the rest."""
        rows, ids, _, dim = _load(database, model_id, model_version, [source_id])
        if dim is None:
            # the reference still builds an (empty) per-source index and stores it
            if self._index is not None:
                self._index.replace_source(source_id, np.zeros((0, self._index.dim), np.float32), np.zeros(0, np.int64))
        else:
            if self._index is None:
                cfg = dict(getattr(self, "_cfg", {}))
                if cfg.get("store") is None:
                    cfg["store"] = default_store(dim, cfg.get("metric", PCV_METRIC_DOT_REF))
                self._index = Index(dim, **cfg)
                self._hidden_on_device = frozenset()
            self._index.replace_source(source_id, rows, ids)
        if source_id not in self._sources:  # search.rs:73-76
            self._sources.append(int(source_id))

    # -- queries ----------------------------------------------------------------
    def search_vector(self, sources: Sequence[int], num_results: int, vector) -> list[SearchItem]:
        """search.rs:157-182.  Infallible in the reference except for NaN scores
        (panic); here errors surface as PcvError."""
        if self._index is None or num_results == 0:
            return []
        self._sync_hidden()
        ids, scores, _, counts = self._index.search(vector, num_results, sources=list(sources))
        return [SearchItem(int(ids[0, i]), float(scores[0, i])) for i in range(int(counts[0]))]

    def search_vectors(self, sources: Sequence[int], num_results: int, vectors) -> list[list[SearchItem]]:
        """Batched form (new; the reference has no batched entry point)."""
        if self._index is None or num_results == 0:
            return [[] for _ in range(len(vectors))]
        self._sync_hidden()
        ids, scores, _, counts = self._index.search(vectors, num_results, sources=list(sources))
        return [[SearchItem(int(ids[b, i]), float(scores[b, i])) for i in range(int(counts[b]))]
                for b in range(ids.shape[0])]

    def search_vector_and_retrieve(self, database, sources: Sequence[int], num_results: int, vector):
        """search.rs:195-247: search, then hydrate the hits from `items`, dropping
        rows skipped/hidden since the build, re-sorted by ascending score."""
        items = self.search_vector(sources, num_results, vector)
        if not items:
            return []
        conn = _open(database)
        marks = ",".join("?" for _ in items)
        sql = ("SELECT id, source_id, external_id, content, name, author, description, modified, last_accessed "
               f"FROM items WHERE skipped is NULL AND hidden_at IS NULL AND id IN ({marks})")  # search.rs:210-212
        by_id = {it.id: it for it in items}
        order = {it.id: i for i, it in enumerate(items)}  # result order: ascending score, ties by the stated rule
        rows = [(dict(zip(("id", "source_id", "external_id", "content", "name", "author", "description", "modified",
                           "last_accessed"), r)), by_id[r[0]]) for r in conn.execute(sql, [it.id for it in items])]
        rows.sort(key=lambda p: order[p[1].id])  # search.rs:245 (ascending score)
        return rows

    def embedding_of(self, item_id: int):
        """The `--like ID` query vector (perceive-cli/cmd/search.rs:64-85), taken from the
        resident matrix; None when the item has no row (the CLI reports "Item not found")."""
        return None if self._index is None else self._index.embedding_of(item_id)

    @property
    def index(self) -> Optional[Index]:
        return self._index

    def close(self) -> None:
        if self._index is not None:
            self._index.close()
            self._index = None
