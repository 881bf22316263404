/*
 * perceive_cuda.h — C ABI of libperceive_cuda (B200 / sm_100a).
 *
 * Drop-in boundary for ONE path of dimfeld/perceive: the nearest-neighbour
 * lookup behind `perceive_core::search::Searcher`
 * (reference: crates/perceive-core/search.rs).  The reference has no FFI for
 * this path (it is a plain Rust struct API over the third-party `hnsw_rs`
 * graph), so every entry point below cites the reference interface it
 * replaces; INTEGRATION.md shows the Rust `extern "C"` block that binds them.
 *
 * Conventions
 *   - every function returns an int32 status (PCV_OK == 0); on failure a
 *     thread-local message is available from pcv_last_error()
 *   - no C++ exception leaves the library: host allocation failure is
 *     PCV_ERR_OOM, anything else unexpected PCV_ERR_STATE
 *   - the library never frees caller memory and never returns memory the
 *     caller must free (a pcv_rowset is library memory behind a handle, released
 *     with pcv_rowset_destroy); all sizes are explicit
 *   - a pcv_index may be used from any thread; searches on one handle are
 *     serialised internally (reference: `Searcher: Send + Sync`, imposed by
 *     crates/perceive-tauri/src-tauri/app_state.rs:75)
 *   - there is NO CPU fallback: without a CUDA device every compute entry
 *     point fails with PCV_ERR_CUDA
 *
 * Result order (the stated tie-break; the reference's own is undefined,
 * search.rs:179 is an unstable sort): similarity descending, ties -> lower
 * doc id first.  `out_scores` carries the reference's *distance*
 * max(0, 1 - dot/dim) (search.rs:274-277), which is therefore ascending.
 */
#ifndef PERCEIVE_CUDA_H
#define PERCEIVE_CUDA_H

#include <stddef.h>
#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define PCV_ABI_VERSION 2

#if defined(__GNUC__)
#define PCV_API __attribute__((visibility("default")))
#else
#define PCV_API
#endif

/* status codes */
#define PCV_OK 0
#define PCV_ERR_INVALID 1      /* bad argument (null, size, k, dim ...)            */
#define PCV_ERR_CUDA 2         /* CUDA runtime/driver error, or no device          */
#define PCV_ERR_NONFINITE 3    /* NaN/Inf in rows or queries (reference panics,    */
                               /* search.rs:179 `partial_cmp().unwrap()`)          */
#define PCV_ERR_OOM 4          /* host or device allocation failed                 */
#define PCV_ERR_UNSUPPORTED 5  /* shape outside what the kernels implement         */
#define PCV_ERR_NCCL 6         /* communicator error                               */
#define PCV_ERR_STATE 7        /* call order violated (e.g. search before rows)    */
#define PCV_ERR_ZERO_NORM 8    /* zero-length row/query under PCV_METRIC_COSINE    */
                               /* (lib.rs:67-77 divides by the norm, no epsilon)   */

/* storage type of the device-resident document matrix
 *   PCV_F32       fp32 rows; searched by the scan kernel (K1), bit-reproducible; batches run
 *                 four queries per pass over the rows
 *   PCV_BF16      bf16 rows (RNE); queries are rounded to bf16 on entry; K1, and the
 *                 tcgen05 kernel (K2) for batches of >= 16 queries
 *   PCV_F32_SPLIT the SAME fp32 values, exactly, held as two 16-bit planes (4 bytes per element):
 *                 hi = the top 16 bits of x (x truncated to bf16), lo = its low 16 bits, so that
 *                 x == (hi << 16) | lo.  Single queries are scanned by K1 over both
 *                 planes; batches (K3) run a tcgen05 FILTER over the hi plane alone, rescore the
 *                 surviving candidates exactly in fp32 (K1's summation order) and prove the
 *                 candidate set complete, falling back to the exact scan for queries where the
 *                 proof fails — results are bit-identical to a PCV_F32 index whatever the data.
 *                 64 <= dim <= 768, PCV_METRIC_DOT_REF only.                               */
typedef enum pcv_dtype { PCV_F32 = 0, PCV_BF16 = 1, PCV_F32_SPLIT = 2 } pcv_dtype;

/* PCV_METRIC_DOT_REF: similarity = dot(q, x) in fp32; reported score is the
 *   reference distance max(0, 1 - dot/dim)        (search.rs:266-279)
 * PCV_METRIC_COSINE : similarity = dot / (|q| |x|), norms computed on the
 *   device from the stored values; reported score = similarity
 *   (crates/perceive-core/lib.rs:67-77)                                    */
typedef enum pcv_metric { PCV_METRIC_DOT_REF = 0, PCV_METRIC_COSINE = 1 } pcv_metric;

/* L2-normalise each row at load time as x / max(|x|, 1e-12)
 * (crates/perceive-core/model/worker.rs:95-103).                            */
#define PCV_FLAG_PRENORMALISE 1u
/* Do not bracket searches with CUDA timing events: pcv_stats.last_search_ms reads 0.  Two event
 * records are two bubbles in the stream; they show when a whole search is ~15 us (a 10k-row corpus,
 * BASELINE config 1) and callers that time searches themselves, or not at all — the reference's
 * Searcher has no such counter — can leave them out.                                              */
#define PCV_FLAG_NO_TIMING 2u

/* synthetic corpus distributions (bench/test support) */
typedef enum pcv_dist {
  PCV_DIST_UNIT_SPHERE = 0, /* approx-gaussian direction, L2-normalised        */
  PCV_DIST_SCALED = 1       /* un-normalised, per-row log-uniform scale (C5)   */
} pcv_dist;

#define PCV_MAX_K 1024u
#define PCV_MAX_DIM 4096u

typedef struct pcv_index pcv_index;

typedef struct pcv_stats {
  uint64_t n_rows;          /* rows resident on this shard                      */
  uint64_t n_rows_global;   /* rows over all shards                             */
  uint32_t dim;             /* logical dimension                                */
  uint32_t dim_padded;      /* stored dimension (multiple of 16 bytes)          */
  uint32_t n_sources;       /* distinct source ids on this shard                */
  uint32_t dtype;           /* pcv_dtype                                        */
  uint64_t matrix_bytes;    /* bytes of the resident document matrix            */
  uint64_t last_scan_bytes; /* algorithmic bytes streamed by the last search    */
  float last_search_ms;     /* device time of the last search (CUDA events); 0 under PCV_FLAG_NO_TIMING */
  uint32_t last_launches;   /* kernels launched by the last search              */
  uint32_t sm_count;
  uint32_t world;           /* shards (1 without a communicator)                */
  uint32_t rank;
  uint32_t last_kernel;     /* 1 = scan (K1), 2 = tcgen05 GEMM (K2 / K3 filter), 0 = none */
  uint32_t last_fallback_queries; /* PCV_F32_SPLIT batches: queries of the last search whose candidate */
                            /* set the filter could not prove complete and the exact scan re-did  */
} pcv_stats;

/* ---- lifecycle ------------------------------------------------------- */

/* Replaces: Searcher construction (search.rs:29-56).  One handle owns the
 * rows of ONE device; multi-GPU = pcv_index_create_multi (one process), or one handle per
 * process + pcv_index_attach_comm / pcv_index_p2p_* (one process per GPU).               */
PCV_API int32_t pcv_index_create(int32_t device, uint32_t dim, pcv_dtype store, pcv_metric metric,
                         uint32_t flags, pcv_index** out);
/* Replaces: the same construction when the corpus spans several GPUs of ONE process — the reference
 * Searcher is one `Send + Sync` object in one process (crates/perceive-tauri/src-tauri/app_state.rs:63-75,
 * crates/perceive-cli/state.rs:28-56).  One handle, one row-range shard per listed device (distinct
 * devices with peer access to one another, at most 16), one worker stream per device; set_rows /
 * replace_source deal rows out in balanced slices; a search runs every shard's local top-k, the shards
 * store their candidates into one another's buffers over NVLink and merge — every other call of this
 * header works on the handle unchanged.  Device-resident queries and results (pcv_search_device) live
 * on devices[0]; rows are numbered shard after shard.  n_devices == 1 is an ordinary index.
 * Not combinable with pcv_index_attach_comm / pcv_index_p2p_* (one handle per process).            */
PCV_API int32_t pcv_index_create_multi(const int32_t* devices, int32_t n_devices, uint32_t dim, pcv_dtype store,
                               pcv_metric metric, uint32_t flags, pcv_index** out);
PCV_API int32_t pcv_index_destroy(pcv_index* idx);

/* Replaces: Searcher::build_sources' load+insert (search.rs:81-155).
 * rows: n x dim fp32 row-major (host); ids: items.id per row; source_ids:
 * items.source_id per row (NULL = all rows in source 0).  Rows are regrouped
 * into contiguous per-source segments ordered by (source_id, id), uploaded,
 * optionally normalised / converted to bf16 on the device.  Replaces all
 * previously loaded rows.  n == 0 is legal (empty index).                  */
PCV_API int32_t pcv_index_set_rows(pcv_index* idx, const float* rows, const int64_t* ids,
                           const int64_t* source_ids, uint64_t n);

/* Replaces: Searcher::rebuild_source (search.rs:58-79): swap (or add) the
 * rows of one source, keep every other segment.  n == 0 removes the source. */
PCV_API int32_t pcv_index_replace_source(pcv_index* idx, int64_t source_id, const float* rows,
                                 const int64_t* ids, uint64_t n);

/* Bench/test support (no reference counterpart): fill the index on the
 * device with rows [first_row, first_row+n) of the deterministic synthetic
 * corpus `seed`; ids = first_row + i + 1, one source (id 0).               */
PCV_API int32_t pcv_index_generate_synthetic(pcv_index* idx, uint64_t n, uint64_t seed, pcv_dist dist,
                                     uint64_t first_row);

/* Same generator for queries (row index = query index), written to HOST
 * memory as fp32, so callers and tests can feed pcv_search.                */
PCV_API int32_t pcv_synthetic_rows_host(uint64_t seed, pcv_dist dist, uint64_t first_row, uint64_t n,
                                uint32_t dim, float* out);

/* Copy stored rows back (as fp32; bf16 rows are widened).  Test support and
 * the `--like ID` lookup (perceive-cli/cmd/search.rs:64-85).               */
PCV_API int32_t pcv_index_get_rows(pcv_index* idx, uint64_t first_row, uint64_t n, float* out_rows,
                           int64_t* out_ids, int64_t* out_source_ids);

/* Row currently holding items.id `id` (*out_row = UINT64_MAX when the id is
 * not on this shard).  With pcv_index_get_rows this is the `--like ID` fetch
 * of perceive-cli/cmd/search.rs:64-85 without the SQL round trip.          */
PCV_API int32_t pcv_index_find_id(pcv_index* idx, int64_t id, uint64_t* out_row);

/* Extension, off until called: install the set of hidden
 * items.id.  The reference keeps `Searcher.hidden` (search.rs:31-34, filled by
 * perceive-cli/cmd/hide.rs:17) but no search method reads it; a row hidden
 * after the build still takes one of the k result slots and is only dropped
 * by the hydrate query (search.rs:210-212), so the caller sees < k items.
 * With a hidden set installed those rows are cut out of the scanned row
 * ranges, so the k best VISIBLE rows come back.  Replaces the previous set;
 * n == 0 restores the reference behaviour.  Ids not resident on this shard
 * are remembered (they apply again after set_rows / replace_source).       */
PCV_API int32_t pcv_index_set_hidden(pcv_index* idx, const int64_t* ids, uint64_t n);

/* ---- native loader ------------------------------------ */

/* Replaces: the SQL + decode half of Searcher::build_sources (search.rs:87-113).
 * Opens the reference's SQLite database read-only, runs the reference's row
 * selection for (model_id, model_version) — live items joined with their
 * embedding, `skipped IS NULL AND hidden_at IS NULL` (search.rs:87-92) —, keeps
 * rows whose source_id is listed (sources == NULL: every source; n_sources == 0
 * with a non-NULL pointer: none, search.rs:107-112), and decodes every BLOB
 * (search.rs:281-286) into ONE row-major fp32 matrix.  The rowset is host
 * memory owned by the library; its views feed pcv_index_set_rows /
 * pcv_index_replace_source.  Rows come in table order (no ORDER BY, as in the
 * reference).  A BLOB whose size is not a whole number of f32 values, or rows of
 * different dimensions, are PCV_ERR_INVALID (the reference panics / builds a
 * broken graph).  libsqlite3 is loaded with dlopen on first use; a host without
 * it gets PCV_ERR_UNSUPPORTED.                                              */
typedef struct pcv_rowset pcv_rowset;
PCV_API int32_t pcv_rowset_from_sqlite(const char* db_path, uint32_t model_id, uint32_t model_version,
                               const int64_t* sources, uint32_t n_sources, pcv_rowset** out);
/* Borrowed views, valid until pcv_rowset_destroy; any out pointer may be NULL.
 * An empty rowset has n = 0 and dim = 0.                                    */
PCV_API int32_t pcv_rowset_view(const pcv_rowset* rs, uint64_t* out_n, uint32_t* out_dim,
                        const float** out_rows, const int64_t** out_ids,
                        const int64_t** out_source_ids);
PCV_API int32_t pcv_rowset_destroy(pcv_rowset* rs);

/* ---- search ---------------------------------------------------------- */

/* Replaces: Searcher::search_vector (search.rs:157-182), batched.
 * queries: B x dim fp32 (host).  sources/n_sources: the source filter of
 * search.rs:166; n_sources == 0 means "no source selected" and returns zero
 * results exactly like the reference; pass sources == NULL to search all.
 * Outputs (host, caller-allocated): out_ids[B*k], out_scores[B*k],
 * out_sims[B*k] (optional, raw similarity), out_counts[B] (results per
 * query, < k when fewer rows are selected; unused slots: id -1, score +inf). */
PCV_API int32_t pcv_search(pcv_index* idx, const float* queries, uint32_t n_queries, uint32_t k,
                   const int64_t* sources, uint32_t n_sources, int64_t* out_ids,
                   float* out_scores, float* out_sims, uint32_t* out_counts);

/* Same, with every buffer already resident on the index's device; enqueued
 * on the index's stream, returns without synchronising.                    */
PCV_API int32_t pcv_search_device(pcv_index* idx, const float* d_queries, uint32_t n_queries, uint32_t k,
                          const int64_t* sources, uint32_t n_sources, int64_t* d_out_ids,
                          float* d_out_scores, float* d_out_sims, uint32_t* d_out_counts);

/* Replaces: the scoring half of Highlighter::highlight
 * (crates/perceive-core/model/highlight.rs:103-127): scores =
 * dot_product(query, chunk encodings) (lib.rs:63-65), then for each document
 * the position of its best chunk.  query: dim fp32; chunks: n_chunks x dim
 * fp32 row-major (host; dim = the index's); doc_chunk_end[n_docs]: cumulative
 * chunk counts, the reference's `document_chunk_boundaries`.  Outputs (host):
 * out_best_chunk[d] = index of the best chunk RELATIVE to document d's first
 * chunk, the last one among equal maxima (itertools position_max_by), -1 for
 * a document without chunks (highlight.rs:122-127 pushes None);
 * out_best_score[d] (optional) its score, 0 when -1; out_scores[n_chunks]
 * (optional) every chunk's score.  Non-finite input is PCV_ERR_NONFINITE
 * (the reference panics on a NaN score, highlight.rs:124).                  */
PCV_API int32_t pcv_index_best_chunks(pcv_index* idx, const float* query, const float* chunks,
                              uint32_t n_chunks, const uint32_t* doc_chunk_end, uint32_t n_docs,
                              int32_t* out_best_chunk, float* out_best_score, float* out_scores);

/* Stream plumbing: run this index's work on a caller-owned cudaStream_t
 * (NULL restores the index's own stream).                                  */
PCV_API int32_t pcv_index_set_stream(pcv_index* idx, void* cuda_stream);
PCV_API int32_t pcv_index_synchronize(pcv_index* idx);
PCV_API int32_t pcv_index_stats(pcv_index* idx, pcv_stats* out);

/* ---- multi-GPU (row-range shards) ------------------------ */

/* 128-byte NCCL unique id, created by rank 0 and distributed by the host.  */
PCV_API int32_t pcv_comm_unique_id(uint8_t out_id[128]);
/* Join the communicator: this handle becomes shard `rank` of `world`.
 * After this, every pcv_search* call is collective: local top-k ->
 * ncclAllGather of the candidates -> merge; every rank gets the result.    */
PCV_API int32_t pcv_index_attach_comm(pcv_index* idx, const uint8_t id[128], int32_t rank,
                              int32_t world);

/* Exchange WITHOUT NCCL ("B200-native alternative"): every rank
 * exports a receive buffer (a CUDA IPC handle, 64 bytes) sized for `max_records`
 * = n_queries * k candidates per shard; the host hands all handles to every rank.
 * After pcv_index_p2p_attach a search whose candidates fit stores them straight
 * into every peer's buffer over NVLink and merges in the same launch (epoch
 * flags, no collective call); larger searches fall back to the NCCL
 * communicator when one is attached.  Every search stays collective.          */
PCV_API int32_t pcv_index_p2p_export(pcv_index* idx, int32_t world, uint32_t max_records,
                             uint8_t out_handle[64]);
PCV_API int32_t pcv_index_p2p_attach(pcv_index* idx, const uint8_t* handles /* world x 64 */,
                             int32_t rank, int32_t world);
/* Unmap the peers' buffers and go back to the NCCL exchange (collective decision of the host:
 * e.g. when mapping failed on ANY rank every rank must detach).                            */
PCV_API int32_t pcv_index_p2p_detach(pcv_index* idx);

/* Merge `n_lists` candidate lists of `k` (sim, id) records each per query —
 * the kernel the all-gather feeds; exposed so logical shards on ONE device
 * can be merged without NCCL (shard-count-invariance tests).  Device ptrs;
 * list l of query q starts at (l * n_queries + q) * k.                      */
PCV_API int32_t pcv_merge_candidates_device(pcv_index* idx, const float* d_sims, const int64_t* d_ids,
                                    uint32_t n_lists, uint32_t n_queries, uint32_t k,
                                    int64_t* d_out_ids, float* d_out_scores, float* d_out_sims,
                                    uint32_t* d_out_counts);

/* ---- embedding wire format -------------------------------------------- */

/* Replaces: deserialize_embedding (search.rs:281-286): little-endian f32,
 * no header.  Unlike the reference (which panics on a trailing partial
 * chunk) a length that is not a multiple of 4 is PCV_ERR_INVALID.          */
PCV_API int32_t pcv_decode_embedding(const uint8_t* blob, size_t blob_len, float* out, size_t out_cap,
                             size_t* out_dim);
/* Bulk form for the loader (the decode loop of build_sources, search.rs:96-113): `n` blobs laid
 * end to end in `blobs`, blob i being lens[i] bytes long; every blob must decode to exactly
 * `dim` floats (PCV_ERR_INVALID names the first offender) and lands in out[i*dim .. (i+1)*dim).
 * One call instead of one heap allocation + one call per row.                 */
PCV_API int32_t pcv_decode_embeddings_bulk(const uint8_t* blobs, const size_t* lens, size_t n, size_t dim,
                                   float* out);
/* Replaces: serialize_embedding (search.rs:288-294).                        */
PCV_API int32_t pcv_encode_embedding(const float* v, size_t dim, uint8_t* out, size_t out_cap);

/* Reference distance for one pair (search.rs:270-278), fp32, host side —
 * used by the host shim to score a single pair without a device call.      */
PCV_API float pcv_distance_from_dot(float dot, uint32_t dim);

/* ---- misc -------------------------------------------------------------- */
PCV_API const char* pcv_last_error(void);
PCV_API uint32_t pcv_abi_version(void);
PCV_API int32_t pcv_device_count(int32_t* out);

#ifdef __cplusplus
}
#endif
#endif /* PERCEIVE_CUDA_H */
