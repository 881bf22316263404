//! perceive-cuda — safe wrapper over libperceive_cuda.so (C ABI: include/perceive_cuda.h, version 2).
//!
//! Replaces the index behind `perceive_core::search::Searcher`
//! (crates/perceive-core/search.rs) with an exact device-resident top-k search on B200.
//! UNBUILT in this repository (no Rust toolchain in the build image); the extern block is held to
//! the header — names, arity and argument types — by tests/test_abi.py, and every call made here
//! is exercised through the same ABI by perceive_b200/_ffi.py and include/perceive_search.hpp.
#![allow(non_camel_case_types)]
use std::ffi::CStr;
use std::os::raw::{c_char, c_void};

#[repr(C)] pub struct pcv_index { _private: [u8; 0] }
#[repr(C)] pub struct pcv_rowset { _private: [u8; 0] }

#[repr(C)] #[derive(Default, Debug, Clone, Copy)]
pub struct pcv_stats {
    pub n_rows: u64, pub n_rows_global: u64, pub dim: u32, pub dim_padded: u32,
    pub n_sources: u32, pub dtype: u32, pub matrix_bytes: u64, pub last_scan_bytes: u64,
    pub last_search_ms: f32, pub last_launches: u32, pub sm_count: u32, pub world: u32,
    pub rank: u32, pub last_kernel: u32, pub last_fallback_queries: u32,
}

pub const PCV_ABI_VERSION: u32 = 2;
pub const PCV_F32: i32 = 0;            pub const PCV_BF16: i32 = 1;    pub const PCV_F32_SPLIT: i32 = 2;
pub const PCV_METRIC_DOT_REF: i32 = 0; pub const PCV_METRIC_COSINE: i32 = 1;
pub const PCV_FLAG_PRENORMALISE: u32 = 1;
pub const PCV_FLAG_NO_TIMING: u32 = 2;
pub const PCV_MAX_K: u32 = 1024;
pub const PCV_MAX_DIM: u32 = 4096;
pub const PCV_DIST_UNIT_SPHERE: i32 = 0; pub const PCV_DIST_SCALED: i32 = 1;

extern "C" {
    pub fn pcv_index_create(device: i32, dim: u32, store: i32, metric: i32, flags: u32,
                            out: *mut *mut pcv_index) -> i32;
    pub fn pcv_index_create_multi(devices: *const i32, n_devices: i32, dim: u32, store: i32,
                                  metric: i32, flags: u32, out: *mut *mut pcv_index) -> i32;
    pub fn pcv_index_destroy(idx: *mut pcv_index) -> i32;
    pub fn pcv_index_set_rows(idx: *mut pcv_index, rows: *const f32, ids: *const i64,
                              source_ids: *const i64, n: u64) -> i32;
    pub fn pcv_index_replace_source(idx: *mut pcv_index, source_id: i64, rows: *const f32,
                                    ids: *const i64, n: u64) -> i32;
    pub fn pcv_index_generate_synthetic(idx: *mut pcv_index, n: u64, seed: u64, dist: i32,
                                        first_row: u64) -> i32;
    pub fn pcv_synthetic_rows_host(seed: u64, dist: i32, first_row: u64, n: u64, dim: u32,
                                   out: *mut f32) -> i32;
    pub fn pcv_index_get_rows(idx: *mut pcv_index, first_row: u64, n: u64, out_rows: *mut f32,
                              out_ids: *mut i64, out_source_ids: *mut i64) -> i32;
    pub fn pcv_index_find_id(idx: *mut pcv_index, id: i64, out_row: *mut u64) -> i32;
    pub fn pcv_index_set_hidden(idx: *mut pcv_index, ids: *const i64, n: u64) -> i32;
    pub fn pcv_rowset_from_sqlite(db_path: *const c_char, model_id: u32, model_version: u32,
                                  sources: *const i64, n_sources: u32, out: *mut *mut pcv_rowset) -> i32;
    pub fn pcv_rowset_view(rs: *const pcv_rowset, out_n: *mut u64, out_dim: *mut u32,
                           out_rows: *mut *const f32, out_ids: *mut *const i64,
                           out_source_ids: *mut *const i64) -> i32;
    pub fn pcv_rowset_destroy(rs: *mut pcv_rowset) -> i32;
    pub fn pcv_search(idx: *mut pcv_index, queries: *const f32, n_queries: u32, k: u32,
                      sources: *const i64, n_sources: u32, out_ids: *mut i64,
                      out_scores: *mut f32, out_sims: *mut f32, out_counts: *mut u32) -> i32;
    pub fn pcv_search_device(idx: *mut pcv_index, d_queries: *const f32, n_queries: u32, k: u32,
                             sources: *const i64, n_sources: u32, d_out_ids: *mut i64,
                             d_out_scores: *mut f32, d_out_sims: *mut f32,
                             d_out_counts: *mut u32) -> i32;
    pub fn pcv_index_best_chunks(idx: *mut pcv_index, query: *const f32, chunks: *const f32,
                                 n_chunks: u32, doc_chunk_end: *const u32, n_docs: u32,
                                 out_best_chunk: *mut i32, out_best_score: *mut f32,
                                 out_scores: *mut f32) -> i32;
    pub fn pcv_index_set_stream(idx: *mut pcv_index, cuda_stream: *mut c_void) -> i32;
    pub fn pcv_index_synchronize(idx: *mut pcv_index) -> i32;
    pub fn pcv_index_stats(idx: *mut pcv_index, out: *mut pcv_stats) -> i32;
    pub fn pcv_comm_unique_id(out_id: *mut u8) -> i32;             // 128 bytes
    pub fn pcv_index_attach_comm(idx: *mut pcv_index, id: *const u8, rank: i32, world: i32) -> i32;
    pub fn pcv_index_p2p_export(idx: *mut pcv_index, world: i32, max_records: u32, out_handle: *mut u8) -> i32; // 64 bytes
    pub fn pcv_index_p2p_attach(idx: *mut pcv_index, handles: *const u8, rank: i32, world: i32) -> i32;
    pub fn pcv_index_p2p_detach(idx: *mut pcv_index) -> i32;
    pub fn pcv_merge_candidates_device(idx: *mut pcv_index, d_sims: *const f32, d_ids: *const i64,
                                       n_lists: u32, n_queries: u32, k: u32, d_out_ids: *mut i64,
                                       d_out_scores: *mut f32, d_out_sims: *mut f32,
                                       d_out_counts: *mut u32) -> i32;
    pub fn pcv_decode_embedding(blob: *const u8, blob_len: usize, out: *mut f32, out_cap: usize,
                                out_dim: *mut usize) -> i32;
    pub fn pcv_decode_embeddings_bulk(blobs: *const u8, lens: *const usize, n: usize, dim: usize,
                                      out: *mut f32) -> i32;
    pub fn pcv_encode_embedding(v: *const f32, dim: usize, out: *mut u8, out_cap: usize) -> i32;
    pub fn pcv_distance_from_dot(dot: f32, dim: u32) -> f32;
    pub fn pcv_last_error() -> *const c_char;
    pub fn pcv_abi_version() -> u32;
    pub fn pcv_device_count(out: *mut i32) -> i32;
}

/// Stand-in address for an EMPTY source filter: the ABI reads NULL as "every source" and
/// (non-NULL, 0) as "no source" (search.rs:166 with an empty slice), and a slice's own pointer
/// may dangle when it is empty.
static NO_SOURCE: [i64; 1] = [0];

fn last_error() -> String {
    unsafe { CStr::from_ptr(pcv_last_error()) }.to_string_lossy().into_owned()
}

fn check(rc: i32) -> eyre::Result<()> {
    if rc == 0 { return Ok(()); }
    Err(eyre::eyre!("libperceive_cuda error {rc}: {}", last_error()))
}

/// Pointer of a slice that the C side may read `len` times; an empty slice's dangling pointer
/// never crosses the boundary.
fn ptr_or_null<T>(s: &[T]) -> *const T { if s.is_empty() { std::ptr::null() } else { s.as_ptr() } }

/// Result of a batched search: row-major `[n_queries][k]`, `counts[q]` live entries per query.
pub struct Hits { pub k: usize, pub ids: Vec<i64>, pub scores: Vec<f32>, pub sims: Vec<f32>, pub counts: Vec<u32> }

/// Safe owner of the device-resident matrix: one GPU, or several GPUs of this process behind ONE
/// handle (`new_multi`).  `Send + Sync`: the library serialises searches on a handle (required by
/// `AsyncBuilder<T: Send + Sync + 'static>`, perceive-tauri app_state.rs:75); `&mut self` on the
/// methods that replace rows mirrors `Searcher::rebuild_source(&mut self, ..)`.
pub struct Index { raw: *mut pcv_index, dim: usize }
unsafe impl Send for Index {}
unsafe impl Sync for Index {}

impl Index {
    pub fn new(device: i32, dim: u32, store: i32, metric: i32, flags: u32) -> eyre::Result<Self> {
        let mut p = std::ptr::null_mut();
        check(unsafe { pcv_index_create(device, dim, store, metric, flags, &mut p) })?;
        Ok(Index { raw: p, dim: dim as usize })
    }
    /// One handle over several GPUs of this process (row-range shards, exchange over NVLink).
    pub fn new_multi(devices: &[i32], dim: u32, store: i32, metric: i32, flags: u32) -> eyre::Result<Self> {
        eyre::ensure!(!devices.is_empty(), "no device listed");
        if devices.len() == 1 { return Self::new(devices[0], dim, store, metric, flags); }
        let mut p = std::ptr::null_mut();
        check(unsafe { pcv_index_create_multi(devices.as_ptr(), devices.len() as i32, dim, store, metric, flags, &mut p) })?;
        Ok(Index { raw: p, dim: dim as usize })
    }
    pub fn dim(&self) -> usize { self.dim }

    /// `rows` is `ids.len() x dim` row-major; `sources` is empty (every row in source 0) or one id per row.
    pub fn set_rows(&mut self, rows: &[f32], ids: &[i64], sources: &[i64]) -> eyre::Result<()> {
        eyre::ensure!(rows.len() == ids.len() * self.dim, "rows holds {} floats, expected {} x {}", rows.len(), ids.len(), self.dim);
        eyre::ensure!(sources.is_empty() || sources.len() == ids.len(), "{} source ids for {} rows", sources.len(), ids.len());
        check(unsafe { pcv_index_set_rows(self.raw, ptr_or_null(rows), ptr_or_null(ids), ptr_or_null(sources), ids.len() as u64) })
    }
    pub fn replace_source(&mut self, source: i64, rows: &[f32], ids: &[i64]) -> eyre::Result<()> {
        eyre::ensure!(rows.len() == ids.len() * self.dim, "rows holds {} floats, expected {} x {}", rows.len(), ids.len(), self.dim);
        check(unsafe { pcv_index_replace_source(self.raw, source, ptr_or_null(rows), ptr_or_null(ids), ids.len() as u64) })
    }
    /// Bench support: rows `[first_row, first_row + n)` of the deterministic synthetic corpus, generated on the device.
    pub fn generate_synthetic(&mut self, n: u64, seed: u64, dist: i32, first_row: u64) -> eyre::Result<()> {
        check(unsafe { pcv_index_generate_synthetic(self.raw, n, seed, dist, first_row) })
    }
    /// Rows whose items.id is in `ids` are skipped by every later search (opt-in; the reference
    /// never consults `Searcher.hidden`, search.rs:34).  An empty slice restores that behaviour.
    pub fn set_hidden(&self, ids: &[i64]) -> eyre::Result<()> {
        check(unsafe { pcv_index_set_hidden(self.raw, ptr_or_null(ids), ids.len() as u64) })
    }
    /// The stored embedding of item `id` (the `--like ID` query, cmd/search.rs:64-85).
    pub fn embedding_of(&self, id: i64) -> eyre::Result<Option<Vec<f32>>> {
        let mut row = u64::MAX;
        check(unsafe { pcv_index_find_id(self.raw, id, &mut row) })?;
        if row == u64::MAX { return Ok(None); }
        let mut v = vec![0f32; self.dim];
        check(unsafe { pcv_index_get_rows(self.raw, row, 1, v.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) })?;
        Ok(Some(v))
    }
    /// Highlighter scoring (model/highlight.rs:103-127): position of the best chunk inside each
    /// document (None without chunks), replacing `dot_product` + `position_max_by`.
    pub fn best_chunks(&self, query: &[f32], chunks: &[f32], doc_chunk_end: &[u32]) -> eyre::Result<Vec<Option<usize>>> {
        eyre::ensure!(query.len() == self.dim, "query of {} floats, index dimension {}", query.len(), self.dim);
        eyre::ensure!(chunks.len() % self.dim == 0, "chunks is not a whole number of {}-d rows", self.dim);
        let n_chunks = chunks.len() / self.dim;
        eyre::ensure!(doc_chunk_end.last().map_or(true, |&e| e as usize <= n_chunks), "doc_chunk_end runs past the chunks");
        let mut best = vec![-1i32; doc_chunk_end.len()];
        check(unsafe { pcv_index_best_chunks(self.raw, query.as_ptr(), ptr_or_null(chunks), n_chunks as u32,
                                             ptr_or_null(doc_chunk_end), doc_chunk_end.len() as u32,
                                             best.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) })?;
        Ok(best.into_iter().map(|b| usize::try_from(b).ok()).collect())
    }
    /// `queries` holds `n x dim` floats; `sources`: None = every source, Some(&[]) = none (search.rs:166).
    pub fn search(&self, queries: &[f32], k: usize, sources: Option<&[i64]>) -> eyre::Result<Hits> {
        eyre::ensure!(self.dim > 0 && queries.len() % self.dim == 0, "queries is not a whole number of {}-d vectors", self.dim);
        let n = queries.len() / self.dim;
        let total = n.checked_mul(k).ok_or_else(|| eyre::eyre!("n_queries x k overflows"))?;
        eyre::ensure!(n <= u32::MAX as usize && k <= u32::MAX as usize, "batch or k too large");
        let mut hits = Hits { k, ids: vec![-1; total], scores: vec![f32::INFINITY; total],
                              sims: vec![f32::NEG_INFINITY; total], counts: vec![0; n] };
        if n == 0 || k == 0 { return Ok(hits); }
        let (sp, sn) = match sources {
            Some(s) if s.is_empty() => (NO_SOURCE.as_ptr(), 0u32),
            Some(s) => (s.as_ptr(), s.len() as u32),
            None => (std::ptr::null(), 0u32),
        };
        check(unsafe { pcv_search(self.raw, queries.as_ptr(), n as u32, k as u32, sp, sn, hits.ids.as_mut_ptr(),
                                  hits.scores.as_mut_ptr(), hits.sims.as_mut_ptr(), hits.counts.as_mut_ptr()) })?;
        Ok(hits)
    }
    pub fn stats(&self) -> eyre::Result<pcv_stats> {
        let mut st = pcv_stats::default();
        check(unsafe { pcv_index_stats(self.raw, &mut st) })?;
        Ok(st)
    }
}
impl Drop for Index { fn drop(&mut self) { unsafe { pcv_index_destroy(self.raw); } } }

/// `n` rows of the deterministic synthetic stream (queries of the bench workloads), on the host.
pub fn synthetic_rows_host(seed: u64, dist: i32, first_row: u64, n: usize, dim: usize) -> eyre::Result<Vec<f32>> {
    let mut out = vec![0f32; n * dim];
    check(unsafe { pcv_synthetic_rows_host(seed, dist, first_row, n as u64, dim as u32, out.as_mut_ptr()) })?;
    Ok(out)
}

/// search.rs:281-286.  Like the reference this PANICS on a blob that is not a whole number of f32
/// values (there it is the `chunk[3]` index on the trailing partial chunk).
pub fn decode_embedding(blob: &[u8]) -> Vec<f32> {
    let mut out = vec![0f32; blob.len() / 4];
    let rc = unsafe { pcv_decode_embedding(ptr_or_null(blob), blob.len(), out.as_mut_ptr(), out.len(), std::ptr::null_mut()) };
    if rc != 0 { panic!("{}", last_error()); }
    out
}
/// Many blobs laid end to end, every one `dim` floats: one call, one allocation (build_sources' decode loop).
pub fn decode_embeddings_bulk(blobs: &[u8], lens: &[usize], dim: usize) -> eyre::Result<Vec<f32>> {
    eyre::ensure!(lens.iter().sum::<usize>() == blobs.len(), "blob lengths do not add up to the buffer");
    let mut out = vec![0f32; lens.len() * dim];
    check(unsafe { pcv_decode_embeddings_bulk(ptr_or_null(blobs), ptr_or_null(lens), lens.len(), dim, out.as_mut_ptr()) })?;
    Ok(out)
}
/// search.rs:288-294.
pub fn encode_embedding(v: &[f32]) -> Vec<u8> {
    let mut out = vec![0u8; v.len() * 4];
    let rc = unsafe { pcv_encode_embedding(ptr_or_null(v), v.len(), out.as_mut_ptr(), out.len()) };
    debug_assert_eq!(rc, 0);
    out
}
/// search.rs:266-279 for one pair, given the dot product.
pub fn distance_from_dot(dot: f32, dim: u32) -> f32 { unsafe { pcv_distance_from_dot(dot, dim) } }

/// Rows of one model read from the reference's SQLite file by the library itself
/// (search.rs:87-113 without a Vec per row).  Borrow the views, hand them to `Index::set_rows`.
pub struct RowSet(*mut pcv_rowset);
unsafe impl Send for RowSet {}  // plain host memory behind the handle
impl RowSet {
    /// `sources`: None = every source; Some(&[]) = none (search.rs:107-112 with an empty list).
    pub fn from_sqlite(path: &std::path::Path, model_id: u32, model_version: u32, sources: Option<&[i64]>) -> eyre::Result<Self> {
        let c = std::ffi::CString::new(path.to_string_lossy().as_bytes())?;
        let mut p = std::ptr::null_mut();
        let (sp, sn) = match sources {
            Some(s) if s.is_empty() => (NO_SOURCE.as_ptr(), 0u32),
            Some(s) => (s.as_ptr(), s.len() as u32),
            None => (std::ptr::null(), 0u32),
        };
        check(unsafe { pcv_rowset_from_sqlite(c.as_ptr(), model_id, model_version, sp, sn, &mut p) })?;
        Ok(RowSet(p))
    }
    /// (rows, ids, source_ids, dim)
    pub fn view(&self) -> (&[f32], &[i64], &[i64], u32) {
        let (mut n, mut dim) = (0u64, 0u32);
        let (mut r, mut i, mut s) = (std::ptr::null(), std::ptr::null(), std::ptr::null());
        unsafe {
            pcv_rowset_view(self.0, &mut n, &mut dim, &mut r, &mut i, &mut s);
            if n == 0 { return (&[], &[], &[], 0); }
            (std::slice::from_raw_parts(r, n as usize * dim as usize),
             std::slice::from_raw_parts(i, n as usize), std::slice::from_raw_parts(s, n as usize), dim)
        }
    }
}
impl Drop for RowSet { fn drop(&mut self) { unsafe { pcv_rowset_destroy(self.0); } } }
