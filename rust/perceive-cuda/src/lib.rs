//! perceive-cuda — safe wrapper over libperceive_cuda.so (C ABI: include/perceive_cuda.h).
//!
//! Replaces the index behind `perceive_core::search::Searcher`
//! (crates/perceive-core/search.rs) with an exact device-resident top-k scan on B200.
//! Reviewed but UNBUILT in this repository (no Rust toolchain in the build image); every call
//! made here is exercised through the same ABI by perceive_b200/_ffi.py and the test-suite.
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
    pub rank: u32, pub last_kernel: u32,
}

pub const PCV_F32: i32 = 0;            pub const PCV_BF16: i32 = 1;    pub const PCV_F32_SPLIT: i32 = 2;
pub const PCV_METRIC_DOT_REF: i32 = 0; pub const PCV_METRIC_COSINE: i32 = 1;
pub const PCV_FLAG_PRENORMALISE: u32 = 1;

extern "C" {
    pub fn pcv_index_create(device: i32, dim: u32, store: i32, metric: i32, flags: u32,
                            out: *mut *mut pcv_index) -> i32;
    pub fn pcv_index_destroy(idx: *mut pcv_index) -> i32;
    pub fn pcv_index_set_rows(idx: *mut pcv_index, rows: *const f32, ids: *const i64,
                              source_ids: *const i64, n: u64) -> i32;
    pub fn pcv_index_replace_source(idx: *mut pcv_index, source_id: i64, rows: *const f32,
                                    ids: *const i64, n: u64) -> i32;
    pub fn pcv_index_get_rows(idx: *mut pcv_index, first_row: u64, n: u64, out_rows: *mut f32,
                              out_ids: *mut i64, out_source_ids: *mut i64) -> i32;
    pub fn pcv_index_find_id(idx: *mut pcv_index, id: i64, out_row: *mut u64) -> i32;
    pub fn pcv_index_set_hidden(idx: *mut pcv_index, ids: *const i64, n: u64) -> i32;
    pub fn pcv_rowset_from_sqlite(db_path: *const std::os::raw::c_char, model_id: u32, model_version: u32,
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
    pub fn pcv_index_generate_synthetic(idx: *mut pcv_index, n: u64, seed: u64, dist: i32,
                                        first_row: u64) -> i32;
    pub fn pcv_synthetic_rows_host(seed: u64, dist: i32, first_row: u64, n: u64, dim: u32,
                                   out: *mut f32) -> i32;
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

/// Safe owner of one device shard.  `Send + Sync`: the library serialises searches on a handle
/// (required by `AsyncBuilder<T: Send + Sync + 'static>`, perceive-tauri app_state.rs:75).
pub struct Index(*mut pcv_index);
unsafe impl Send for Index {}
unsafe impl Sync for Index {}

/// Stand-in address for an EMPTY source filter: the ABI reads NULL as "every source" and
/// (non-NULL, 0) as "no source" (search.rs:166 with an empty slice), and a slice's own pointer
/// may dangle when it is empty.
static NO_SOURCE: [i64; 1] = [0];

fn check(rc: i32) -> eyre::Result<()> {
    if rc == 0 { return Ok(()); }
    let msg = unsafe { CStr::from_ptr(pcv_last_error()) }.to_string_lossy().into_owned();
    Err(eyre::eyre!("libperceive_cuda error {rc}: {msg}"))
}

impl Index {
    pub fn new(device: i32, dim: u32, store: i32, metric: i32, flags: u32) -> eyre::Result<Self> {
        let mut p = std::ptr::null_mut();
        check(unsafe { pcv_index_create(device, dim, store, metric, flags, &mut p) })?;
        Ok(Index(p))
    }
    pub fn set_rows(&mut self, rows: &[f32], ids: &[i64], sources: &[i64]) -> eyre::Result<()> {
        check(unsafe { pcv_index_set_rows(self.0, rows.as_ptr(), ids.as_ptr(), sources.as_ptr(),
                                          ids.len() as u64) })
    }
    pub fn replace_source(&mut self, source: i64, rows: &[f32], ids: &[i64]) -> eyre::Result<()> {
        check(unsafe { pcv_index_replace_source(self.0, source, rows.as_ptr(), ids.as_ptr(),
                                                ids.len() as u64) })
    }
    /// Rows whose items.id is in `ids` are skipped by every later search (opt-in; the reference
    /// never consults `Searcher.hidden`, search.rs:34).  An empty slice restores that behaviour.
    pub fn set_hidden(&mut self, ids: &[i64]) -> eyre::Result<()> {
        check(unsafe { pcv_index_set_hidden(self.0, ids.as_ptr(), ids.len() as u64) })
    }
    /// The stored embedding of item `id` (the `--like ID` query, cmd/search.rs:64-85).
    pub fn embedding_of(&self, id: i64, dim: usize) -> eyre::Result<Option<Vec<f32>>> {
        let mut row = u64::MAX;
        check(unsafe { pcv_index_find_id(self.0, id, &mut row) })?;
        if row == u64::MAX { return Ok(None); }
        let mut v = vec![0f32; dim];
        check(unsafe { pcv_index_get_rows(self.0, row, 1, v.as_mut_ptr(), std::ptr::null_mut(),
                                          std::ptr::null_mut()) })?;
        Ok(Some(v))
    }
    /// Highlighter scoring (model/highlight.rs:103-127): position of the best chunk inside each
    /// document (None without chunks), replacing `dot_product` + `position_max_by`.
    pub fn best_chunks(&self, query: &[f32], chunks: &[f32], doc_chunk_end: &[u32])
        -> eyre::Result<Vec<Option<usize>>> {
        let n_chunks = (chunks.len() / query.len().max(1)) as u32;
        let mut best = vec![-1i32; doc_chunk_end.len()];
        check(unsafe { pcv_index_best_chunks(self.0, query.as_ptr(), chunks.as_ptr(), n_chunks,
                                             doc_chunk_end.as_ptr(), doc_chunk_end.len() as u32,
                                             best.as_mut_ptr(), std::ptr::null_mut(), std::ptr::null_mut()) })?;
        Ok(best.into_iter().map(|b| if b < 0 { None } else { Some(b as usize) }).collect())
    }
    /// (ids, reference distances) per query, best first.
    pub fn search(&self, queries: &[f32], n_queries: u32, k: u32, sources: Option<&[i64]>)
        -> eyre::Result<(Vec<i64>, Vec<f32>, Vec<u32>)> {
        let n = (n_queries * k) as usize;
        let (mut ids, mut scores, mut counts) = (vec![-1i64; n], vec![f32::INFINITY; n], vec![0u32; n_queries as usize]);
        let (sp, sn) = match sources { Some(s) => (if s.is_empty() { NO_SOURCE.as_ptr() } else { s.as_ptr() }, s.len() as u32),
                                        None => (std::ptr::null(), 0) };
        check(unsafe { pcv_search(self.0, queries.as_ptr(), n_queries, k, sp, sn, ids.as_mut_ptr(),
                                  scores.as_mut_ptr(), std::ptr::null_mut(), counts.as_mut_ptr()) })?;
        Ok((ids, scores, counts))
    }
}
/// Rows of one model read from the reference's SQLite file by the library itself
/// (search.rs:87-113 without a Vec per row).  Borrow the views, hand them to `Index::set_rows`.
pub struct RowSet(*mut pcv_rowset);
unsafe impl Send for RowSet {}  // plain host memory behind the handle
impl RowSet {
    pub fn from_sqlite(path: &std::path::Path, model_id: u32, model_version: u32, sources: &[i64])
        -> eyre::Result<Self> {
        let c = std::ffi::CString::new(path.to_string_lossy().as_bytes())?;
        let mut p = std::ptr::null_mut();
        let sp = if sources.is_empty() { NO_SOURCE.as_ptr() } else { sources.as_ptr() };
        check(unsafe { pcv_rowset_from_sqlite(c.as_ptr(), model_id, model_version, sp,
                                              sources.len() as u32, &mut p) })?;
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
impl Drop for Index { fn drop(&mut self) { unsafe { pcv_index_destroy(self.0); } } }