//! Drop-in body for crates/perceive-core/search.rs: every `pub` item of the reference keeps its
//! signature (search.rs:18-35,38,58,157,184,195,249,262,281,288); the HNSW graphs are replaced by
//! one `perceive_cuda::Index`.  UNBUILT here (see rust/perceive-cuda/Cargo.toml); the Python twin
//! is perceive_b200/searcher.py, tested in tests/test_searcher_sqlite.py.  Bodies elided with
//! `/* … */` are the reference's own unchanged code.
pub struct SearchItem { pub id: i64, pub score: f32 }          // search.rs:18-22, unchanged

pub struct Searcher {
    index: Option<perceive_cuda::Index>,                         // was: Vec<SourceSearch> of hnsw graphs
    sources: Vec<i64>,
    pub hidden: HashSet<i64>,                                    // search.rs:34, still public, still unread by search
    model: (u32, u32),
}

impl Searcher {
    pub fn build(database: &Database, model_id: u32, model_version: u32) -> Result<Searcher, eyre::Report> {
        let conn = database.read_pool.get()?;
        let sources = /* SELECT id FROM sources — search.rs:45-48 */;
        let (rows, ids, srcs, dim) = load_rows(&conn, model_id, model_version, &sources)?;   // search.rs:87-113
        let index = match dim { 0 => None, d => {
            let mut ix = perceive_cuda::Index::new(0, d as u32, PCV_F32, PCV_METRIC_DOT_REF, 0)?;
            ix.set_rows(&rows, &ids, &srcs)?; Some(ix) } };
        Ok(Searcher { index, sources, hidden: HashSet::default(), model: (model_id, model_version) })
    }

    pub fn rebuild_source(&mut self, database: &Database, source_id: i64, model_id: u32, model_version: u32)
        -> Result<(), eyre::Report> {                                                        // search.rs:58-79
        let conn = database.read_pool.get()?;
        let (rows, ids, _, dim) = load_rows(&conn, model_id, model_version, &[source_id])?;
        match (&mut self.index, dim) {
            (Some(ix), _) => ix.replace_source(source_id, &rows, &ids)?,
            (None, 0) => {}
            (None, d) => { let mut ix = perceive_cuda::Index::new(0, d as u32, PCV_F32, PCV_METRIC_DOT_REF, 0)?;
                           ix.set_rows(&rows, &ids, &vec![source_id; ids.len()])?; self.index = Some(ix); }
        }
        if !self.sources.contains(&source_id) { self.sources.push(source_id); }             // search.rs:73-76
        Ok(())
    }

    /// search.rs:157-182.  Infallible signature kept: a library failure panics with its message
    /// (the reference panics on NaN scores at search.rs:179).
    pub fn search_vector(&self, sources: &[i64], num_results: usize, vector: Vec<f32>) -> Vec<SearchItem> {
        let Some(ix) = &self.index else { return vec![] };
        if num_results == 0 { return vec![] }
        let (ids, scores, counts) = ix.search(&vector, 1, num_results as u32, Some(sources))
            .unwrap_or_else(|e| panic!("{e}"));
        (0..counts[0] as usize).map(|i| SearchItem { id: ids[i], score: scores[i] }).collect()
    }

    /// NEW (no reference counterpart): batched search used by `perceive bench`.
    pub fn search_vectors(&self, sources: &[i64], num_results: usize, vectors: &[f32], n: usize)
        -> Vec<Vec<SearchItem>> { /* one pcv_search call with n_queries = n */ }

    // search(), search_vector_and_retrieve(), search_and_retrieve(): bodies unchanged
    // (search.rs:184-259) — they only call search_vector / encode_query and SQLite.
}

pub fn deserialize_embedding(value: &[u8]) -> Vec<f32> { /* pcv_decode_embedding; panics like the reference on len % 4 != 0 */ }
pub fn serialize_embedding(embedding: &[f32]) -> Vec<u8> { /* pcv_encode_embedding */ }
