//! Replacement for crates/perceive-core/search.rs: the per-source `hnsw_rs` graphs become ONE
//! device-resident exact index (`perceive_cuda::Index`, libperceive_cuda.so).  Every `pub` item of
//! the reference keeps its name and signature — `SearchItem`, `Searcher::{build, rebuild_source,
//! search_vector, search, search_vector_and_retrieve, search_and_retrieve}`, `pub hidden`,
//! `encode_query`, `NdArrayDistance`, `deserialize_embedding`, `serialize_embedding`
//! (search.rs:18-35,38,58,157,184,195,249,262,267,281,288) — so perceive-cli and perceive-tauri
//! compile unchanged.  `search_vectors` (batched) is new and used by `perceive bench`.
//!
//! Cargo.toml of perceive-core: add `perceive-cuda = { path = "../perceive-cuda" }`; `hnsw_rs`,
//! `ndarray` and `blas-src` are no longer needed by this file.
//!
//! UNBUILT in this repository (no Rust toolchain in the image).  The same logic is executed, test
//! for test, by include/perceive_search.hpp (C++) and perceive_b200/searcher.py (Python).
use std::collections::HashMap;
use std::rc::Rc;

use ahash::HashSet;
use perceive_cuda::{Index, PCV_F32_SPLIT, PCV_FLAG_NO_TIMING, PCV_METRIC_DOT_REF};
use rusqlite::Connection;
use time::OffsetDateTime;

use crate::{
    db::{Database, DbError},
    model::Model,
    Item, ItemMetadata,
};

#[derive(Debug, Copy, Clone)]
pub struct SearchItem {
    pub id: i64,
    /// The reference's distance `max(0, 1 - dot/len)`: lower is better.
    pub score: f32,
}

/// GPUs the index is spread over: `PERCEIVE_CUDA_DEVICES=0,1,2,3` (default: device 0).  More than one
/// device still gives ONE `Searcher` in this one process (`pcv_index_create_multi`).
fn configured_devices() -> Vec<i32> {
    std::env::var("PERCEIVE_CUDA_DEVICES")
        .ok()
        .map(|v| v.split(',').filter_map(|d| d.trim().parse().ok()).collect::<Vec<i32>>())
        .filter(|v| !v.is_empty())
        .unwrap_or_else(|| vec![0])
}

pub struct Searcher {
    /// None until the first row arrives (the dimension is only known then).
    index: Option<Index>,
    /// Source ids this searcher knows, in the reference's order (build order, then pushes).
    sources: Vec<i64>,
    /// The search structure is built only from non-hidden items; ids hidden afterwards are kept here,
    /// exactly like the reference (search.rs:31-34) — which never reads the set while searching.
    pub hidden: HashSet<i64>,
    /// Opt-in extension: when set, `hidden` ids are cut out of the scanned rows, so a search
    /// returns `num_results` VISIBLE items instead of a list the hydrate query then shortens.
    pub filter_hidden: bool,
}

/// One model's live rows: a single row-major matrix instead of a `Vec<f32>` per row.
struct LoadedRows {
    dim: usize,
    rows: Vec<f32>,
    ids: Vec<i64>,
    source_ids: Vec<i64>,
}

impl Searcher {
    pub fn build(
        database: &Database,
        model_id: u32,
        model_version: u32,
    ) -> Result<Searcher, eyre::Report> {
        let conn = database.read_pool.get()?;
        let sources: Vec<i64> = conn
            .prepare("SELECT id FROM sources")?
            .query_map([], |row| row.get(0))?
            .collect::<Result<_, _>>()?;

        let loaded = Self::load_rows(&conn, model_id, model_version, &sources)?;
        let index = match loaded.dim {
            0 => None,
            _ => Some(Self::make_index(&loaded)?),
        };
        Ok(Searcher {
            index,
            sources,
            hidden: HashSet::default(),
            filter_hidden: false,
        })
    }

    pub fn rebuild_source(
        &mut self,
        database: &Database,
        source_id: i64,
        model_id: u32,
        model_version: u32,
    ) -> Result<(), eyre::Report> {
        let conn = database.read_pool.get()?;
        let loaded = Self::load_rows(&conn, model_id, model_version, &[source_id])?;
        match self.index.as_mut() {
            // swap (or add, or empty) this source's segment; every other source stays resident
            Some(index) => index.replace_source(source_id, &loaded.rows, &loaded.ids)?,
            None if loaded.dim > 0 => self.index = Some(Self::make_index(&loaded)?),
            None => {}
        }
        if !self.sources.contains(&source_id) {
            self.sources.push(source_id);
        }
        Ok(())
    }

    /// fp32 values kept exactly (two 16-bit planes): single queries are scanned, batches go through
    /// the tensor cores, and both return the same bits as a plain fp32 scan.
    fn make_index(loaded: &LoadedRows) -> Result<Index, eyre::Report> {
        let mut index = Index::new_multi(
            &configured_devices(),
            loaded.dim as u32,
            PCV_F32_SPLIT,
            PCV_METRIC_DOT_REF,
            PCV_FLAG_NO_TIMING, // no per-search CUDA events: nothing here reads last_search_ms
        )?;
        index.set_rows(&loaded.rows, &loaded.ids, &loaded.source_ids)?;
        Ok(index)
    }

    /// The reference's row selection (search.rs:87-92): live items of the model that have an
    /// embedding, restricted to `sources`.  BLOBs are collected end to end and decoded by one
    /// library call into one matrix.
    fn load_rows(
        conn: &Connection,
        model_id: u32,
        model_version: u32,
        sources: &[i64],
    ) -> Result<LoadedRows, eyre::Report> {
        let mut stmt = conn.prepare(
            "SELECT items.id, items.source_id, ie.embedding \
             FROM items JOIN item_embeddings ie \
               ON ie.model_id = ?1 AND ie.model_version = ?2 AND ie.item_id = items.id \
             WHERE items.skipped IS NULL AND items.hidden_at IS NULL",
        )?;
        let mut blobs: Vec<u8> = Vec::new();
        let mut lens: Vec<usize> = Vec::new();
        let mut ids = Vec::new();
        let mut source_ids = Vec::new();
        let mut query = stmt.query([model_id, model_version])?;
        while let Some(row) = query.next()? {
            let source_id: i64 = row.get(1)?;
            if !sources.contains(&source_id) {
                continue;
            }
            let blob = row.get_ref(2)?.as_blob().map_err(DbError::query)?;
            blobs.extend_from_slice(blob);
            lens.push(blob.len());
            ids.push(row.get::<_, i64>(0)?);
            source_ids.push(source_id);
        }
        let dim = lens.first().map_or(0, |bytes| bytes / 4);
        let rows = if ids.is_empty() {
            Vec::new()
        } else {
            // every BLOB must be exactly dim floats: an error here, where the reference would build a broken graph
            perceive_cuda::decode_embeddings_bulk(&blobs, &lens, dim)?
        };
        Ok(LoadedRows { dim, rows, ids, source_ids })
    }

    fn push_hidden(&self, index: &Index) {
        let mut ids: Vec<i64> = if self.filter_hidden {
            self.hidden.iter().copied().collect()
        } else {
            Vec::new()
        };
        ids.sort_unstable();
        index
            .set_hidden(&ids)
            .unwrap_or_else(|e| panic!("perceive-cuda: {e}"));
    }

    /// search.rs:157-182.  The signature is infallible, as in the reference (which panics on a NaN
    /// score, search.rs:179): a library failure — no device, non-finite query — panics with its message.
    pub fn search_vector(
        &self,
        sources: &[i64],
        num_results: usize,
        vector: Vec<f32>,
    ) -> Vec<SearchItem> {
        self.search_vectors(sources, num_results, &vector)
            .into_iter()
            .next()
            .unwrap_or_default()
    }

    /// NEW (the reference searches one vector at a time): `vectors` holds whole query vectors back to
    /// back; one device call scores all of them.  Result `q` is what `search_vector` returns for query `q`.
    pub fn search_vectors(
        &self,
        sources: &[i64],
        num_results: usize,
        vectors: &[f32],
    ) -> Vec<Vec<SearchItem>> {
        let Some(index) = self.index.as_ref() else {
            return Vec::new();
        };
        if num_results == 0 || vectors.is_empty() {
            return Vec::new();
        }
        self.push_hidden(index);
        let hits = index
            .search(vectors, num_results, Some(sources))
            .unwrap_or_else(|e| panic!("perceive-cuda: {e}"));
        hits.counts
            .iter()
            .enumerate()
            .map(|(q, &count)| {
                let first = q * hits.k;
                (first..first + count as usize)
                    .map(|i| SearchItem { id: hits.ids[i], score: hits.scores[i] })
                    .collect()
            })
            .collect()
    }

    pub fn search(
        &self,
        model: &Model,
        sources: &[i64],
        num_results: usize,
        query: &str,
    ) -> Vec<SearchItem> {
        self.search_vector(sources, num_results, encode_query(model, query))
    }

    /// search.rs:195-247: the hits, hydrated from `items`.  Rows hidden or skipped since the build
    /// drop out here (search.rs:210-212); the result stays in ascending-score order (search.rs:245).
    pub fn search_vector_and_retrieve(
        &self,
        database: &Database,
        sources: &[i64],
        num_results: usize,
        vector: Vec<f32>,
    ) -> Result<Vec<(Item, SearchItem)>, DbError> {
        let hits = self.search_vector(sources, num_results, vector);
        let by_id: HashMap<i64, SearchItem> = hits.iter().map(|hit| (hit.id, *hit)).collect();
        let wanted = Rc::new(
            hits.iter()
                .map(|hit| rusqlite::types::Value::from(hit.id))
                .collect::<Vec<_>>(),
        );

        let conn = database.read_pool.get()?;
        let mut stmt = conn.prepare_cached(
            "SELECT id, source_id, external_id, content, name, author, description, modified, last_accessed \
             FROM items WHERE skipped IS NULL AND hidden_at IS NULL AND id IN rarray(?)",
        )?;
        let timestamp = |t: Option<i64>| t.map(|t| OffsetDateTime::from_unix_timestamp(t).unwrap());
        let mut found = stmt
            .query_map([wanted], |row| {
                Ok(Item {
                    id: row.get(0)?,
                    source_id: row.get(1)?,
                    external_id: row.get(2)?,
                    hash: None,
                    content: row.get(3)?,
                    raw_content: None,
                    process_version: 0,
                    metadata: ItemMetadata {
                        name: row.get(4)?,
                        author: row.get(5)?,
                        description: row.get(6)?,
                        mtime: timestamp(row.get(7)?),
                        atime: timestamp(row.get(8)?),
                    },
                    skipped: None,
                })
            })?
            .map(|item| {
                let item = item?;
                let hit = by_id[&item.id];
                Ok((item, hit))
            })
            .collect::<Result<Vec<_>, DbError>>()?;

        // ascending distance; equal distances keep the library's order (lower id first)
        found.sort_by(|a, b| a.1.score.total_cmp(&b.1.score).then(a.1.id.cmp(&b.1.id)));
        Ok(found)
    }

    pub fn search_and_retrieve(
        &self,
        database: &Database,
        model: &Model,
        sources: &[i64],
        num_results: usize,
        query: &str,
    ) -> Result<Vec<(Item, SearchItem)>, DbError> {
        self.search_vector_and_retrieve(database, sources, num_results, encode_query(model, query))
    }

    /// The `--like ID` query vector (perceive-cli/cmd/search.rs:64-85) without the SQL round trip.
    pub fn embedding_of(&self, item_id: i64) -> Option<Vec<f32>> {
        let index = self.index.as_ref()?;
        index.embedding_of(item_id).unwrap_or_else(|e| panic!("perceive-cuda: {e}"))
    }
}

pub fn encode_query(model: &Model, query: &str) -> Vec<f32> {
    let mut encoded = Vec::from(model.encode(&[query]).unwrap());
    encoded.pop().unwrap()
}

/// Kept as a `pub` type for source compatibility (search.rs:266-279); nothing implements
/// `hnsw_rs::dist::Distance` any more.  `eval` is the same quantity, computed from the dot product
/// the way the device does.
#[derive(Clone)]
pub struct NdArrayDistance {}

impl NdArrayDistance {
    pub fn eval(&self, va: &[f32], vb: &[f32]) -> f32 {
        let dot: f32 = va.iter().zip(vb).map(|(a, b)| a * b).sum();
        perceive_cuda::distance_from_dot(dot, va.len() as u32)
    }
}

pub fn deserialize_embedding(value: &[u8]) -> Vec<f32> {
    perceive_cuda::decode_embedding(value)
}

pub fn serialize_embedding(embedding: &[f32]) -> Vec<u8> {
    perceive_cuda::encode_embedding(embedding)
}
