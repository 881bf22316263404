// perceive_search.hpp — C++17 host-side mirror of `perceive_core::search` over the C ABI.
//
// The reference's host code is Rust (crates/perceive-core/search.rs); this image has no Rust
// toolchain, so the compiled-language twin of the shim in rust/search_shim.rs is this header:
// same names, argument meaning and error behaviour as the reference's `pub` items, so a test
// written against it reads like a test of the reference.  Header-only; link libperceive_cuda.so.
//
//   reference (search.rs)                                   here
//   SearchItem {id: i64, score: f32}            :18-22      perceive::SearchItem
//   Searcher::build(db, model_id, model_version) :38-56     Searcher::build(db_path, ...)
//   Searcher::rebuild_source(&mut self, db, ..)  :58-79     Searcher::rebuild_source(db_path, ...)
//   Searcher::search_vector(&self, sources, n, v) :157-182  Searcher::search_vector(...)
//   pub hidden: HashSet<i64>                      :34       Searcher::hidden
//   deserialize_embedding / serialize_embedding  :281-294   perceive::deserialize_embedding / serialize_embedding
//   (new) batched search                                    Searcher::search_vectors
//   (cmd/search.rs:64-85 `--like`)                          Searcher::embedding_of
//   (model/highlight.rs:103-127)                            Searcher::best_chunks
//
// Not mirrored: search / search_and_retrieve / *_and_retrieve (search.rs:184-259) — they call the
// BERT encoder and the hydrate SQL on the caller's connection pool, both unchanged by this work.
//
// Errors: the reference returns Result<_, eyre::Report> from build / rebuild_source and is
// infallible in search_vector (it panics on a NaN score, search.rs:179).  Here every failure is a
// perceive::Error carrying the library's status code and message; there is no CPU fallback, so
// without a CUDA device build() throws with PCV_ERR_CUDA.
#pragma once
#include <algorithm>
#include <cstdint>
#include <stdexcept>
#include <string>
#include <unordered_set>
#include <utility>
#include <vector>

#include "perceive_cuda.h"

namespace perceive {

struct Error : std::runtime_error {
  int32_t code;
  Error(int32_t c, const std::string& what) : std::runtime_error(what), code(c) {}
};

inline void check(int32_t rc) {
  if (rc != PCV_OK) throw Error(rc, std::string("libperceive_cuda error ") + std::to_string(rc) + ": " + pcv_last_error());
}

struct SearchItem {  // search.rs:18-22; score = the reference distance max(0, 1 - dot/len), lower is better
  int64_t id;
  float score;
};

inline std::vector<float> deserialize_embedding(const uint8_t* value, size_t len) {  // search.rs:281-286
  std::vector<float> out(len / 4);
  check(pcv_decode_embedding(value, len, out.data(), out.size(), nullptr));  // len % 4 != 0: the reference panics
  return out;
}

inline std::vector<uint8_t> serialize_embedding(const std::vector<float>& embedding) {  // search.rs:288-294
  std::vector<uint8_t> out(embedding.size() * 4);
  check(pcv_encode_embedding(embedding.data(), embedding.size(), out.data(), out.size()));
  return out;
}

struct Options {  // what the reference hard-codes: one device, fp32 values, its own distance
  int32_t device = 0;
  /// more than one entry: the corpus is sharded over these GPUs of this one process (pcv_index_create_multi);
  /// the Searcher stays ONE Send + Sync object, as in the reference (app_state.rs:63-75)
  std::vector<int32_t> devices;
  /// fp32 values kept exactly as two 16-bit planes: one query is the exact scan, a batch goes through the
  /// tensor-core filter, both return the bits a PCV_F32 index returns.  create() falls back to PCV_F32 for
  /// dimensions that layout does not support.
  pcv_dtype store = PCV_F32_SPLIT;
  pcv_metric metric = PCV_METRIC_DOT_REF;
  /// the reference's Searcher has no timing counter: searches are not bracketed by CUDA events (a 10k-row search
  /// is ~15 us of kernel; the two event records cost 5 us more)
  uint32_t flags = PCV_FLAG_NO_TIMING;
};

class Searcher {
 public:
  /// search.rs:31-34: ids hidden after the build.  Like the reference, search_vector does not read
  /// it unless `filter_hidden` is set; the hydrate SQL drops such rows later.
  std::unordered_set<int64_t> hidden;
  bool filter_hidden = false;

  Searcher() = default;
  Searcher(const Searcher&) = delete;
  Searcher& operator=(const Searcher&) = delete;
  Searcher(Searcher&& o) noexcept { *this = std::move(o); }
  Searcher& operator=(Searcher&& o) noexcept {
    if (this != &o) {
      reset();
      hidden = std::move(o.hidden);
      filter_hidden = o.filter_hidden;
      index_ = o.index_;
      dim_ = o.dim_;
      sources_ = std::move(o.sources_);
      opt_ = o.opt_;
      hidden_sent_ = std::move(o.hidden_sent_);
      o.index_ = nullptr;
    }
    return *this;
  }
  ~Searcher() { reset(); }

  /// search.rs:38-56: every live item of every source that has an embedding for this model.
  static Searcher build(const std::string& db_path, uint32_t model_id, uint32_t model_version, Options opt = {}) {
    Searcher s;
    s.opt_ = opt;
    RowSet rs(db_path, model_id, model_version, nullptr, 0);
    if (rs.n) {
      s.create(rs.dim);
      check(pcv_index_set_rows(s.index_, rs.rows, rs.ids, rs.source_ids, rs.n));
      for (uint64_t i = 0; i < rs.n; ++i) s.note_source(rs.source_ids[i]);
    }
    return s;
  }

  /// search.rs:58-79: replace (or add) one source's rows, keep every other source.
  void rebuild_source(const std::string& db_path, int64_t source_id, uint32_t model_id, uint32_t model_version) {
    RowSet rs(db_path, model_id, model_version, &source_id, 1);
    if (!index_ && rs.n) create(rs.dim);
    if (index_) check(pcv_index_replace_source(index_, source_id, rs.rows, rs.ids, rs.n));
    note_source(source_id);  // search.rs:73-76
  }

  /// search.rs:157-182: the `num_results` nearest items of the listed sources, ascending score.
  std::vector<SearchItem> search_vector(const std::vector<int64_t>& sources, size_t num_results,
                                        const std::vector<float>& vector) const {
    auto all = search_vectors(sources, num_results, vector, 1);
    return all.empty() ? std::vector<SearchItem>{} : std::move(all[0]);
  }

  /// New (the reference searches one vector at a time): `n` queries stored back to back.
  std::vector<std::vector<SearchItem>> search_vectors(const std::vector<int64_t>& sources, size_t num_results,
                                                      const std::vector<float>& vectors, size_t n) const {
    std::vector<std::vector<SearchItem>> out(n);
    if (!index_ || num_results == 0 || n == 0) return out;
    if (vectors.size() != n * dim_) throw Error(PCV_ERR_INVALID, "query dimension does not match the index");
    sync_hidden();
    const uint32_t k = (uint32_t)num_results;
    std::vector<int64_t> ids(n * k);
    std::vector<float> scores(n * k);
    std::vector<uint32_t> counts(n);
    static const int64_t none = 0;  // an empty filter selects nothing (search.rs:166), NULL would mean "all"
    check(pcv_search(index_, vectors.data(), (uint32_t)n, k, sources.empty() ? &none : sources.data(),
                     (uint32_t)sources.size(), ids.data(), scores.data(), nullptr, counts.data()));
    for (size_t q = 0; q < n; ++q)
      for (uint32_t i = 0; i < counts[q]; ++i) out[q].push_back(SearchItem{ids[q * k + i], scores[q * k + i]});
    return out;
  }

  /// The `--like ID` query vector (perceive-cli/cmd/search.rs:64-85); empty when the item has no row.
  std::vector<float> embedding_of(int64_t item_id) const {
    if (!index_) return {};
    uint64_t row = UINT64_MAX;
    check(pcv_index_find_id(index_, item_id, &row));
    if (row == UINT64_MAX) return {};
    std::vector<float> v(dim_);
    check(pcv_index_get_rows(index_, row, 1, v.data(), nullptr, nullptr));
    return v;
  }

  /// Highlighter scoring (model/highlight.rs:103-127): position of the best chunk inside each
  /// document, -1 for a document without chunks.  Needs a built index (its device and dimension).
  std::vector<int32_t> best_chunks(const std::vector<float>& query, const std::vector<float>& chunks,
                                   const std::vector<uint32_t>& doc_chunk_end) const {
    std::vector<int32_t> best(doc_chunk_end.size(), -1);
    if (!index_ || best.empty()) return best;
    check(pcv_index_best_chunks(index_, query.data(), chunks.data(), (uint32_t)(chunks.size() / dim_), doc_chunk_end.data(),
                                (uint32_t)doc_chunk_end.size(), best.data(), nullptr, nullptr));
    return best;
  }

  const std::vector<int64_t>& sources() const { return sources_; }
  uint32_t dim() const { return dim_; }
  pcv_index* handle() const { return index_; }

 private:
  struct RowSet {  // pcv_rowset_from_sqlite: the SQL + decode half of build_sources (search.rs:87-113)
    pcv_rowset* h = nullptr;
    uint64_t n = 0;
    uint32_t dim = 0;
    const float* rows = nullptr;
    const int64_t* ids = nullptr;
    const int64_t* source_ids = nullptr;
    RowSet(const std::string& path, uint32_t model_id, uint32_t model_version, const int64_t* sources, uint32_t n_sources) {
      check(pcv_rowset_from_sqlite(path.c_str(), model_id, model_version, sources, n_sources, &h));
      pcv_rowset_view(h, &n, &dim, &rows, &ids, &source_ids);
    }
    RowSet(const RowSet&) = delete;
    RowSet& operator=(const RowSet&) = delete;
    ~RowSet() { pcv_rowset_destroy(h); }
  };

  void create(uint32_t dim) {
    const pcv_dtype store = (opt_.store == PCV_F32_SPLIT && (dim < 64 || dim > 768 || opt_.metric != PCV_METRIC_DOT_REF)) ? PCV_F32 : opt_.store;
    if (opt_.devices.size() > 1)
      check(pcv_index_create_multi(opt_.devices.data(), (int32_t)opt_.devices.size(), dim, store, opt_.metric, opt_.flags, &index_));
    else
      check(pcv_index_create(opt_.devices.empty() ? opt_.device : opt_.devices[0], dim, store, opt_.metric, opt_.flags, &index_));
    dim_ = dim;
    hidden_sent_.clear();
  }
  void note_source(int64_t s) {
    for (int64_t have : sources_)
      if (have == s) return;
    sources_.push_back(s);
  }
  void sync_hidden() const {
    std::vector<int64_t> ids;
    if (filter_hidden) ids.assign(hidden.begin(), hidden.end());
    std::sort(ids.begin(), ids.end());
    if (ids == hidden_sent_) return;  // callers mutate `hidden` directly (hide.rs:17): compare, then re-send
    check(pcv_index_set_hidden(index_, ids.data(), ids.size()));
    hidden_sent_.swap(ids);
  }
  void reset() {
    if (index_) pcv_index_destroy(index_);
    index_ = nullptr;
  }

  pcv_index* index_ = nullptr;
  uint32_t dim_ = 0;
  std::vector<int64_t> sources_;
  Options opt_;
  mutable std::vector<int64_t> hidden_sent_;  // what pcv_index_set_hidden last received
};

}  // namespace perceive
