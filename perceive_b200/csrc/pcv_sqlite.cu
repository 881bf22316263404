// pcv_sqlite.cu — native reader for the reference's embedding table (host only, no kernels).
//
// Replaces the load half of Searcher::build_sources (crates/perceive-core/search.rs:87-113):
// run the reference's join of `items` and `item_embeddings` for one (model_id, model_version),
// keep rows of the listed sources, decode every BLOB (little-endian f32, search.rs:281-286)
// straight into ONE contiguous row-major matrix — the reference allocates a Vec<f32> per row.
// The result ("rowset") is what pcv_index_set_rows / pcv_index_replace_source take.
//
// libsqlite3 is resolved with dlopen at first use (like NCCL): the library has no link-time
// dependency on it, and a host without it gets PCV_ERR_UNSUPPORTED, not a load failure.
#include <dlfcn.h>

#include <cstdarg>
#include <cstdint>
#include <cstdio>
#include <cstring>
#include <mutex>
#include <new>
#include <string>
#include <type_traits>
#include <vector>

#include "../../include/perceive_cuda.h"

int32_t pcv_internal_fail(int32_t code, const char* msg);  // pcv_api.cu (not exported)

namespace {

int32_t failf(int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  return pcv_internal_fail(code, buf);
}

// The handful of sqlite3 C API entry points used, declared by hand (no sqlite3.h in the image;
// these signatures are part of SQLite's frozen public ABI).
struct sqlite3;
struct sqlite3_stmt;
constexpr int kSqliteOk = 0, kSqliteRow = 100, kSqliteDone = 101, kSqliteBlob = 4;
constexpr int kOpenReadOnly = 0x1, kOpenUri = 0x40;

struct SqliteApi {
  void* h = nullptr;
  int (*open_v2)(const char*, sqlite3**, int, const char*) = nullptr;
  int (*close)(sqlite3*) = nullptr;
  int (*prepare_v2)(sqlite3*, const char*, int, sqlite3_stmt**, const char**) = nullptr;
  int (*bind_int64)(sqlite3_stmt*, int, long long) = nullptr;
  int (*step)(sqlite3_stmt*) = nullptr;
  long long (*column_int64)(sqlite3_stmt*, int) = nullptr;
  const void* (*column_blob)(sqlite3_stmt*, int) = nullptr;
  int (*column_bytes)(sqlite3_stmt*, int) = nullptr;
  int (*column_type)(sqlite3_stmt*, int) = nullptr;
  int (*finalize)(sqlite3_stmt*) = nullptr;
  const char* (*errmsg)(sqlite3*) = nullptr;
  int (*busy_timeout)(sqlite3*, int) = nullptr;
  bool ok = false;
};

const SqliteApi& sqlite_api() {
  static SqliteApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    for (const char* name : {"libsqlite3.so.0", "libsqlite3.so"}) {
      api.h = dlopen(name, RTLD_NOW | RTLD_LOCAL);
      if (api.h) break;
    }
    if (!api.h) return;
    bool all = true;
    auto sym = [&](auto& fn, const char* name) {
      fn = reinterpret_cast<std::remove_reference_t<decltype(fn)>>(dlsym(api.h, name));
      all = all && fn != nullptr;
    };
    sym(api.open_v2, "sqlite3_open_v2");
    sym(api.close, "sqlite3_close");
    sym(api.prepare_v2, "sqlite3_prepare_v2");
    sym(api.bind_int64, "sqlite3_bind_int64");
    sym(api.step, "sqlite3_step");
    sym(api.column_int64, "sqlite3_column_int64");
    sym(api.column_blob, "sqlite3_column_blob");
    sym(api.column_bytes, "sqlite3_column_bytes");
    sym(api.column_type, "sqlite3_column_type");
    sym(api.finalize, "sqlite3_finalize");
    sym(api.errmsg, "sqlite3_errmsg");
    sym(api.busy_timeout, "sqlite3_busy_timeout");
    api.ok = all;
  });
  return api;
}

// Same row selection as search.rs:87-92: live items (not skipped, not hidden) that have an
// embedding for this model.  No ORDER BY: pcv_index_set_rows regroups by (source_id, id).
const char* kLoadSql =
    "SELECT i.id, i.source_id, e.embedding FROM items AS i "
    "JOIN item_embeddings AS e ON e.model_id = ?1 AND e.model_version = ?2 AND e.item_id = i.id "
    "WHERE i.skipped IS NULL AND i.hidden_at IS NULL";

}  // namespace

struct pcv_rowset {
  std::vector<float> rows;
  std::vector<int64_t> ids, source_ids;
  uint32_t dim = 0;
};

extern "C" {

int32_t pcv_rowset_from_sqlite(const char* db_path, uint32_t model_id, uint32_t model_version, const int64_t* sources,
                               uint32_t n_sources, pcv_rowset** out) {
  if (out) *out = nullptr;
  if (!db_path || !out) return failf(PCV_ERR_INVALID, "null argument");
  const bool all = sources == nullptr;  // same convention as pcv_search: NULL = every source
  const SqliteApi& s = sqlite_api();
  if (!s.ok) return failf(PCV_ERR_UNSUPPORTED, "libsqlite3.so.0 could not be loaded: the native loader is unavailable on this host");
  sqlite3* db = nullptr;
  if (s.open_v2(db_path, &db, kOpenReadOnly | kOpenUri, nullptr) != kSqliteOk) {
    const std::string why = db ? s.errmsg(db) : "out of memory";
    if (db) s.close(db);
    return failf(PCV_ERR_INVALID, "cannot open %s read-only: %s", db_path, why.c_str());
  }
  s.busy_timeout(db, 5000);  // the indexer may hold the write lock (WAL readers are not blocked)
  sqlite3_stmt* st = nullptr;
  if (s.prepare_v2(db, kLoadSql, -1, &st, nullptr) != kSqliteOk) {
    const std::string why = s.errmsg(db);
    s.close(db);
    return failf(PCV_ERR_INVALID, "%s does not hold the reference schema (items, item_embeddings): %s", db_path, why.c_str());
  }
  s.bind_int64(st, 1, (long long)model_id);
  s.bind_int64(st, 2, (long long)model_version);
  pcv_rowset* rs = nullptr;
  int32_t rc = PCV_OK;
  int step = kSqliteDone;
  try {
  rs = new pcv_rowset();
  while ((step = s.step(st)) == kSqliteRow) {
    const int64_t id = s.column_int64(st, 0), src = s.column_int64(st, 1);
    bool wanted = all;  // search.rs:107-112: rows of sources that are not listed are dropped
    for (uint32_t i = 0; i < n_sources && !wanted; ++i) wanted = sources[i] == src;
    if (!wanted) continue;
    // type first: SQLite leaves sqlite3_column_type undefined once a column_* call has converted the value
    const int ctype = s.column_type(st, 2);
    const uint8_t* blob = static_cast<const uint8_t*>(s.column_blob(st, 2));
    const int bytes = s.column_bytes(st, 2);
    if (ctype != kSqliteBlob || bytes <= 0 || bytes % 4 != 0) {
      rc = failf(PCV_ERR_INVALID, "embedding of item %lld is not a BLOB of whole f32 values (%d bytes)", (long long)id, bytes);
      break;
    }
    const uint32_t dim = (uint32_t)bytes / 4;
    if (rs->dim == 0) {
      if (dim > PCV_MAX_DIM) { rc = failf(PCV_ERR_UNSUPPORTED, "embedding dimension %u above %u", dim, PCV_MAX_DIM); break; }
      rs->dim = dim;
    } else if (dim != rs->dim) {
      rc = failf(PCV_ERR_INVALID, "embedding of item %lld has %u values, earlier rows have %u", (long long)id, dim, rs->dim);
      break;
    }
    const size_t at = rs->rows.size();
    rs->rows.resize(at + dim);
    rc = pcv_decode_embedding(blob, (size_t)bytes, rs->rows.data() + at, dim, nullptr);
    if (rc != PCV_OK) break;
    rs->ids.push_back(id);
    rs->source_ids.push_back(src);
  }
  } catch (const std::bad_alloc&) {  // no exception may cross the C boundary
    rc = failf(PCV_ERR_OOM, "out of host memory while reading %s", db_path);
  }
  if (rc == PCV_OK && step != kSqliteDone) rc = failf(PCV_ERR_INVALID, "reading %s failed: %s", db_path, s.errmsg(db));
  s.finalize(st);
  s.close(db);
  if (rc != PCV_OK) { delete rs; return rc; }
  *out = rs;
  return PCV_OK;
}

int32_t pcv_rowset_view(const pcv_rowset* rs, uint64_t* out_n, uint32_t* out_dim, const float** out_rows,
                        const int64_t** out_ids, const int64_t** out_source_ids) {
  if (!rs) return failf(PCV_ERR_INVALID, "null rowset");
  if (out_n) *out_n = rs->ids.size();
  if (out_dim) *out_dim = rs->dim;
  if (out_rows) *out_rows = rs->rows.data();
  if (out_ids) *out_ids = rs->ids.data();
  if (out_source_ids) *out_source_ids = rs->source_ids.data();
  return PCV_OK;
}

int32_t pcv_rowset_destroy(pcv_rowset* rs) {
  delete rs;
  return PCV_OK;
}

}  // extern "C"
