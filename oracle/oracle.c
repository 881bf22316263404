/*
 * oracle.c — TEST INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * CPU restatement of the one path of dimfeld/perceive this repository
 * replaces: the nearest-neighbour lookup of perceive_core::search::Searcher.
 * Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline /
 * --impl reference legs may load this file; the product (perceive_b200/)
 * never does.
 *
 * PARITY UNPINNED: the reference has no tests, fixtures or golden vectors for
 * search (7 unit tests, none on this path) and cannot be
 * built here (no Rust toolchain; hnsw_rs 0.1.17, ndarray 0.15.6, tch 0.10.1
 * are un-vendored registry dependencies, Cargo.lock:1748,2437,4261).  The
 * reference's search is an APPROXIMATE HNSW walk; what is restated here is
 * the exact arithmetic on that path applied to EVERY row:
 *   - the distance functor   crates/perceive-core/search.rs:266-279
 *   - per-source top-k, concat, ascending sort, truncate   search.rs:157-182
 *   - the embedding BLOB codec   search.rs:281-294
 *   - cosine with norms   crates/perceive-core/lib.rs:63-77
 *   - encoder-side normalisation   crates/perceive-core/model/worker.rs:95-103
 * The oracle is pinned instead by known-answer vectors derived by hand from
 * those definitions (tests/golden/, tests/test_oracle.py) and by a numpy
 * float64 twin (oracle/oracle.py).
 *
 * Ordering (stated, since the reference's sort is unstable, search.rs:179):
 * similarity descending, ties -> lower id first; the reported score is the
 * reference distance, hence ascending.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

/* ---- BLOB codec: search.rs:281-294 ------------------------------------ */
ORC_API int orc_decode_embedding(const uint8_t* blob, size_t len, float* out) {
  /* reference: value.chunks(4).map(f32::from_le_bytes) — a trailing partial
   * chunk indexes chunk[3] out of bounds and panics; reported here as -1 */
  if (len % 4 != 0) return -1;
  for (size_t i = 0; i < len / 4; ++i) {
    uint32_t b = (uint32_t)blob[4 * i] | ((uint32_t)blob[4 * i + 1] << 8) |
                 ((uint32_t)blob[4 * i + 2] << 16) | ((uint32_t)blob[4 * i + 3] << 24);
    memcpy(out + i, &b, 4);
  }
  return 0;
}
ORC_API void orc_encode_embedding(const float* v, size_t dim, uint8_t* out) {
  for (size_t i = 0; i < dim; ++i) {
    uint32_t b;
    memcpy(&b, v + i, 4);
    out[4 * i] = (uint8_t)b; out[4 * i + 1] = (uint8_t)(b >> 8);
    out[4 * i + 2] = (uint8_t)(b >> 16); out[4 * i + 3] = (uint8_t)(b >> 24);
  }
}

/* ---- distance functor: search.rs:270-278 -------------------------------- */
ORC_API float orc_distance_from_dot(float dot, uint32_t len) {
  float result = 1.0f - (dot / (float)len); /* search.rs:275 */
  return result > 0.0f ? result : 0.0f;     /* search.rs:277 result.max(0.0) */
}

/* fp32 dot, plain sequential accumulation (one of the orders ndarray/BLAS
 * sdot may use; the reference does not define the order, search.rs:274) */
static float dot_seq(const float* a, const float* b, uint32_t d) {
  float acc = 0.0f;
  for (uint32_t i = 0; i < d; ++i) acc = fmaf(a[i], b[i], acc);
  return acc;
}
static double dot_f64(const float* a, const float* b, uint32_t d) {
  double acc = 0.0;
  for (uint32_t i = 0; i < d; ++i) acc += (double)a[i] * (double)b[i];
  return acc;
}

/* fp32 dot in summation order "v1" — the order the device scan documents
 * (perceive_b200/csrc/pcv_scan.cuh): rows are cut into 16-byte chunks of epc
 * elements; lane g of lpr owns chunks g, g+lpr, ...; per lane one fmaf chain
 * per element position; a pairwise tree over the epc chains; then an xor
 * butterfly lpr/2..1 across lanes.  Restated here so fp32 results can be
 * compared bit for bit. */
static void v1_shape(uint32_t d_padded, uint32_t epc, uint32_t* lpr, uint32_t* nj) {
  uint32_t chunks = d_padded / epc, l = 8;
  while (l < 32 && (chunks + l - 1) / l > 12) l <<= 1;
  *lpr = l;
  *nj = (chunks + l - 1) / l;
}
static float dot_v1(const float* q, const float* x, uint32_t d, uint32_t epc) {
  uint32_t d_padded = (d + epc - 1) / epc * epc, lpr, nj;
  v1_shape(d_padded, epc, &lpr, &nj);
  const uint32_t chunks = d_padded / epc;
  float lane[32];
  for (uint32_t g = 0; g < lpr; ++g) {
    float acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (uint32_t j = 0; j < nj; ++j) {
      const uint32_t c = g + j * lpr;
      if (c >= chunks) continue;
      for (uint32_t e = 0; e < epc; ++e) {
        const uint32_t i = c * epc + e;
        const float qv = i < d ? q[i] : 0.0f, xv = i < d ? x[i] : 0.0f;
        acc[e] = fmaf(qv, xv, acc[e]);
      }
    }
    lane[g] = epc == 4 ? (acc[0] + acc[1]) + (acc[2] + acc[3])
                       : ((acc[0] + acc[1]) + (acc[2] + acc[3])) + ((acc[4] + acc[5]) + (acc[6] + acc[7]));
  }
  for (uint32_t off = lpr / 2; off >= 1; off >>= 1) {
    float nxt[32];
    for (uint32_t g = 0; g < lpr; ++g) nxt[g] = lane[g] + lane[g ^ off];
    memcpy(lane, nxt, sizeof(float) * lpr);
  }
  return lane[0];
}
ORC_API float orc_dot(const float* a, const float* b, uint32_t d, int mode, uint32_t epc) {
  if (mode == 0) return (float)dot_f64(a, b, d);
  if (mode == 1) return dot_seq(a, b, d);
  return dot_v1(a, b, d, epc);
}
ORC_API double orc_dot_f64(const float* a, const float* b, uint32_t d) { return dot_f64(a, b, d); }

/* ---- encoder-side normalisation: model/worker.rs:95-103 ---------------------
 * x / max(norm2(x), 1e-12).  |x|^2 summed in the order the device load kernel
 * documents (32 interleaved fmaf chains + 16..1 butterfly). */
static float sumsq_32(const float* x, uint32_t d) {
  float part[32];
  for (int l = 0; l < 32; ++l) part[l] = 0.0f;
  for (uint32_t c = 0; c < d; ++c) part[c & 31] = fmaf(x[c], x[c], part[c & 31]);
  for (int off = 16; off >= 1; off >>= 1) {
    float nxt[32];
    for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ off];
    memcpy(part, nxt, sizeof part);
  }
  return part[0];
}
ORC_API void orc_normalise_rows(float* rows, uint64_t n, uint32_t d) {
  for (uint64_t r = 0; r < n; ++r) {
    float* x = rows + r * (uint64_t)d;
    const float nrm = fmaxf(sqrtf(sumsq_32(x, d)), 1e-12f); /* clamp_min(1e-12), worker.rs:99 */
    for (uint32_t c = 0; c < d; ++c) x[c] = x[c] / nrm;
  }
}

/* bf16 round-to-nearest-even of an fp32 array (the stored values of a bf16 index) */
ORC_API void orc_round_bf16(float* v, uint64_t n) {
  for (uint64_t i = 0; i < n; ++i) {
    uint32_t b;
    memcpy(&b, v + i, 4);
    if ((b & 0x7fffffffu) <= 0x7f800000u) b += 0x7fffu + ((b >> 16) & 1u);
    b &= 0xffff0000u;
    memcpy(v + i, &b, 4);
  }
}

/* ---- synthetic corpus (restatement of perceive_b200/csrc/pcv_synth.cuh) ---- */
static uint64_t mix64(uint64_t z) {
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ull;
  z ^= z >> 27; z *= 0x94d049bb133111ebull;
  z ^= z >> 31;
  return z;
}
static uint64_t synth_hash(uint64_t seed, uint64_t row, uint32_t col) {
  return mix64((seed * 0x9e3779b97f4a7c15ull) ^ mix64(row * 0xd1b54a32d192ed03ull + (uint64_t)col + 1ull));
}
static float synth_gauss(uint64_t seed, uint64_t row, uint32_t col) {
  const uint64_t h = synth_hash(seed, row, col);
  const int32_t s = (int32_t)((h & 0xffffu) + ((h >> 16) & 0xffffu) + ((h >> 32) & 0xffffu) + (h >> 48));
  return (float)(s - 131070) * 2.64290613e-5f;
}
ORC_API void orc_synth_rows(uint64_t seed, int dist, uint64_t first_row, uint64_t n, uint32_t d, float* out) {
#pragma omp parallel for schedule(static)
  for (int64_t r = 0; r < (int64_t)n; ++r) {
    const uint64_t row = first_row + (uint64_t)r;
    float* x = out + (uint64_t)r * d;
    for (uint32_t c = 0; c < d; ++c) x[c] = synth_gauss(seed, row, c);
    if (dist == 0) {
      const float nrm = fmaxf(sqrtf(sumsq_32(x, d)), 1e-12f);
      for (uint32_t c = 0; c < d; ++c) x[c] = x[c] / nrm;
    } else {
      const uint64_t h = synth_hash(seed ^ 0x5ca1ab1e0ddba11ull, row, 0xffffffffu);
      const int e = (int)(h % 5ull) - 2;
      const float mant = 1.0f + (float)((h >> 40) & 0xffffu) * (1.0f / 65536.0f);
      const float s = ldexpf(mant, e);
      for (uint32_t c = 0; c < d; ++c) x[c] = x[c] * s;
    }
  }
}

/* ---- exact search: search.rs:157-182 applied to every row ---------------- */
typedef struct { double sim; int64_t id; } cand_t;
static int cand_better(const cand_t* a, const cand_t* b) { /* a ranks before b */
  if (a->sim != b->sim) return a->sim > b->sim;
  return a->id < b->id;
}
static int cand_cmp(const void* pa, const void* pb) {
  const cand_t* a = (const cand_t*)pa; const cand_t* b = (const cand_t*)pb;
  if (cand_better(a, b)) return -1;
  if (cand_better(b, a)) return 1;
  return 0;
}

/* similarity of one (query,row) pair.
 * metric 0: dot (ranking key of the reference distance, search.rs:274)
 * metric 1: cosine, lib.rs:67-77: rows divided by their L2 norm, then dot
 * mode 0: float64 truth; 1: fp32 sequential; 2: fp32 order v1 (device order) */
static double pair_sim(const float* q, const float* x, uint32_t d, int metric, int mode, uint32_t epc) {
  if (mode == 0) {
    double dot = dot_f64(q, x, d);
    if (metric == 1) dot /= sqrt(dot_f64(q, q, d)) * sqrt(dot_f64(x, x, d));
    return dot;
  }
  if (mode == 1) {
    float dot = dot_seq(q, x, d);
    if (metric == 1) dot = dot / (sqrtf(dot_seq(x, x, d)) * sqrtf(dot_seq(q, q, d)));
    return (double)dot;
  }
  float dot = dot_v1(q, x, d, epc);
  if (metric == 1) dot = dot / (sqrtf(dot_v1(x, x, d, epc)) * sqrtf(dot_v1(q, q, d, epc)));
  return (double)(dot + 0.0f);
}

/* out arrays sized k; returns number of results (<= k).
 * sources == NULL: all sources; otherwise only rows whose source is listed
 * (search.rs:166).  Shape follows the reference: a top-k per source
 * (search.rs:167-176), concatenated (:177), sorted (:179), truncated (:180). */
ORC_API uint32_t orc_search(const float* rows, uint64_t n, uint32_t d, const int64_t* ids, const int64_t* source_ids,
                            const int64_t* sources, uint32_t n_sources, const float* query, uint32_t k,
                            int metric, int mode, uint32_t epc, int64_t* out_ids, float* out_scores,
                            double* out_sims) {
  /* distinct sources present */
  int64_t* uniq = (int64_t*)malloc(sizeof(int64_t) * (n ? n : 1));
  uint64_t n_uniq = 0;
  for (uint64_t r = 0; r < n; ++r) {
    const int64_t s = source_ids ? source_ids[r] : 0;
    uint64_t j = 0;
    for (; j < n_uniq; ++j) if (uniq[j] == s) break;
    if (j == n_uniq) uniq[n_uniq++] = s;
  }
  cand_t* all = (cand_t*)malloc(sizeof(cand_t) * (size_t)(n_uniq * k + 1));
  uint64_t n_all = 0;
  cand_t* sims = (cand_t*)malloc(sizeof(cand_t) * (n ? n : 1));
  for (uint64_t u = 0; u < n_uniq; ++u) {
    int selected = sources == NULL;
    for (uint32_t i = 0; i < n_sources && !selected; ++i) selected = sources[i] == uniq[u];
    if (!selected) continue;
    uint64_t m = 0;
    for (uint64_t r = 0; r < n; ++r) {
      if ((source_ids ? source_ids[r] : 0) != uniq[u]) continue;
      sims[m].id = (int64_t)r; /* row index for now */
      ++m;
    }
#pragma omp parallel for schedule(static)
    for (int64_t i = 0; i < (int64_t)m; ++i) {
      const uint64_t r = (uint64_t)sims[i].id;
      sims[i].sim = pair_sim(query, rows + r * (uint64_t)d, d, metric, mode, epc);
      sims[i].id = ids[r];
    }
    qsort(sims, m, sizeof(cand_t), cand_cmp);
    for (uint64_t i = 0; i < m && i < k; ++i) all[n_all++] = sims[i];
  }
  qsort(all, n_all, sizeof(cand_t), cand_cmp);
  const uint32_t cnt = (uint32_t)(n_all < k ? n_all : k);
  for (uint32_t i = 0; i < cnt; ++i) {
    out_ids[i] = all[i].id;
    if (out_sims) out_sims[i] = all[i].sim;
    out_scores[i] = metric == 1 ? (float)all[i].sim : orc_distance_from_dot((float)all[i].sim, d);
  }
  free(uniq); free(all); free(sims);
  return cnt;
}

