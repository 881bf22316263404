/*
 * baseline.c — TEST/BENCH INFRASTRUCTURE, NOT PRODUCT CODE.
 *
 * The timed CPU arm of bench.py (cpu_baseline, --impl reference): the
 * reference's exact scoring (crates/perceive-core/search.rs:266-279) applied as
 * a full scan, written the way a CPU deployment would be.  kind = "port": the
 * reference's own search (an hnsw_rs graph walk) cannot be built here.
 * Compiled separately from oracle.c so it may use fused multiply-add
 * contraction and SIMD reductions; results are checked against oracle.c
 * within tolerance, never used as the parity oracle themselves.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>
#ifdef _OPENMP
#include <omp.h>
#endif

#define ORC_API __attribute__((visibility("default")))

static float distance_from_dot(float dot, uint32_t len) { /* search.rs:275-277 */
  float result = 1.0f - (dot / (float)len);
  return result > 0.0f ? result : 0.0f;
}

/* ---- timed CPU baseline ------------------------------------------------------
 * The same scan written the way a CPU implementation of the reference's
 * scoring would be deployed: OpenMP over row blocks (the reference fans out
 * with rayon, search.rs:163-166), SIMD fp32 dot (its ndarray dot dispatches to
 * BLAS sdot, search.rs:274), per-thread top-k, then concat + sort + truncate
 * (search.rs:177-181).  One source, dense ids (id = id_base + row). */
__attribute__((target_clones("avx512f", "avx2", "default")))
static float dot_simd(const float* a, const float* b, uint32_t d) {
  float acc = 0.0f;
#pragma omp simd reduction(+ : acc)
  for (uint32_t i = 0; i < d; ++i) acc += a[i] * b[i];
  return acc;
}
typedef struct { float sim; int64_t id; } fcand_t;
static int fbetter(float s1, int64_t i1, float s2, int64_t i2) { return s1 > s2 || (s1 == s2 && i1 < i2); }
static int fcand_cmp(const void* pa, const void* pb) {
  const fcand_t* a = (const fcand_t*)pa; const fcand_t* b = (const fcand_t*)pb;
  if (fbetter(a->sim, a->id, b->sim, b->id)) return -1;
  if (fbetter(b->sim, b->id, a->sim, a->id)) return 1;
  return 0;
}
ORC_API int orc_max_threads(void) {
#ifdef _OPENMP
  return omp_get_max_threads();
#else
  return 1;
#endif
}
ORC_API uint32_t orc_search_fast(const float* rows, uint64_t n, uint32_t d, int64_t id_base, const float* query,
                                 uint32_t k, int threads, int64_t* out_ids, float* out_scores, float* out_sims) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  fcand_t* all = (fcand_t*)malloc(sizeof(fcand_t) * (size_t)threads * k);
  uint32_t* cnts = (uint32_t*)calloc((size_t)threads, sizeof(uint32_t));
#pragma omp parallel num_threads(threads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num(), nt = omp_get_num_threads();
#else
    const int t = 0, nt = 1;
#endif
    fcand_t* best = all + (size_t)t * k; /* sorted, best first */
    uint32_t cnt = 0;
    const uint64_t r0 = n * (uint64_t)t / nt, r1 = n * (uint64_t)(t + 1) / nt;
    for (uint64_t r = r0; r < r1; ++r) {
      const float s = dot_simd(query, rows + r * (uint64_t)d, d);
      const int64_t id = id_base + (int64_t)r;
      if (cnt == k && !fbetter(s, id, best[k - 1].sim, best[k - 1].id)) continue;
      uint32_t pos = cnt < k ? cnt : k - 1;
      while (pos > 0 && fbetter(s, id, best[pos - 1].sim, best[pos - 1].id)) { best[pos] = best[pos - 1]; --pos; }
      best[pos].sim = s; best[pos].id = id;
      if (cnt < k) ++cnt;
    }
    cnts[t] = cnt;
  }
  /* concat + sort + truncate (search.rs:177-181) */
  uint32_t n_all = 0;
  for (int t = 0; t < threads; ++t) {
    if (n_all != (uint32_t)t * k) memmove(all + n_all, all + (size_t)t * k, sizeof(fcand_t) * cnts[t]);
    n_all += cnts[t];
  }
  qsort(all, n_all, sizeof(fcand_t), fcand_cmp);
  const uint32_t cnt = n_all < k ? n_all : k;
  for (uint32_t i = 0; i < cnt; ++i) {
    out_ids[i] = all[i].id;
    if (out_sims) out_sims[i] = all[i].sim;
    out_scores[i] = distance_from_dot(all[i].sim, d);
  }
  free(all); free(cnts);
  return cnt;
}

/* ---- batched CPU baseline ------------------------------------------------------
 * A batch of queries against the same rows, the way a CPU deployment would run the
 * reference's scoring for many queries at once: not B matrix-vector products (each one
 * re-streams the corpus from memory) but ONE blocked pass, sgemm-style — a block of rows
 * is scored against every query while it sits in cache, with a register tile of
 * RB rows x QB queries of 16-lane fp32 accumulators (24 zmm under AVX-512).  Per-thread
 * per-query top-k, then concat + sort + truncate per query (search.rs:177-181).
 * Scores are fp32 dots (SIMD order, FMA contraction allowed): a timed baseline, never
 * the parity oracle. */
typedef float v16f __attribute__((vector_size(64), aligned(4)));
#define RB 4
#define QB 6
__attribute__((target_clones("avx512f", "avx2", "default")))
static void tile_dots(const float* rows, const float* qs, uint32_t d, uint32_t nr, uint32_t nq, float* out /*[RB][QB]*/) {
  v16f acc[RB][QB];
  for (int r = 0; r < RB; ++r)
    for (int q = 0; q < QB; ++q) acc[r][q] = (v16f){0};
  const uint32_t d16 = d & ~15u;
  for (uint32_t c = 0; c < d16; c += 16) {
    v16f xv[RB], qv[QB];
    for (int r = 0; r < RB; ++r) memcpy(&xv[r], rows + (size_t)(r < (int)nr ? r : 0) * d + c, 64);
    for (int q = 0; q < QB; ++q) memcpy(&qv[q], qs + (size_t)(q < (int)nq ? q : 0) * d + c, 64);
    for (int r = 0; r < RB; ++r)
      for (int q = 0; q < QB; ++q) acc[r][q] += xv[r] * qv[q];
  }
  for (int r = 0; r < RB; ++r)
    for (int q = 0; q < QB; ++q) {
      float s = 0.0f;
      for (int l = 0; l < 16; ++l) s += acc[r][q][l];
      if (r < (int)nr && q < (int)nq)
        for (uint32_t c = d16; c < d; ++c) s += rows[(size_t)r * d + c] * qs[(size_t)q * d + c];
      out[r * QB + q] = s;
    }
}
ORC_API void orc_search_batch_fast(const float* rows, uint64_t n, uint32_t d, int64_t id_base, const float* queries,
                                   uint32_t n_queries, uint32_t k, int threads, int64_t* out_ids, float* out_scores,
                                   float* out_sims, uint32_t* out_counts) {
#ifdef _OPENMP
  if (threads <= 0) threads = omp_get_max_threads();
#else
  threads = 1;
#endif
  fcand_t* all = (fcand_t*)malloc(sizeof(fcand_t) * (size_t)threads * n_queries * k);
  uint32_t* cnts = (uint32_t*)calloc((size_t)threads * n_queries, sizeof(uint32_t));
  const uint64_t BLK = 256; /* rows per cache block: 256 x d x 4 B = 384 KB at d = 384 (L2-resident) */
  const uint64_t n_blocks = (n + BLK - 1) / BLK;
#pragma omp parallel num_threads(threads)
  {
#ifdef _OPENMP
    const int t = omp_get_thread_num();
#else
    const int t = 0;
#endif
    fcand_t* mine = all + (size_t)t * n_queries * k;
    uint32_t* mcnt = cnts + (size_t)t * n_queries;
    float tile[RB * QB];
#pragma omp for schedule(dynamic, 4)
    for (uint64_t b = 0; b < n_blocks; ++b) {
      const uint64_t b0 = b * BLK, b1 = b0 + BLK < n ? b0 + BLK : n;
      for (uint32_t q0 = 0; q0 < n_queries; q0 += QB) {
        const uint32_t nq = n_queries - q0 < QB ? n_queries - q0 : QB;
        for (uint64_t r0 = b0; r0 < b1; r0 += RB) {
          const uint32_t nr = (uint32_t)(b1 - r0 < RB ? b1 - r0 : RB);
          tile_dots(rows + r0 * (uint64_t)d, queries + (size_t)q0 * d, d, nr, nq, tile);
          for (uint32_t r = 0; r < nr; ++r)
            for (uint32_t q = 0; q < nq; ++q) {
              const float s = tile[r * QB + q];
              const int64_t id = id_base + (int64_t)(r0 + r);
              fcand_t* best = mine + (size_t)(q0 + q) * k;
              uint32_t cnt = mcnt[q0 + q];
              if (cnt == k && !fbetter(s, id, best[k - 1].sim, best[k - 1].id)) continue;
              uint32_t pos = cnt < k ? cnt : k - 1;
              while (pos > 0 && fbetter(s, id, best[pos - 1].sim, best[pos - 1].id)) { best[pos] = best[pos - 1]; --pos; }
              best[pos].sim = s; best[pos].id = id;
              if (cnt < k) mcnt[q0 + q] = cnt + 1;
            }
        }
      }
    }
  }
  /* per query: concat the threads' lists, sort, truncate (search.rs:177-181) */
  fcand_t* cat = (fcand_t*)malloc(sizeof(fcand_t) * (size_t)threads * k);
  for (uint32_t q = 0; q < n_queries; ++q) {
    uint32_t n_all = 0;
    for (int t = 0; t < threads; ++t) {
      const uint32_t c = cnts[(size_t)t * n_queries + q];
      memcpy(cat + n_all, all + ((size_t)t * n_queries + q) * k, sizeof(fcand_t) * c);
      n_all += c;
    }
    qsort(cat, n_all, sizeof(fcand_t), fcand_cmp);
    const uint32_t cnt = n_all < k ? n_all : k;
    for (uint32_t i = 0; i < k; ++i) {
      out_ids[(size_t)q * k + i] = i < cnt ? cat[i].id : -1;
      if (out_sims) out_sims[(size_t)q * k + i] = i < cnt ? cat[i].sim : -INFINITY;
      out_scores[(size_t)q * k + i] = i < cnt ? distance_from_dot(cat[i].sim, d) : INFINITY;
    }
    if (out_counts) out_counts[q] = cnt;
  }
  free(cat); free(all); free(cnts);
}
