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
