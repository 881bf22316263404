// pcv_load.cuh — load-time kernels: validate, (optionally) L2-normalise, convert
// fp32 rows into the stored layout; and the cross-shard candidate merge (K5).
#pragma once
#include <math_constants.h>
#include "pcv_common.cuh"
#include "pcv_synth.cuh"

namespace pcv {

#define PCV_LOADFLAG_NONFINITE 1u
#define PCV_LOADFLAG_ZERONORM 2u

// One warp per row.  src: n x dim fp32 dense.  dst: n x dim_padded T (zero
// padded).  normalise: x / max(|x|, 1e-12), the form of
// crates/perceive-core/model/worker.rs:95-103; |x|^2 is summed in the same fixed
// order as the synthetic generator (lane l takes columns l, l+32, ... with
// fmaf, then a 16..1 xor butterfly) so the oracle can mirror it bit for bit.
template <typename T>
__global__ void load_rows_kernel(const float* __restrict__ src, T* __restrict__ dst, uint64_t n,
                                 uint32_t dim, uint32_t dim_padded, int normalise, int check_zero,
                                 unsigned int* __restrict__ flags) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < n; r += nwarps) {
    const float* in = src + r * (uint64_t)dim;
    float part = 0.0f;
    bool bad = false;
    for (uint32_t c = lane; c < dim; c += 32) {
      const float x = in[c];
      bad |= !isfinite(x);
      part = fmaf(x, x, part);
    }
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) part = part + __shfl_xor_sync(PCV_FULL_MASK, part, off);
    const float div = normalise ? fmaxf(sqrtf(part), 1e-12f) : 1.0f;
    T* out = dst + r * (uint64_t)dim_padded;
    bool nonzero = false;
    for (uint32_t c = lane; c < dim_padded; c += 32) {
      float x = 0.0f;
      if (c < dim) x = normalise ? (in[c] / div) : in[c];
      if constexpr (sizeof(T) == 4) {
        out[c] = x;
        nonzero |= (x != 0.0f);
      } else {
        const uint16_t h = f32_to_bf16_rne(x);
        out[c] = h;
        nonzero |= ((h & 0x7fffu) != 0);
      }
    }
    if (__any_sync(PCV_FULL_MASK, bad) && lane == 0) atomicOr(flags, PCV_LOADFLAG_NONFINITE);
    if (check_zero && !__any_sync(PCV_FULL_MASK, nonzero) && lane == 0) atomicOr(flags, PCV_LOADFLAG_ZERONORM);
  }
}

// K5 — merge candidate lists from `n_lists` shards (the buffer an
// ncclAllGather produces).  One warp per query; lane l tracks the head of
// list l.  Lists are sorted (sim desc, id asc) and padded with
// (-inf, INT64_MAX).  Mirrors the concat + sort + truncate of
// crates/perceive-core/search.rs:177-181 across shards instead of sources.
__global__ void merge_candidates_kernel(const float* __restrict__ sims, size_t sims_list_stride,
                                        const int64_t* __restrict__ ids, size_t ids_list_stride,
                                        uint32_t n_lists, uint32_t n_queries, uint32_t k,
                                        uint32_t dim, int cosine, int64_t* __restrict__ out_ids,
                                        float* __restrict__ out_scores, float* __restrict__ out_sims,
                                        uint32_t* __restrict__ out_counts) {
  const int lane = threadIdx.x & 31;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n_queries) return;
  const float* ls = sims + (size_t)lane * sims_list_stride + (size_t)q * k;
  const int64_t* li = ids + (size_t)lane * ids_list_stride + (size_t)q * k;
  uint32_t head = 0;
  uint32_t count = 0;
  for (uint32_t e = 0; e < k; ++e) {
    float s = -CUDART_INF_F;
    int64_t id = INT64_MAX;
    if ((uint32_t)lane < n_lists && head < k) {
      s = ls[head];
      id = li[head];
    }
    uint32_t o = f32_to_ordered(s);
    if (id == INT64_MAX) o = 0u;  // padding never wins over a real candidate
    // warp arg-best over (o desc, id asc)
    uint32_t bo = o;
    int64_t bid = id;
    int bl = lane;
#pragma unroll
    for (int off = 16; off >= 1; off >>= 1) {
      const uint32_t oo = __shfl_xor_sync(PCV_FULL_MASK, bo, off);
      const int64_t oid = __shfl_xor_sync(PCV_FULL_MASK, (long long)bid, off);
      const int ol = __shfl_xor_sync(PCV_FULL_MASK, bl, off);
      const bool take = (oo > bo) || (oo == bo && (oid < bid || (oid == bid && ol < bl)));
      if (take) { bo = oo; bid = oid; bl = ol; }
    }
    const bool live = (bid != INT64_MAX);
    if (lane == bl && live) ++head;
    if (lane == 0) {
      const size_t w = (size_t)q * k + e;
      const float bs = live ? ordered_to_f32(bo) : -CUDART_INF_F;
      out_ids[w] = live ? bid : (int64_t)-1;
      if (out_sims) out_sims[w] = bs;
      if (out_scores) out_scores[w] = live ? (cosine ? bs : ref_distance(bs, dim)) : CUDART_INF_F;
    }
    count += live ? 1u : 0u;
  }
  if (out_counts && lane == 0) out_counts[q] = count;
}

// zero-padded copy of queries: src n x dim -> dst n x stride.  round_bf16: the
// index stores bf16, and a bf16 index computes on bf16 operands on BOTH sides
// (queries rounded to nearest-even here), so the scan (K1) and the tensor-core
// path (K2) score exactly the same numbers.
__global__ void pad_queries_kernel(const float* __restrict__ src, float* __restrict__ dst, uint32_t n,
                                   uint32_t dim, uint32_t stride, int round_bf16) {
  const size_t total = (size_t)n * stride;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(i / stride), c = (uint32_t)(i % stride);
    float x = (c < dim) ? src[(size_t)r * dim + c] : 0.0f;
    if (round_bf16) x = bf16_to_f32(f32_to_bf16_rne(x));
    dst[i] = x;
  }
}

}  // namespace pcv
