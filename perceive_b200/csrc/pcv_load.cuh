// pcv_load.cuh — load-time kernels: validate, (optionally) L2-normalise, convert
// fp32 rows into the stored layout; and the cross-shard candidate merge (K5).
#pragma once
#include <math_constants.h>
#include "pcv_common.cuh"
#include "pcv_exchange.cuh"
#include "pcv_synth.cuh"

namespace pcv {

#define PCV_LOADFLAG_NONFINITE 1u
#define PCV_LOADFLAG_ZERONORM 2u

// One warp per row.  src: n x dim fp32 dense.  dst: n x dim_padded T (zero
// padded).  normalise: x / max(|x|, 1e-12), the form of
// crates/perceive-core/model/worker.rs:95-103; |x|^2 is summed in the same fixed
// order as the synthetic generator (lane l takes columns l, l+32, ... with
// fmaf, then a 16..1 xor butterfly) so the oracle can mirror it bit for bit.
// PCV_F32_SPLIT (T = uint16_t, dst_lo non-null): the fp32 value is kept EXACTLY as two 16-bit
// planes, hi = its top 16 bits (x truncated to bf16) and lo = its low 16 bits, x == (hi << 16) | lo;
// `stats` collects max |x|^2 and max |x - hi|^2 over the rows (the filter margin of pcv_rescore.cuh).
template <typename T>
__global__ void load_rows_kernel(const float* __restrict__ src, T* __restrict__ dst, T* __restrict__ dst_lo, uint64_t n,
                                 uint32_t dim, uint32_t dim_padded, int normalise, int check_zero,
                                 unsigned int* __restrict__ flags, unsigned int* __restrict__ stats) {
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
    float ss = 0.0f;  // |stored row|^2 in fp32, as the cosine kernels will compute it
    float xx = 0.0f, ee = 0.0f;
    for (uint32_t c = lane; c < dim_padded; c += 32) {
      float x = 0.0f;
      if (c < dim) x = normalise ? (in[c] / div) : in[c];
      if constexpr (sizeof(T) == 4) {
        out[c] = x;
        ss = fmaf(x, x, ss);
      } else if (dst_lo) {
        const uint32_t bits = __float_as_uint(x);
        const uint32_t h = split_hi_bits(bits);
        out[c] = (uint16_t)h;
        dst_lo[r * (uint64_t)dim_padded + c] = (uint16_t)bits;
        const float e = x - __uint_as_float(h << 16);
        xx = fmaf(x, x, xx);
        ee = fmaf(e, e, ee);
      } else {
        const uint16_t h = f32_to_bf16_rne(x);
        out[c] = h;
        const float hv = bf16_to_f32(h);
        ss = fmaf(hv, hv, ss);
      }
    }
    if (stats && dst_lo) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        xx += __shfl_xor_sync(PCV_FULL_MASK, xx, off);
        ee += __shfl_xor_sync(PCV_FULL_MASK, ee, off);
      }
      if (lane == 0 && !bad) {
        atomicMax(stats, __float_as_uint(xx));
        atomicMax(stats + 1, __float_as_uint(ee));
      }
    }
    if (__any_sync(PCV_FULL_MASK, bad) && lane == 0) atomicOr(flags, PCV_LOADFLAG_NONFINITE);
    if (check_zero) {
      // cosine divides by the norm with no epsilon (lib.rs:67-77): a row whose fp32 sum of squares underflows to
      // 0 (|x| below ~1e-19) or overflows would turn every similarity into NaN / 0 silently — refuse it here
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) ss += __shfl_xor_sync(PCV_FULL_MASK, ss, off);
      if (!(ss > 0.0f && ss < CUDART_INF_F) && lane == 0) atomicOr(flags, PCV_LOADFLAG_ZERONORM);
    }
  }
}

// the buffer an ncclAllGather produces -> final results
__global__ void merge_candidates_kernel(const float* __restrict__ sims, size_t sims_list_stride,
                                        const int64_t* __restrict__ ids, size_t ids_list_stride,
                                        uint32_t n_lists, uint32_t n_queries, uint32_t k,
                                        uint32_t dim, int cosine, int64_t* __restrict__ out_ids,
                                        float* __restrict__ out_scores, float* __restrict__ out_sims,
                                        uint32_t* __restrict__ out_counts) {
  const int lane = threadIdx.x & 31;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n_queries) return;
  merge_lists_warp(sims, sims_list_stride, ids, ids_list_stride, n_lists, q, k, dim, cosine, out_ids, out_scores,
                   out_sims, out_counts, lane);
}

__global__ void __launch_bounds__(256) p2p_exchange_merge_kernel(const P2PParams p) {
  __shared__ int s_last;
  const size_t half = p2p_half_bytes(p.world, p.cap) * (p.epoch & 1u);
  const size_t ids_off = (size_t)p.world * p.cap * 4;
  const size_t flags_off = (size_t)p.world * p.cap * 12;
  const uint32_t n_rec = p.n_queries * p.k;
  // ---- 1. store this shard's candidates into every rank's buffer (peer memory, NVLink) ----
  for (uint32_t dst = 0; dst < p.world; ++dst) {
    uint8_t* base = p.peer[dst] + half;
    float* d_sims = reinterpret_cast<float*>(base) + (size_t)p.rank * p.cap;
    int64_t* d_ids = reinterpret_cast<int64_t*>(base + ids_off) + (size_t)p.rank * p.cap;
    for (uint32_t i = blockIdx.x * blockDim.x + threadIdx.x; i < n_rec; i += gridDim.x * blockDim.x) {
      d_sims[i] = p.s_sims[i];
      d_ids[i] = p.s_ids[i];
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const unsigned prev = atomicAdd(p.done_ctr, 1u);
    s_last = (prev == gridDim.x - 1);
  }
  __syncthreads();
  if (s_last) {
    // every CTA's stores are ordered before this point: publish the epoch to every rank
    __threadfence_system();
    if (threadIdx.x < p.world) {
      unsigned int* flag = reinterpret_cast<unsigned int*>(p.peer[threadIdx.x] + half + flags_off) + p.rank;
      st_release_sys_u32(flag, p.epoch);
    }
    if (threadIdx.x == 0) *p.done_ctr = 0u;
  }
  // ---- 2. wait until every shard's candidates have landed here, then merge ----------------
  const int lane = threadIdx.x & 31;
  const uint8_t* mine = p.peer[p.rank] + half;
  const unsigned int* flags = reinterpret_cast<const unsigned int*>(mine + flags_off);
  if ((uint32_t)lane < p.world) {
    uint32_t spins = 0;
    while (ld_acquire_sys_u32(flags + lane) != p.epoch) {
      if (++spins > (1u << 27)) __trap();  // a peer died: fail the launch instead of hanging the GPU
    }
  }
  __syncwarp();
  const float* r_sims = reinterpret_cast<const float*>(mine);
  const int64_t* r_ids = reinterpret_cast<const int64_t*>(mine + ids_off);
  const uint32_t warps = (gridDim.x * blockDim.x) >> 5;
  for (uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; q < p.n_queries; q += warps)
    merge_lists_warp(r_sims, p.cap, r_ids, p.cap, p.world, q, p.k, p.dim, p.cosine, p.out_ids, p.out_scores,
                     p.out_sims, p.out_counts, lane);
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

// ---------------------------------------------------------------------------
// Highlighter scoring (crates/perceive-core/model/highlight.rs:103-127): one
// query against the chunk encodings of a handful of documents, then the best
// chunk of each document.  One CTA per document; each warp takes chunks of its
// document in turn (lane-strided fp32 FMA, butterfly reduce), then warp 0 picks
// the maximum.  itertools' position_max_by keeps the LAST of equal maxima.
// chunk_end[d] = chunks before the end of document d (cumulative).
// ---------------------------------------------------------------------------
__global__ void __launch_bounds__(128) best_chunk_kernel(const float* __restrict__ query, const float* __restrict__ chunks,
                                                          uint32_t dim, const uint32_t* __restrict__ chunk_end,
                                                          float* __restrict__ scores, int32_t* __restrict__ best,
                                                          float* __restrict__ best_score) {
  const uint32_t d = blockIdx.x;
  const uint32_t c0 = d ? chunk_end[d - 1] : 0u, c1 = chunk_end[d];
  const uint32_t lane = threadIdx.x & 31u, warp = threadIdx.x >> 5;
  for (uint32_t c = c0 + warp; c < c1; c += 4) {
    const float* row = chunks + (size_t)c * dim;
    float acc = 0.0f;
    for (uint32_t i = lane; i < dim; i += 32) acc = fmaf(query[i], row[i], acc);
#pragma unroll
    for (int off = 16; off; off >>= 1) acc += __shfl_xor_sync(PCV_FULL_MASK, acc, off);
    if (lane == 0) scores[c] = acc;
  }
  __syncthreads();
  if (warp != 0) return;
  float bs = 0.0f;
  int32_t bi = -1;
  for (uint32_t c = c0 + lane; c < c1; c += 32) {
    const float s = scores[c];
    if (bi < 0 || s >= bs) { bs = s; bi = (int32_t)(c - c0); }  // ascending c: >= keeps the last maximum
  }
#pragma unroll
  for (int off = 16; off; off >>= 1) {
    const float os = __shfl_xor_sync(PCV_FULL_MASK, bs, off);
    const int32_t oi = __shfl_xor_sync(PCV_FULL_MASK, bi, off);
    if (oi >= 0 && (bi < 0 || os > bs || (os == bs && oi > bi))) { bs = os; bi = oi; }
  }
  if (lane == 0) { best[d] = bi; best_score[d] = bs; }
}

}  // namespace pcv
