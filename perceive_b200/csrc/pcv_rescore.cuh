// pcv_rescore.cuh — K3, second half: exact fp32 rescoring of the tensor-core filter's candidates.
//
// BASELINE config 4 asks for fp32 results from a batched search (256 queries x 10^8 rows).  A
// PCV_F32_SPLIT index holds every fp32 value x as two 16-bit planes: hi = the top 16 bits of x (x truncated
// to bf16) and lo = its low 16 bits.  A batched search is
//   1. filter  — the tcgen05 kernel (pcv_gemm.cu) over the hi plane ALONE (2 of the 4 bytes per
//      element, one MMA per element) keeps the kf best rows per query by t = bf16(q) . hi(x);
//   2. rescore — this file: the kf candidates of a query are rebuilt exactly from both planes and
//      scored in fp32 in the scan kernel's summation order (pcv_scan.cuh "v1"), so the similarity of a
//      row is bit-identical to what K1 computes for it (the quantity NdArrayDistance::eval defines,
//      crates/perceive-core/search.rs:266-279); the k best by (fp32 similarity, lower id) are emitted;
//   3. proof   — |q.x - bf16(q).hi(x)| <= |q - bf16(q)| |x| + |bf16(q)| |x - hi(x)| (Cauchy-Schwarz)
//      <= margin(q), computed from the query and two maxima taken over the stored rows at load time,
//      plus a bound on fp32 accumulation error on both sides.  A row outside the candidate set has
//      t <= t_kf (the smallest candidate score), hence an exact similarity <= t_kf + margin.  When that
//      is below the k-th best rescored similarity the result is provably the exact top-k.  Otherwise
//      (all-equal scores, adversarially clustered rows) the query is appended to a fallback list that
//      one GROUPED launch of the exact scan (K1 over both planes) then resolves.
// Results are therefore bit-identical to an fp32 index searched by K1, whatever the data.
#pragma once
#include <math_constants.h>
#include "pcv_common.cuh"
#include "pcv_scan.cuh"
#include "pcv_synth.cuh"
#include "pcv_topk.cuh"

namespace pcv {

// candidates kept by the filter for a result size of k (<= 128, the tensor path's limit)
__host__ __device__ __forceinline__ uint32_t split_filter_k(uint32_t k) {
  const uint32_t want = 2u * k + 28u;
  return want < 48u ? 48u : (want > 128u ? 128u : want);
}

// Per-query filter margin.  stats[0], stats[1]: bit patterns of max |x|^2 and max |x - hi(x)|^2 over
// the stored rows (non-negative floats order like their bits; maintained by the load kernels).
// One warp per query; fp32 sums inflated by 2^-10 so they stay upper bounds.
__global__ void split_query_margin_kernel(const float* __restrict__ q, uint32_t n_queries, uint32_t stride,
                                          const unsigned int* __restrict__ stats, float* __restrict__ margin) {
  const int lane = threadIdx.x & 31;
  const uint32_t w = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (w >= n_queries) return;
  float qq = 0.0f, hh = 0.0f, ee = 0.0f;
  for (uint32_t c = lane; c < stride; c += 32) {
    const float x = q[(size_t)w * stride + c];
    const float h = bf16_to_f32(f32_to_bf16_rne(x));
    const float e = x - h;
    qq = fmaf(x, x, qq);
    hh = fmaf(h, h, hh);
    ee = fmaf(e, e, ee);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) {
    qq += __shfl_xor_sync(PCV_FULL_MASK, qq, off);
    hh += __shfl_xor_sync(PCV_FULL_MASK, hh, off);
    ee += __shfl_xor_sync(PCV_FULL_MASK, ee, off);
  }
  if (lane == 0) {
    const float up = 1.0009765625f;  // 1 + 2^-10
    const float xmax = sqrtf(__uint_as_float(stats[0])) * up;
    const float emax = sqrtf(__uint_as_float(stats[1])) * up;
    const float data = (sqrtf(ee) * xmax + sqrtf(hh) * emax) * up;
    // accumulation error of the tensor-core dot and of the fp32 rescoring: <= stride * 2^-22 |q||x|
    const float arith = (float)stride * 2.384185791015625e-7f * sqrtf(qq) * xmax;
    margin[w] = data + arith;
  }
}

struct RescoreParams {
  const uint64_t* cand;  // [n_queries][kf] filter keys (tensor-core score, local rank), sorted descending, 0 = empty
  uint32_t kf, k, n_queries;
  uint32_t q_offset;     // this launch's first query: cand / margin are indexed from 0, queries / outputs from q_offset
  const uint8_t* hi;
  const uint8_t* lo;
  uint32_t plane_row_bytes;  // dim_padded * 2
  uint32_t d_chunks;         // 16-byte chunks of the fp32 row (dim_padded / 4)
  uint32_t lpr_log2;         // lanes per row, as the scan kernel picks them for this dimension
  const float* queries;      // fp32 [q_offset + n_queries][q_stride], zero padded
  uint32_t q_stride;
  const float* margin;       // [n_queries]
  const uint32_t* row_of_lrank;
  const int64_t* ids;
  int64_t id_base;
  uint32_t emit_mode, dim;
  int64_t* out_ids;
  float* out_scores;
  float* out_sims;
  uint32_t* out_counts;
  uint32_t* fb_list;   // queries whose candidate set could not be proven complete
  uint32_t* fb_count;
};

// One CTA per query, 8 warps; 32/LPR candidates per warp step, LPR lanes per candidate row.
template <int NJ>
__global__ void __launch_bounds__(256) rescore_exact_kernel(const RescoreParams p) {
  __shared__ uint64_t s_keys[128], s_selk[128], s_out[128];
  __shared__ BlockSelectScratch s_sc;
  const uint32_t ql = blockIdx.x, q = p.q_offset + blockIdx.x;
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int lpr_log2 = (int)p.lpr_log2;
  const int LPR = 1 << lpr_log2;
  const int g = lane & (LPR - 1);
  const int rsub = lane >> lpr_log2;
  const int RPI = 32 >> lpr_log2;
  const uint32_t kf = p.kf, k = p.k;
  const uint64_t* cand = p.cand + (size_t)ql * kf;

  float qv[NJ][4];
#pragma unroll
  for (int j = 0; j < NJ; ++j) {
    const uint32_t c = (uint32_t)g + (uint32_t)j * LPR;
#pragma unroll
    for (int e = 0; e < 4; ++e) qv[j][e] = (c < p.d_chunks) ? __ldg(p.queries + (size_t)q * p.q_stride + c * 4 + e) : 0.0f;
  }
  if (threadIdx.x < 128) s_keys[threadIdx.x] = 0ull;
  __syncthreads();

  for (uint32_t base = (uint32_t)(warp * RPI); base < kf; base += 8u * RPI) {  // warp-uniform bound
    const uint32_t j = base + (uint32_t)rsub;
    const uint64_t key = (j < kf) ? __ldcg(reinterpret_cast<const unsigned long long*>(cand) + j) : 0ull;
    const bool live = key != 0ull;
    const uint32_t lr = live ? key_lrank(key) : 0u;
    const uint32_t row = live ? (p.row_of_lrank ? __ldg(p.row_of_lrank + lr) : lr) : 0u;
    const uint8_t* xh = p.hi + (size_t)row * p.plane_row_bytes;
    const uint8_t* xl = p.lo + (size_t)row * p.plane_row_bytes;
    float acc[4] = {0.0f, 0.0f, 0.0f, 0.0f};
#pragma unroll
    for (int jj = 0; jj < NJ; ++jj) {
      const uint32_t c = (uint32_t)g + (uint32_t)jj * LPR;
      if (c < p.d_chunks) {
        const uint2 h = __ldg(reinterpret_cast<const uint2*>(xh) + c);
        const uint2 l = __ldg(reinterpret_cast<const uint2*>(xl) + c);
        float x[4];
        Chunk<SplitF32>::unpack(h, l, x);
#pragma unroll
        for (int e = 0; e < 4; ++e) acc[e] = fmaf(qv[jj][e], x[e], acc[e]);
      }
    }
    const float sim = group_sum(tree_sum<4>(acc), lpr_log2);  // the scan kernel's order: bit-identical to K1
    if (live && g == 0) s_keys[j] = make_key(sim, lr);
  }
  __syncthreads();
  const uint32_t count = block_select_sorted(s_keys, kf, k, s_selk, s_out, s_sc);
  for (uint32_t e = threadIdx.x; e < k; e += 256) {
    const uint64_t key = s_out[e];
    const bool live = key != 0ull;
    float sim = -CUDART_INF_F;
    int64_t id = (p.emit_mode == 1) ? INT64_MAX : (int64_t)-1;
    if (live) {
      sim = key_sim(key);
      const uint32_t lr = key_lrank(key);
      const uint32_t row = p.row_of_lrank ? p.row_of_lrank[lr] : lr;
      id = p.ids ? p.ids[row] : p.id_base + (int64_t)row;
    }
    const size_t o = (size_t)q * k + e;
    p.out_ids[o] = id;
    if (p.out_sims) p.out_sims[o] = sim;
    if (p.out_scores) p.out_scores[o] = live ? ref_distance(sim, p.dim) : CUDART_INF_F;
  }
  if (threadIdx.x == 0) {
    if (p.out_counts) p.out_counts[q] = count;
    // proof of completeness (see the header): rows outside the candidate set score <= t_kf + margin
    const uint64_t last = __ldcg(reinterpret_cast<const unsigned long long*>(cand) + (kf - 1));
    const bool full = last != 0ull;  // fewer than kf candidates: every selected row was rescored
    const bool proven = !full || (count == k && key_sim(last) + p.margin[ql] < key_sim(s_out[k - 1]));
    if (!proven) p.fb_list[atomicAdd(p.fb_count, 1u)] = q;
  }
}

}  // namespace pcv
