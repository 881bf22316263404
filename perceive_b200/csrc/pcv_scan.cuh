// pcv_scan.cuh — K1: memory-bound exact scan with fused warp top-k (sm_100a).
//
// Replaces the per-query work of Searcher::search_vector
// (crates/perceive-core/search.rs:157-182): instead of walking one HNSW graph
// per source and calling NdArrayDistance::eval (search.rs:266-279) on the
// visited nodes, every selected row is scored exactly and the k best are kept.
//
// Shape of the kernel (HBM-bound: N*row_bytes streamed once per launch)
//   * persistent grid, one CTA per SM, SCAN_WARPS warps, no CTA-wide sync in
//     the streaming loop: every warp owns a private ring of `nslots` shared
//     memory slots and an mbarrier per slot.  Lane 0 feeds the ring with 1-D
//     bulk async copies (TMA engine, cp.async.bulk ... mbarrier::complete_tx),
//     one copy = one tile of `tile_rows` whole rows (contiguous in HBM).
//   * tiles are dealt round-robin to the grid's warps, so at any instant the
//     chip streams one contiguous window of the matrix.
//   * scoring: LPR lanes share a row (LPR = 8/16/32 by dimension), each lane
//     owns the 16-byte chunks g, g+LPR, ... of the row, reads them with
//     conflict-free LDS.128, multiplies with its query slice held in registers
//     (or shared memory for wider batches) and the partial sums are combined in
//     a FIXED order (documented below) so a row's fp32 score is bit-identical
//     whatever warp, CTA, shard or GPU scores it.
//   * top-k: one compare against the warp's k-th key per row; winners enter a
//     register-resident sorted list (pcv_topk.cuh).  Warp lists -> CTA list ->
//     global partials; the last CTA to finish merges the partials and writes
//     the final ids/scores itself (no second launch).
//
// fp32 summation order, v1 (the oracle mirrors it exactly; oracle/oracle.c
// `orc_dot_v1`): EPC = elements per 16-byte chunk (4 fp32 / 8 bf16).
//   lane g of LPR: acc[c] = fma(q[e], x[e], acc[c]) over chunks j = 0..NJ-1 in
//   order, c = element within chunk;  s = pairwise tree over acc[0..EPC);
//   then s += shfl_xor(s, off) for off = LPR/2 ... 1.
//
// PCV_F32_SPLIT rows (T = SplitF32): the SAME fp32 values held as two 16-bit planes — hi = the
// top 16 bits of x (x truncated to bf16), lo = the low 16 bits — so that the tensor-core filter
// (pcv_gemm.cu) can stream the hi plane alone.  A tile is two bulk copies (hi rows, lo rows); x is
// rebuilt exactly as (hi << 16) | lo (one PRMT) and summed in the fp32 order above, so a split
// index and an fp32 index return bit-identical results.
//
// GROUPED = true: one launch walks a LIST of queries NB at a time (the list and its length live
// in device memory), re-streaming the matrix per group.  Used for batches over fp32 rows and for
// the queries the split filter could not prove complete (pcv_rescore.cuh), whose number the host
// does not know when it enqueues the launch.
#pragma once
#include <math_constants.h>
#include <type_traits>
#include "pcv_common.cuh"
#include "pcv_exchange.cuh"
#include "pcv_topk.cuh"

namespace pcv {

constexpr int SCAN_WARPS = 8;
constexpr int SCAN_MAX_GROUPS = 64;  // groups of NB queries one GROUPED launch may walk (bounds the partial lists)
constexpr int SCAN_THREADS = SCAN_WARPS * 32;
constexpr int SCAN_MAX_SLOTS = 8;
constexpr int SCAN_MAX_NB = 8;

// keys the last CTA's head-threshold selection may gather before it falls back to the radix select (pcv_topk.cuh)
__host__ __device__ inline uint32_t scan_sel_cap(uint32_t k) { return 4u * k; }

struct SplitF32 {};  // storage tag: fp32 values as a hi (top 16 bits = truncated bf16) and a lo (low 16 bits) plane

struct ScanParams {
  const uint8_t* rows;           // stored matrix, row-major, row_bytes per row (split: the hi plane)
  const uint8_t* rows_lo;        // split: the lo plane (plane rows are row_bytes / 2 apart)
  uint32_t row_bytes;            // multiple of 16 (split: bytes of the fp32 row the two planes encode)
  uint32_t d_chunks;             // 16-byte chunks per row
  uint32_t lpr_log2;             // log2(lanes per row)
  uint32_t tile_iters;           // iterations (of 32/LPR rows) per tile
  uint32_t tile_rows;            // (32/LPR) * tile_iters
  uint32_t slot_bytes;           // tile_rows * row_bytes
  uint32_t nslots;               // ring depth per warp
  const uint32_t* range_prefix;  // [n_ranges+1] tiles before range r
  const uint2* ranges;           // [n_ranges] (row_begin, row_end)
  uint32_t n_ranges;
  uint2 range0;                  // ranges[0], for n_ranges == 1
  uint32_t total_tiles;
  const float* queries;          // [NB][q_stride] fp32, zero padded
  uint32_t q_stride;             // elements
  uint32_t nb;                   // live queries in this launch (<= NB)
  uint32_t k;
  uint32_t dim;                  // logical dimension (reference distance divisor)
  uint32_t emit_mode;            // 0 final results, 1 (sim,id) candidates, 2 candidates delivered straight into every
                                 // shard's receive buffer + merge, all by the last CTA (xchg; k <= 128, one group)
  uint32_t l2_evict_first;
  const uint32_t* lrank_of_row;  // nullable: identity
  const uint32_t* row_of_lrank;  // nullable: identity
  const int64_t* ids;            // nullable: id = id_base + row
  int64_t id_base;
  uint64_t* partial;             // [grid][NB][k]
  unsigned int* done;            // last-block counter (self-resetting)
  int64_t* out_ids;              // [NB][k]
  float* out_scores;             // [NB][k]  (nullable in candidate mode)
  float* out_sims;               // [NB][k]  (nullable)
  uint32_t* out_counts;          // [NB]     (nullable in candidate mode)
  // GROUPED launches: queries / out_* are the bases of ALL queries; group g scores the queries
  // q_list[g*NB .. g*NB+NB) (q_list null: g*NB + b) of n_listed (q_count non-null: *q_count) entries
  const uint32_t* q_list;
  const uint32_t* q_count;
  uint32_t n_listed;
  uint32_t group_begin, group_count;  // this launch handles groups [group_begin, group_begin + group_count)
  ExchangeTarget xchg;                // emit_mode 2
};

template <typename T> struct Chunk;
template <> struct Chunk<float> {
  static constexpr int EPC = 4;
  static __device__ __forceinline__ void unpack(const uint4& r, float* f) {
    f[0] = __uint_as_float(r.x); f[1] = __uint_as_float(r.y);
    f[2] = __uint_as_float(r.z); f[3] = __uint_as_float(r.w);
  }
};
template <> struct Chunk<uint16_t> {  // bf16 bits
  static constexpr int EPC = 8;
  static __device__ __forceinline__ void unpack(const uint4& r, float* f) {
    f[0] = __uint_as_float(r.x << 16); f[1] = __uint_as_float(r.x & 0xffff0000u);
    f[2] = __uint_as_float(r.y << 16); f[3] = __uint_as_float(r.y & 0xffff0000u);
    f[4] = __uint_as_float(r.z << 16); f[5] = __uint_as_float(r.z & 0xffff0000u);
    f[6] = __uint_as_float(r.w << 16); f[7] = __uint_as_float(r.w & 0xffff0000u);
  }
};

template <> struct Chunk<SplitF32> {
  static constexpr int EPC = 4;
  // hi: two words of two bf16 each (the top halves); lo: the matching low halves.  x = (hi << 16) | lo
  static __device__ __forceinline__ void unpack(const uint2& hi, const uint2& lo, float* f) {
    f[0] = __uint_as_float(__byte_perm(lo.x, hi.x, 0x5410));
    f[1] = __uint_as_float(__byte_perm(lo.x, hi.x, 0x7632));
    f[2] = __uint_as_float(__byte_perm(lo.y, hi.y, 0x5410));
    f[3] = __uint_as_float(__byte_perm(lo.y, hi.y, 0x7632));
  }
};

template <int EPC>
__device__ __forceinline__ float tree_sum(const float* a) {
  if constexpr (EPC == 4) return (a[0] + a[1]) + (a[2] + a[3]);
  else return ((a[0] + a[1]) + (a[2] + a[3])) + ((a[4] + a[5]) + (a[6] + a[7]));
}
__device__ __forceinline__ float group_sum(float s, int lpr_log2) {
  if (lpr_log2 >= 5) s += __shfl_xor_sync(PCV_FULL_MASK, s, 16);
  if (lpr_log2 >= 4) s += __shfl_xor_sync(PCV_FULL_MASK, s, 8);
  s += __shfl_xor_sync(PCV_FULL_MASK, s, 4);
  s += __shfl_xor_sync(PCV_FULL_MASK, s, 2);
  s += __shfl_xor_sync(PCV_FULL_MASK, s, 1);
  return s;
}

// tile index -> (first row, number of rows)
__device__ __forceinline__ void tile_rows_of(const ScanParams& p, uint32_t t, uint32_t& row0,
                                             uint32_t& nrows) {
  if (p.n_ranges == 1) {  // an unfiltered scan: the range rides in the parameters, the first copy waits for no load
    row0 = p.range0.x + t * p.tile_rows;
    nrows = min(p.tile_rows, p.range0.y - row0);
    return;
  }
  uint32_t r = 0;
  if (p.n_ranges > 1) {
    uint32_t lo = 0, hi = p.n_ranges;  // prefix[lo] <= t < prefix[hi]
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(p.range_prefix + mid) <= t) lo = mid; else hi = mid;
    }
    r = lo;
  }
  const uint2 rg = __ldg(p.ranges + r);
  row0 = rg.x + (t - __ldg(p.range_prefix + r)) * p.tile_rows;
  nrows = min(p.tile_rows, rg.y - row0);
}

template <typename T, int NJ, int NB, int KPL, bool COSINE, bool GROUPED = false>
__global__ void __launch_bounds__(SCAN_THREADS, 1) scan_kernel(const __grid_constant__ ScanParams p) {
  constexpr int EPC = Chunk<T>::EPC;
  constexpr bool SPLIT = std::is_same<T, SplitF32>::value;
  constexpr bool QREG = (NB * NJ * EPC <= 96);  // query slice in registers, else shared memory
  extern __shared__ __align__(128) uint8_t smem[];
  __shared__ int s_last;
  __shared__ BlockSelectScratch s_sel;

  const int lane = threadIdx.x & 31;
  const int warp = threadIdx.x >> 5;
  const int lpr_log2 = (int)p.lpr_log2;
  const int LPR = 1 << lpr_log2;
  const int g = lane & (LPR - 1);
  const int rsub = lane >> lpr_log2;
  const int RPI = 32 >> lpr_log2;
  const int k = (int)p.k;

  const uint32_t ring_bytes = p.nslots * p.slot_bytes;
  uint8_t* ring = smem + (size_t)warp * ring_bytes;
  uint8_t* after_rings = smem + (size_t)SCAN_WARPS * ring_bytes;
  uint64_t* bars = reinterpret_cast<uint64_t*>(after_rings) + warp * SCAN_MAX_SLOTS;
  float* q_smem = reinterpret_cast<float*>(after_rings + SCAN_WARPS * SCAN_MAX_SLOTS * sizeof(uint64_t));
  // split rows: a slot holds the tile's hi rows, then (at a fixed offset) its lo rows
  const uint32_t plane_row_bytes = SPLIT ? p.row_bytes / 2 : p.row_bytes;
  const uint32_t lo_off = SPLIT ? p.tile_rows * plane_row_bytes : 0u;

  if (lane == 0) {
    for (uint32_t s = 0; s < p.nslots; ++s) mbar_init(smem_u32(bars + s), 1);
    mbar_fence_init();
    fence_proxy_async_smem();
  }

  const uint32_t gw = blockIdx.x * SCAN_WARPS + warp;
  const uint32_t GW = gridDim.x * SCAN_WARPS;
  uint64_t policy = 0;
  if (p.l2_evict_first) policy = l2_policy_evict_first();

  auto issue = [&](uint32_t t, uint32_t slot) {
    uint32_t row0, nrows;
    tile_rows_of(p, t, row0, nrows);
    const uint32_t bytes = nrows * plane_row_bytes;
    const uint32_t bar = smem_u32(bars + slot);
    mbar_arrive_expect_tx(bar, SPLIT ? 2 * bytes : bytes);
    const uint32_t dst = smem_u32(ring + (size_t)slot * p.slot_bytes);
    const uint8_t* src = p.rows + (size_t)row0 * plane_row_bytes;
    if (p.l2_evict_first) bulk_g2s_hint(dst, src, bytes, bar, policy);
    else bulk_g2s(dst, src, bytes, bar);
    if constexpr (SPLIT) {
      const uint8_t* src_lo = p.rows_lo + (size_t)row0 * plane_row_bytes;
      if (p.l2_evict_first) bulk_g2s_hint(dst + lo_off, src_lo, bytes, bar, policy);
      else bulk_g2s(dst + lo_off, src_lo, bytes, bar);
    }
  };

  // ring position: persists across the groups of a GROUPED launch (every slot is drained between groups)
  uint32_t slot = 0, parity = 0;
  uint32_t n_total = 0, grp_end = 1;
  if constexpr (GROUPED) {
    n_total = p.q_count ? __ldcg(p.q_count) : p.n_listed;
    grp_end = min(p.group_begin + p.group_count, (n_total + NB - 1) / NB);
  }

#pragma unroll 1
  for (uint32_t grp = GROUPED ? p.group_begin : 0u; grp < grp_end; ++grp) {
  // ---- which queries -------------------------------------------------------
  uint32_t nb_live = p.nb;
  uint32_t qidx[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) qidx[b] = (uint32_t)b;
  if constexpr (GROUPED) {
    nb_live = min((uint32_t)NB, n_total - grp * NB);
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      const uint32_t e = grp * NB + (uint32_t)b;
      qidx[b] = ((uint32_t)b < nb_live) ? (p.q_list ? __ldcg(p.q_list + e) : e) : 0u;
    }
  }
  const uint32_t grp_local = GROUPED ? grp - p.group_begin : 0u;
  uint64_t* partial = p.partial + (size_t)grp_local * gridDim.x * NB * k;
  unsigned int* done = p.done + grp_local;

  // ---- query slice --------------------------------------------------------
  float q[QREG ? NB : 1][QREG ? NJ : 1][EPC];
  if constexpr (QREG) {
#pragma unroll
    for (int b = 0; b < NB; ++b)
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const uint32_t c = (uint32_t)g + (uint32_t)j * LPR;
#pragma unroll
        for (int e = 0; e < EPC; ++e)
          q[b][j][e] = (c < p.d_chunks && (uint32_t)b < nb_live) ? __ldg(p.queries + (size_t)qidx[b] * p.q_stride + c * EPC + e) : 0.0f;
      }
  } else {
    for (uint32_t i = threadIdx.x; i < NB * p.q_stride; i += SCAN_THREADS) {
      const uint32_t b = i / p.q_stride, c = i - b * p.q_stride;
      uint32_t qi = b;
      if constexpr (GROUPED) {
        const uint32_t e = grp * NB + b;
        qi = (b < nb_live) ? (p.q_list ? __ldcg(p.q_list + e) : e) : 0u;
      }
      q_smem[i] = (b < nb_live) ? __ldg(p.queries + (size_t)qi * p.q_stride + c) : 0.0f;
    }
  }
  __syncthreads();  // barriers initialised, q_smem filled

  float qnorm[NB];
  if constexpr (COSINE) {
#pragma unroll
    for (int b = 0; b < NB; ++b) {
      float a[EPC];
#pragma unroll
      for (int e = 0; e < EPC; ++e) a[e] = 0.0f;
#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const uint32_t c = (uint32_t)g + (uint32_t)j * LPR;
#pragma unroll
        for (int e = 0; e < EPC; ++e) {
          float qv;
          if constexpr (QREG) qv = q[b][j][e];
          else qv = (c < p.d_chunks) ? q_smem[(size_t)b * p.q_stride + c * EPC + e] : 0.0f;
          a[e] = fmaf(qv, qv, a[e]);
        }
      }
      qnorm[b] = sqrtf(group_sum(tree_sum<EPC>(a), lpr_log2));
    }
  }

  WarpList<KPL> list[NB];
  uint64_t thrk[NB];
  float thrf[NB];
#pragma unroll
  for (int b = 0; b < NB; ++b) {
    list[b].clear();
    thrk[b] = 0ull;
    thrf[b] = -CUDART_INF_F;
  }

  // ---- streaming loop -------------------------------------------------------
  if (lane == 0) {
    uint32_t sl = slot;
    for (uint32_t s = 0; s < p.nslots; ++s) {
      const uint64_t t = (uint64_t)gw + (uint64_t)s * GW;
      if (t < p.total_tiles) issue((uint32_t)t, sl);
      if (++sl == p.nslots) sl = 0;
    }
  }

  for (uint64_t t64 = gw; t64 < p.total_tiles; t64 += GW) {
    const uint32_t t = (uint32_t)t64;
    uint32_t row0, nrows;
    tile_rows_of(p, t, row0, nrows);
    mbar_wait(smem_u32(bars + slot), parity);
    const uint8_t* tile = ring + (size_t)slot * p.slot_bytes;

    for (uint32_t it = 0; it < p.tile_iters; ++it) {
      const uint32_t rit = it * RPI + rsub;  // row within tile (this lane's group)
      if (it * RPI >= nrows) break;          // warp-uniform
      const uint8_t* xrow = tile + (size_t)min(rit, nrows - 1) * plane_row_bytes;
      float acc[NB][EPC];
      float axx[EPC];
#pragma unroll
      for (int b = 0; b < NB; ++b)
#pragma unroll
        for (int e = 0; e < EPC; ++e) acc[b][e] = 0.0f;
#pragma unroll
      for (int e = 0; e < EPC; ++e) axx[e] = 0.0f;

#pragma unroll
      for (int j = 0; j < NJ; ++j) {
        const uint32_t c = (uint32_t)g + (uint32_t)j * LPR;
        if (c < p.d_chunks) {
          float x[EPC];
          if constexpr (SPLIT) {
            const uint2 hi = reinterpret_cast<const uint2*>(xrow)[c];
            const uint2 lo = reinterpret_cast<const uint2*>(xrow + lo_off)[c];
            Chunk<T>::unpack(hi, lo, x);
          } else {
            const uint4 raw = reinterpret_cast<const uint4*>(xrow)[c];
            Chunk<T>::unpack(raw, x);
          }
          if constexpr (COSINE) {
#pragma unroll
            for (int e = 0; e < EPC; ++e) axx[e] = fmaf(x[e], x[e], axx[e]);
          }
          if constexpr (QREG) {
#pragma unroll
            for (int b = 0; b < NB; ++b)
#pragma unroll
              for (int e = 0; e < EPC; ++e) acc[b][e] = fmaf(q[b][j][e], x[e], acc[b][e]);
          } else {
#pragma unroll
            for (int b = 0; b < NB; ++b) {
              const float4* qp = reinterpret_cast<const float4*>(q_smem + (size_t)b * p.q_stride + c * EPC);
#pragma unroll
              for (int h = 0; h < EPC / 4; ++h) {
                const float4 qv = qp[h];
                acc[b][4 * h + 0] = fmaf(qv.x, x[4 * h + 0], acc[b][4 * h + 0]);
                acc[b][4 * h + 1] = fmaf(qv.y, x[4 * h + 1], acc[b][4 * h + 1]);
                acc[b][4 * h + 2] = fmaf(qv.z, x[4 * h + 2], acc[b][4 * h + 2]);
                acc[b][4 * h + 3] = fmaf(qv.w, x[4 * h + 3], acc[b][4 * h + 3]);
              }
            }
          }
        }
      }
      float xnorm = 1.0f;
      if constexpr (COSINE) xnorm = sqrtf(group_sum(tree_sum<EPC>(axx), lpr_log2));
      const bool valid = (rit < nrows) && (g == 0);
#pragma unroll
      for (int b = 0; b < NB; ++b) {
        float sim = group_sum(tree_sum<EPC>(acc[b]), lpr_log2);
        if constexpr (COSINE) sim = sim / (xnorm * qnorm[b]);
        unsigned m = __ballot_sync(PCV_FULL_MASK, valid && (sim >= thrf[b]) && ((uint32_t)b < nb_live));
        while (m) {
          const int src = __ffs(m) - 1;
          m &= m - 1;
          const float s = __shfl_sync(PCV_FULL_MASK, sim, src);
          const uint32_t row = row0 + it * RPI + ((uint32_t)src >> lpr_log2);
          const uint32_t lr = p.lrank_of_row ? __ldg(p.lrank_of_row + row) : row;
          const uint64_t key = make_key(s, lr);
          if (key > thrk[b]) {
            list[b].insert(key, lane);
            thrk[b] = list[b].at(k - 1);
            thrf[b] = thrk[b] ? key_sim(thrk[b]) : -CUDART_INF_F;
          }
        }
      }
    }

    // refill this slot with the tile nslots rounds ahead
    __syncwarp();
    if (lane == 0) {
      const uint64_t t2 = t64 + (uint64_t)p.nslots * GW;
      if (t2 < p.total_tiles) {
        fence_proxy_async_smem();
        issue((uint32_t)t2, slot);
      }
    }
    if (++slot == p.nslots) { slot = 0; parity ^= 1u; }
  }

  // ---- CTA merge: 8 warp lists -> 1 -----------------------------------------
  __syncthreads();  // every ring is drained: shared memory is reusable
  uint64_t* stage = reinterpret_cast<uint64_t*>(smem);  // [NB][warp][k]
#pragma unroll
  for (int b = 0; b < NB; ++b) list[b].store(stage + ((size_t)b * SCAN_WARPS + warp) * k, k, lane);
  __syncthreads();
  if constexpr (KPL <= 2) {
    // k <= 64: every key ranks itself among the 8 k keys of its query by counting — all 256 threads for a
    // few hundred cycles, where one warp walking eight lists insert by insert was a third of a short scan
#pragma unroll 1
    for (int b = 0; b < NB; ++b) {
      if ((uint32_t)b >= nb_live) break;  // block-uniform; the last CTA never reads those
      block_merge_lists(stage + (size_t)b * SCAN_WARPS * k, SCAN_WARPS, (uint32_t)k, partial + ((size_t)blockIdx.x * NB + b) * k);
    }
  } else if constexpr (KPL == 4) {
    // 65 <= k <= 128: one block-wide select over the 8 lists (a warp-list merge is ~100 dependent
    // inserts per list at this size)
    uint64_t* sel = stage + (size_t)NB * SCAN_WARPS * k;
    uint64_t* out = sel + k;
#pragma unroll 1
    for (int b = 0; b < NB; ++b) {
      block_select_sorted(stage + (size_t)b * SCAN_WARPS * k, (uint32_t)(SCAN_WARPS * k), (uint32_t)k, sel, out, s_sel);
      for (int e = threadIdx.x; e < k; e += SCAN_THREADS) partial[((size_t)blockIdx.x * NB + b) * k + e] = out[e];
      __syncthreads();
    }
  } else {
    for (int b = warp; b < NB; b += SCAN_WARPS) {
      WarpList<KPL> m;
      m.clear();
      for (int w2 = 0; w2 < SCAN_WARPS; ++w2) m.merge_sorted(stage + ((size_t)b * SCAN_WARPS + w2) * k, k, k, lane);
      m.store(partial + ((size_t)blockIdx.x * NB + b) * k, k, lane);
    }
  }

  // ---- last CTA merges the grid's partial lists and emits the result --------
  // release: the CTA's stores happen-before the barrier, thread 0's fence is cumulative over them; acquire: thread
  // 0's fence after the counter, the barrier, then reads that go to L2 (ld.global.cg).  One fence per CTA, not 256.
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence();
    const unsigned prev = atomicAdd(done, 1u);
    s_last = (prev == gridDim.x - 1);
    __threadfence();
  }
  __syncthreads();
  if (s_last) {

  if constexpr (KPL <= 4) {
    // k <= 128: pull every CTA's k keys into shared memory in one parallel sweep and select
    // block-wide (radix select + rank by counting) instead of walking 148 lists, one dependent
    // L2 load after another.
    const uint32_t n = gridDim.x * (uint32_t)k;
    uint64_t* keys = reinterpret_cast<uint64_t*>(smem);  // [n], then sel[sel_cap], out[k]
    uint64_t* sel = keys + n;
    const uint32_t sel_cap = scan_sel_cap((uint32_t)k);
    uint64_t* out = sel + sel_cap;
#pragma unroll 1
    for (int b = 0; b < NB; ++b) {
      if ((uint32_t)b >= nb_live) break;  // block-uniform
      __syncthreads();
      // eight loads in flight per thread before the first store: one L2 round trip for 148 x 10 keys, not three
      for (uint32_t base = threadIdx.x; base < n; base += 8u * SCAN_THREADS) {
        uint64_t v[8];
#pragma unroll
        for (uint32_t u = 0; u < 8u; ++u) {
          const uint32_t i = base + u * SCAN_THREADS;
          const uint32_t c = i / (uint32_t)k, e = i - c * (uint32_t)k;
          v[u] = i < n ? __ldcg(reinterpret_cast<const unsigned long long*>(partial) + ((size_t)c * NB + b) * k + e) : 0ull;
        }
#pragma unroll
        for (uint32_t u = 0; u < 8u; ++u) {
          const uint32_t i = base + u * SCAN_THREADS;
          if (i < n) keys[i] = v[u];
        }
      }
      __syncthreads();
      const uint32_t count = block_select_lists(keys, gridDim.x, (uint32_t)k, sel, sel_cap, out, s_sel);
      for (uint32_t e = threadIdx.x; e < (uint32_t)k; e += SCAN_THREADS) {
        const uint64_t key = out[e];
        const bool live = key != 0ull;
        float sim = -CUDART_INF_F;
        int64_t id = (p.emit_mode != 0) ? INT64_MAX : (int64_t)-1;
        if (live) {
          sim = key_sim(key);
          const uint32_t lr = key_lrank(key);
          const uint32_t row = p.row_of_lrank ? p.row_of_lrank[lr] : lr;
          id = p.ids ? p.ids[row] : p.id_base + (int64_t)row;
        }
        const size_t o = (size_t)qidx[b] * k + e;
        if (p.emit_mode == 2) {
          // sharded search, fused exchange: this shard's candidate goes straight into every shard's receive
          // buffer (peer memory over NVLink); the merged result is written further down by this same CTA
          for (uint32_t dst = 0; dst < p.xchg.world; ++dst) {
            float* ps;
            int64_t* pi;
            exchange_slot(p.xchg, dst, (uint32_t)o, ps, pi);
            *ps = sim;
            *pi = id;
          }
          continue;
        }
        p.out_ids[o] = id;
        if (p.out_sims) p.out_sims[o] = sim;
        if (p.out_scores) {
          float sc = CUDART_INF_F;
          if (live) sc = COSINE ? sim : ref_distance(sim, p.dim);
          p.out_scores[o] = sc;
        }
      }
      if (p.out_counts && p.emit_mode != 2 && threadIdx.x == 0) p.out_counts[qidx[b]] = count;
    }
    if (p.emit_mode == 2)
      exchange_publish_and_merge(p.xchg, nb_live, (uint32_t)k, p.dim, COSINE ? 1 : 0, p.out_ids, p.out_scores, p.out_sims, p.out_counts);
  } else {
  // k > 128: warp lists, one query after another (rare: the keys of 148 CTAs x 1024 do not fit)
#pragma unroll 1
  for (int b = 0; b < NB; ++b) {
    WarpList<KPL> m;
    m.clear();
    for (uint32_t c = warp; c < gridDim.x; c += SCAN_WARPS)
      m.merge_sorted_cg(partial + ((size_t)c * NB + b) * k, k, k, lane);
    m.store(stage + ((size_t)b * SCAN_WARPS + warp) * k, k, lane);
  }
  __syncthreads();
  for (int b = warp; b < NB; b += SCAN_WARPS) {
    if ((uint32_t)b >= nb_live) continue;
    WarpList<KPL> m;
    m.clear();
    for (int w2 = 0; w2 < SCAN_WARPS; ++w2) m.merge_sorted(stage + ((size_t)b * SCAN_WARPS + w2) * k, k, k, lane);
    uint32_t count = 0;
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
      const int e = s * 32 + lane;
      const uint64_t key = m.v[s];
      const bool live = (e < k) && (key != 0ull);
      count += __popc(__ballot_sync(PCV_FULL_MASK, live));
      if (e < k) {
        float sim = -CUDART_INF_F;
        int64_t id = (p.emit_mode == 1) ? INT64_MAX : (int64_t)-1;
        if (live) {
          sim = key_sim(key);
          const uint32_t lr = key_lrank(key);
          const uint32_t row = p.row_of_lrank ? p.row_of_lrank[lr] : lr;
          id = p.ids ? p.ids[row] : p.id_base + (int64_t)row;
        }
        const size_t o = (size_t)qidx[b] * k + e;
        p.out_ids[o] = id;
        if (p.out_sims) p.out_sims[o] = sim;
        if (p.out_scores) {
          float sc = CUDART_INF_F;
          if (live) sc = COSINE ? sim : ref_distance(sim, p.dim);
          p.out_scores[o] = sc;
        }
      }
    }
    if (p.out_counts && lane == 0) p.out_counts[qidx[b]] = count;
  }
  }
  if (threadIdx.x == 0) *done = 0u;
  }  // s_last
  if constexpr (GROUPED) __syncthreads();  // shared memory (staging / keys) is the next group's ring
  }  // groups
}

// bytes of dynamic shared memory a launch needs
inline size_t scan_smem_bytes(const ScanParams& p, int nb_template, bool q_in_smem, int grid) {
  size_t ring = (size_t)SCAN_WARPS * p.nslots * p.slot_bytes;
  size_t bars = (size_t)SCAN_WARPS * SCAN_MAX_SLOTS * sizeof(uint64_t);
  size_t q = q_in_smem ? (size_t)nb_template * p.q_stride * sizeof(float) : 0;
  size_t stage = ((size_t)SCAN_WARPS * nb_template * p.k + 2 * (size_t)p.k) * sizeof(uint64_t);
  // last-CTA block select (k <= 128): every CTA's k keys of one query + sel[k] + out[k]
  size_t select = p.k <= 128 ? ((size_t)grid * p.k + scan_sel_cap(p.k) + (size_t)p.k) * sizeof(uint64_t) : 0;
  size_t total = ring + bars + q;
  total = total > stage ? total : stage;
  return total > select ? total : select;
}

}  // namespace pcv
