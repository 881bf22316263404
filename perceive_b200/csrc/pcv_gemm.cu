// pcv_gemm.cu — K2: batched exact search on the 5th-gen tensor cores (sm_100a).
//
// Replaces, for a BATCH of queries, the per-query work of Searcher::search_vector
// (crates/perceive-core/search.rs:157-182): every selected row is scored against
// every query (the dot of NdArrayDistance::eval, search.rs:271-274) and the k best
// per query are kept.  The reference has no batched entry point;
// this is the `search_vectors` path behind pcv_search with n_queries >= 16 on a
// bf16 index.
//
// Shape of the kernel (tensor-bound: 2*B*N*d flops, rows streamed ~once from HBM)
//   * persistent grid, one CTA per SM, 6 warps with fixed roles:
//       warp 0      TMA producer (one lane): query tile + document tiles
//       warp 1      MMA issuer (one lane): tcgen05.mma kind::f16, M=128 N=128 K=16
//       warps 2..5  epilogue: tcgen05.ld the fp32 scores out of TMEM, filter, append
//   * QUERY-STATIONARY: a CTA keeps one 128-query tile (all of K, <= 96 KB, 128B-swizzled
//     K-major) resident in shared memory and streams 128-row document tiles through an
//     8-stage TMA ring (16 KB per stage = 128 rows x 64 bf16).  Work items (query tile,
//     document tile) are dealt in query-tile-major order in equal contiguous shares, so
//     CTAs on different query tiles walk the documents in the same order at the same
//     time: a document tile is fetched from HBM once and served to the other query
//     tiles from L2.
//   * four fp32 accumulators of 128 columns fill the 512 TMEM columns: the MMA issuer
//     runs up to three tiles ahead of the epilogue.
//   * fused top-k: thread r of the epilogue owns query r of the tile (TMEM lane r) and
//     keeps that query's running threshold in a register.  A score costs one max/compare;
//     the rare survivor is appended as a u64 ranking key (pcv_common.cuh) to the
//     (CTA, query) candidate buffer in global memory.  A buffer that would overflow is
//     reduced to its k best by the warp (exact; only adversarial inputs get here).
//   * thresholds come from a geometric multi-pass schedule on the host side: pass p
//     covers tiles [T_p, 64*T_p) with the k-th best similarity of everything before T_p
//     as the entry threshold, so each pass appends O(k) candidates per buffer; a small
//     select kernel folds the candidates into the running per-query top-k between passes
//     and emits the final ids/scores.  The B x N score matrix is never written.
//
// Numerics: both operands are bf16 (products exact in fp32), fp32 accumulation inside
// the tensor core in an order the hardware does not specify -> compared with the
// float64 oracle on the same bf16 values under a stated tolerance, not bit for bit.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <cstdlib>
#include <cstring>

#include "pcv_common.cuh"
#include "pcv_gemm_launch.cuh"
#include "pcv_synth.cuh"
#include "pcv_topk.cuh"

namespace pcv {

namespace {

constexpr int G_BM = 128;      // queries per tile (UMMA M, TMEM lanes)
constexpr int G_BN = 128;      // document rows per tile (UMMA N, TMEM columns per accumulator)
constexpr int G_BK = 64;       // bf16 elements per K block = one 128-byte swizzle row
constexpr int G_MAX_KB = 6;    // resident K blocks of the query tile (dim_padded <= 384)
constexpr int G_STAGES = 8;    // document ring depth (K blocks)
constexpr int G_ACC = 4;       // TMEM accumulators
constexpr int G_THREADS = 192;
constexpr uint32_t G_KB_BYTES = G_BM * G_BK * 2;     // 16 KB
constexpr uint32_t G_STAGE_BYTES = G_BN * G_BK * 2;  // 16 KB
constexpr uint32_t G_SMEM_A = G_MAX_KB * G_KB_BYTES;
constexpr uint32_t G_SMEM_B = G_STAGES * G_STAGE_BYTES;
constexpr uint32_t G_NBARS = 2 * G_STAGES + 2 + 2 * G_ACC;
constexpr uint32_t G_SMEM_BYTES = G_SMEM_A + G_SMEM_B + G_NBARS * 8 + 16 + 1024;  // + alignment slack
static_assert(G_SMEM_BYTES <= 232448, "K2 shared memory budget");

struct GemmParams {
  CUtensorMap tmap_q;  // [m_tiles*128][dim_padded] bf16, box 64 x 128, SWIZZLE_128B
  CUtensorMap tmap_x;  // [n_rows][dim_padded] bf16, same box
  const uint2* ranges;
  const uint32_t* range_prefix;
  uint32_t n_ranges;
  uint32_t tile_begin, n_tiles;  // this pass covers document tiles [tile_begin, tile_begin + n_tiles)
  uint32_t m_tiles, n_queries, kb, k;
  uint32_t seg_max, cand_cap;
  uint64_t* cand;       // [grid][seg_max][128][cand_cap]
  uint32_t* cand_cnt;   // [grid][seg_max][128]
  const float* thr;     // [n_queries] entry thresholds (nullable: -inf)
  const uint32_t* lrank_of_row;
};

__device__ __forceinline__ void gemm_tile_rows(const GemmParams& p, uint32_t t, uint32_t& row0, uint32_t& nrows) {
  uint32_t r = 0;
  if (p.n_ranges > 1) {
    uint32_t lo = 0, hi = p.n_ranges;
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (__ldg(p.range_prefix + mid) <= t) lo = mid; else hi = mid;
    }
    r = lo;
  }
  const uint2 rg = __ldg(p.ranges + r);
  row0 = rg.x + (t - __ldg(p.range_prefix + r)) * GEMM_TILE_ROWS;
  nrows = min(GEMM_TILE_ROWS, rg.y - row0);
}

__device__ __forceinline__ uint64_t l2_policy_evict_normal() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_normal.b64 %0, 1.0;" : "=l"(p));
  return p;
}

__global__ void __launch_bounds__(G_THREADS, 1) gemm_topk_kernel(const __grid_constant__ GemmParams p) {
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_a = smem;
  uint8_t* smem_b = smem + G_SMEM_A;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_SMEM_A + G_SMEM_B);
  uint64_t* bar_full = bars;                     // [G_STAGES] TMA -> MMA
  uint64_t* bar_empty = bars + G_STAGES;         // [G_STAGES] MMA -> TMA
  uint64_t* bar_a_full = bars + 2 * G_STAGES;    // query tile landed
  uint64_t* bar_a_free = bars + 2 * G_STAGES + 1;  // every MMA reading the query tile retired
  uint64_t* bar_tfull = bars + 2 * G_STAGES + 2;   // [G_ACC] MMA -> epilogue
  uint64_t* bar_tempty = bar_tfull + G_ACC;        // [G_ACC] epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + G_NBARS);

  // warp index through a shuffle: the compiler then knows every role branch is warp-uniform and
  // keeps descriptors / barrier addresses in uniform registers (no per-lane waterfall around UTCHMMA)
  const int warp = __shfl_sync(PCV_FULL_MASK, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta = blockIdx.x;
  const uint64_t n_items = (uint64_t)p.m_tiles * p.n_tiles;
  const uint64_t i0 = (uint64_t)cta * n_items / gridDim.x;
  const uint64_t i1 = (uint64_t)(cta + 1) * n_items / gridDim.x;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G_STAGES; ++s) {
      mbar_init(smem_u32(bar_full + s), 1);
      mbar_init(smem_u32(bar_empty + s), 1);
    }
    mbar_init(smem_u32(bar_a_full), 1);
    mbar_init(smem_u32(bar_a_free), 1);
    for (int a = 0; a < G_ACC; ++a) {
      mbar_init(smem_u32(bar_tfull + a), 1);
      mbar_init(smem_u32(bar_tempty + a), 4);
    }
    mbar_fence_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    tc_alloc(smem_u32(tmem_slot), 512);
    tc_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);

  if (warp == 0) {
    // ===================== TMA producer (whole warp walks the loop, one elected lane issues) =====
    if (elect_one_sync()) {
      tma_prefetch_desc(&p.tmap_q);
      tma_prefetch_desc(&p.tmap_x);
    }
    const uint64_t pol_q = l2_policy_evict_last();
    const uint64_t pol_x = l2_policy_evict_normal();
    const uint32_t a_base = smem_u32(smem_a), b_base = smem_u32(smem_b);
    const uint32_t full0 = smem_u32(bar_full), empty0 = smem_u32(bar_empty);
    uint32_t stage = 0, phase = 0;
    int64_t cur_m = -1;
    uint32_t n_switch = 0;
    for (uint64_t it = i0; it < i1; ++it) {
      const uint32_t m = (uint32_t)(it / p.n_tiles);
      const uint32_t t = (uint32_t)(it % p.n_tiles) + p.tile_begin;
      if ((int64_t)m != cur_m) {
        if (cur_m >= 0) mbar_wait_bounded(smem_u32(bar_a_free), (n_switch - 1) & 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(smem_u32(bar_a_full), p.kb * G_KB_BYTES);
          for (uint32_t kb = 0; kb < p.kb; ++kb)
            tma_load_2d(a_base + kb * G_KB_BYTES, &p.tmap_q, smem_u32(bar_a_full), (int32_t)(kb * G_BK),
                        (int32_t)(m * G_BM), pol_q);
        }
        __syncwarp();
        cur_m = m;
        ++n_switch;
      }
      uint32_t row0, nrows;
      gemm_tile_rows(p, t, row0, nrows);
      for (uint32_t kb = 0; kb < p.kb; ++kb) {
        mbar_wait_bounded(empty0 + stage * 8, phase ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(full0 + stage * 8, G_STAGE_BYTES);
          tma_load_2d(b_base + stage * G_STAGE_BYTES, &p.tmap_x, full0 + stage * 8, (int32_t)(kb * G_BK),
                      (int32_t)row0, pol_x);
        }
        __syncwarp();
        if (++stage == G_STAGES) { stage = 0; phase ^= 1u; }
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer (whole warp walks the loop, one elected lane issues) =====
    constexpr uint32_t idesc = umma_idesc_bf16_f32(G_BM, G_BN);
    const uint64_t a_desc0 = umma_desc_k_sw128(smem_u32(smem_a));
    const uint64_t b_desc0 = umma_desc_k_sw128(smem_u32(smem_b));
    const uint32_t full0 = smem_u32(bar_full), empty0 = smem_u32(bar_empty);
    const uint32_t tfull0 = smem_u32(bar_tfull), tempty0 = smem_u32(bar_tempty);
    uint32_t stage = 0, phase = 0, acc = 0, acc_par = 0;
    int64_t cur_m = -1;
    uint32_t n_switch = 0;
    for (uint64_t it = i0; it < i1; ++it) {
      const uint32_t m = (uint32_t)(it / p.n_tiles);
      if ((int64_t)m != cur_m) {
        mbar_wait_bounded(smem_u32(bar_a_full), n_switch & 1u);
        cur_m = m;
        ++n_switch;
      }
      mbar_wait_bounded(tempty0 + acc * 8, acc_par ^ 1u);
      tc_fence_after();
      const uint32_t d_tmem = tmem_base + acc * G_BN;
      for (uint32_t kb = 0; kb < p.kb; ++kb) {
        mbar_wait_bounded(full0 + stage * 8, phase);
        tc_fence_after();
        if (elect_one_sync()) {
          // descriptor start-address field counts 16-byte units: advance by adding to the low word
          const uint64_t a_desc = a_desc0 + (uint64_t)((kb * G_KB_BYTES) >> 4);
          const uint64_t b_desc = b_desc0 + (uint64_t)((stage * G_STAGE_BYTES) >> 4);
#pragma unroll
          for (uint32_t j = 0; j < G_BK / 16; ++j)
            tc_mma_bf16(d_tmem, a_desc + j * 2, b_desc + j * 2, idesc, (kb | j) != 0u);
          tc_commit(empty0 + stage * 8);  // frees the ring slot once these MMAs retire
        }
        __syncwarp();
        if (++stage == G_STAGES) { stage = 0; phase ^= 1u; }
      }
      const bool last_of_m = (it + 1 == i1) || ((uint32_t)((it + 1) / p.n_tiles) != m);
      if (elect_one_sync()) {
        tc_commit(tfull0 + acc * 8);
        if (last_of_m) tc_commit(smem_u32(bar_a_free));
      }
      __syncwarp();
      if (++acc == G_ACC) { acc = 0; acc_par ^= 1u; }
    }
  } else {
    // ===================== epilogue: fused top-k filter =====================
    const int quarter = warp & 3;  // TMEM lanes this warp may read
    const int row = quarter * 32 + lane;
    const int k = (int)p.k;
    uint32_t acc = 0, acc_par = 0;
    int64_t cur_m = -1;
    int seg = -1;
    float thr = CUDART_INF_F;
    uint32_t cnt = 0;
    size_t slot = 0;
    uint64_t* buf = nullptr;
    for (uint64_t it = i0; it < i1; ++it) {
      const uint32_t m = (uint32_t)(it / p.n_tiles);
      const uint32_t t = (uint32_t)(it % p.n_tiles) + p.tile_begin;
      if ((int64_t)m != cur_m) {
        if (cur_m >= 0) p.cand_cnt[slot] = cnt;
        ++seg;
        cur_m = m;
        slot = ((size_t)cta * p.seg_max + (size_t)seg) * G_BM + (size_t)row;
        buf = p.cand + slot * p.cand_cap;
        cnt = 0;
        const uint32_t q = m * G_BM + (uint32_t)row;
        thr = (q < p.n_queries) ? (p.thr ? __ldg(p.thr + q) : -CUDART_INF_F) : CUDART_INF_F;
      }
      uint32_t row0, nrows;
      gemm_tile_rows(p, t, row0, nrows);
      mbar_wait_bounded(smem_u32(bar_tfull + acc), acc_par);
      tc_fence_after();
      const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * G_BN;
#pragma unroll 1
      for (uint32_t c0 = 0; c0 < (uint32_t)G_BN; c0 += 32) {
        if (c0 >= nrows) break;  // warp-uniform
        uint32_t v[32];
        tc_ld_32x32b_x32(taddr + c0, v);
        tc_wait_ld();
#pragma unroll
        for (int g = 0; g < 4; ++g) {
          float mx = __uint_as_float(v[8 * g]);
#pragma unroll
          for (int e = 1; e < 8; ++e) mx = fmaxf(mx, __uint_as_float(v[8 * g + e]));
          if (mx >= thr) {
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const float s = __uint_as_float(v[8 * g + e]);
              const uint32_t col = c0 + 8 * g + e;
              if (s >= thr && col < nrows) {
                const uint32_t r = row0 + col;
                const uint32_t lr = p.lrank_of_row ? __ldg(p.lrank_of_row + r) : r;
                buf[cnt++] = make_key(s, lr);
              }
            }
          }
        }
      }
      tc_fence_before();
      __syncwarp();
      if (lane == 0) mbar_arrive(smem_u32(bar_tempty + acc));
      if (++acc == G_ACC) { acc = 0; acc_par ^= 1u; }

      // a buffer that could overflow during the next tile is cut back to its k best
      unsigned need = __ballot_sync(PCV_FULL_MASK, cnt + (uint32_t)G_BN > p.cand_cap);
      while (need) {
        const int L = __ffs(need) - 1;
        need &= need - 1;
        const uint64_t* base = reinterpret_cast<const uint64_t*>(shfl_u64(reinterpret_cast<uint64_t>(buf), L));
        const int n = (int)__shfl_sync(PCV_FULL_MASK, cnt, L);
        __threadfence_block();
        WarpList<4> wl;
        wl.clear();
        wl.merge_unsorted(base, n, k, lane);
        __syncwarp();
        wl.store(const_cast<uint64_t*>(base), k, lane);
        const uint64_t kth = wl.at(k - 1);
        __syncwarp();
        if (lane == L) {
          cnt = (uint32_t)k;
          thr = fmaxf(thr, key_sim(kth));
        }
      }
    }
    if (cur_m >= 0) p.cand_cnt[slot] = cnt;
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_dealloc(tmem_base, 512);
  }
}

// ---------------------------------------------------------------------------
// select: fold one pass's candidate buffers into the running per-query top-k
// (sorted keys) and, on the last pass, emit ids / scores.  One CTA per query.
// ---------------------------------------------------------------------------
struct SelectParams {
  const uint64_t* cand;
  const uint32_t* cand_cnt;
  uint32_t grid_gemm, seg_max, cand_cap, m_tiles, n_tiles, k;
  uint64_t* topk;     // [n_queries][k] running result (in/out)
  int has_prev;
  float* thr;         // [n_queries] out: k-th similarity so far (or -inf)
  int emit;
  uint32_t emit_mode, dim;
  const uint32_t* row_of_lrank;
  const int64_t* ids;
  int64_t id_base;
  int64_t* out_ids;
  float* out_scores;
  float* out_sims;
  uint32_t* out_counts;
};

constexpr int SEL_WARPS = 8;

template <int KPL>
__global__ void __launch_bounds__(SEL_WARPS * 32) gemm_select_kernel(const SelectParams p) {
  extern __shared__ uint64_t sel_stage[];  // [SEL_WARPS][k]
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t q = blockIdx.x;
  const uint32_t m = q / G_BM, row = q % G_BM;
  const int k = (int)p.k;
  const uint64_t n_items = (uint64_t)p.m_tiles * p.n_tiles;

  WarpList<KPL> wl;
  wl.clear();
  if (warp == 0 && p.has_prev) wl.merge_sorted(p.topk + (size_t)q * k, k, k, lane);
  for (uint32_t c = warp; c < p.grid_gemm; c += SEL_WARPS) {
    const uint64_t i0 = (uint64_t)c * n_items / p.grid_gemm;
    const uint64_t i1 = (uint64_t)(c + 1) * n_items / p.grid_gemm;
    if (i0 >= i1) continue;
    const uint32_t m_first = (uint32_t)(i0 / p.n_tiles), m_last = (uint32_t)((i1 - 1) / p.n_tiles);
    if (m < m_first || m > m_last) continue;
    const size_t slot = ((size_t)c * p.seg_max + (m - m_first)) * G_BM + row;
    const int n = (int)p.cand_cnt[slot];
    wl.merge_unsorted(p.cand + slot * p.cand_cap, n, k, lane);
  }
  wl.store(sel_stage + (size_t)warp * k, k, lane);
  __syncthreads();
  if (warp != 0) return;
  WarpList<KPL> out;
  out.clear();
  for (int w2 = 0; w2 < SEL_WARPS; ++w2) out.merge_sorted(sel_stage + (size_t)w2 * k, k, k, lane);
  out.store(p.topk + (size_t)q * k, k, lane);
  const uint64_t kth = out.at(k - 1);
  if (lane == 0) p.thr[q] = kth ? key_sim(kth) : -CUDART_INF_F;
  if (!p.emit) return;
  uint32_t count = 0;
#pragma unroll
  for (int s = 0; s < KPL; ++s) {
    const int e = s * 32 + lane;
    const uint64_t key = out.v[s];
    const bool live = (e < k) && (key != 0ull);
    count += __popc(__ballot_sync(PCV_FULL_MASK, live));
    if (e < k) {
      float sim = -CUDART_INF_F;
      int64_t id = (p.emit_mode == 1) ? INT64_MAX : (int64_t)-1;
      if (live) {
        sim = key_sim(key);
        const uint32_t lr = key_lrank(key);
        const uint32_t r = p.row_of_lrank ? p.row_of_lrank[lr] : lr;
        id = p.ids ? p.ids[r] : p.id_base + (int64_t)r;
      }
      const size_t o = (size_t)q * k + e;
      p.out_ids[o] = id;
      if (p.out_sims) p.out_sims[o] = sim;
      if (p.out_scores) p.out_scores[o] = live ? ref_distance(sim, p.dim) : CUDART_INF_F;
    }
  }
  if (p.out_counts && lane == 0) p.out_counts[q] = count;
}

// fp32 queries (bf16-representable values) -> bf16 [m_tiles*128][dim_padded], zero padded rows
__global__ void queries_to_bf16_kernel(const float* __restrict__ src, uint16_t* __restrict__ dst, uint32_t n_queries,
                                       uint32_t rows_padded, uint32_t dim_padded) {
  const size_t total = (size_t)rows_padded * dim_padded;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += (size_t)gridDim.x * blockDim.x) {
    const uint32_t r = (uint32_t)(i / dim_padded);
    dst[i] = (r < n_queries) ? f32_to_bf16_rne(src[i]) : (uint16_t)0;
  }
}

typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                  const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                  CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = [] {
    void* f = nullptr;
    cudaDriverEntryPointQueryResult q;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &f, cudaEnableDefault, &q) != cudaSuccess ||
        q != cudaDriverEntryPointSuccess)
      f = nullptr;
    return reinterpret_cast<EncodeTiledFn>(f);
  }();
  return fn;
}

// [rows][dim_padded] bf16 row-major, box = 64 elements x 128 rows, 128-byte swizzle, zero fill
bool make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint32_t dim_padded) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {dim_padded, rows};
  const cuuint64_t strides[1] = {(cuuint64_t)dim_padded * 2};
  const cuuint32_t box[2] = {(cuuint32_t)G_BK, (cuuint32_t)G_BN};
  const cuuint32_t estr[2] = {1, 1};
  return fn(out, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, const_cast<void*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) == CUDA_SUCCESS;
}

template <typename T>
cudaError_t reserve(T*& p, size_t& cap, size_t n) {
  if (n <= cap) return cudaSuccess;
  if (p) cudaFree(p);
  p = nullptr;
  cap = 0;
  cudaError_t e = cudaMalloc((void**)&p, n * sizeof(T));
  if (e == cudaSuccess) cap = n;
  return e;
}

uint32_t env_u32(const char* name, uint32_t dflt) {
  const char* e = getenv(name);
  return e ? (uint32_t)strtoul(e, nullptr, 10) : dflt;
}

}  // namespace

void GemmWorkspace::release() {
  if (d_q_bf16) cudaFree(d_q_bf16);
  if (d_cand) cudaFree(d_cand);
  if (d_cand_cnt) cudaFree(d_cand_cnt);
  if (d_topk) cudaFree(d_topk);
  if (d_thr) cudaFree(d_thr);
  d_q_bf16 = nullptr;
  d_cand = nullptr;
  d_cand_cnt = nullptr;
  d_topk = nullptr;
  d_thr = nullptr;
  q_cap = cand_cap = cnt_cap = topk_cap = thr_cap = 0;
}

bool gemm_path_applicable(bool bf16_rows, bool cosine, uint32_t dim_padded, uint32_t n_queries, uint32_t k,
                          uint64_t selected_rows, uint64_t n_rows) {
  if (!bf16_rows || cosine) return false;
  if (dim_padded < (uint32_t)G_BK || dim_padded > (uint32_t)(G_MAX_KB * G_BK)) return false;
  if (k > 128) return false;
  if (n_rows >= 0x7fffff00ull) return false;  // TMA coordinates are int32
  if (n_queries < env_u32("PCV_GEMM_MIN_BATCH", 16)) return false;
  if (selected_rows < env_u32("PCV_GEMM_MIN_ROWS", 4096)) return false;
  return encode_tiled_fn() != nullptr;
}

const char* gemm_search(GemmWorkspace& ws, const GemmCall& c, uint32_t* launches, cudaError_t* err) {
  *err = cudaSuccess;
  uint32_t nl = 0;
  const uint32_t m_tiles = (c.n_queries + G_BM - 1) / G_BM;
  const uint32_t rows_padded = m_tiles * G_BM;
  const uint32_t kb = (c.dim_padded + G_BK - 1) / G_BK;
  const uint32_t cand_cap = std::max<uint32_t>(512u, env_u32("PCV_GEMM_CAND_CAP", 1024));
  const uint32_t ratio = std::max<uint32_t>(2u, env_u32("PCV_GEMM_PASS_RATIO", 64));
  // PCV_GEMM_MAX_CTAS: test knob — fewer CTAs means more tiles per candidate buffer (forces the overflow path)
  const uint32_t sms = std::max<uint32_t>(1u, std::min<uint32_t>((uint32_t)c.sm_count, env_u32("PCV_GEMM_MAX_CTAS", 1u << 20)));
  const uint32_t k = c.k;

#define GCHK(call, what)            \
  do {                              \
    *err = (call);                  \
    if (*err != cudaSuccess) return what; \
  } while (0)

  static bool attr_done[64] = {};
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63]) {
    GCHK(cudaFuncSetAttribute(gemm_topk_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)G_SMEM_BYTES),
         "cudaFuncSetAttribute(gemm_topk_kernel)");
    attr_done[dev & 63] = true;
  }

  // queries -> bf16, padded to whole tiles
  GCHK(reserve(ws.d_q_bf16, ws.q_cap, (size_t)rows_padded * c.dim_padded * 2),
       "query buffer allocation");
  queries_to_bf16_kernel<<<std::min<uint32_t>(1024u, (rows_padded * c.dim_padded + 255) / 256), 256, 0, c.stream>>>(
      c.queries, (uint16_t*)ws.d_q_bf16, c.n_queries, rows_padded, c.dim_padded);
  GCHK(cudaGetLastError(), "queries_to_bf16_kernel launch");
  ++nl;
  GCHK(reserve(ws.d_topk, ws.topk_cap, (size_t)c.n_queries * k), "top-k buffer allocation");
  GCHK(reserve(ws.d_thr, ws.thr_cap, (size_t)c.n_queries), "threshold buffer allocation");

  GemmParams gp;
  memset(&gp, 0, sizeof gp);
  if (!make_tmap(&gp.tmap_q, ws.d_q_bf16, rows_padded, c.dim_padded) ||
      !make_tmap(&gp.tmap_x, c.rows, c.n_rows, c.dim_padded)) {
    *err = cudaErrorInvalidValue;
    return "cuTensorMapEncodeTiled";
  }
  gp.ranges = c.d_ranges;
  gp.range_prefix = c.d_range_prefix;
  gp.n_ranges = c.n_ranges;
  gp.m_tiles = m_tiles;
  gp.n_queries = c.n_queries;
  gp.kb = kb;
  gp.k = k;
  gp.cand_cap = cand_cap;
  gp.lrank_of_row = c.lrank_of_row;

  // geometric pass schedule over the document tiles
  const uint32_t T = c.total_tiles;
  uint32_t first = (uint32_t)std::max<uint64_t>(1, (uint64_t)(cand_cap / 2) * sms / m_tiles / GEMM_TILE_ROWS);
  uint32_t tb = 0;
  bool has_prev = false;
  while (tb < T || (T == 0 && !has_prev)) {
    uint64_t te64 = (tb == 0) ? first : (uint64_t)tb * ratio;
    uint32_t te = (uint32_t)std::min<uint64_t>(te64, T);
    if ((uint64_t)(T - te) * 4 < te) te = T;  // do not leave a sliver for a pass of its own
    const uint32_t nt = te - tb;
    const uint64_t items = (uint64_t)m_tiles * nt;
    const uint32_t grid = (uint32_t)std::max<uint64_t>(1, std::min<uint64_t>(sms, items));
    const uint64_t per_cta = (items + grid - 1) / grid;
    const uint32_t seg_max = nt ? (uint32_t)std::min<uint64_t>(m_tiles, (per_cta + nt - 2) / nt + 1) : 1u;
    const size_t n_slots = (size_t)grid * seg_max * G_BM;
    GCHK(reserve(ws.d_cand, ws.cand_cap, n_slots * cand_cap), "candidate buffer allocation");
    GCHK(reserve(ws.d_cand_cnt, ws.cnt_cap, n_slots), "candidate counter allocation");
    GCHK(cudaMemsetAsync(ws.d_cand_cnt, 0, n_slots * sizeof(uint32_t), c.stream), "cudaMemsetAsync");
    gp.tile_begin = tb;
    gp.n_tiles = nt;
    gp.seg_max = seg_max;
    gp.cand = ws.d_cand;
    gp.cand_cnt = ws.d_cand_cnt;
    gp.thr = has_prev ? ws.d_thr : nullptr;
    if (nt) {
      gemm_topk_kernel<<<grid, G_THREADS, G_SMEM_BYTES, c.stream>>>(gp);
      GCHK(cudaGetLastError(), "gemm_topk_kernel launch");
      ++nl;
    }
    SelectParams sp;
    memset(&sp, 0, sizeof sp);
    sp.cand = ws.d_cand;
    sp.cand_cnt = ws.d_cand_cnt;
    sp.grid_gemm = grid;
    sp.seg_max = seg_max;
    sp.cand_cap = cand_cap;
    sp.m_tiles = m_tiles;
    sp.n_tiles = nt;
    sp.k = k;
    sp.topk = ws.d_topk;
    sp.has_prev = has_prev ? 1 : 0;
    sp.thr = ws.d_thr;
    sp.emit = (te == T) ? 1 : 0;
    sp.emit_mode = c.emit_mode;
    sp.dim = c.dim;
    sp.row_of_lrank = c.row_of_lrank;
    sp.ids = c.ids;
    sp.id_base = c.id_base;
    sp.out_ids = c.out_ids;
    sp.out_scores = c.out_scores;
    sp.out_sims = c.out_sims;
    sp.out_counts = c.out_counts;
    const size_t sel_smem = (size_t)SEL_WARPS * k * sizeof(uint64_t);
    if (k <= 32) gemm_select_kernel<1><<<c.n_queries, SEL_WARPS * 32, sel_smem, c.stream>>>(sp);
    else gemm_select_kernel<4><<<c.n_queries, SEL_WARPS * 32, sel_smem, c.stream>>>(sp);
    GCHK(cudaGetLastError(), "gemm_select_kernel launch");
    ++nl;
    has_prev = true;
    tb = te;
    if (T == 0) break;
  }
#undef GCHK
  if (launches) *launches = nl;
  return nullptr;
}

}  // namespace pcv
