// pcv_gemm.cu — K2: batched exact search on the 5th-gen tensor cores (sm_100a).
//
// Replaces, for a BATCH of queries, the per-query work of Searcher::search_vector
// (crates/perceive-core/search.rs:157-182): every selected row is scored against
// every query (the dot of NdArrayDistance::eval, search.rs:271-274) and the k best
// per query are kept.  The reference has no batched entry point;
// this is the `search_vectors` path behind pcv_search with n_queries >= 16 on a
// bf16 index (K2), and — run with `keys_only` over the hi plane of a PCV_F32_SPLIT
// index — the FILTER in front of the exact fp32 rescoring (K3, pcv_rescore.cuh).
//
// Shape of the kernel (tensor-bound: 2*B*N*d flops; every document row leaves HBM once)
//   * persistent grid, one CTA per SM, 7 warps with fixed roles:
//       warp 0      TMA producer for query K-blocks   (whole warp walks the loop, one lane issues)
//       warp 1      MMA issuer: tcgen05.mma kind::f16, M=128 N=128 K=16, fp32 accumulators in TMEM
//       warps 2..5  epilogue: tcgen05.ld the scores out of TMEM, threshold filter, append
//       warp 6      TMA producer for document K-blocks
//   * DOCUMENT-STATIONARY: a CTA owns a contiguous share of the pass's 128-row document
//     tiles.  A tile's K-blocks (16 KB each: 128 rows x 64 bf16, 128B-swizzled K-major) sit
//     in an 8-slot ring in shared memory while EVERY 128-query tile is multiplied against
//     them; the slots are handed back, K-block by K-block, during the last query tile, so
//     the next document tile streams in underneath.  Each document byte is read from HBM
//     exactly once per pass.  Query K-blocks stream through a 6-stage ring from L2 (the
//     whole query batch is < 1 MB and every CTA cycles through the same tiles).
//   * four fp32 accumulators of 128 columns fill the 512 TMEM columns: the MMA issuer
//     runs up to three (query tile, document tile) items ahead of the epilogue.
//   * fused top-k: thread r of the epilogue owns query r of the current query tile (TMEM
//     lane r) with that query's running threshold.  A score costs one max/compare; the
//     rare survivor is appended as a u64 ranking key (pcv_common.cuh) to the (CTA, query)
//     candidate buffer in global memory.  A buffer that could overflow is cut back to its
//     k best by the warp (exact; only adversarial inputs get here).
//   * thresholds come from a geometric pass schedule on the host side: pass p covers
//     tiles [T_p, r*T_p) with the k-th best similarity of everything before T_p as the
//     entry threshold, so a pass appends O(k) candidates per query per CTA; a radix-select
//     kernel folds the candidates into the running per-query top-k between passes (and resets
//     the counters it consumed) and emits the final ids/scores.  The B x N score matrix is
//     never written.  The kernels of one search are launched as a programmatic-dependent chain.
//
// Numerics: both operands are bf16 (products exact in fp32), fp32 accumulation inside
// the tensor core in an order the hardware does not specify -> compared with the
// float64 oracle on the same bf16 values under a stated tolerance, not bit for bit.
#include <cuda.h>
#include <cuda_runtime.h>
#include <math_constants.h>

#include <algorithm>
#include <atomic>
#include <cstdlib>
#include <cstring>
#include <utility>

#include "pcv_common.cuh"
#include "pcv_gemm_launch.cuh"
#include "pcv_synth.cuh"
#include "pcv_topk.cuh"

namespace pcv {

namespace {

// Two tile shapes share one kernel (template parameter SHAPE):
//   SHAPE_BF16   bf16 rows, dim <= 384 (K2, and K3's filter over the hi plane of split rows): 128 document
//                rows per tile, 8-slot document ring of 16 KB.
//   SHAPE_WIDE   bf16 rows, 384 < dim <= 768 (config 5): 12 K blocks per row, 96 document rows per
//                tile (N = 192 over a CTA pair), 13-slot document ring of 12 KB, 4-stage query ring.
// (A 64-row / 16-slot "streaming" shape for batches that use a document tile only once was measured on
// config 4 and lost to SHAPE_BF16, 18.3 ms against 16.8 ms per batch on one GPU: the pass is bound by the
// power cap, and N = 128 MMAs cost more shared-memory operand reads per flop than N = 256 ones.)
constexpr int SHAPE_BF16 = 0, SHAPE_WIDE = 2;
constexpr int G_BM = 128;       // queries per tile (UMMA M, TMEM lanes)
constexpr int G_BK = 64;        // bf16 elements per K block = one 128-byte swizzle row
constexpr int G_ACC = 4;        // TMEM accumulators (128 columns apart)
constexpr int G_ACC_COLS = 128;
constexpr int G_THREADS = 224;  // 7 warps
constexpr uint32_t G_PLANE_BYTES = G_BM * G_BK * 2;  // 16 KB: one 128-row K block
constexpr uint32_t G_SMEM_RINGS = 224 * 1024;        // document ring + query ring (split per shape, below)
constexpr int G_MAX_XSLOTS = 13;
constexpr int G_MAX_QSTAGES = 6;
constexpr uint32_t G_NBARS = 2 * G_MAX_XSLOTS + 2 * G_MAX_QSTAGES + 2 * G_ACC;
template <int SHAPE> struct GemmShape {
  // SHAPE_WIDE: a 768-d tile is 12 K blocks; 96 document rows (N = 192 over a CTA pair) is the most that fits next to a
  // 4-stage query ring — measured against 64 rows / 6 stages on config 5's shard (profiles/README.md, round 2).
  static constexpr int BN = SHAPE == SHAPE_BF16 ? 128 : 96;          // document rows per tile (UMMA N)
  static constexpr int MAX_KB = SHAPE == SHAPE_WIDE ? 12 : 6;        // K blocks per row
  static constexpr int XSLOTS = SHAPE == SHAPE_WIDE ? 13 : 8;        // current tile's K blocks + prefetch
  static constexpr int QSTAGES = SHAPE == SHAPE_WIDE ? 4 : 6;        // query ring depth (K blocks)
  static constexpr uint32_t QSTAGE_BYTES = G_PLANE_BYTES;
  static constexpr uint32_t XSLOT_BYTES = BN * G_BK * 2;
  static constexpr uint32_t SMEM_X = XSLOTS * XSLOT_BYTES;           // document ring region (slots are 1024-byte multiples)
  static constexpr uint32_t SMEM_Q = QSTAGES * QSTAGE_BYTES;
  static_assert(XSLOTS >= MAX_KB + 1, "document ring must hold one tile plus prefetch");
  static_assert(XSLOTS <= G_MAX_XSLOTS && QSTAGES <= G_MAX_QSTAGES, "barrier arrays");
  static_assert(XSLOT_BYTES % 1024 == 0 && SMEM_X + SMEM_Q <= G_SMEM_RINGS, "ring regions");
};
constexpr uint32_t G_SMEM_BYTES = G_SMEM_RINGS + G_NBARS * 8 + 16 + 1024;  // + alignment slack
static_assert(G_SMEM_BYTES <= 232448, "K2 shared memory budget");

struct GemmParams {
  CUtensorMap tmap_q;   // [m_tiles*128][dim_padded] bf16, box 64 x 128, SWIZZLE_128B
  CUtensorMap tmap_x;   // [n_rows][dim_padded] bf16, box 64 x BN
  uint32_t tile_rows;   // document rows per tile (BN)
  const uint2* ranges;
  const uint32_t* range_prefix;
  uint32_t n_ranges;
  uint32_t tile_begin, n_tiles;  // this pass covers document tiles [tile_begin, tile_begin + n_tiles)
  uint32_t m_tiles, n_queries, kb, k;
  uint32_t cand_cap;
  uint64_t* cand;        // [grid][m_tiles][128][cand_cap] ranking keys
  uint32_t* cand_cnt;    // [grid][m_tiles][128]
  uint32_t* thr_state;   // [grid][m_tiles][128] running threshold, order-preserving image of the f32
  const float* thr;      // [n_queries] entry thresholds (nullable: -inf)
  const uint32_t* lrank_of_row;
  const float* x_inv_norm;  // cosine: 1/|row| per stored row (padded by one tile); null = dot product
  uint32_t n_rows_total;    // rows of the matrix (TMA zero-fills beyond): coordinate of the dummy tile
  // bootstrap pass (pair kernel only): nothing is appended; every (document tile, query) writes the
  // maximum score of the tile to boot_max[tile][query] and the k-th largest of those becomes the first
  // threshold (k tiles' maxima are k distinct rows, so it never exceeds the true k-th best score)
  float* boot_max;          // [n_tiles][boot_qp], null outside the bootstrap pass
  uint32_t boot_qp;         // m_tiles * 128
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
  row0 = rg.x + (t - __ldg(p.range_prefix + r)) * p.tile_rows;
  nrows = min(p.tile_rows, rg.y - row0);
}

// KB_T: K blocks per row when known at compile time (6 = 384-d: the MMA issue loop unrolls and the
// query ring stage equals the K block, so every descriptor is a constant offset); 0 = runtime value.
template <int KB_T, int SHAPE>
__global__ void __launch_bounds__(G_THREADS, 1) gemm_topk_kernel(const __grid_constant__ GemmParams p) {
  using SH = GemmShape<SHAPE>;
  constexpr int G_BN = SH::BN;
  constexpr int G_QSTAGES = SH::QSTAGES;
  constexpr int G_XSLOTS = SH::XSLOTS;
  constexpr uint32_t G_XSLOT_BYTES = SH::XSLOT_BYTES;
  static_assert(KB_T == 0 || KB_T % G_QSTAGES == 0, "static K-block count must be a multiple of the query ring depth");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);
  uint8_t* smem_x = smem;
  uint8_t* smem_q = smem + SH::SMEM_X;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_SMEM_RINGS);
  uint64_t* bar_xfull = bars;                        // [G_XSLOTS]  TMA -> MMA
  uint64_t* bar_xempty = bar_xfull + G_MAX_XSLOTS;   // [G_XSLOTS]  MMA -> TMA (after the last query tile)
  uint64_t* bar_qfull = bar_xempty + G_MAX_XSLOTS;   // [G_QSTAGES] TMA -> MMA
  uint64_t* bar_qempty = bar_qfull + G_MAX_QSTAGES;  // [G_QSTAGES] MMA -> TMA
  uint64_t* bar_tfull = bar_qempty + G_MAX_QSTAGES;  // [G_ACC]     MMA -> epilogue
  uint64_t* bar_tempty = bar_tfull + G_ACC;          // [G_ACC]     epilogue -> MMA
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + G_NBARS);

  // warp index through a shuffle: the compiler then knows every role branch is warp-uniform and
  // keeps descriptors / barrier addresses in uniform registers (no per-lane waterfall around UTCHMMA)
  const int warp = __shfl_sync(PCV_FULL_MASK, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta = blockIdx.x;
  // this CTA's contiguous share of the pass's document tiles
  const uint32_t g0 = (uint32_t)((uint64_t)cta * p.n_tiles / gridDim.x);
  const uint32_t g1 = (uint32_t)((uint64_t)(cta + 1) * p.n_tiles / gridDim.x);
  const uint32_t t0 = 0, t1 = g1 - g0;
  auto tile_of = [&](uint32_t i, uint32_t& row0, uint32_t& nrows) { gemm_tile_rows(p, g0 + i + p.tile_begin, row0, nrows); };
  const uint32_t m_tiles = p.m_tiles;
  const uint32_t KB = KB_T ? (uint32_t)KB_T : p.kb;

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < G_XSLOTS; ++s) {
      mbar_init(smem_u32(bar_xfull + s), 1);
      mbar_init(smem_u32(bar_xempty + s), 1);
    }
    for (int s = 0; s < G_QSTAGES; ++s) {
      mbar_init(smem_u32(bar_qfull + s), 1);
      mbar_init(smem_u32(bar_qempty + s), 1);
    }
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
  pdl_launch_dependents();
  pdl_wait();  // thresholds, counters and queries come from the kernels before this one

  if (warp == 0) {
    // ===================== query producer =====================
    if (elect_one_sync()) {
      tma_prefetch_desc(&p.tmap_q);
    }
    const uint64_t pol = l2_policy_evict_last();
    const uint32_t q_base = smem_u32(smem_q), full0 = smem_u32(bar_qfull), empty0 = smem_u32(bar_qempty);
    uint32_t stage = 0, phase = 0;
    bool ready = false;  // result of the probe issued one step earlier
    for (uint32_t t = t0; t < t1; ++t)
      for (uint32_t m = 0; m < m_tiles; ++m)
        for (uint32_t kb = 0; kb < KB; ++kb) {
          if (!ready) mbar_wait_bounded(empty0 + stage * 8, phase ^ 1u);
          uint32_t nstage = stage + 1, nphase = phase;
          if (nstage == G_QSTAGES) { nstage = 0; nphase ^= 1u; }
          ready = mbar_test(empty0 + nstage * 8, nphase ^ 1u);  // probe the next stage while this one is issued
          if (elect_one_sync()) {
            mbar_arrive_expect_tx(full0 + stage * 8, SH::QSTAGE_BYTES);
            tma_load_2d(q_base + stage * SH::QSTAGE_BYTES, &p.tmap_q, full0 + stage * 8, (int32_t)(kb * G_BK),
                        (int32_t)(m * G_BM), pol);
          }
          __syncwarp();
          stage = nstage;
          phase = nphase;
        }
  } else if (warp == 6) {
    // ===================== document producer =====================
    if (elect_one_sync()) {
      tma_prefetch_desc(&p.tmap_x);
    }
    const uint64_t pol = l2_policy_evict_first();
    const uint32_t x_base = smem_u32(smem_x), full0 = smem_u32(bar_xfull), empty0 = smem_u32(bar_xempty);
    uint32_t slot = 0, phase = 0;
    bool ready = false;
    for (uint32_t t = t0; t < t1; ++t) {
      uint32_t row0, nrows;
      tile_of(t, row0, nrows);
      for (uint32_t kb = 0; kb < KB; ++kb) {
        if (!ready) mbar_wait_bounded(empty0 + slot * 8, phase ^ 1u);
        uint32_t nslot = slot + 1, nphase = phase;
        if (nslot == G_XSLOTS) { nslot = 0; nphase ^= 1u; }
        ready = mbar_test(empty0 + nslot * 8, nphase ^ 1u);
        if (elect_one_sync()) {
          mbar_arrive_expect_tx(full0 + slot * 8, G_XSLOT_BYTES);
          tma_load_2d(x_base + slot * G_XSLOT_BYTES, &p.tmap_x, full0 + slot * 8, (int32_t)(kb * G_BK), (int32_t)row0,
                      pol);
        }
        __syncwarp();
        slot = nslot;
        phase = nphase;
      }
    }
  } else if (warp == 1) {
    // ===================== MMA issuer =====================
    constexpr uint32_t idesc = umma_idesc_bf16_f32(G_BM, G_BN);
    const uint64_t q_desc0 = umma_desc_k_sw128(smem_u32(smem_q));
    const uint64_t x_desc0 = umma_desc_k_sw128(smem_u32(smem_x));
    const uint32_t qfull0 = smem_u32(bar_qfull), qempty0 = smem_u32(bar_qempty);
    const uint32_t xfull0 = smem_u32(bar_xfull), xempty0 = smem_u32(bar_xempty);
    const uint32_t tfull0 = smem_u32(bar_tfull), tempty0 = smem_u32(bar_tempty);
    uint32_t qstage = 0, qphase = 0, acc = 0, acc_par = 0;
    uint32_t xslot_tile = 0, xphase_tile = 0;  // ring position of the current tile's K block 0
    for (uint32_t t = t0; t < t1; ++t) {
      for (uint32_t m = 0; m < m_tiles; ++m) {
        mbar_wait_bounded(tempty0 + acc * 8, acc_par ^ 1u);
        tc_fence_after();
        const uint32_t d_tmem = tmem_base + acc * G_ACC_COLS;
        const bool last_m = (m + 1 == m_tiles);
        uint32_t xslot = xslot_tile, xphase = xphase_tile;
        // `ready`: the probe of this K step's barriers, issued while the previous step's MMAs were
        // going out (a barrier round trip costs ~100 cycles; a K step is only 128-384 cycles of MMA)
        bool ready = false;
#pragma unroll
        for (uint32_t kb = 0; kb < (KB_T ? (uint32_t)KB_T : KB); ++kb) {
          const uint32_t qs = KB_T ? kb % G_QSTAGES : qstage;  // compile-time stage when K blocks are static
          const uint32_t qph = KB_T ? (qphase ^ ((kb / G_QSTAGES) & 1u)) : qphase;
          if (!ready) {
            if (m == 0) mbar_wait_bounded(xfull0 + xslot * 8, xphase);
            mbar_wait_bounded(qfull0 + qs * 8, qph);
          }
          tc_fence_after();
          // next step's ring positions, probed now
          uint32_t nqs, nqph, nxslot = xslot + 1, nxphase = xphase;
          if (KB_T) {
            nqs = (kb + 1) % G_QSTAGES;
            nqph = qphase ^ (((kb + 1) / G_QSTAGES) & 1u);
          } else {
            nqs = qstage + 1;
            nqph = qphase;
            if (nqs == (uint32_t)G_QSTAGES) { nqs = 0; nqph ^= 1u; }
          }
          if (nxslot == (uint32_t)G_XSLOTS) { nxslot = 0; nxphase ^= 1u; }
          if (kb + 1 < (KB_T ? (uint32_t)KB_T : KB)) {
            ready = mbar_test(qfull0 + nqs * 8, nqph);
            if (m == 0) ready = mbar_test(xfull0 + nxslot * 8, nxphase) && ready;
          } else {
            ready = false;
          }
          if (elect_one_sync()) {
            // descriptor start-address field counts 16-byte units: advance by adding to the low word
            const uint64_t a_desc = q_desc0 + (uint64_t)((qs * SH::QSTAGE_BYTES) >> 4);
            const uint64_t b_desc = x_desc0 + (uint64_t)((xslot * G_XSLOT_BYTES) >> 4);
#pragma unroll
            for (uint32_t j = 0; j < G_BK / 16; ++j)
              tc_mma_bf16(d_tmem, a_desc + j * 2, b_desc + j * 2, idesc, (kb | j) != 0u);
            tc_commit(qempty0 + qs * 8);                 // query stage is free once these retire
            if (last_m) tc_commit(xempty0 + xslot * 8);  // last query tile: hand the document slot back
          }
          __syncwarp();
          if (!KB_T) { qstage = nqs; qphase = nqph; }
          xslot = nxslot;
          xphase = nxphase;
        }
        if (KB_T) qphase ^= (uint32_t)((KB_T / G_QSTAGES) & 1);  // trips round the query ring per item
        if (elect_one_sync()) tc_commit(tfull0 + acc * 8);
        __syncwarp();
        if (++acc == G_ACC) { acc = 0; acc_par ^= 1u; }
        if (last_m) { xslot_tile = xslot; xphase_tile = xphase; }
      }
    }
  } else {
    // ===================== epilogue: fused top-k filter =====================
    const int quarter = warp & 3;  // TMEM lanes this warp may read
    const int row = quarter * 32 + lane;
    const int k = (int)p.k;
    const size_t slot0 = (size_t)cta * m_tiles * G_BM + (size_t)row;
    // running thresholds of this CTA's (query tile, row) pairs start at the pass's entry threshold
    for (uint32_t m = 0; m < m_tiles; ++m) {
      const uint32_t q = m * G_BM + (uint32_t)row;
      const float th = (q < p.n_queries) ? (p.thr ? __ldg(p.thr + q) : -CUDART_INF_F) : CUDART_INF_F;
      p.thr_state[slot0 + (size_t)m * G_BM] = f32_to_ordered(th);
    }
    uint32_t acc = 0, acc_par = 0;
    const uint32_t tfull0 = smem_u32(bar_tfull), tempty0 = smem_u32(bar_tempty);
    const float* __restrict__ xinv = p.x_inv_norm;
    // per-(query tile, row) state lives in global memory (L2); the NEXT item's state is fetched
    // while the current item is processed, so its latency never sits on the epilogue's critical path
    uint32_t cnt_next = p.cand_cnt[slot0];
    uint32_t thr_next = p.thr_state[slot0];
    for (uint32_t t = t0; t < t1; ++t) {
      uint32_t row0, nrows;
      tile_of(t, row0, nrows);
      for (uint32_t m = 0; m < m_tiles; ++m) {
        const size_t slot = slot0 + (size_t)m * G_BM;
        uint32_t cnt = cnt_next;
        float thr = ordered_to_f32(thr_next);
        if (m_tiles > 1) {
          const size_t nslot = slot0 + (size_t)((m + 1 == m_tiles) ? 0u : m + 1) * G_BM;
          cnt_next = p.cand_cnt[nslot];
          thr_next = p.thr_state[nslot];
        }
        uint64_t* buf = p.cand + slot * p.cand_cap;
        const uint32_t cnt_in = cnt;
        mbar_wait_bounded(tfull0 + acc * 8, acc_par);
        tc_fence_after();
        const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * G_ACC_COLS;
        // fast path: the whole 128-column accumulator row into registers, one max tree
        // (independent chains), one compare per row
        float rmx;
        uint32_t gm = 0;  // bit g: columns [8g, 8g+8) hold a score >= threshold
        {
          constexpr int NCH = G_BN / 32;  // 32-column chunks per accumulator row
          uint32_t v[NCH][32];
#pragma unroll
          for (int c = 0; c < NCH; ++c) tc_ld_32x32b_x32(taddr + 32 * c, v[c]);
          tc_wait_ld();
          if (xinv) {
            // cosine (crates/perceive-core/lib.rs:67-77): rank by dot / |row|; the query's own norm is a
            // positive constant of this thread's row and is applied when the result is emitted
#pragma unroll
            for (int c = 0; c < NCH; ++c)
#pragma unroll
              for (int e = 0; e < 32; ++e)
                v[c][e] = __float_as_uint(__uint_as_float(v[c][e]) * __ldg(xinv + row0 + 32 * c + e));
          }
          float gmx[4 * NCH];
#pragma unroll
          for (int c = 0; c < NCH; ++c)
#pragma unroll
            for (int g = 0; g < 4; ++g) {
              float mx = __uint_as_float(v[c][8 * g]);
#pragma unroll
              for (int e = 1; e < 8; ++e) mx = fmaxf(mx, __uint_as_float(v[c][8 * g + e]));
              gmx[4 * c + g] = mx;
            }
          rmx = gmx[0];
#pragma unroll
          for (int i = 1; i < 4 * NCH; ++i) rmx = fmaxf(rmx, gmx[i]);
          if (__any_sync(PCV_FULL_MASK, rmx >= thr)) {
#pragma unroll
            for (int i = 0; i < 4 * NCH; ++i) gm |= (gmx[i] >= thr ? 1u : 0u) << i;
          }
        }
        // slow path, kept SMALL on purpose (a fully unrolled version thrashes the instruction
        // cache): re-read only the 8-column groups some lane needs, in a loop
        uint32_t wm = __reduce_or_sync(PCV_FULL_MASK, gm);
        while (wm) {
          const uint32_t g = (uint32_t)__ffs(wm) - 1u;
          wm &= wm - 1u;
          uint32_t w8[8];
          tc_ld_32x32b_x8(taddr + 8u * g, w8);
          tc_wait_ld();
          if ((gm >> g) & 1u) {
            // mask of this lane's survivors among the 8 columns (branch-free), then one trip per set
            // bit: the per-element branches of a straight unrolled version dominated the dense passes
            float sc8[8];
            uint32_t em = 0;
#pragma unroll
            for (int e = 0; e < 8; ++e) {
              const uint32_t col = 8u * g + (uint32_t)e;
              float sc = __uint_as_float(w8[e]);
              if (xinv) sc *= __ldg(xinv + row0 + col);
              sc8[e] = sc;
              em |= ((sc >= thr && col < nrows) ? 1u : 0u) << e;
            }
            while (em) {
              const uint32_t e = (uint32_t)__ffs(em) - 1u;
              em &= em - 1u;
              float sc = sc8[0];
#pragma unroll
              for (int j = 1; j < 8; ++j) sc = (e == (uint32_t)j) ? sc8[j] : sc;  // select chain, no local memory
              const uint32_t r = row0 + 8u * g + e;
              const uint32_t lr = p.lrank_of_row ? __ldg(p.lrank_of_row + r) : r;
              buf[cnt++] = make_key(sc, lr);
            }
          }
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(tempty0 + acc * 8);
        if (++acc == G_ACC) { acc = 0; acc_par ^= 1u; }

        // a buffer that could overflow during its next tile is cut back to its k best
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
            p.thr_state[slot] = f32_to_ordered(thr);
          }
        }
        if (cnt != cnt_in) p.cand_cnt[slot] = cnt;
        if (m_tiles == 1) { cnt_next = cnt; thr_next = f32_to_ordered(thr); }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 0) {
    __syncwarp();
    tc_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace pcv
#include "pcv_gemm_pair.cuh"
namespace pcv {
namespace {

// ---------------------------------------------------------------------------
// select: fold one pass's candidate buffers into the running per-query top-k
// (sorted keys) and, on the last pass, emit ids / scores.  One CTA per query:
// candidates are gathered into shared memory, the k-th largest key is found by
// an 8-round byte-wise radix select, the k survivors are ranked by counting.
// ---------------------------------------------------------------------------
struct SelectParams {
  const uint64_t* cand;
  uint32_t* cand_cnt;    // consumed here: every counter read is reset to 0 for the next pass (no memset between passes)
  uint32_t grid_gemm, cand_cap, m_tiles, k;
  uint32_t smem_keys;  // capacity of the shared-memory key array
  uint64_t* topk;      // [n_queries][k] running result (in/out), sorted descending, 0 = empty
  int has_prev;
  float* thr;          // [n_queries] out: k-th similarity so far (or -inf)
  int emit;
  const float* q_scale;  // cosine: 1/|query| applied to the emitted similarity (null = 1)
  int cosine;            // reported score = similarity (cosine) or the reference distance
  uint32_t emit_mode, dim;
  const uint32_t* row_of_lrank;
  const int64_t* ids;
  int64_t id_base;
  int64_t* out_ids;
  float* out_scores;
  float* out_sims;
  uint32_t* out_counts;
};

constexpr int SEL_THREADS = 256;
constexpr int SEL_MAX_CTAS = 256;  // gemm grid never exceeds the SM count (148)

__global__ void __launch_bounds__(SEL_THREADS) gemm_select_kernel(const SelectParams p) {
  extern __shared__ uint64_t sel_keys[];  // [smem_keys]
  __shared__ uint32_t s_off[SEL_MAX_CTAS + 1];
  __shared__ uint32_t s_hist[256];
  __shared__ uint64_t s_sel[128];
  __shared__ uint64_t s_out[128];
  __shared__ uint64_t s_prefix;
  __shared__ uint32_t s_need, s_nsel, s_live, s_exact;
  const uint32_t tid = threadIdx.x;
  const uint32_t q = blockIdx.x;
  const uint32_t m = q / G_BM, row = q % G_BM;
  const uint32_t k = p.k;
  const uint32_t G = p.grid_gemm;
  pdl_launch_dependents();
  pdl_wait();  // the pass that filled the candidate buffers

  // --- gather ----------------------------------------------------------------
  // s_off[c] .. s_off[c+1]: where CTA c's candidates land; slots 0..k-1 hold the running result.
  // Block-wide exclusive scan of the G (<= 256) counts: one load per thread, two barriers.
  __shared__ uint32_t s_wsum[SEL_THREADS / 32];
  {
    uint32_t n_mine = 0u;
    if (tid < G) {
      uint32_t* cnt = p.cand_cnt + ((size_t)tid * p.m_tiles + m) * G_BM + row;
      n_mine = *cnt;
      if (n_mine) *cnt = 0u;  // this query's counters are only ever read here
    }
    uint32_t incl = n_mine;
#pragma unroll
    for (int off = 1; off < 32; off <<= 1) {
      const uint32_t o = __shfl_up_sync(PCV_FULL_MASK, incl, off);
      if ((int)(tid & 31u) >= off) incl += o;
    }
    if ((tid & 31u) == 31u) s_wsum[tid >> 5] = incl;
    __syncthreads();
    uint32_t base = p.has_prev ? k : 0u;
    for (uint32_t w = 0; w < (tid >> 5); ++w) base += s_wsum[w];
    if (tid == 0) s_off[0] = p.has_prev ? k : 0u;
    if (tid < G) s_off[tid + 1] = base + incl;  // end offset of CTA tid
  }
  __syncthreads();
  const uint32_t n_prev = s_off[0];
  const uint32_t n_total = s_off[G];
  // Expected sizes fit shared memory (a pass appends O(k) keys per query); an adversarial pass that
  // does not is selected straight from global memory — same algorithm, slower reads.
  const bool in_smem = n_total <= p.smem_keys;
  const uint64_t* prev = p.topk + (size_t)q * k;
  auto key_at = [&](uint32_t i) -> uint64_t {
    if (in_smem) return sel_keys[i];
    if (i < n_prev) return prev[i];
    uint32_t lo = 0, hi = G;  // s_off[lo] <= i < s_off[hi]
    while (hi - lo > 1) {
      const uint32_t mid = (lo + hi) >> 1;
      if (s_off[mid] <= i) lo = mid; else hi = mid;
    }
    const uint64_t* src = p.cand + (((size_t)lo * p.m_tiles + m) * G_BM + row) * p.cand_cap;
    return __ldcg(reinterpret_cast<const unsigned long long*>(src) + (i - s_off[lo]));
  };
  if (in_smem) {
    for (uint32_t i = tid; i < n_prev; i += SEL_THREADS) sel_keys[i] = prev[i];
    const uint32_t warp = tid >> 5, lane = tid & 31;
    // four buffers per warp step: their first loads are issued together (a buffer rarely holds more
    // than 32 keys), so a warp pays the memory round trip G/32 times instead of G/8 times
    constexpr uint32_t GB = 4;
    for (uint32_t c0 = warp * GB; c0 < G; c0 += (SEL_THREADS / 32) * GB) {
      uint64_t v[GB];
      uint32_t b[GB], e[GB];
      const unsigned long long* src[GB];
#pragma unroll
      for (uint32_t j = 0; j < GB; ++j) {
        const uint32_t c = c0 + j;
        b[j] = (c < G) ? s_off[c] : 0u;
        e[j] = (c < G) ? s_off[c + 1] : 0u;
        src[j] = reinterpret_cast<const unsigned long long*>(p.cand + (((size_t)c * p.m_tiles + m) * G_BM + row) * p.cand_cap);
        v[j] = (b[j] + lane < e[j]) ? __ldcg(src[j] + lane) : 0ull;
      }
#pragma unroll
      for (uint32_t j = 0; j < GB; ++j) {
        if (b[j] + lane < e[j]) sel_keys[b[j] + lane] = v[j];
        for (uint32_t i = b[j] + lane + 32; i < e[j]; i += 32) sel_keys[i] = __ldcg(src[j] + (i - b[j]));
      }
    }
  }
  if (tid == 0) { s_nsel = 0; s_prefix = 0ull; s_live = 0; s_exact = 0; }
  __syncthreads();

  // --- radix select: the k-th largest key (keys are distinct; 0 = empty) ---------
  uint32_t live = 0;
  for (uint32_t i = tid; i < n_total; i += SEL_THREADS) live += key_at(i) != 0ull;
  if (live) atomicAdd(&s_live, live);
  __syncthreads();
  const uint32_t n_live = s_live;
  const uint32_t want = min(k, n_live);
  uint64_t kth = 1ull;  // n_live <= k: every non-empty key survives
  if (n_live > k) {
    if (tid == 0) s_need = k;
    uint64_t mask = 0ull;
    for (int shift = 56; shift >= 0; shift -= 8) {
      s_hist[tid] = 0;
      __syncthreads();
      const uint64_t prefix = s_prefix;
      for (uint32_t i = tid; i < n_total; i += SEL_THREADS) {
        const uint64_t key = key_at(i);
        if (key != 0ull && (key & mask) == prefix) atomicAdd(&s_hist[(uint32_t)(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid < 32) {
        // warp 0: suffix sums over the 256 bins, 8 bins per lane, top bin first
        uint32_t loc[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { loc[j] = s_hist[255 - (tid * 8 + j)]; sum += loc[j]; }
        uint32_t incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t o = __shfl_up_sync(PCV_FULL_MASK, incl, off);
          if ((int)tid >= off) incl += o;
        }
        uint32_t before = incl - sum;  // keys in bins above this lane's bins
        const uint32_t need = s_need;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (before < need && before + loc[j] >= need) {
            s_prefix = prefix | ((uint64_t)(255 - (tid * 8 + j)) << shift);
            s_need = need - before;
            s_exact = (before + loc[j] == need) ? 1u : 0u;  // the bin holds exactly what is still needed
          }
          before += loc[j];
        }
      }
      mask |= 0xffull << shift;
      __syncthreads();
      if (s_exact) break;  // every key >= prefix (low bits zero) is a survivor: later rounds change nothing
    }
    kth = s_prefix;
  }
  // --- collect the survivors, rank them by counting -----------------------------
  for (uint32_t i = tid; i < n_total; i += SEL_THREADS) {
    const uint64_t key = key_at(i);
    if (key != 0ull && key >= kth) {
      const uint32_t pos = atomicAdd(&s_nsel, 1u);
      if (pos < 128u) s_sel[pos] = key;
    }
  }
  __syncthreads();
  const uint32_t nsel = min(s_nsel, want);
  if (tid < 128) s_out[tid] = 0ull;
  __syncthreads();
  if (tid < nsel) {
    const uint64_t mine = s_sel[tid];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < nsel; ++j) rank += s_sel[j] > mine;
    s_out[rank] = mine;
  }
  __syncthreads();
  if (tid < k) p.topk[(size_t)q * k + tid] = s_out[tid];
  if (tid == 0) p.thr[q] = (nsel == k) ? key_sim(s_out[k - 1]) : -CUDART_INF_F;
  if (!p.emit) return;
  if (tid < k) {
    const uint64_t key = s_out[tid];
    const bool is_live = key != 0ull;
    float sim = -CUDART_INF_F;
    int64_t id = (p.emit_mode == 1) ? INT64_MAX : (int64_t)-1;
    if (is_live) {
      sim = key_sim(key);
      if (p.q_scale) sim *= p.q_scale[q];
      const uint32_t lr = key_lrank(key);
      const uint32_t r = p.row_of_lrank ? p.row_of_lrank[lr] : lr;
      id = p.ids ? p.ids[r] : p.id_base + (int64_t)r;
    }
    const size_t o = (size_t)q * k + tid;
    p.out_ids[o] = id;
    if (p.out_sims) p.out_sims[o] = sim;
    if (p.out_scores) p.out_scores[o] = is_live ? (p.cosine ? sim : ref_distance(sim, p.dim)) : CUDART_INF_F;
  }
  if (p.out_counts && tid == 0) p.out_counts[q] = nsel;
}

// bootstrap threshold: k-th largest of the n_tiles tile maxima of one query (one CTA per query).
// Fewer than k finite maxima leave the threshold at -inf.
__global__ void __launch_bounds__(256) gemm_boot_threshold_kernel(const float* __restrict__ boot_max, uint32_t n_tiles,
                                                                   uint32_t qp, uint32_t k, float* __restrict__ thr) {
  extern __shared__ uint64_t boot_keys[];  // [n_tiles] keys, [k] survivors, [k] sorted
  __shared__ BlockSelectScratch sc;
  const uint32_t q = blockIdx.x;
  pdl_launch_dependents();
  pdl_wait();  // the bootstrap pass wrote boot_max
  for (uint32_t i = threadIdx.x; i < n_tiles; i += 256) {
    const float mx = __ldcg(boot_max + (size_t)i * qp + q);
    boot_keys[i] = (mx > -CUDART_INF_F) ? make_key(mx, i) : 0ull;  // a partial tile wrote -inf: no vote
  }
  __syncthreads();
  uint64_t* sel = boot_keys + n_tiles;
  uint64_t* out = sel + k;
  const uint32_t nsel = block_select_sorted(boot_keys, n_tiles, k, sel, out, sc);
  if (threadIdx.x == 0) thr[q] = (nsel == k) ? key_sim(out[k - 1]) : -CUDART_INF_F;
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

// cosine: 1/|q| of the bf16-rounded query (fp32 accumulation), one warp per query
__global__ void query_inv_norms_kernel(const uint16_t* __restrict__ q_bf16, float* __restrict__ out, uint32_t n_queries,
                                       uint32_t dim_padded) {
  const int lane = threadIdx.x & 31;
  const uint32_t q = (blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  if (q >= n_queries) return;
  float acc = 0.0f;
  for (uint32_t c = lane; c < dim_padded; c += 32) {
    const float x = bf16_to_f32(q_bf16[(size_t)q * dim_padded + c]);
    acc = fmaf(x, x, acc);
  }
#pragma unroll
  for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(PCV_FULL_MASK, acc, off);
  if (lane == 0) out[q] = 1.0f / sqrtf(acc);
}

// cosine: 1/|row| of every stored bf16 row, computed on the device from the stored values
// ("norms computed in-kernel", BASELINE config 5); one warp per row, fp32 accumulation
__global__ void row_inv_norms_kernel(const uint16_t* __restrict__ rows, float* __restrict__ out, uint64_t n_rows,
                                     uint32_t dim_padded, uint64_t n_out) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < n_out; r += nwarps) {
    float acc = 0.0f;
    if (r < n_rows) {
      const uint4* src = reinterpret_cast<const uint4*>(rows + r * dim_padded);
      for (uint32_t c = lane; c < dim_padded / 8; c += 32) {
        const uint4 u = src[c];
        const uint32_t w[4] = {u.x, u.y, u.z, u.w};
#pragma unroll
        for (int i = 0; i < 4; ++i) {
          const float a = __uint_as_float(w[i] << 16), b = __uint_as_float(w[i] & 0xffff0000u);
          acc = fmaf(a, a, acc);
          acc = fmaf(b, b, acc);
        }
      }
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) acc += __shfl_xor_sync(PCV_FULL_MASK, acc, off);
    }
    if (lane == 0) out[r] = (r < n_rows) ? 1.0f / sqrtf(acc) : 0.0f;  // padding rows score 0
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

// [rows][dim_padded] bf16 with a row pitch in bytes, box = 64 elements x box_rows, 128-byte swizzle, zero fill
bool make_tmap(CUtensorMap* out, const void* base, uint64_t rows, uint32_t dim_padded, uint64_t pitch_bytes,
               uint32_t box_rows) {
  EncodeTiledFn fn = encode_tiled_fn();
  if (!fn) return false;
  const cuuint64_t dims[2] = {dim_padded, rows};
  const cuuint64_t strides[1] = {(cuuint64_t)pitch_bytes};
  const cuuint32_t box[2] = {(cuuint32_t)G_BK, (cuuint32_t)box_rows};
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

// Launch with the programmatic-stream-serialisation attribute (see pdl_wait in pcv_common.cuh): the kernels of one
// search form a chain pass -> select -> pass -> ...; each may begin its prologue under its predecessor's tail.
template <typename... KArgs, typename... Args>
cudaError_t launch_chain(void (*kern)(KArgs...), uint32_t grid, uint32_t block, size_t smem, cudaStream_t st, bool pdl,
                         Args&&... args) {
  cudaLaunchConfig_t cfg = {};
  cfg.gridDim = dim3(grid);
  cfg.blockDim = dim3(block);
  cfg.dynamicSmemBytes = smem;
  cfg.stream = st;
  cudaLaunchAttribute at[1];
  at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  at[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = at;
  cfg.numAttrs = pdl ? 1u : 0u;
  return cudaLaunchKernelEx(&cfg, kern, std::forward<Args>(args)...);
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
  if (d_qinv) cudaFree(d_qinv);
  if (d_boot) cudaFree(d_boot);
  d_boot = nullptr;
  boot_cap = 0;
  d_qinv = nullptr;
  qinv_cap = 0;
  d_q_bf16 = nullptr;
  d_cand = nullptr;
  d_cand_cnt = nullptr;
  d_topk = nullptr;
  d_thr = nullptr;
  q_cap = cand_cap = cnt_cap = topk_cap = thr_cap = 0;
}

bool gemm_path_applicable(bool cosine, uint32_t dim_padded, uint32_t n_queries, uint32_t k, uint64_t selected_rows,
                          uint64_t n_rows) {
  (void)cosine;
  if (dim_padded < (uint32_t)G_BK || dim_padded > GemmShape<SHAPE_WIDE>::MAX_KB * G_BK) return false;
  if (k > 128) return false;
  if (n_rows >= 0x7fffff00ull) return false;  // TMA coordinates are int32
  // The scan (K1) costs one HBM pass per 4 queries and ~no fixed overhead; the tensor path costs
  // ~0.5 ms of threshold passes whatever the size.  Small batches and small corpora stay on K1; from 16
  // queries — or 4 queries over >= 2M rows — the tensor path wins (measured on B200, 4 queries, bf16 x
  // 384: K1 1.09 ms at 2M rows, i.e. ~5.4 ms at 10M; the tensor path 1.36 ms at 10M rows).
  const uint32_t min_batch = env_u32("PCV_GEMM_MIN_BATCH", 16);
  const bool big = selected_rows >= (2u << 20) && n_queries >= std::min<uint32_t>(min_batch, 4u);
  if (n_queries < min_batch && !big) return false;
  if (selected_rows < env_u32("PCV_GEMM_MIN_ROWS", 4096)) return false;
  return encode_tiled_fn() != nullptr;
}

uint32_t gemm_tile_rows(uint32_t dim_padded) { return dim_padded > 384 ? 96u : 128u; }

cudaError_t gemm_row_inv_norms(const uint8_t* rows, uint64_t n_rows, uint32_t dim_padded, float* out, uint64_t n_out,
                               int sm_count, cudaStream_t stream) {
  const int blocks = (int)std::min<uint64_t>((n_out + 7) / 8, (uint64_t)sm_count * 16);
  row_inv_norms_kernel<<<std::max(blocks, 1), 256, 0, stream>>>(reinterpret_cast<const uint16_t*>(rows), out, n_rows,
                                                               dim_padded, n_out);
  return cudaGetLastError();
}

const char* gemm_search(GemmWorkspace& ws, const GemmCall& c, uint32_t* launches, cudaError_t* err) {
  *err = cudaSuccess;
  uint32_t nl = 0;
  const uint32_t m_tiles = (c.n_queries + G_BM - 1) / G_BM;
  const uint32_t rows_padded = (m_tiles + 1) / 2 * 2 * G_BM;  // pair mode walks query tiles two at a time
  const uint32_t kb = (c.dim_padded + G_BK - 1) / G_BK;
  const uint32_t k = c.k;
  // candidate buffer per (CTA, query): must keep a whole tile of head-room above k
  const int shape = c.dim_padded > 384 ? SHAPE_WIDE : SHAPE_BF16;
  // pair mode (2-CTA MMA) appends up to 2*BN keys per item: 256 for the bf16 shape
  const bool pair_ok = !env_u32("PCV_GEMM_NO_PAIR", 0);
  // (a buffer must hold k <= 128 kept keys plus one item's worth of appends: 2 * BN in pair mode)
  const uint32_t cand_cap = std::max<uint32_t>(pair_ok ? (shape == SHAPE_BF16 ? 512u : 384u) : 256u, env_u32("PCV_GEMM_CAND_CAP", 256));
  const uint32_t tile_rows = gemm_tile_rows(c.dim_padded);
  // pass schedule: tiles seen grow by `ratio_early` per pass until `dense_tiles`, then one last pass
  const uint32_t ratio = std::max<uint32_t>(2u, env_u32("PCV_GEMM_PASS_RATIO", 4));
  const uint32_t dense_tiles = env_u32("PCV_GEMM_DENSE_TILES", 8192);
  // PCV_GEMM_MAX_CTAS: test knob — fewer CTAs means more tiles per candidate buffer (forces the overflow path)
  const uint32_t sms = std::max<uint32_t>(1u, std::min<uint32_t>(std::min<uint32_t>((uint32_t)c.sm_count, (uint32_t)SEL_MAX_CTAS), env_u32("PCV_GEMM_MAX_CTAS", 1u << 20)));

#define GCHK(call, what)            \
  do {                              \
    *err = (call);                  \
    if (*err != cudaSuccess) return what; \
  } while (0)

  // shared memory of the select kernel: pass 0 hands it first*128 keys per query, later passes O(k)
  const size_t sel_smem = 6144 * sizeof(uint64_t);
  static std::atomic<bool> attr_done[64];  // handles on different threads may search at the same time: the attribute
                                            // calls are idempotent, the flag must not be a data race
  int dev = 0;
  cudaGetDevice(&dev);
  if (!attr_done[dev & 63].load(std::memory_order_acquire)) {
#define PCV_SET_SMEM(KBT, SHP)                                                                                  \
  GCHK(cudaFuncSetAttribute(gemm_topk_kernel<KBT, SHP>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                            (int)G_SMEM_BYTES),                                                                  \
       "cudaFuncSetAttribute(gemm_topk_kernel)")
    PCV_SET_SMEM(0, SHAPE_BF16);
    PCV_SET_SMEM(6, SHAPE_BF16);
    PCV_SET_SMEM(0, SHAPE_WIDE);
    PCV_SET_SMEM(12, SHAPE_WIDE);
#undef PCV_SET_SMEM
#define PCV_SET_SMEM(KBT, SHP)                                                                                  \
  GCHK(cudaFuncSetAttribute(gemm_topk_pair_kernel<KBT, SHP>, cudaFuncAttributeMaxDynamicSharedMemorySize,       \
                            (int)GP_SMEM_BYTES),                                                                 \
       "cudaFuncSetAttribute(gemm_topk_pair_kernel)")
    PCV_SET_SMEM(0, SHAPE_BF16);
    PCV_SET_SMEM(6, SHAPE_BF16);
    PCV_SET_SMEM(0, SHAPE_WIDE);
    PCV_SET_SMEM(12, SHAPE_WIDE);
#undef PCV_SET_SMEM
    GCHK(cudaFuncSetAttribute(gemm_select_kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, 200 * 1024),
         "cudaFuncSetAttribute(gemm_select_kernel)");
    attr_done[dev & 63].store(true, std::memory_order_release);
  }

  // queries -> bf16 (round to nearest even), padded to whole tiles
  const size_t q_plane = (size_t)rows_padded * c.dim_padded * 2;
  GCHK(reserve(ws.d_q_bf16, ws.q_cap, q_plane), "query buffer allocation");
  {
    const uint32_t blocks = std::min<uint32_t>(1024u, (rows_padded * c.dim_padded + 255) / 256);
    queries_to_bf16_kernel<<<blocks, 256, 0, c.stream>>>(c.queries, (uint16_t*)ws.d_q_bf16, c.n_queries, rows_padded,
                                                        c.dim_padded);
  }
  GCHK(cudaGetLastError(), "query conversion kernel launch");
  ++nl;
  if (c.cosine) {
    GCHK(reserve(ws.d_qinv, ws.qinv_cap, (size_t)c.n_queries), "query norm buffer allocation");
    query_inv_norms_kernel<<<(c.n_queries + 7) / 8, 256, 0, c.stream>>>((const uint16_t*)ws.d_q_bf16, ws.d_qinv,
                                                                        c.n_queries, c.dim_padded);
    GCHK(cudaGetLastError(), "query_inv_norms_kernel launch");
    ++nl;
  }
  GCHK(reserve(ws.d_topk, ws.topk_cap, (size_t)c.n_queries * k), "top-k buffer allocation");
  GCHK(reserve(ws.d_thr, ws.thr_cap, (size_t)c.n_queries), "threshold buffer allocation");
  const size_t n_slots = (size_t)sms * m_tiles * G_BM;
  GCHK(reserve(ws.d_cand, ws.cand_cap, n_slots * cand_cap), "candidate buffer allocation");
  {
    const uint32_t* before = ws.d_cand_cnt;
    GCHK(reserve(ws.d_cand_cnt, ws.cnt_cap, 2 * n_slots), "candidate counter allocation");
    if (ws.d_cand_cnt != before) ws.cnt_clean = false;
  }

  GemmParams gp;
  memset(&gp, 0, sizeof gp);
  // document rows: a bf16 matrix with a row pitch of row_bytes
  const bool ok = make_tmap(&gp.tmap_x, c.rows, c.n_rows, c.dim_padded, c.row_bytes, tile_rows) &&
                  make_tmap(&gp.tmap_q, ws.d_q_bf16, rows_padded, c.dim_padded, (uint64_t)c.dim_padded * 2, G_BM);
  if (!ok) {
    *err = cudaErrorInvalidValue;
    return "cuTensorMapEncodeTiled";
  }
  gp.n_rows_total = (uint32_t)c.n_rows;
  gp.tile_rows = tile_rows;
  gp.ranges = c.d_ranges;
  gp.range_prefix = c.d_range_prefix;
  gp.n_ranges = c.n_ranges;
  gp.m_tiles = m_tiles;
  gp.n_queries = c.n_queries;
  gp.kb = kb;
  gp.k = k;
  gp.cand_cap = cand_cap;
  gp.cand = ws.d_cand;
  gp.cand_cnt = ws.d_cand_cnt;
  gp.thr_state = ws.d_cand_cnt + n_slots;
  gp.lrank_of_row = c.lrank_of_row;
  gp.x_inv_norm = c.cosine ? c.x_inv_norm : nullptr;

  const bool pdl = !env_u32("PCV_NO_PDL", 0);
  auto launch_pair = [&](uint32_t grid) -> cudaError_t {
    if (shape == SHAPE_BF16)
      return kb == 6 ? launch_chain(gemm_topk_pair_kernel<6, SHAPE_BF16>, grid, G_THREADS, GP_SMEM_BYTES, c.stream, pdl, gp)
                     : launch_chain(gemm_topk_pair_kernel<0, SHAPE_BF16>, grid, G_THREADS, GP_SMEM_BYTES, c.stream, pdl, gp);
    return kb == 12 ? launch_chain(gemm_topk_pair_kernel<12, SHAPE_WIDE>, grid, G_THREADS, GP_SMEM_BYTES, c.stream, pdl, gp)
                    : launch_chain(gemm_topk_pair_kernel<0, SHAPE_WIDE>, grid, G_THREADS, GP_SMEM_BYTES, c.stream, pdl, gp);
  };
  auto launch_single = [&](uint32_t grid) -> cudaError_t {
    if (shape == SHAPE_BF16)
      return kb == 6 ? launch_chain(gemm_topk_kernel<6, SHAPE_BF16>, grid, G_THREADS, G_SMEM_BYTES, c.stream, pdl, gp)
                     : launch_chain(gemm_topk_kernel<0, SHAPE_BF16>, grid, G_THREADS, G_SMEM_BYTES, c.stream, pdl, gp);
    return kb == 12 ? launch_chain(gemm_topk_kernel<12, SHAPE_WIDE>, grid, G_THREADS, G_SMEM_BYTES, c.stream, pdl, gp)
                    : launch_chain(gemm_topk_kernel<0, SHAPE_WIDE>, grid, G_THREADS, G_SMEM_BYTES, c.stream, pdl, gp);
  };
  // The per-(CTA, query) candidate counters must be zero when a pass starts.  The select kernel resets every
  // counter it consumes, so between passes — and between searches — nothing has to be cleared; only a fresh
  // allocation, or a search that failed between a pass and its select, needs the memset.
  if (!ws.cnt_clean) GCHK(cudaMemsetAsync(ws.d_cand_cnt, 0, n_slots * sizeof(uint32_t), c.stream), "cudaMemsetAsync");
  ws.cnt_clean = false;

  // Pass schedule over the document tiles.  Large corpora start with a BOOTSTRAP pass: the first
  // `boot_tiles` tiles are scored once with nothing appended — each (tile, query) only writes its
  // maximum — and the k-th largest maximum per query is the entry threshold of the first real pass,
  // which then covers boot_tiles x ratio tiles (the bootstrap tiles again: 0.3 % of config 3).  That
  // replaces the three threshold-less / dense-candidate passes the geometric schedule needs to get
  // going (0.35 ms of a config-3 batch).  Small corpora, filtered searches (several row ranges) and
  // the single-CTA kernel keep the plain geometric schedule: pass 0 takes every score of a few tiles.
  const uint32_t T = c.total_tiles;
  const uint32_t first = std::max<uint32_t>(1u, std::min<uint32_t>(env_u32("PCV_GEMM_FIRST_TILES", 32), sms));
  const uint32_t boot_tiles = env_u32("PCV_GEMM_BOOT_TILES", 1024);
  const bool boot = pair_ok && sms >= 2 && c.n_ranges == 1 && boot_tiles >= 4 * k && boot_tiles <= 4096 &&
                    (uint64_t)T >= (uint64_t)boot_tiles * ratio;
  const uint32_t after_boot = env_u32("PCV_GEMM_AFTER_BOOT_TILES", 0);  // tuning: end of the first pass after a bootstrap
  uint32_t tb = 0;
  bool has_prev = false;
  bool has_thr = false;
  if (boot) {
    const uint32_t qp = m_tiles * G_BM;
    GCHK(reserve(ws.d_boot, ws.boot_cap, (size_t)boot_tiles * qp), "bootstrap buffer allocation");
    uint32_t grid = std::min<uint32_t>(sms, boot_tiles);
    grid -= grid % 2;
    // the epilogue still reads its counters (all stay zero: nothing passes a +inf threshold)
    gp.tile_begin = 0;
    gp.n_tiles = boot_tiles;
    gp.thr = nullptr;
    gp.boot_max = ws.d_boot;
    gp.boot_qp = qp;
    GCHK(launch_pair(grid), "gemm_topk_pair_kernel (bootstrap) launch");
    gp.boot_max = nullptr;
    GCHK(launch_chain(gemm_boot_threshold_kernel, c.n_queries, 256, ((size_t)boot_tiles + 2 * k) * sizeof(uint64_t), c.stream, pdl,
                      (const float*)ws.d_boot, boot_tiles, qp, k, ws.d_thr),
         "gemm_boot_threshold_kernel launch");
    nl += 2;
    has_thr = true;
  }
  for (;;) {
    uint64_t te64 = (tb == 0) ? (boot ? (after_boot ? (uint64_t)after_boot : (uint64_t)boot_tiles * ratio) : first)
                              : (tb >= dense_tiles ? (uint64_t)T : (uint64_t)tb * ratio);
    uint32_t te = (uint32_t)std::min<uint64_t>(te64, T);
    if ((uint64_t)(T - te) * 4 < te) te = T;  // do not leave a sliver for a pass of its own
    const uint32_t nt = te - tb;
    uint32_t grid = std::max<uint32_t>(1u, std::min<uint32_t>(sms, nt));
    const bool pair = pair_ok && grid >= 2;  // 2-CTA MMA over clusters of two
    if (pair) grid -= grid % 2;
    gp.tile_begin = tb;
    gp.n_tiles = nt;
    gp.thr = (has_prev || has_thr) ? ws.d_thr : nullptr;
    if (nt && pair) {
      GCHK(launch_pair(grid), "gemm_topk_pair_kernel launch");
      ++nl;
    }
    if (nt && !pair) {  // a single CTA's worth of tiles (or PCV_GEMM_NO_PAIR): the one-CTA kernel
      GCHK(launch_single(grid), "gemm_topk_kernel launch");
      ++nl;
    }
    SelectParams sp;
    memset(&sp, 0, sizeof sp);
    sp.cand = ws.d_cand;
    sp.cand_cnt = ws.d_cand_cnt;
    sp.grid_gemm = nt ? grid : 0u;
    sp.cand_cap = cand_cap;
    sp.m_tiles = m_tiles;
    sp.k = k;
    sp.smem_keys = (uint32_t)(sel_smem / sizeof(uint64_t));
    sp.topk = ws.d_topk;
    sp.has_prev = has_prev ? 1 : 0;
    sp.thr = ws.d_thr;
    sp.emit = (te == T && !c.keys_only) ? 1 : 0;
    sp.q_scale = c.cosine ? ws.d_qinv : nullptr;
    sp.cosine = c.cosine ? 1 : 0;
    sp.emit_mode = c.emit_mode;
    sp.dim = c.dim;
    sp.row_of_lrank = c.row_of_lrank;
    sp.ids = c.ids;
    sp.id_base = c.id_base;
    sp.out_ids = c.out_ids;
    sp.out_scores = c.out_scores;
    sp.out_sims = c.out_sims;
    sp.out_counts = c.out_counts;
    GCHK(launch_chain(gemm_select_kernel, c.n_queries, SEL_THREADS, sel_smem, c.stream, pdl, sp), "gemm_select_kernel launch");
    ++nl;
    has_prev = true;
    tb = te;
    if (tb >= T) break;
  }
#undef GCHK
  ws.cnt_clean = true;  // every pass was followed by its select: all counters are back to zero
  if (launches) *launches = nl;
  return nullptr;
}

}  // namespace pcv
