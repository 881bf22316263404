// pcv_gemm_pair.cuh — K2 (and K3's filter pass) with 2-CTA tensor-core instructions (tcgen05.mma.cta_group::2).
// Included by pcv_gemm.cu after GemmParams / GemmShape / gemm_tile_rows.
//
// Why: the query ring is latency-bound, not bandwidth-bound.  A ring stage makes one round trip
// (MMA retires -> commit -> producer wakes -> TMA -> L2 -> barrier -> issue), roughly 2000 cycles,
// and shared memory holds only 6 stages.  One stage feeds 128 x BN MACs per K16,
// i.e. 128-384 cycles of tensor work, so the single-CTA kernel tops out at
// stages x step_time / round_trip (measured: 0.83 of the sustained bf16 rate at 384-d, 0.54 at 768-d).  In pair mode one stage of EACH CTA's query tile is multiplied against the
// document rows of BOTH CTAs (N = 2*BN), so a stage lasts twice as long and the same ring covers
// twice the latency.
//
// A pair (cluster of 2, same TPC) owns a contiguous share of the pass's document tiles, taken two
// at a time: CTA r keeps tile 2i+r resident (its half of the B operand).  Query tiles are taken
// two at a time too: CTA r streams query tile 2j+r (its half of the A operand, M = 256 over the
// pair).  Only the leader (rank 0) issues MMAs; accumulators land in both CTAs' TMEM (each CTA
// gets the scores of ITS 128 queries against all 2*BN rows) and both CTAs run the epilogue.
//   both CTAs' TMA loads (cp.async.bulk.tensor ... cta_group::2) complete their bytes on the LEADER's
//   "full" barriers (the barrier address with the pair's peer bit cleared), which therefore expect
//   the bytes of both halves; the leader's commits are multicast to both CTAs' "empty" / "tmem
//   full" barriers; the peer's epilogue warps arrive remotely on the leader's "tmem empty" barrier.
#pragma once

namespace pcv {
namespace {

constexpr uint32_t GP_NBARS = G_NBARS;
constexpr uint32_t PCV_PEER_BIT_MASK = 0xFEFFFFFFu;  // shared-window address bit that selects the odd CTA of a pair
constexpr uint32_t GP_SMEM_BYTES = G_SMEM_RINGS + GP_NBARS * 8 + 16 + 1024;
static_assert(GP_SMEM_BYTES <= 232448, "pair kernel shared memory budget");

__device__ __forceinline__ uint32_t mapa_u32(uint32_t local_addr, uint32_t cta_rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(local_addr), "r"(cta_rank));
  return r;
}
__device__ __forceinline__ void mbar_arrive_remote(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// 2-D tiled load whose completion bytes are credited to a barrier of the pair's LEADER CTA
__device__ __forceinline__ void tma2_load_2d(uint32_t dst_smem, const void* tmap, uint32_t leader_bar, int32_t crd0,
                                             int32_t crd1, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint"
      " [%0], [%1, {%3, %4}], [%2], %5;" ::"r"(dst_smem),
      "l"(tmap), "r"(leader_bar), "r"(crd0), "r"(crd1), "l"(policy)
      : "memory");
}
__device__ __forceinline__ void tc2_alloc(uint32_t dst_smem, uint32_t cols) {
  asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(dst_smem), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc2_relinquish() {
  asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
}
__device__ __forceinline__ void tc2_dealloc(uint32_t taddr, uint32_t cols) {
  asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(taddr), "r"(cols) : "memory");
}
__device__ __forceinline__ void tc2_mma_bf16(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc,
                                             uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc2_commit_multicast(uint32_t bar, uint16_t cta_mask) {
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(bar),
               "h"(cta_mask)
               : "memory");
}

template <int KB_T, int SHAPE>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(G_THREADS, 1)
    gemm_topk_pair_kernel(const __grid_constant__ GemmParams p) {
  using SH = GemmShape<SHAPE>;
  constexpr int BN = SH::BN;            // document rows per CTA per item
  constexpr int N2 = 2 * BN;            // UMMA N: both CTAs' rows
  constexpr int QSTAGES = SH::QSTAGES;
  constexpr int XSLOTS = SH::XSLOTS;
  constexpr uint32_t XSLOT_BYTES = SH::XSLOT_BYTES;
  constexpr int NACC = 512 / N2;        // TMEM accumulators
  static_assert(KB_T == 0 || KB_T % QSTAGES == 0, "static K-block count must be a multiple of the query ring depth");
  extern __shared__ uint8_t smem_raw[];
  const uint32_t raw_addr = smem_u32(smem_raw);
  uint8_t* smem = smem_raw + ((1024u - (raw_addr & 1023u)) & 1023u);  // same offset in both CTAs
  uint8_t* smem_x = smem;
  uint8_t* smem_q = smem + SH::SMEM_X;
  uint64_t* bars = reinterpret_cast<uint64_t*>(smem + G_SMEM_RINGS);
  uint64_t* bar_xfull = bars;                        // [XSLOTS]  own TMA -> leader MMA / peer forwarder
  uint64_t* bar_xempty = bar_xfull + G_MAX_XSLOTS;   // [XSLOTS]  leader MMA -> both producers
  uint64_t* bar_qfull = bar_xempty + G_MAX_XSLOTS;   // [QSTAGES]
  uint64_t* bar_qempty = bar_qfull + G_MAX_QSTAGES;  // [QSTAGES]
  uint64_t* bar_tfull = bar_qempty + G_MAX_QSTAGES;  // [NACC]    leader MMA -> both epilogues
  uint64_t* bar_tempty = bar_tfull + G_ACC;          // [NACC]    both epilogues -> leader MMA (leader's copy)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(bars + GP_NBARS);

  const int warp = __shfl_sync(PCV_FULL_MASK, (int)(threadIdx.x >> 5), 0);
  const int lane = threadIdx.x & 31;
  const uint32_t cta = blockIdx.x;
  const uint32_t crank = cluster_ctarank();
  const uint32_t n_pairs = gridDim.x / 2, pair = cta / 2;
  const uint32_t g0 = (uint32_t)((uint64_t)pair * p.n_tiles / n_pairs);
  const uint32_t g1 = (uint32_t)((uint64_t)(pair + 1) * p.n_tiles / n_pairs);
  const uint32_t rounds = (g1 - g0 + 1) / 2;  // document tiles are taken two at a time
  // tile of CTA `r` in round i: g0 + 2i + r; past the share = dummy tile (zero rows)
  auto tile_of = [&](uint32_t i, uint32_t r, uint32_t& row0, uint32_t& nrows) {
    const uint32_t t = g0 + 2 * i + r;
    if (t < g1) {
      gemm_tile_rows(p, t + p.tile_begin, row0, nrows);
    } else {
      row0 = p.n_rows_total;  // wholly out of bounds: TMA fills zeros, nothing is appended
      nrows = 0;
    }
  };
  const uint32_t m_tiles = p.m_tiles;          // real query tiles
  const uint32_t m_pairs = (m_tiles + 1) / 2;  // the query buffer is padded to an even tile count
  const uint32_t KB = KB_T ? (uint32_t)KB_T : p.kb;
  // One query tile pair whose K blocks all fit the query ring (batches of up to 256 queries, dim <= 384):
  // the queries are loaded ONCE and stay in shared memory for the whole pass instead of being re-streamed
  // from L2 for every document tile — half the L2->SM traffic of a pass that is power-capped, not stalled.
  const bool q_stationary = (m_pairs == 1) && (KB <= (uint32_t)QSTAGES);

  if (warp == 1 && lane == 0) {
    for (int s = 0; s < XSLOTS; ++s) {
      mbar_init(smem_u32(bar_xfull + s), 1);
      mbar_init(smem_u32(bar_xempty + s), 1);
    }
    for (int s = 0; s < QSTAGES; ++s) {
      mbar_init(smem_u32(bar_qfull + s), 1);
      mbar_init(smem_u32(bar_qempty + s), 1);
    }
    for (int a = 0; a < NACC; ++a) {
      mbar_init(smem_u32(bar_tfull + a), 1);
      mbar_init(smem_u32(bar_tempty + a), 8);  // 4 epilogue warps of each CTA
    }
    mbar_fence_init();
    fence_proxy_async_smem();
  }
  if (warp == 0) {
    tc2_alloc(smem_u32(tmem_slot), 512);
    tc2_relinquish();
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // both CTAs' barriers and TMEM exist before any cross-CTA signal
  tc_fence_after();
  const uint32_t tmem_base = *reinterpret_cast<volatile uint32_t*>(tmem_slot);
  pdl_launch_dependents();  // the select kernel behind this pass may start being scheduled
  pdl_wait();               // thresholds, counters and queries come from the kernels before this one

  if (warp == 0) {
    // ===================== query producer: this CTA's query tile of every pair of tiles ==========
    if (elect_one_sync()) {
      tma_prefetch_desc(&p.tmap_q);
    }
    const uint64_t pol = l2_policy_evict_last();
    const uint32_t q_base = smem_u32(smem_q), full0 = smem_u32(bar_qfull), empty0 = smem_u32(bar_qempty);
    uint32_t stage = 0, phase = 0;
    bool ready = false;
    for (uint32_t t = 0; t < (q_stationary ? min(rounds, 1u) : rounds); ++t)
      for (uint32_t mp = 0; mp < m_pairs; ++mp) {
        const uint32_t m = 2 * mp + crank;
        for (uint32_t kb = 0; kb < KB; ++kb) {
          if (!ready) mbar_wait_bounded(empty0 + stage * 8, phase ^ 1u);
          uint32_t nstage = stage + 1, nphase = phase;
          if (nstage == (uint32_t)QSTAGES) { nstage = 0; nphase ^= 1u; }
          ready = mbar_test(empty0 + nstage * 8, nphase ^ 1u);
          if (elect_one_sync()) {
            const uint32_t lbar = (full0 + stage * 8) & PCV_PEER_BIT_MASK;  // the leader's barrier, from either CTA
            if (crank == 0) mbar_arrive_expect_tx(full0 + stage * 8, 2 * SH::QSTAGE_BYTES);  // both CTAs' stages
            tma2_load_2d(q_base + stage * SH::QSTAGE_BYTES, &p.tmap_q, lbar, (int32_t)(kb * G_BK), (int32_t)(m * G_BM), pol);
          }
          __syncwarp();
          stage = nstage;
          phase = nphase;
        }
      }
  } else if (warp == 6) {
    // ===================== document producer: this CTA's half of every pair of tiles ============
    if (elect_one_sync()) {
      tma_prefetch_desc(&p.tmap_x);
    }
    const uint64_t pol = l2_policy_evict_first();
    const uint32_t x_base = smem_u32(smem_x), full0 = smem_u32(bar_xfull), empty0 = smem_u32(bar_xempty);
    uint32_t slot = 0, phase = 0;
    bool ready = false;
    for (uint32_t t = 0; t < rounds; ++t) {
      uint32_t row0, nrows;
      tile_of(t, crank, row0, nrows);
      for (uint32_t kb = 0; kb < KB; ++kb) {
        if (!ready) mbar_wait_bounded(empty0 + slot * 8, phase ^ 1u);
        uint32_t nslot = slot + 1, nphase = phase;
        if (nslot == (uint32_t)XSLOTS) { nslot = 0; nphase ^= 1u; }
        ready = mbar_test(empty0 + nslot * 8, nphase ^ 1u);
        if (elect_one_sync()) {
          const uint32_t lbar = (full0 + slot * 8) & PCV_PEER_BIT_MASK;
          if (crank == 0) mbar_arrive_expect_tx(full0 + slot * 8, 2 * XSLOT_BYTES);
          tma2_load_2d(x_base + slot * XSLOT_BYTES, &p.tmap_x, lbar, (int32_t)(kb * G_BK), (int32_t)row0, pol);
        }
        __syncwarp();
        slot = nslot;
        phase = nphase;
      }
    }
  } else if (warp == 1) {
    const uint32_t qfull0 = smem_u32(bar_qfull), xfull0 = smem_u32(bar_xfull);
    if (crank == 0) {
      // ===================== MMA issuer (leader): M = 256 over the pair, N = 2*BN ==================
      constexpr uint32_t idesc = umma_idesc_bf16_f32(256, N2);
      const uint64_t q_desc0 = umma_desc_k_sw128(smem_u32(smem_q));
      const uint64_t x_desc0 = umma_desc_k_sw128(smem_u32(smem_x));
      const uint32_t qempty0 = smem_u32(bar_qempty), xempty0 = smem_u32(bar_xempty);
      const uint32_t tfull0 = smem_u32(bar_tfull), tempty0 = smem_u32(bar_tempty);
      uint32_t qstage = 0, qphase = 0, acc = 0, acc_par = 0;
      uint32_t xslot_tile = 0, xphase_tile = 0;
      for (uint32_t t = 0; t < rounds; ++t) {
        for (uint32_t mp = 0; mp < m_pairs; ++mp) {
          mbar_wait_bounded(tempty0 + acc * 8, acc_par ^ 1u);
          tc_fence_after();
          const uint32_t d_tmem = tmem_base + acc * N2;
          const bool last_m = (mp + 1 == m_pairs);
          uint32_t xslot = xslot_tile, xphase = xphase_tile;
          bool ready = false;
          const bool q_wait = !q_stationary || t == 0;  // stationary queries: filled during the first round only
          if (!KB_T && q_stationary) { qstage = 0; qphase = 0; }
#pragma unroll
          for (uint32_t kb = 0; kb < (KB_T ? (uint32_t)KB_T : KB); ++kb) {
            const uint32_t qs = KB_T ? kb % QSTAGES : qstage;
            const uint32_t qph = KB_T ? (qphase ^ ((kb / QSTAGES) & 1u)) : qphase;
            if (!ready) {
              if (mp == 0) mbar_wait_bounded(xfull0 + xslot * 8, xphase);
              if (q_wait) mbar_wait_bounded(qfull0 + qs * 8, qph);
            }
            tc_fence_after();
            uint32_t nqs, nqph, nxslot = xslot + 1, nxphase = xphase;
            if (KB_T) {
              nqs = (kb + 1) % QSTAGES;
              nqph = qphase ^ (((kb + 1) / QSTAGES) & 1u);
            } else {
              nqs = qstage + 1;
              nqph = qphase;
              if (nqs == (uint32_t)QSTAGES) { nqs = 0; nqph ^= 1u; }
            }
            if (nxslot == (uint32_t)XSLOTS) { nxslot = 0; nxphase ^= 1u; }
            if (kb + 1 < (KB_T ? (uint32_t)KB_T : KB)) {
              ready = q_wait ? mbar_test(qfull0 + nqs * 8, nqph) : true;
              if (mp == 0) ready = mbar_test(xfull0 + nxslot * 8, nxphase) && ready;
            } else {
              ready = false;
            }
            if (elect_one_sync()) {
              const uint64_t a_desc = q_desc0 + (uint64_t)((qs * SH::QSTAGE_BYTES) >> 4);
              const uint64_t b_desc = x_desc0 + (uint64_t)((xslot * XSLOT_BYTES) >> 4);
#pragma unroll
              for (uint32_t j = 0; j < G_BK / 16; ++j)
                tc2_mma_bf16(d_tmem, a_desc + j * 2, b_desc + j * 2, idesc, (kb | j) != 0u);
              if (!q_stationary) tc2_commit_multicast(qempty0 + qs * 8, 3);  // both CTAs' query stage is free
              if (last_m) tc2_commit_multicast(xempty0 + xslot * 8, 3);  // both CTAs' document slot is free
            }
            __syncwarp();
            if (!KB_T) { qstage = nqs; qphase = nqph; }
            xslot = nxslot;
            xphase = nxphase;
          }
          if (KB_T && !q_stationary) qphase ^= (uint32_t)((KB_T / QSTAGES) & 1);
          if (elect_one_sync()) tc2_commit_multicast(tfull0 + acc * 8, 3);  // both epilogues may read
          __syncwarp();
          if (++acc == (uint32_t)NACC) { acc = 0; acc_par ^= 1u; }
          if (last_m) { xslot_tile = xslot; xphase_tile = xphase; }
        }
      }
    }
  } else {
    // ===================== epilogue (both CTAs): this CTA's 128 queries vs 2*BN rows ===============
    const int quarter = warp & 3;
    const int row = quarter * 32 + lane;
    const int k = (int)p.k;
    const size_t slot0 = (size_t)cta * m_tiles * G_BM + (size_t)row;
    for (uint32_t m = crank; m < m_tiles; m += 2) {  // this CTA only ever sees query tiles of its parity
      const uint32_t q = m * G_BM + (uint32_t)row;
      // bootstrap pass: +inf, nothing passes the filter; only the tile maxima are written
      const float th = (q < p.n_queries && !p.boot_max) ? (p.thr ? __ldg(p.thr + q) : -CUDART_INF_F) : CUDART_INF_F;
      p.thr_state[slot0 + (size_t)m * G_BM] = f32_to_ordered(th);
    }
    uint32_t acc = 0, acc_par = 0;
    const uint32_t tfull0 = smem_u32(bar_tfull);
    const uint32_t tempty_leader = mapa_u32(smem_u32(bar_tempty), 0);  // the leader's copy, from either CTA
    const float* __restrict__ xinv = p.x_inv_norm;
    const uint32_t my_items = (m_tiles > crank) ? (m_tiles - crank + 1) / 2 : 0;  // real query tiles of this CTA
    uint32_t cnt_next = my_items ? p.cand_cnt[slot0 + (size_t)crank * G_BM] : 0u;
    uint32_t thr_next = my_items ? p.thr_state[slot0 + (size_t)crank * G_BM] : f32_to_ordered(CUDART_INF_F);
    for (uint32_t t = 0; t < rounds; ++t) {
      uint32_t row0h[2], nrowsh[2];
      tile_of(t, 0, row0h[0], nrowsh[0]);
      tile_of(t, 1, row0h[1], nrowsh[1]);
      for (uint32_t mp = 0; mp < m_pairs; ++mp) {
        const uint32_t m = 2 * mp + crank;
        const bool real = m < m_tiles;  // the padding query tile of an odd batch: drain TMEM only
        const size_t slot = slot0 + (size_t)(real ? m : crank) * G_BM;
        uint32_t cnt = cnt_next;
        float thr = real ? ordered_to_f32(thr_next) : CUDART_INF_F;
        if (my_items > 1 || !real) {
          uint32_t nm = m + 2;
          if (nm >= m_tiles) nm = crank;
          if (nm < m_tiles) {
            const size_t nslot = slot0 + (size_t)nm * G_BM;
            cnt_next = p.cand_cnt[nslot];
            thr_next = p.thr_state[nslot];
          }
        }
        uint64_t* buf = p.cand + slot * p.cand_cap;
        const uint32_t cnt_in = cnt;
        mbar_wait_bounded(tfull0 + acc * 8, acc_par);
        tc_fence_after();
        const uint32_t tbase = tmem_base + ((uint32_t)(quarter * 32) << 16) + acc * N2;
#pragma unroll 1
        for (int h = 0; h < 2; ++h) {  // columns [h*BN, (h+1)*BN): the document tile of CTA h
          const uint32_t taddr = tbase + h * BN;
          const uint32_t row0 = row0h[h], nrows = nrowsh[h];
          float rmx;
          uint32_t gm = 0;
          {
            constexpr int NCH = BN / 32;
            uint32_t v[NCH][32];
#pragma unroll
            for (int c = 0; c < NCH; ++c) tc_ld_32x32b_x32(taddr + 32 * c, v[c]);
            tc_wait_ld();
            if (xinv) {
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
          if (p.boot_max) {
            const uint32_t tl = g0 + 2u * t + (uint32_t)h;  // tile index inside this pass
            if (real && tl < g1)
              p.boot_max[(size_t)tl * p.boot_qp + m * G_BM + (uint32_t)row] = (nrows == (uint32_t)BN) ? rmx : -CUDART_INF_F;
          }
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
        }
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_remote(tempty_leader + acc * 8);
        if (++acc == (uint32_t)NACC) { acc = 0; acc_par ^= 1u; }

        unsigned need = __ballot_sync(PCV_FULL_MASK, real && cnt + (uint32_t)N2 > p.cand_cap);
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
        if (real && cnt != cnt_in) p.cand_cnt[slot] = cnt;
        if (my_items == 1 && real) { cnt_next = cnt; thr_next = f32_to_ordered(thr); }
      }
    }
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();  // neither CTA leaves (or frees TMEM) while the other can still signal it
  if (warp == 0) {
    __syncwarp();
    tc2_dealloc(tmem_base, 512);
  }
}

}  // namespace
}  // namespace pcv
