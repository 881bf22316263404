// pcv_exchange.cuh — cross-shard exchange of top-k candidates (K5): the merge of sorted (sim, id) lists, the
// layout of the peer-memory receive buffers, and the device routine by which a PRODUCER kernel (the scan's last
// CTA) delivers its candidates straight into its peers' buffers and merges — so a sharded single-query search
// is ONE launch ("from the top-k epilogue + flag").
#pragma once
#include <math_constants.h>
#include "pcv_common.cuh"

namespace pcv {

// K5 — merge candidate lists from `n_lists` shards.  One warp per query; lane l
// tracks the head of list l.  Lists are sorted (sim desc, id asc) and padded with
// (-inf, INT64_MAX).  Mirrors the concat + sort + truncate of
// crates/perceive-core/search.rs:177-181 across shards instead of sources.
__device__ __forceinline__ void merge_lists_warp(const float* __restrict__ sims, size_t sims_list_stride,
                                                 const int64_t* __restrict__ ids, size_t ids_list_stride,
                                                 uint32_t n_lists, uint32_t q, uint32_t k, uint32_t dim, int cosine,
                                                 int64_t* __restrict__ out_ids, float* __restrict__ out_scores,
                                                 float* __restrict__ out_sims, uint32_t* __restrict__ out_counts,
                                                 int lane) {
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

// ---------------------------------------------------------------------------
// K5p — the same exchange WITHOUT NCCL: candidates travel as plain stores into
// peer memory over NVLink (buffers mapped with CUDA IPC), completion is a
// release-store of the search's epoch into the peer's flag word, and the merge
// runs in the same launch once every shard's flag shows the epoch.  One launch
// replaces ncclAllGather + merge_candidates_kernel ("B200-native
// alternative": every peer is one uniform NVSwitch hop away and the payload is
// B*k*12 bytes, so the exchange is latency-, not bandwidth-bound).
// Receive buffer of one rank, per epoch parity: sims[world][cap] f32,
// ids[world][cap] i64, flags[world] u32.
// ---------------------------------------------------------------------------
#define PCV_P2P_MAX_WORLD 16

struct P2PParams {
  const int64_t* s_ids;  // this shard's candidates (emit_mode 1 output of K1 / K2)
  const float* s_sims;
  uint32_t n_queries, k, dim;
  int cosine;
  uint32_t rank, world, cap;  // cap: records per list in the receive buffers
  uint32_t epoch;
  uint8_t* peer[PCV_P2P_MAX_WORLD];  // receive buffer of every rank (this rank's own included)
  unsigned int* done_ctr;            // local: CTAs that finished their stores
  int64_t* out_ids;
  float* out_scores;
  float* out_sims;
  uint32_t* out_counts;
};

__host__ __device__ __forceinline__ size_t p2p_half_bytes(uint32_t world, uint32_t cap) {
  return ((size_t)world * cap * 12 + (size_t)world * 4 + 127) / 128 * 128;
}
__device__ __forceinline__ void st_release_sys_u32(unsigned int* p, unsigned int v) {
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned int ld_acquire_sys_u32(const unsigned int* p) {
  unsigned int v;
  asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}


// Where a producer kernel delivers its candidates in a sharded search (ScanParams::xchg, emit_mode 2).
struct ExchangeTarget {
  uint8_t* peer[PCV_P2P_MAX_WORLD];  // receive buffer of every rank (this rank's own included)
  uint32_t rank, world, cap, epoch;
};

// Address of record `i` of list `rank` in rank `dst`'s receive buffer (epoch parity selects the half).
__device__ __forceinline__ void exchange_slot(const ExchangeTarget& x, uint32_t dst, uint32_t i, float*& sim, int64_t*& id) {
  uint8_t* base = x.peer[dst] + p2p_half_bytes(x.world, x.cap) * (x.epoch & 1u);
  sim = reinterpret_cast<float*>(base) + (size_t)x.rank * x.cap + i;
  id = reinterpret_cast<int64_t*>(base + (size_t)x.world * x.cap * 4) + (size_t)x.rank * x.cap + i;
}

// Called by EVERY thread of the one CTA that stored this shard's n_queries * k candidates through
// exchange_slot(): publish the epoch to every rank, wait until every shard's candidates have landed here
// (bounded spin: a dead peer fails the launch instead of hanging the GPU), merge.  blockDim.x >= 32 * n_queries
// is not required: warps take queries in turn.
__device__ __forceinline__ void exchange_publish_and_merge(const ExchangeTarget& x, uint32_t n_queries, uint32_t k, uint32_t dim,
                                                           int cosine, int64_t* out_ids, float* out_scores, float* out_sims,
                                                           uint32_t* out_counts) {
  const size_t half = p2p_half_bytes(x.world, x.cap) * (x.epoch & 1u);
  const size_t ids_off = (size_t)x.world * x.cap * 4;
  const size_t flags_off = (size_t)x.world * x.cap * 12;
  __threadfence_system();
  __syncthreads();  // every thread's peer stores are ordered before the flags
  if (threadIdx.x < x.world) {
    unsigned int* flag = reinterpret_cast<unsigned int*>(x.peer[threadIdx.x] + half + flags_off) + x.rank;
    st_release_sys_u32(flag, x.epoch);
  }
  const int lane = threadIdx.x & 31;
  const uint8_t* mine = x.peer[x.rank] + half;
  const unsigned int* flags = reinterpret_cast<const unsigned int*>(mine + flags_off);
  if ((uint32_t)lane < x.world) {
    uint32_t spins = 0;
    while (ld_acquire_sys_u32(flags + lane) != x.epoch) {
      if (++spins > (1u << 27)) __trap();
    }
  }
  __syncwarp();
  const float* r_sims = reinterpret_cast<const float*>(mine);
  const int64_t* r_ids = reinterpret_cast<const int64_t*>(mine + ids_off);
  for (uint32_t q = threadIdx.x >> 5; q < n_queries; q += blockDim.x >> 5)
    merge_lists_warp(r_sims, x.cap, r_ids, x.cap, x.world, q, k, dim, cosine, out_ids, out_scores, out_sims, out_counts, lane);
}

}  // namespace pcv
