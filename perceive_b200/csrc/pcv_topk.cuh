// pcv_topk.cuh — warp-resident sorted top-k list (registers + shuffles).
//
// The fused top-k of every kernel in this library is built from one primitive:
// a descending-sorted list of u64 ranking keys (pcv_common.cuh) spread over the
// 32 lanes of a warp, KPL entries per lane (entry e lives in lane e%32, slot
// e/32), capacity 32*KPL >= k.  A candidate only reaches insert() after it has
// beaten the k-th key, so the common path per scored row is one compare.
// Replaces the per-source hnsw.search + concat + sort + truncate of
// crates/perceive-core/search.rs:163-181 with an exact selection.
#pragma once
#include "pcv_common.cuh"

namespace pcv {

template <int KPL>
struct WarpList {
  uint64_t v[KPL];

  __device__ __forceinline__ void clear() {
#pragma unroll
    for (int s = 0; s < KPL; ++s) v[s] = 0ull;
  }

  // Insert warp-uniform key x (distinct from every stored key).  Entries that x
  // beats move one position down; the last entry of the capacity falls off.
  __device__ __forceinline__ void insert(uint64_t x, int lane) {
    uint64_t carry = ~0ull;  // virtual predecessor of entry 0: beats everything
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
      const uint64_t old = v[s];
      uint64_t prev = shfl_up_u64(old, 1);
      if (lane == 0) prev = carry;
      carry = shfl_u64(old, 31);
      if (x > old) v[s] = (prev > x) ? x : prev;
    }
  }

  // key at sorted position pos (warp-uniform pos < 32*KPL)
  __device__ __forceinline__ uint64_t at(int pos) const {
    uint64_t t = 0ull;
#pragma unroll
    for (int s = 0; s < KPL; ++s)
      if (s == (pos >> 5)) t = v[s];
    return shfl_u64(t, pos & 31);
  }

  // Merge `n` keys stored DESCENDING at src (global or shared memory) into the
  // list, keeping only the best k.  Because src is sorted the scan stops at the
  // first key that does not beat the current k-th key.
  // CG = true reads through L2 (ld.global.cg): src was written by other CTAs of
  // the same launch.
  template <bool CG>
  __device__ __forceinline__ void merge_sorted_impl(const uint64_t* src, int n, int k, int lane) {
    uint64_t thr = at(k - 1);
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      uint64_t x = 0ull;
      if (i < n) {
        if constexpr (CG) x = __ldcg(reinterpret_cast<const unsigned long long*>(src) + i);
        else x = src[i];
      }
      unsigned m = __ballot_sync(PCV_FULL_MASK, x > thr);
      if (m == 0u) break;
      while (m) {
        const int src_lane = __ffs(m) - 1;
        m &= m - 1;
        const uint64_t xi = shfl_u64(x, src_lane);
        if (xi > thr) {
          insert(xi, lane);
          thr = at(k - 1);
        } else {
          m = 0u;  // sorted: nothing after this one can pass
        }
      }
      // a chunk that ended on a rejected key ends the whole scan
      const uint64_t last = shfl_u64(x, 31);
      if (!(last > thr)) break;
    }
  }

  __device__ __forceinline__ void merge_sorted(const uint64_t* src, int n, int k, int lane) {
    merge_sorted_impl<false>(src, n, k, lane);
  }
  __device__ __forceinline__ void merge_sorted_cg(const uint64_t* src, int n, int k, int lane) {
    merge_sorted_impl<true>(src, n, k, lane);
  }

  // Merge `n` keys in ARBITRARY order at src (global memory written by another
  // launch or by this warp; read through L2).  Empty slots are key 0.
  __device__ __forceinline__ void merge_unsorted(const uint64_t* src, int n, int k, int lane) {
    uint64_t thr = at(k - 1);
    for (int base = 0; base < n; base += 32) {
      const int i = base + lane;
      uint64_t x = 0ull;
      if (i < n) x = __ldcg(reinterpret_cast<const unsigned long long*>(src) + i);
      unsigned m = __ballot_sync(PCV_FULL_MASK, x > thr);
      while (m) {
        const int src_lane = __ffs(m) - 1;
        m &= m - 1;
        const uint64_t xi = shfl_u64(x, src_lane);
        if (xi > thr) {
          insert(xi, lane);
          thr = at(k - 1);
        }
      }
    }
  }

  // store the first k entries to dst[0..k)
  __device__ __forceinline__ void store(uint64_t* dst, int k, int lane) const {
#pragma unroll
    for (int s = 0; s < KPL; ++s) {
      const int e = s * 32 + lane;
      if (e < k) dst[e] = v[s];
    }
  }
};

}  // namespace pcv
