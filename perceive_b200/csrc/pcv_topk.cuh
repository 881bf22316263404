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

// ---------------------------------------------------------------------------
// Block-wide exact top-k of n keys held in shared memory (256 threads): an 8-round
// byte-wise radix select finds the k-th largest key, the survivors are ranked by
// counting.  Keys are distinct, 0 = empty slot.  Used where many short candidate
// lists meet (the last CTA of the scan): every key is loaded once, in parallel,
// instead of walking the lists one dependent load after another.
// ---------------------------------------------------------------------------
struct BlockSelectScratch {
  uint32_t hist[256];
  unsigned long long prefix;
  uint32_t need, nsel, live, exact;
};

// keys[n] (shared), sel[k] and out[k] (shared scratch / result, out sorted descending, 0-padded).
// Returns the number of non-empty results (<= k).  All 256 threads must call.
__device__ __forceinline__ uint32_t block_select_sorted(const uint64_t* keys, uint32_t n, uint32_t k, uint64_t* sel,
                                                        uint64_t* out, BlockSelectScratch& sc) {
  const uint32_t tid = threadIdx.x;
  constexpr uint32_t NT = 256;
  if (tid == 0) { sc.nsel = 0; sc.prefix = 0ull; sc.live = 0; sc.need = k; sc.exact = 0; }
  for (uint32_t i = tid; i < k; i += NT) out[i] = 0ull;
  __syncthreads();
  uint32_t live = 0;
  for (uint32_t i = tid; i < n; i += NT) live += keys[i] != 0ull;
  if (live) atomicAdd(&sc.live, live);
  __syncthreads();
  const uint32_t n_live = sc.live;
  const uint32_t want = min(k, n_live);
  uint64_t kth = 1ull;  // n_live <= k: every non-empty key survives
  if (n_live > k) {
    uint64_t mask = 0ull;
    for (int shift = 56; shift >= 0; shift -= 8) {
      sc.hist[tid] = 0;
      __syncthreads();
      const uint64_t prefix = sc.prefix;
      for (uint32_t i = tid; i < n; i += NT) {
        const uint64_t key = keys[i];
        if (key != 0ull && (key & mask) == prefix) atomicAdd(&sc.hist[(uint32_t)(key >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (tid < 32) {
        // warp 0: counts of the bins above each lane's 8 bins, top bin first
        uint32_t loc[8], sum = 0;
#pragma unroll
        for (int j = 0; j < 8; ++j) { loc[j] = sc.hist[255 - (tid * 8 + j)]; sum += loc[j]; }
        uint32_t incl = sum;
#pragma unroll
        for (int off = 1; off < 32; off <<= 1) {
          const uint32_t o = __shfl_up_sync(PCV_FULL_MASK, incl, off);
          if ((int)tid >= off) incl += o;
        }
        uint32_t before = incl - sum;
        const uint32_t need = sc.need;
        __syncwarp();
#pragma unroll
        for (int j = 0; j < 8; ++j) {
          if (before < need && before + loc[j] >= need) {
            sc.prefix = prefix | ((uint64_t)(255 - (tid * 8 + j)) << shift);
            sc.need = need - before;
            sc.exact = (before + loc[j] == need) ? 1u : 0u;  // the bin holds exactly what is still needed
          }
          before += loc[j];
        }
      }
      mask |= 0xffull << shift;
      __syncthreads();
      if (sc.exact) break;  // every key >= prefix (low bits zero) is a survivor: later rounds change nothing
    }
    kth = sc.prefix;
  }
  for (uint32_t i = tid; i < n; i += NT) {
    const uint64_t key = keys[i];
    if (key != 0ull && key >= kth) {
      const uint32_t pos = atomicAdd(&sc.nsel, 1u);
      if (pos < k) sel[pos] = key;
    }
  }
  __syncthreads();
  const uint32_t nsel = min(sc.nsel, want);
  for (uint32_t i = tid; i < nsel; i += NT) {
    const uint64_t mine = sel[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < nsel; ++j) rank += sel[j] > mine;
    out[rank] = mine;
  }
  __syncthreads();
  return nsel;
}

// ---------------------------------------------------------------------------
// The same selection for the shape the scan produces — n_lists lists of k keys, each sorted
// descending and 0-padded — when the lists far outnumber k (148 CTAs, k = 10): the k-th largest
// list HEAD h is a lower bound of the k-th largest key overall (k heads are >= h), so only keys >= h
// can be in the result, and on data dealt evenly to the lists there are barely more than k of them.
// Heads are ranked by counting, the survivors gathered and ranked by counting: four barriers, no
// radix rounds.  More than sel_cap survivors (the best rows all sit in a few lists): the radix
// select above does the job.  sel holds sel_cap (>= k) keys.  All 256 threads must call.
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t block_select_lists(const uint64_t* keys, uint32_t n_lists, uint32_t k, uint64_t* sel,
                                                       uint32_t sel_cap, uint64_t* out, BlockSelectScratch& sc) {
  const uint32_t tid = threadIdx.x;
  constexpr uint32_t NT = 256;
  const uint32_t n = n_lists * k;
  if (n_lists < 4u * k) return block_select_sorted(keys, n, k, sel, out, sc);  // block-uniform
  if (tid == 0) { sc.nsel = 0; sc.prefix = 0ull; }
  for (uint32_t i = tid; i < k; i += NT) out[i] = 0ull;
  __syncthreads();
  for (uint32_t c = tid; c < n_lists; c += NT) {
    const uint64_t mine = keys[(size_t)c * k];
    if (mine == 0ull) continue;  // an empty list; fewer than k non-empty heads leave h = 0 (every key survives)
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n_lists; ++j) rank += keys[(size_t)j * k] > mine;  // non-empty keys are distinct
    if (rank == k - 1) sc.prefix = mine;
  }
  __syncthreads();
  const uint64_t h = sc.prefix;
  for (uint32_t i = tid; i < n; i += NT) {
    const uint64_t key = keys[i];
    if (key != 0ull && key >= h) {
      const uint32_t pos = atomicAdd(&sc.nsel, 1u);
      if (pos < sel_cap) sel[pos] = key;
    }
  }
  __syncthreads();
  const uint32_t nsel = sc.nsel;
  if (nsel > sel_cap) {
    __syncthreads();  // everyone has read nsel before the radix select resets the scratch
    return block_select_sorted(keys, n, k, sel, out, sc);
  }
  for (uint32_t i = tid; i < nsel; i += NT) {
    const uint64_t mine = sel[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < nsel; ++j) rank += sel[j] > mine;
    if (rank < k) out[rank] = mine;
  }
  __syncthreads();
  return min(nsel, k);
}

// Merge n_lists (<= 8) sorted lists of k keys (shared memory, list after list) into dst[0..k), best first:
// every key's rank among all of them by counting.  Keys are distinct but for the empty ones (0), which the
// index orders.  256 threads; the caller synchronises before reading dst if it lives in shared memory.
__device__ __forceinline__ void block_merge_lists(const uint64_t* lists, uint32_t n_lists, uint32_t k, uint64_t* dst) {
  const uint32_t n = n_lists * k;
  for (uint32_t i = threadIdx.x; i < n; i += 256u) {
    const uint64_t mine = lists[i];
    uint32_t rank = 0;
    for (uint32_t j = 0; j < n; ++j) {
      const uint64_t o = lists[j];
      rank += (o > mine) || (o == mine && j < i);
    }
    if (rank < k) dst[rank] = mine;
  }
}

}  // namespace pcv
