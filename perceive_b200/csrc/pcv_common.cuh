// pcv_common.cuh — shared device/host helpers for libperceive_cuda (sm_100a).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

#define PCV_WARP 32
#define PCV_FULL_MASK 0xffffffffu

namespace pcv {

// ---------------------------------------------------------------------------
// Ranking keys.  Every top-k structure in the library orders candidates by ONE
// unsigned 64-bit key:  hi32 = order-preserving image of the fp32 similarity,
// lo32 = ~local_rank, where local_rank orders this shard's rows by doc id.
// Larger key == better, so "similarity descending, ties -> lower id first"
// (the tie-break stated in include/perceive_cuda.h; the reference's own is an
// unstable sort, search.rs:179) is a plain integer max.  Key 0 == empty slot.
// ---------------------------------------------------------------------------
__host__ __device__ __forceinline__ uint32_t f32_to_ordered(float f) {
  f = f + 0.0f;  // -0.0 -> +0.0 so equal floats get equal images
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } cv; cv.f = f; uint32_t b = cv.u;
#endif
  return b ^ ((b & 0x80000000u) ? 0xffffffffu : 0x80000000u);
}
__host__ __device__ __forceinline__ float ordered_to_f32(uint32_t o) {
  uint32_t b = o ^ ((o & 0x80000000u) ? 0x80000000u : 0xffffffffu);
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } cv; cv.u = b; return cv.f;
#endif
}
__host__ __device__ __forceinline__ uint64_t make_key(float sim, uint32_t lrank) {
  return ((uint64_t)f32_to_ordered(sim) << 32) | (uint64_t)(0xffffffffu - lrank);
}
__host__ __device__ __forceinline__ float key_sim(uint64_t key) { return ordered_to_f32((uint32_t)(key >> 32)); }
__host__ __device__ __forceinline__ uint32_t key_lrank(uint64_t key) { return 0xffffffffu - (uint32_t)key; }

// reference distance, crates/perceive-core/search.rs:274-277 (fp32 throughout)
__host__ __device__ __forceinline__ float ref_distance(float dot, uint32_t dim) {
  float r = 1.0f - (dot / (float)dim);
  return r > 0.0f ? r : 0.0f;
}

#ifdef __CUDACC__
// ---------------------------------------------------------------------------
// PTX wrappers: mbarrier + 1-D bulk async copy (TMA engine, SASS UBLKCP).
// ---------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) {
  return (uint32_t)__cvta_generic_to_shared(p);
}
__device__ __forceinline__ void mbar_init(uint32_t bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(bar), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_fence_init() {
  asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void fence_proxy_async_smem() {
  asm volatile("fence.proxy.async.shared::cta;" ::: "memory");
}
__device__ __forceinline__ void mbar_arrive_expect_tx(uint32_t bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(bar), "r"(bytes)
               : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint32_t bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(bar) : "memory");
}
__device__ __forceinline__ bool mbar_try_wait(uint32_t bar, uint32_t parity) {
  uint32_t ok;
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
      "selp.u32 %0, 1, 0, p;\n\t}"
      : "=r"(ok)
      : "r"(bar), "r"(parity)
      : "memory");
  return ok != 0;
}
__device__ __forceinline__ void mbar_wait(uint32_t bar, uint32_t parity) {
  while (!mbar_try_wait(bar, parity)) {
  }
}
// global -> shared bulk copy, completion signalled on an mbarrier (complete_tx).
// dst/src 16-byte aligned, bytes a multiple of 16.
__device__ __forceinline__ void bulk_g2s(uint32_t dst_smem, const void* src, uint32_t bytes,
                                         uint32_t bar) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::
          "r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar)
      : "memory");
}
__device__ __forceinline__ void bulk_g2s_hint(uint32_t dst_smem, const void* src, uint32_t bytes,
                                              uint32_t bar, uint64_t policy) {
  asm volatile(
      "cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes.L2::cache_hint [%0], "
      "[%1], %2, [%3], %4;" ::"r"(dst_smem),
      "l"(src), "r"(bytes), "r"(bar), "l"(policy)
      : "memory");
}
__device__ __forceinline__ uint64_t l2_policy_evict_first() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_first.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t l2_policy_evict_last() {
  uint64_t p;
  asm volatile("createpolicy.fractional.L2::evict_last.b64 %0, 1.0;" : "=l"(p));
  return p;
}
__device__ __forceinline__ uint64_t shfl_u64(uint64_t v, int src) {
  return __shfl_sync(PCV_FULL_MASK, (unsigned long long)v, src);
}
__device__ __forceinline__ uint64_t shfl_up_u64(uint64_t v, int delta) {
  return __shfl_up_sync(PCV_FULL_MASK, (unsigned long long)v, delta);
}
#endif  // __CUDACC__

}  // namespace pcv
