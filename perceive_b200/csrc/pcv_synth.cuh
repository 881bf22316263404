// pcv_synth.cuh — deterministic counter-based synthetic corpus.
//
// value(seed,row,col) is a pure function built from integer hashing and IEEE
// fp32 ops that round identically on host and device (explicit fmaf, sqrtf,
// division; no transcendental), so a 100M-row corpus can be generated on the
// device and any row regenerated on the host bit-for-bit.
//   g(seed,row,col) = Irwin-Hall(4) approx-gaussian from one 64-bit hash
//   UNIT_SPHERE: x = g / max(|g|, 1e-12)  (normalisation form of
//                crates/perceive-core/model/worker.rs:95-103)
//   SCALED     : x = g * s(row), s = mantissa in [1,2) times 2^e, e in [-2,2]
// The sum of squares uses a fixed order: lane l of 32 accumulates columns
// l, l+32, ... sequentially with fmaf, then a 16/8/4/2/1 xor-butterfly.
#pragma once
#include <math.h>
#include <stdint.h>
#include "pcv_common.cuh"

namespace pcv {

__host__ __device__ __forceinline__ uint64_t mix64(uint64_t z) {
  z ^= z >> 30; z *= 0xbf58476d1ce4e5b9ull;
  z ^= z >> 27; z *= 0x94d049bb133111ebull;
  z ^= z >> 31;
  return z;
}
__host__ __device__ __forceinline__ uint64_t synth_hash(uint64_t seed, uint64_t row, uint32_t col) {
  return mix64((seed * 0x9e3779b97f4a7c15ull) ^ mix64(row * 0xd1b54a32d192ed03ull + (uint64_t)col + 1ull));
}
__host__ __device__ __forceinline__ float synth_gauss(uint64_t seed, uint64_t row, uint32_t col) {
  const uint64_t h = synth_hash(seed, row, col);
  const int32_t s = (int32_t)((h & 0xffffu) + ((h >> 16) & 0xffffu) + ((h >> 32) & 0xffffu) + (h >> 48));
  // Irwin-Hall(4) of 16-bit uniforms: mean 131070, sd 65536/sqrt(3)
  return (float)(s - 131070) * 2.64290613e-5f;
}
__host__ __device__ __forceinline__ float synth_row_scale(uint64_t seed, uint64_t row) {
  const uint64_t h = synth_hash(seed ^ 0x5ca1ab1e0ddba11ull, row, 0xffffffffu);
  const int e = (int)(h % 5ull) - 2;
  const float mant = 1.0f + (float)((h >> 40) & 0xffffu) * (1.0f / 65536.0f);
  return ldexpf(mant, e);
}

// host reference of one row in the exact device order (used by
// pcv_synthetic_rows_host; the oracle carries its own restatement)
inline void synth_row_host(uint64_t seed, int dist, uint64_t row, uint32_t dim, float* out) {
  float part[32];
  for (int l = 0; l < 32; ++l) part[l] = 0.0f;
  for (uint32_t c = 0; c < dim; ++c) {
    const float g = synth_gauss(seed, row, c);
    out[c] = g;
    part[c & 31] = fmaf(g, g, part[c & 31]);
  }
  if (dist == 0) {
    for (int off = 16; off >= 1; off >>= 1) {
      float nxt[32];
      for (int l = 0; l < 32; ++l) nxt[l] = part[l] + part[l ^ off];
      for (int l = 0; l < 32; ++l) part[l] = nxt[l];
    }
    const float nrm = fmaxf(sqrtf(part[0]), 1e-12f);
    for (uint32_t c = 0; c < dim; ++c) out[c] = out[c] / nrm;
  } else {
    const float s = synth_row_scale(seed, row);
    for (uint32_t c = 0; c < dim; ++c) out[c] = out[c] * s;
  }
}

__host__ __device__ __forceinline__ uint16_t f32_to_bf16_rne(float f) {
#ifdef __CUDA_ARCH__
  uint32_t b = __float_as_uint(f);
#else
  union { float f; uint32_t u; } cv; cv.f = f; uint32_t b = cv.u;
#endif
  if ((b & 0x7fffffffu) > 0x7f800000u) return (uint16_t)((b >> 16) | 0x40u);  // quiet NaN
  b += 0x7fffu + ((b >> 16) & 1u);
  return (uint16_t)(b >> 16);
}
__host__ __device__ __forceinline__ float bf16_to_f32(uint16_t h) {
  const uint32_t b = (uint32_t)h << 16;
#ifdef __CUDA_ARCH__
  return __uint_as_float(b);
#else
  union { float f; uint32_t u; } cv; cv.u = b; return cv.f;
#endif
}

#ifdef __CUDACC__
// One warp per row.  Writes row `first_row + r` of the corpus into the stored
// layout (dim_padded elements per row, zero padded), fp32 or bf16.
template <typename T>
__global__ void synth_rows_kernel(T* __restrict__ rows, T* __restrict__ rows_lo, uint64_t n, uint32_t dim, uint32_t dim_padded,
                                  uint64_t seed, int dist, uint64_t first_row, unsigned int* __restrict__ stats) {
  const int lane = threadIdx.x & 31;
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint64_t nwarps = ((uint64_t)gridDim.x * blockDim.x) >> 5;
  for (uint64_t r = warp; r < n; r += nwarps) {
    const uint64_t row = first_row + r;
    float part = 0.0f;
    for (uint32_t c = lane; c < dim; c += 32) {
      const float g = synth_gauss(seed, row, c);
      part = fmaf(g, g, part);
    }
    float mul = 1.0f, div = 1.0f;
    if (dist == 0) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) part = part + __shfl_xor_sync(PCV_FULL_MASK, part, off);
      div = fmaxf(sqrtf(part), 1e-12f);
    } else {
      mul = synth_row_scale(seed, row);
    }
    T* out = rows + r * (uint64_t)dim_padded;
    float xx = 0.0f, ee = 0.0f;
    for (uint32_t c = lane; c < dim_padded; c += 32) {
      float x = 0.0f;
      if (c < dim) {
        const float g = synth_gauss(seed, row, c);
        x = (dist == 0) ? (g / div) : (g * mul);
      }
      if constexpr (sizeof(T) == 4) {
        out[c] = x;
      } else if (rows_lo) {  // PCV_F32_SPLIT: hi = top 16 bits, lo = low 16 bits (pcv_load.cuh)
        const uint32_t bits = __float_as_uint(x);
        const uint32_t h = split_hi_bits(bits);
        out[c] = (uint16_t)h;
        rows_lo[r * (uint64_t)dim_padded + c] = (uint16_t)bits;
        const float e = x - __uint_as_float(h << 16);
        xx = fmaf(x, x, xx);
        ee = fmaf(e, e, ee);
      } else {
        out[c] = f32_to_bf16_rne(x);
      }
    }
    if (stats && rows_lo) {
#pragma unroll
      for (int off = 16; off >= 1; off >>= 1) {
        xx += __shfl_xor_sync(PCV_FULL_MASK, xx, off);
        ee += __shfl_xor_sync(PCV_FULL_MASK, ee, off);
      }
      if (lane == 0) {
        atomicMax(stats, __float_as_uint(xx));
        atomicMax(stats + 1, __float_as_uint(ee));
      }
    }
  }
}
#endif

}  // namespace pcv
