// pcv_gemm_launch.cuh — host interface of K2, the tcgen05/TMEM batched search
// (bf16 rows x bf16 queries, fp32 accumulate, top-k fused in the epilogue).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcv {

struct GemmWorkspace {
  void* d_q_bf16 = nullptr;      // queries converted to bf16 (padded to the M tile)
  size_t q_cap = 0;
  uint64_t* d_partial = nullptr; // per-CTA candidate lists
  size_t partial_cap = 0;
  void* tmaps = nullptr;         // cached tensor maps (host)
  bool tmaps_valid = false;
  void invalidate() { tmaps_valid = false; }
  void release();
};

struct GemmCall {
  const uint8_t* rows;
  uint64_t n_rows;
  uint32_t dim_padded, dim;
  const uint2* ranges;  // host
  uint32_t n_ranges;
  const float* queries;  // device fp32, padded rows of dim_padded
  uint32_t n_queries, k;
  bool cosine;
  uint32_t emit_mode;
  const uint32_t* lrank_of_row;
  const uint32_t* row_of_lrank;
  const int64_t* ids;
  int64_t id_base;
  int64_t* out_ids;
  float* out_scores;
  float* out_sims;
  uint32_t* out_counts;
  int sm_count;
  cudaStream_t stream;
};

bool gemm_path_applicable(bool bf16_rows, bool cosine, uint32_t dim_padded, uint32_t n_queries, uint32_t k,
                          uint64_t selected_rows);
cudaError_t gemm_search(GemmWorkspace& ws, const GemmCall& call, uint32_t* launches);

}  // namespace pcv
