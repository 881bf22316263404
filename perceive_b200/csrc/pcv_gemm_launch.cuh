// pcv_gemm_launch.cuh — host interface of K2, the tcgen05/TMEM batched search
// (bf16 rows x bf16 queries, fp32 accumulate, top-k selection fused into the
// epilogue so the B x N score matrix never exists in memory).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>

namespace pcv {

struct GemmWorkspace {
  uint8_t* d_q_bf16 = nullptr;  // queries as bf16, padded to a multiple of 128 rows
  size_t q_cap = 0;          // bytes
  uint64_t* d_cand = nullptr;  // per (CTA, query-tile segment, query) candidate buffers
  size_t cand_cap = 0;         // keys
  uint32_t* d_cand_cnt = nullptr;
  size_t cnt_cap = 0;          // entries
  bool cnt_clean = false;      // every counter is zero (the select kernel resets what it consumes)
  uint64_t* d_topk = nullptr;  // running top-k keys per query (between passes)
  float* d_thr = nullptr;      // running k-th similarity per query
  size_t topk_cap = 0;         // keys
  size_t thr_cap = 0;
  float* d_qinv = nullptr;     // cosine: 1/|query|
  size_t qinv_cap = 0;
  float* d_boot = nullptr;     // bootstrap pass: per (document tile, query) maximum score
  size_t boot_cap = 0;
  void release();
};

struct GemmCall {
  const uint8_t* rows;  // bf16 matrix (a bf16 index, or the hi plane of a PCV_F32_SPLIT index)
  uint64_t row_bytes;   // row pitch
  bool keys_only;       // filter mode (K3): nothing is emitted; the k best ranking keys (tensor-core score,
                        // local rank) of every query are left in GemmWorkspace::d_topk, sorted descending
  uint64_t n_rows;
  uint32_t dim_padded, dim;
  const uint2* d_ranges;          // device: selected row ranges
  const uint32_t* d_range_prefix; // device: 128-row tiles before range r
  uint32_t n_ranges;
  uint32_t total_tiles;           // tiles of gemm_tile_rows(dim_padded) rows over all ranges
  const float* queries;           // device fp32 [n_queries][dim_padded]; rounded to bf16 (RNE) on the way in
  uint32_t n_queries, k;
  bool cosine;                    // PCV_METRIC_COSINE: dot / (|q| |row|)
  const float* x_inv_norm;        // cosine: device, 1/|row| per stored row, padded by one tile
  uint32_t emit_mode;             // 0 final results, 1 (sim,id) candidates
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

// document rows per tile (UMMA N): 128 for rows up to 384-d, 96 for wider rows
uint32_t gemm_tile_rows(uint32_t dim_padded);
// 1/|row| of every stored bf16 row (n_out >= n_rows entries; the tail is zero padding)
cudaError_t gemm_row_inv_norms(const uint8_t* rows, uint64_t n_rows, uint32_t dim_padded, float* out, uint64_t n_out,
                               int sm_count, cudaStream_t stream);

bool gemm_path_applicable(bool cosine, uint32_t dim_padded, uint32_t n_queries, uint32_t k, uint64_t selected_rows,
                          uint64_t n_rows);
// nullptr on success, else a static description of what failed (with *err set)
const char* gemm_search(GemmWorkspace& ws, const GemmCall& call, uint32_t* launches, cudaError_t* err);

}  // namespace pcv
