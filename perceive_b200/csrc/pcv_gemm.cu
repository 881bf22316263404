// pcv_gemm.cu — K2: tcgen05/TMEM batched search.  (Round-1 placeholder: the
// dispatcher reports the path as not applicable, so large batches run as
// repeated K1 scans until the tensor-core kernel lands.)
#include "pcv_gemm_launch.cuh"

namespace pcv {

void GemmWorkspace::release() {
  if (d_q_bf16) cudaFree(d_q_bf16);
  if (d_partial) cudaFree(d_partial);
  d_q_bf16 = nullptr;
  d_partial = nullptr;
  q_cap = partial_cap = 0;
  tmaps_valid = false;
}

bool gemm_path_applicable(bool, bool, uint32_t, uint32_t, uint32_t, uint64_t) { return false; }

cudaError_t gemm_search(GemmWorkspace&, const GemmCall&, uint32_t* launches) {
  if (launches) *launches = 0;
  return cudaErrorNotSupported;
}

}  // namespace pcv
