// pcv_scan_launch.cuh — host-side launch table for the K1 scan variants.
// Each (storage type, metric) pair is instantiated in its own translation unit
// (pcv_scan_inst.cu compiled with -DPCV_T / -DPCV_COS / -DPCV_TAG) so the build
// can run them in parallel.
#pragma once
#include <cuda_runtime.h>
#include "pcv_scan.cuh"

namespace pcv {

typedef cudaError_t (*scan_launch_fn)(const ScanParams& p, int grid, size_t smem, cudaStream_t st);

struct ScanVariant {
  int nj, nb, kpl;
  bool grouped;
  bool q_in_smem;
  scan_launch_fn fn;
};

// nullptr when the combination is not instantiated
const ScanVariant* scan_lookup_f32_dot(int nj, int nb, int kpl, bool grouped);
const ScanVariant* scan_lookup_f32_cos(int nj, int nb, int kpl, bool grouped);
const ScanVariant* scan_lookup_bf16_dot(int nj, int nb, int kpl, bool grouped);
const ScanVariant* scan_lookup_bf16_cos(int nj, int nb, int kpl, bool grouped);
const ScanVariant* scan_lookup_split_dot(int nj, int nb, int kpl, bool grouped);

}  // namespace pcv
