// pcv_scan_inst.cu — explicit instantiations of the K1 scan kernel for one
// (storage type, metric) pair.  Compile with
//   -DPCV_T=float|uint16_t|SplitF32 -DPCV_COS=false|true -DPCV_TAG=f32_dot|... [-DPCV_GROUPED]
#include <atomic>

#include "pcv_scan_launch.cuh"

#ifndef PCV_T
#error "PCV_T / PCV_COS / PCV_TAG must be defined"
#endif

namespace pcv {
namespace {

template <int NJ, int NB, int KPL, bool GROUPED>
cudaError_t launch(const ScanParams& p, int grid, size_t smem, cudaStream_t st) {
  auto kern = scan_kernel<PCV_T, NJ, NB, KPL, PCV_COS, GROUPED>;
  static std::atomic<unsigned long long> attr_done{0ull};  // per-device bitmask (searches may come from several threads)
  int dev = 0;
  cudaGetDevice(&dev);
  if (!((attr_done.load(std::memory_order_acquire) >> (dev & 63)) & 1ull)) {
    cudaError_t e = cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize, 232448 - 2048);
    if (e != cudaSuccess) return e;
    attr_done.fetch_or(1ull << (dev & 63), std::memory_order_release);
  }
  kern<<<grid, SCAN_THREADS, smem, st>>>(p);
  return cudaGetLastError();
}

using pcv::SplitF32;
constexpr bool qsmem(int nj, int nb) { return nb * nj * Chunk<PCV_T>::EPC > 96; }

#define V(NJ, NB, KPL) {NJ, NB, KPL, false, qsmem(NJ, NB), &launch<NJ, NB, KPL, false>}
#define VG(NJ, NB, KPL) {NJ, NB, KPL, true, qsmem(NJ, NB), &launch<NJ, NB, KPL, true>}
const ScanVariant kVariants[] = {
    V(6, 1, 1),  V(6, 1, 4),  V(6, 1, 32),  V(6, 2, 1),  V(6, 2, 4),  V(6, 4, 1),  V(6, 4, 4),
    V(12, 1, 1), V(12, 1, 4), V(12, 1, 32), V(12, 2, 1), V(12, 2, 4), V(12, 4, 1), V(12, 4, 4),
#ifdef PCV_GROUPED
    // one launch walks a device-resident list of queries, four at a time (batches over fp32 rows; the
    // queries a PCV_F32_SPLIT filter could not prove complete)
    VG(6, 4, 1), VG(6, 4, 4), VG(12, 4, 1), VG(12, 4, 4),
#endif
};
#undef V
#undef VG

}  // namespace

#define PCV_CAT2(a, b) a##b
#define PCV_CAT(a, b) PCV_CAT2(a, b)
const ScanVariant* PCV_CAT(scan_lookup_, PCV_TAG)(int nj, int nb, int kpl, bool grouped) {
  for (const ScanVariant& v : kVariants)
    if (v.nj == nj && v.nb == nb && v.kpl == kpl && v.grouped == grouped) return &v;
  return nullptr;
}

}  // namespace pcv
