// pcv_api.cu — C ABI of libperceive_cuda (see include/perceive_cuda.h).
//
// Host side of the device-resident exact search that replaces
// perceive_core::search::Searcher's index (crates/perceive-core/search.rs).
// No torch, no CPU fallback: every compute entry point needs a CUDA device.
#include <cuda_runtime.h>
#include <dlfcn.h>
#include <nccl.h>  // types only: the library is resolved lazily with dlopen (see NcclApi)
#include <nvtx3/nvToolsExt.h>  // header-only NVTX v3: ranges cost a no-op call unless a profiler is attached

#include <algorithm>
#include <atomic>
#include <chrono>
#include <cmath>
#include <condition_variable>
#include <cstdarg>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <exception>
#include <functional>
#include <mutex>
#include <new>
#include <numeric>
#include <string>
#include <thread>
#include <vector>

#include "../../include/perceive_cuda.h"
#include "pcv_common.cuh"
#include "pcv_gemm_launch.cuh"
#include "pcv_load.cuh"
#include "pcv_rescore.cuh"
#include "pcv_scan_launch.cuh"
#include "pcv_synth.cuh"

namespace {

thread_local std::string g_err;

// NVTX range for the life of a scope: load / search / exchange show up by name on an Nsight timeline
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

int32_t fail(int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_err = buf;
  return code;
}

#define CU(call)                                                                          \
  do {                                                                                    \
    cudaError_t e__ = (call);                                                             \
    if (e__ != cudaSuccess)                                                               \
      return fail(e__ == cudaErrorMemoryAllocation ? PCV_ERR_OOM : PCV_ERR_CUDA,          \
                  "%s failed: %s (%s:%d)", #call, cudaGetErrorString(e__), __FILE__, __LINE__); \
  } while (0)

// NCCL is bound at run time, on the first communicator call.  A link-time
// dependency would pin whichever libnccl.so.2 the loader finds first and break
// hosts that later load a different build under the same SONAME (PyTorch
// bundles its own); dlopen by SONAME reuses the copy already in the process.
struct NcclApi {
  ncclResult_t (*GetUniqueId)(ncclUniqueId*) = nullptr;
  ncclResult_t (*CommInitRank)(ncclComm_t*, int, ncclUniqueId, int) = nullptr;
  ncclResult_t (*AllGather)(const void*, void*, size_t, ncclDataType_t, ncclComm_t, cudaStream_t) = nullptr;
  ncclResult_t (*CommDestroy)(ncclComm_t) = nullptr;
  const char* (*GetErrorString)(ncclResult_t) = nullptr;
  bool ok = false;
  std::string err;
};
NcclApi& nccl_api() {
  static NcclApi api;
  static std::once_flag once;
  std::call_once(once, [] {
    void* h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
    if (!h) h = dlopen("libnccl.so", RTLD_NOW | RTLD_GLOBAL);
    if (!h) { api.err = std::string("cannot load libnccl.so.2: ") + dlerror(); return; }
    api.GetUniqueId = (decltype(api.GetUniqueId))dlsym(h, "ncclGetUniqueId");
    api.CommInitRank = (decltype(api.CommInitRank))dlsym(h, "ncclCommInitRank");
    api.AllGather = (decltype(api.AllGather))dlsym(h, "ncclAllGather");
    api.CommDestroy = (decltype(api.CommDestroy))dlsym(h, "ncclCommDestroy");
    api.GetErrorString = (decltype(api.GetErrorString))dlsym(h, "ncclGetErrorString");
    api.ok = api.GetUniqueId && api.CommInitRank && api.AllGather && api.CommDestroy && api.GetErrorString;
    if (!api.ok) api.err = "libnccl.so.2 lacks a required symbol";
  });
  return api;
}

#define NC(call)                                                                      \
  do {                                                                                \
    ncclResult_t r__ = (call);                                                        \
    if (r__ != ncclSuccess)                                                           \
      return fail(PCV_ERR_NCCL, "%s failed: %s (%s:%d)", #call, nccl_api().GetErrorString(r__), \
                  __FILE__, __LINE__);                                                \
  } while (0)

struct Segment {
  int64_t source_id;
  uint64_t begin, end;  // rows [begin, end)
};

template <typename T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;  // elements
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 16;
    cudaError_t e = cudaMalloc((void**)&p, want * sizeof(T));
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFree(p);
    p = nullptr;
    cap = 0;
  }
};

struct PinBuf {
  uint8_t* p = nullptr;
  size_t cap = 0;
  cudaError_t reserve(size_t n) {
    if (n <= cap) return cudaSuccess;
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
    size_t want = n + n / 4 + 256;
    cudaError_t e = cudaMallocHost((void**)&p, want);
    if (e == cudaSuccess) cap = want;
    return e;
  }
  void release() {
    if (p) cudaFreeHost(p);
    p = nullptr;
    cap = 0;
  }
};

// debugging / A-B switches: set to anything but "0" to take effect
bool env_flag(const char* name) {
  const char* e = getenv(name);
  return e && *e && strcmp(e, "0") != 0;
}

size_t elem_size(pcv_dtype t) { return t == PCV_BF16 ? 2 : 4; }  // PCV_F32_SPLIT: two 16-bit planes

// words of the small device control block every index owns (d_ctl)
enum : uint32_t {
  CTL_SCAN_DONE = 0,    // [SCAN_MAX_GROUPS] last-CTA counters of the scan kernel (self-resetting)
  CTL_P2P_DONE = 64,    // last-CTA counter of the peer exchange kernel
  CTL_LOAD_FLAGS = 65,  // PCV_LOADFLAG_* raised by the load kernels
  CTL_XMAX2 = 66,       // PCV_F32_SPLIT: max |x|^2 over the stored rows (float bits)
  CTL_EMAX2 = 67,       //                max |x - hi(x)|^2
  CTL_FB_COUNT = 68,    // queries the last split search handed to the exact fallback scan
  CTL_WORDS = 96
};
static_assert(pcv::SCAN_MAX_GROUPS <= 64, "control block layout");

}  // namespace

// One worker thread per shard of a single-process many-GPU handle: each is bound to its device once, and a search
// hands every shard's share of a phase to its worker at the same time instead of walking the devices from the calling
// thread (8 GPUs: ~100 launches per batched search, one after another, were 0.26 ms of a 2.5 ms batch; a single
// query paid ~100 us of serial launches for a 60 us scan).  Workers — and the caller waiting for them — spin for up to
// 300 us before they sleep on a condition variable: back-to-back phases and searches find everybody awake.
struct ShardPool {
  std::vector<std::thread> threads;
  std::mutex m;
  std::condition_variable cv_work, cv_done;
  std::atomic<uint64_t> ticket{0};
  std::atomic<int> pending{0};
  std::atomic<bool> quit{false};
  const std::function<int32_t(int)>* task = nullptr;
  std::vector<int32_t> rc;
  std::vector<std::string> err;
};

struct pcv_index {
  int device = 0;
  uint32_t dim = 0, dim_padded = 0;
  pcv_dtype store = PCV_F32;
  pcv_metric metric = PCV_METRIC_DOT_REF;
  uint32_t flags = 0;
  std::mutex mu;
  cudaStream_t own_stream = nullptr, stream = nullptr;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr;
  bool ev_valid = false;
  bool timing = true;     // searches bracketed by ev0 / ev1 (PCV_FLAG_NO_TIMING clears it)
  bool searched = false;  // a search has been enqueued since the rows last changed
  int sm_count = 0;

  // resident matrix
  uint8_t* d_rows = nullptr;
  uint64_t n_rows = 0;
  size_t row_bytes = 0;
  int64_t* d_ids = nullptr;  // null: id = id_base + row
  int64_t id_base = 1;
  uint32_t* d_lrank_of_row = nullptr;
  uint32_t* d_row_of_lrank = nullptr;
  std::vector<int64_t> h_ids;  // per row (sorted order); empty when dense
  std::vector<Segment> segs;
  // rows excluded from every search (pcv_index_set_hidden): cut out of the row ranges on the host
  std::vector<int64_t> hidden_ids;    // sorted, unique
  std::vector<uint32_t> hidden_rows;  // the ones resident on this shard, sorted
  bool hidden_dirty = false;          // hidden_rows must be recomputed (ids or matrix changed)

  // workspace
  DevBuf<uint64_t> partial;
  unsigned int* d_done = nullptr;  // control block, CTL_WORDS words
  DevBuf<float> q_pad;
  // selected row ranges + tile prefix as the kernels read them; two cached sets because the scan and
  // the tensor path tile the same ranges differently (a split search uses both in one call)
  struct RangeSet {
    DevBuf<uint32_t> range_prefix;
    DevBuf<uint2> ranges;
    std::vector<uint32_t> h_range_prefix;  // what is currently uploaded
    std::vector<uint2> h_ranges;
    uint32_t tile_rows = 0;
    uint32_t total_tiles = 0;
  } rs[2];  // [0] scan tiling, [1] tensor-path tiling
  DevBuf<float> margin;     // split filter: per-query margin
  DevBuf<uint32_t> fb_list; // split filter: queries for the exact fallback scan
  DevBuf<uint8_t> o_pack;  // [ids | scores | sims | counts] of the last host-buffer search
  DevBuf<float> q_in;
  PinBuf pin;
  unsigned int* d_flags = nullptr;

  // K2 (tcgen05 GEMM) workspace
  pcv::GemmWorkspace gemm;
  float* d_xinv = nullptr;  // cosine on the tensor path: 1/|row|, computed on the device when first needed

  // multi-GPU
  ncclComm_t comm = nullptr;
  int rank = 0, world = 1;
  DevBuf<uint8_t> cand_send, cand_recv;
  // peer-memory exchange (no NCCL): receive buffers of every rank mapped through CUDA IPC
  uint8_t* p2p_local = nullptr;
  uint8_t* p2p_peer[PCV_P2P_MAX_WORLD] = {};
  uint32_t p2p_world = 0, p2p_cap = 0;
  bool p2p_attached = false;
  bool p2p_in_process = false;  // peers are shards of the same process (peer access, no IPC handles)
  uint32_t p2p_epoch = 0;
  bool shard_failed = false;    // a collective search failed part-way: out of step with the peers
  bool fused_exchange = false;  // the search being enqueued delivers its candidates from the scan kernel itself
  // single-process, many-GPU handle (pcv_index_create_multi): this object is only the front; the rows live in
  // one ordinary one-device shard per GPU, searched together
  std::vector<pcv_index*> shards;
  std::vector<uint64_t> shard_row0;  // first global row of each shard (+ total at the end)
  ShardPool* pool = nullptr;         // one worker thread per shard (n_shards > 1)
  // what the shards' scans were last prepared for (scan_prepare): a repeated single-query search skips that phase
  uint64_t rows_epoch = 1, prep_epoch = 0;
  uint32_t prep_k = 0;
  bool prep_all = false;
  std::vector<int64_t> prep_sources;

  // stats
  uint64_t last_scan_bytes = 0;
  uint32_t last_launches = 0;
  uint32_t last_kernel = 0;
  bool last_used_filter = false;

  // PCV_F32_SPLIT: the two planes of the resident matrix
  uint8_t* hi_plane() const { return d_rows; }
  uint8_t* lo_plane() const { return d_rows + n_rows * (row_bytes / 2); }
};

namespace {

void free_matrix(pcv_index* ix) {
  if (ix->d_rows) cudaFree(ix->d_rows);
  if (ix->d_ids) cudaFree(ix->d_ids);
  if (ix->d_lrank_of_row) cudaFree(ix->d_lrank_of_row);
  if (ix->d_row_of_lrank) cudaFree(ix->d_row_of_lrank);
  if (ix->d_xinv) cudaFree(ix->d_xinv);
  ix->d_xinv = nullptr;
  ix->d_rows = nullptr;
  ix->d_ids = nullptr;
  ix->d_lrank_of_row = nullptr;
  ix->d_row_of_lrank = nullptr;
  ix->n_rows = 0;
  ix->h_ids.clear();
  ix->segs.clear();
  ix->hidden_rows.clear();
  ix->hidden_dirty = true;
  for (auto& r : ix->rs) { r.h_ranges.clear(); r.h_range_prefix.clear(); }
  if (ix->d_done) cudaMemsetAsync(ix->d_done + CTL_XMAX2, 0, 2 * sizeof(unsigned int), ix->stream);
}

// Upload `n` fp32 rows (host, gathered through perm when given) into rows [row_off, row_off + n) of the
// matrix at d_base (stored layout; `alloc_rows` rows in all, which places the lo plane of a split
// matrix) via double-buffered pinned staging + the load kernel.
int32_t upload_rows(pcv_index* ix, const float* rows, const uint64_t* perm, uint64_t n, uint8_t* d_base,
                    uint64_t row_off, uint64_t alloc_rows) {
  if (n == 0) return PCV_OK;
  NvtxRange nvtx("pcv:load_rows (pinned staging -> H2D -> validate/normalise/convert)");
  const uint32_t dim = ix->dim;
  const size_t in_row = (size_t)dim * 4;
  const uint64_t chunk_rows = std::max<uint64_t>(1, std::min<uint64_t>(n, (32u << 20) / in_row));
  uint8_t* pinned[2] = {nullptr, nullptr};
  float* d_stage[2] = {nullptr, nullptr};
  cudaEvent_t done[2] = {nullptr, nullptr};
  int32_t rc = PCV_OK;
  auto cleanup = [&]() {
    for (int i = 0; i < 2; ++i) {
      if (pinned[i]) cudaFreeHost(pinned[i]);
      if (d_stage[i]) cudaFree(d_stage[i]);
      if (done[i]) cudaEventDestroy(done[i]);
    }
  };
  for (int i = 0; i < 2 && rc == PCV_OK; ++i) {
    if (cudaMallocHost((void**)&pinned[i], chunk_rows * in_row) != cudaSuccess ||
        cudaMalloc((void**)&d_stage[i], chunk_rows * in_row) != cudaSuccess ||
        cudaEventCreateWithFlags(&done[i], cudaEventDisableTiming) != cudaSuccess)
      rc = fail(PCV_ERR_OOM, "staging allocation failed: %s", cudaGetErrorString(cudaGetLastError()));
  }
  if (rc != PCV_OK) { cleanup(); return rc; }
  const int normalise = (ix->flags & PCV_FLAG_PRENORMALISE) ? 1 : 0;
  const int check_zero = (ix->metric == PCV_METRIC_COSINE) ? 1 : 0;
  int buf = 0;
  for (uint64_t r0 = 0; r0 < n; r0 += chunk_rows, buf ^= 1) {
    const uint64_t nr = std::min(chunk_rows, n - r0);
    cudaEventSynchronize(done[buf]);
    if (perm) {
      for (uint64_t r = 0; r < nr; ++r)
        memcpy(pinned[buf] + r * in_row, rows + perm[r0 + r] * (size_t)dim, in_row);
    } else {
      memcpy(pinned[buf], rows + r0 * (size_t)dim, nr * in_row);
    }
    cudaError_t e = cudaMemcpyAsync(d_stage[buf], pinned[buf], nr * in_row, cudaMemcpyHostToDevice, ix->stream);
    if (e != cudaSuccess) { rc = fail(PCV_ERR_CUDA, "H2D failed: %s", cudaGetErrorString(e)); break; }
    const int threads = 256;
    const int blocks = (int)std::min<uint64_t>((nr + 7) / 8, (uint64_t)ix->sm_count * 8);
    if (ix->store == PCV_F32) {
      uint8_t* dst = d_base + (row_off + r0) * ix->row_bytes;
      pcv::load_rows_kernel<float><<<blocks, threads, 0, ix->stream>>>(d_stage[buf], (float*)dst, nullptr, nr, dim, ix->dim_padded, normalise, check_zero, ix->d_flags, nullptr);
    } else {
      const size_t prb = (size_t)ix->dim_padded * 2;  // bytes of one row of one 16-bit plane
      uint8_t* dst = d_base + (row_off + r0) * prb;
      uint8_t* dst_lo = ix->store == PCV_F32_SPLIT ? d_base + (alloc_rows + row_off + r0) * prb : nullptr;
      pcv::load_rows_kernel<uint16_t><<<blocks, threads, 0, ix->stream>>>(d_stage[buf], (uint16_t*)dst, (uint16_t*)dst_lo, nr, dim, ix->dim_padded, normalise, check_zero, ix->d_flags, ix->d_done + CTL_XMAX2);
    }
    e = cudaGetLastError();
    if (e != cudaSuccess) { rc = fail(PCV_ERR_CUDA, "load kernel launch failed: %s", cudaGetErrorString(e)); break; }
    cudaEventRecord(done[buf], ix->stream);
  }
  cudaError_t e = cudaStreamSynchronize(ix->stream);
  if (rc == PCV_OK && e != cudaSuccess) rc = fail(PCV_ERR_CUDA, "load failed: %s", cudaGetErrorString(e));
  cleanup();
  return rc;
}

int32_t check_load_flags(pcv_index* ix) {
  unsigned int f = 0;
  CU(cudaMemcpy(&f, ix->d_flags, sizeof f, cudaMemcpyDeviceToHost));
  CU(cudaMemset(ix->d_flags, 0, sizeof f));
  if (f & PCV_LOADFLAG_NONFINITE) return fail(PCV_ERR_NONFINITE, "non-finite value in document rows");
  if (f & PCV_LOADFLAG_ZERONORM) return fail(PCV_ERR_ZERO_NORM, "document row with a zero (or unrepresentable) norm under the cosine metric");
  return PCV_OK;
}

// Build the id->rank tables for rows whose ids are `ids` (row order).  Rows are ordered by (source,
// id); the ranking key needs an order by id alone, which differs as soon as two sources interleave their
// ids.  Both outputs stay null when the ids already ascend (identity).  Nothing of the index is touched.
int32_t make_rank_tables(const std::vector<int64_t>& ids, uint32_t** d_lrank_of_row, uint32_t** d_row_of_lrank) {
  *d_lrank_of_row = *d_row_of_lrank = nullptr;
  const uint64_t n = ids.size();
  if (n == 0 || std::is_sorted(ids.begin(), ids.end())) return PCV_OK;
  std::vector<uint32_t> row_of(n), lrank_of(n);
  std::iota(row_of.begin(), row_of.end(), 0u);
  const int64_t* idp = ids.data();
  std::stable_sort(row_of.begin(), row_of.end(), [idp](uint32_t a, uint32_t b) { return idp[a] < idp[b]; });
  for (uint64_t r = 0; r < n; ++r) lrank_of[row_of[r]] = (uint32_t)r;
  cudaError_t e = cudaMalloc((void**)d_lrank_of_row, n * 4);
  if (e == cudaSuccess) e = cudaMalloc((void**)d_row_of_lrank, n * 4);
  if (e == cudaSuccess) e = cudaMemcpy(*d_lrank_of_row, lrank_of.data(), n * 4, cudaMemcpyHostToDevice);
  if (e == cudaSuccess) e = cudaMemcpy(*d_row_of_lrank, row_of.data(), n * 4, cudaMemcpyHostToDevice);
  if (e != cudaSuccess) {
    if (*d_lrank_of_row) cudaFree(*d_lrank_of_row);
    if (*d_row_of_lrank) cudaFree(*d_row_of_lrank);
    *d_lrank_of_row = *d_row_of_lrank = nullptr;
    return fail(e == cudaErrorMemoryAllocation ? PCV_ERR_OOM : PCV_ERR_CUDA, "rank table upload failed: %s", cudaGetErrorString(e));
  }
  return PCV_OK;
}

struct ScanPlan {
  uint32_t lpr_log2, nj, tile_iters, tile_rows, slot_bytes, nslots;
};

int32_t plan_scan(const pcv_index* ix, ScanPlan& pl) {
  const uint32_t d_chunks = (uint32_t)(ix->row_bytes / 16);
  uint32_t lpr_log2 = 3;
  if (const char* e = getenv("PCV_SCAN_LPR_LOG2")) lpr_log2 = std::min(5u, std::max(3u, (uint32_t)atoi(e)));
  while (lpr_log2 < 5 && (d_chunks + (1u << lpr_log2) - 1) / (1u << lpr_log2) > 12) ++lpr_log2;
  const uint32_t per_lane = (d_chunks + (1u << lpr_log2) - 1) >> lpr_log2;
  if (per_lane > 12)
    return fail(PCV_ERR_UNSUPPORTED, "dimension %u too large for the scan kernel (max %u bytes per row)", ix->dim, 12u * 32u * 16u);
  pl.lpr_log2 = lpr_log2;
  pl.nj = per_lane <= 6 ? 6 : 12;
  const uint32_t rpi = 32u >> lpr_log2;
  // two planes = two bulk copies per tile: 12 KB tiles keep each copy at the 6 KB the fp32 scan was tuned to
  // (measured on config 2 over split rows: 0.97 of the HBM peak against 0.87 with 6 KB tiles)
  uint32_t tile_bytes = ix->store == PCV_F32_SPLIT ? 12288 : 6144;
  if (const char* e = getenv("PCV_SCAN_TILE_BYTES")) tile_bytes = (uint32_t)atoi(e);
  uint32_t iters = std::max<uint32_t>(1, (uint32_t)(tile_bytes / (rpi * ix->row_bytes)));
  // dynamic shared memory: 8 private rings + mbarriers + (for NB=4) the query block
  const size_t budget = (size_t)(232448 - 2048 - pcv::SCAN_WARPS * pcv::SCAN_MAX_SLOTS * 8 - 4 * (size_t)ix->dim_padded * 4) / pcv::SCAN_WARPS;
  while (iters > 1 && (size_t)iters * rpi * ix->row_bytes * 2 > budget) --iters;
  pl.tile_iters = iters;
  pl.tile_rows = iters * rpi;
  pl.slot_bytes = (uint32_t)(pl.tile_rows * ix->row_bytes);
  // Ring depth: measured on B200 (tools/tune_scan.py, profiles/r1_scan_tuning.md) a
  // shallow ring wins — 2 slots x 6 KB x 8 warps = 96 KB in flight per SM already
  // covers the HBM latency-bandwidth product, and deeper rings widen the window
  // of DRAM pages open at once (6.28 TB/s at 2 slots vs 5.84 TB/s at 4).
  const uint32_t max_slots = (uint32_t)std::min<size_t>(pcv::SCAN_MAX_SLOTS, budget / pl.slot_bytes);
  uint32_t nslots = std::min<uint32_t>(max_slots, 2u);
  if (const char* e = getenv("PCV_SCAN_NSLOTS")) nslots = std::min<uint32_t>(max_slots, (uint32_t)atoi(e));
  if (nslots < 1) return fail(PCV_ERR_UNSUPPORTED, "row of %zu bytes does not fit the scan ring", ix->row_bytes);
  pl.nslots = nslots;
  return PCV_OK;
}

// Index of the first NaN/Inf in v[0..n), or n.  Exponent-bits test in blocks without an early
// exit inside the block, so the compiler vectorises it (a 1024 x 384 batch is 393k values).
size_t first_nonfinite(const float* v, size_t n) {
  constexpr size_t BLK = 1024;
  for (size_t b = 0; b < n; b += BLK) {
    const size_t e = std::min(n, b + BLK);
    uint32_t bad = 0;
    for (size_t i = b; i < e; ++i) {
      uint32_t bits;
      memcpy(&bits, v + i, 4);
      bad |= ((bits & 0x7f800000u) == 0x7f800000u) ? 1u : 0u;
    }
    if (bad)
      for (size_t i = b; i < e; ++i)
        if (!std::isfinite(v[i])) return i;
  }
  return n;
}

// Row holding items.id `id`, or -1.  Ids ascend inside each source segment.
int64_t find_row(const pcv_index* ix, int64_t id) {
  if (ix->h_ids.empty()) {
    const int64_t r = id - ix->id_base;
    return (r >= 0 && (uint64_t)r < ix->n_rows) ? r : -1;
  }
  for (const Segment& s : ix->segs) {
    const int64_t* b = ix->h_ids.data() + s.begin;
    const int64_t* e = ix->h_ids.data() + s.end;
    const int64_t* it = std::lower_bound(b, e, id);
    if (it != e && *it == id) return (int64_t)(it - ix->h_ids.data());
  }
  return -1;
}

void resolve_hidden(pcv_index* ix) {
  ix->hidden_rows.clear();
  for (int64_t id : ix->hidden_ids) {
    const int64_t r = find_row(ix, id);
    if (r >= 0) ix->hidden_rows.push_back((uint32_t)r);
  }
  std::sort(ix->hidden_rows.begin(), ix->hidden_rows.end());
  ix->hidden_dirty = false;
}

// Translate the source filter into row ranges + tile prefix (set `which`: 0 scan tiling, 1 tensor-path
// tiling), upload if changed.
int32_t prepare_ranges(pcv_index* ix, const int64_t* sources, uint32_t n_sources, bool all, uint32_t tile_rows, int which) {
  pcv_index::RangeSet& R = ix->rs[which];
  std::vector<uint2> rg;
  for (const Segment& s : ix->segs) {
    bool sel = all;
    if (!sel)
      for (uint32_t i = 0; i < n_sources; ++i)
        if (sources[i] == s.source_id) { sel = true; break; }
    if (!sel || s.end == s.begin) continue;
    if (!rg.empty() && rg.back().y == (uint32_t)s.begin) rg.back().y = (uint32_t)s.end;
    else rg.push_back(make_uint2((uint32_t)s.begin, (uint32_t)s.end));
  }
  if (ix->hidden_dirty) resolve_hidden(ix);
  if (!ix->hidden_rows.empty()) {  // cut the hidden rows out: a kernel never sees them
    std::vector<uint2> cut;
    const std::vector<uint32_t>& hr = ix->hidden_rows;
    for (const uint2& r : rg) {
      uint32_t b = r.x;
      for (auto it = std::lower_bound(hr.begin(), hr.end(), r.x); it != hr.end() && *it < r.y; ++it) {
        if (*it > b) cut.push_back(make_uint2(b, *it));
        b = *it + 1;
      }
      if (b < r.y) cut.push_back(make_uint2(b, r.y));
    }
    rg.swap(cut);
  }
  std::vector<uint32_t> prefix(rg.size() + 1, 0u);
  for (size_t i = 0; i < rg.size(); ++i) {
    const uint64_t tiles = ((uint64_t)(rg[i].y - rg[i].x) + tile_rows - 1) / tile_rows;
    const uint64_t tot = (uint64_t)prefix[i] + tiles;
    if (tot > 0xfffffff0ull) return fail(PCV_ERR_UNSUPPORTED, "too many tiles");
    prefix[i + 1] = (uint32_t)tot;
  }
  if (rg.empty()) rg.push_back(make_uint2(0u, 0u));
  bool same = R.tile_rows == tile_rows && rg.size() == R.h_ranges.size() && prefix == R.h_range_prefix;
  if (same)
    for (size_t i = 0; i < rg.size(); ++i)
      if (rg[i].x != R.h_ranges[i].x || rg[i].y != R.h_ranges[i].y) { same = false; break; }
  R.total_tiles = prefix.back();
  if (same) return PCV_OK;
  CU(R.ranges.reserve(rg.size()));
  CU(R.range_prefix.reserve(prefix.size()));
  // synchronous small copies: filters change rarely (cached otherwise)
  CU(cudaMemcpyAsync(R.ranges.p, rg.data(), rg.size() * sizeof(uint2), cudaMemcpyHostToDevice, ix->stream));
  CU(cudaMemcpyAsync(R.range_prefix.p, prefix.data(), prefix.size() * 4, cudaMemcpyHostToDevice, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  R.h_ranges = rg;
  R.h_range_prefix = prefix;
  R.tile_rows = tile_rows;
  return PCV_OK;
}

const pcv::ScanVariant* lookup_scan(const pcv_index* ix, int nj, int nb, int kpl, bool grouped) {
  const bool cos = ix->metric == PCV_METRIC_COSINE;
  if (ix->store == PCV_F32_SPLIT) return cos ? nullptr : pcv::scan_lookup_split_dot(nj, nb, kpl, grouped);
  if (ix->store == PCV_F32) return cos ? pcv::scan_lookup_f32_cos(nj, nb, kpl, grouped) : pcv::scan_lookup_f32_dot(nj, nb, kpl, grouped);
  return cos ? pcv::scan_lookup_bf16_cos(nj, nb, kpl, grouped) : pcv::scan_lookup_bf16_dot(nj, nb, kpl, grouped);
}

struct SearchOut {
  uint32_t emit_mode;
  int64_t* ids;
  float* scores;
  float* sims;
  uint32_t* counts;
  const pcv::ExchangeTarget* xchg = nullptr;  // emit_mode 2 (when the scan can deliver through it; see enqueue_scan)
};

// K1: exact scan of the selected rows for `n_queries` padded device queries.  With a query list
// (d_q_list / d_q_count, device memory) only the listed queries are searched — their number is not
// known to the host — and results land at the listed positions of the output arrays.
int32_t enqueue_scan(pcv_index* ix, const float* d_q_padded, uint32_t n_queries, uint32_t k, const int64_t* sources,
                     uint32_t n_sources, bool all, uint64_t sel_rows, const SearchOut& o, const uint32_t* d_q_list,
                     const uint32_t* d_q_count) {
  ScanPlan pl;
  int32_t rc = plan_scan(ix, pl);
  if (rc != PCV_OK) return rc;
  rc = prepare_ranges(ix, sources, n_sources, all, pl.tile_rows, 0);
  if (rc != PCV_OK) return rc;
  const pcv_index::RangeSet& R = ix->rs[0];

  const int kpl = k <= 32 ? 1 : (k <= 128 ? 4 : 32);
  // queries per pass: 4 when their slices fit the lane's registers or the rows are fp32 (the
  // shared-memory query path then still halves the passes); bf16 rows carry 8 elements per chunk,
  // 4 queries would spill the slices to shared memory and run LDS-bound (measured 4x slower per
  // byte), so 2 per pass there
  int nb = 1;
  if (n_queries >= 2 && kpl <= 4) nb = (ix->store == PCV_BF16 && pl.nj * 8 * 4 > 96) ? 2 : 4;
  if (d_q_list && kpl <= 4) nb = 4;
  if (const char* e = getenv("PCV_SCAN_NB")) nb = atoi(e);
  // one GROUPED launch walks up to SCAN_MAX_GROUPS groups of 4 queries (fp32 and split rows, k <= 128)
  const pcv::ScanVariant* gvar = (nb == 4 && (n_queries > 4 || d_q_list)) ? lookup_scan(ix, (int)pl.nj, nb, kpl, true) : nullptr;
  if (d_q_list && !gvar) return fail(PCV_ERR_UNSUPPORTED, "no grouped scan variant for nj=%u kpl=%d", pl.nj, kpl);
  const pcv::ScanVariant* var = gvar ? gvar : lookup_scan(ix, (int)pl.nj, nb, kpl, false);
  if (!var) return fail(PCV_ERR_UNSUPPORTED, "no scan variant for nj=%u nb=%d kpl=%d", pl.nj, nb, kpl);

  int grid = (int)std::min<uint64_t>((uint64_t)ix->sm_count, ((uint64_t)R.total_tiles + pcv::SCAN_WARPS - 1) / pcv::SCAN_WARPS);
  if (grid < 1) grid = 1;
  const uint32_t n_groups = (n_queries + (uint32_t)nb - 1) / (uint32_t)nb;
  const uint32_t groups_per_launch = gvar ? std::min<uint32_t>(n_groups, (uint32_t)pcv::SCAN_MAX_GROUPS) : 1u;
  cudaError_t ce = ix->partial.reserve((size_t)grid * nb * k * groups_per_launch);
  if (ce != cudaSuccess) return fail(PCV_ERR_OOM, "workspace allocation failed: %s", cudaGetErrorString(ce));

  pcv::ScanParams p;
  memset(&p, 0, sizeof p);
  p.rows = ix->d_rows;
  p.rows_lo = ix->store == PCV_F32_SPLIT ? ix->lo_plane() : nullptr;
  p.row_bytes = (uint32_t)ix->row_bytes;
  p.d_chunks = (uint32_t)(ix->row_bytes / 16);
  p.lpr_log2 = pl.lpr_log2;
  p.tile_iters = pl.tile_iters;
  p.tile_rows = pl.tile_rows;
  p.slot_bytes = pl.slot_bytes;
  p.nslots = pl.nslots;
  p.range_prefix = R.range_prefix.p;
  p.ranges = R.ranges.p;
  p.n_ranges = (uint32_t)R.h_ranges.size();
  if (p.n_ranges == 1) p.range0 = R.h_ranges[0];
  p.total_tiles = R.total_tiles;
  p.q_stride = ix->dim_padded;
  p.k = k;
  p.dim = ix->dim;
  p.emit_mode = o.emit_mode;
  p.l2_evict_first = (sel_rows * ix->row_bytes > (96ull << 20)) ? 1u : 0u;
  if (const char* e = getenv("PCV_SCAN_L2HINT")) p.l2_evict_first = (uint32_t)atoi(e);
  p.lrank_of_row = ix->d_lrank_of_row;
  p.row_of_lrank = ix->d_row_of_lrank;
  p.ids = ix->d_ids;
  p.id_base = ix->id_base;
  p.partial = ix->partial.p;
  p.done = ix->d_done + CTL_SCAN_DONE;
  if (o.emit_mode == 2) p.xchg = *o.xchg;

  const size_t smem = pcv::scan_smem_bytes(p, nb, var->q_in_smem, grid);
  if (smem > 232448 - 2048) return fail(PCV_ERR_UNSUPPORTED, "scan needs %zu bytes of shared memory", smem);

  if (gvar) {
    p.queries = d_q_padded;
    p.nb = (uint32_t)nb;
    p.out_ids = o.ids;
    p.out_scores = o.scores;
    p.out_sims = o.sims;
    p.out_counts = o.counts;
    p.q_list = d_q_list;
    p.q_count = d_q_count;
    p.n_listed = n_queries;
    for (uint32_t g0 = 0; g0 < n_groups; g0 += groups_per_launch) {
      p.group_begin = g0;
      p.group_count = groups_per_launch;
      cudaError_t e = var->fn(p, grid, smem, ix->stream);
      if (e != cudaSuccess) return fail(PCV_ERR_CUDA, "scan launch failed: %s", cudaGetErrorString(e));
      ix->last_launches += 1;
    }
  } else {
    for (uint32_t q0 = 0; q0 < n_queries; q0 += (uint32_t)nb) {
      p.queries = d_q_padded + (size_t)q0 * ix->dim_padded;
      p.nb = std::min<uint32_t>((uint32_t)nb, n_queries - q0);
      p.out_ids = o.ids + (size_t)q0 * k;
      p.out_scores = o.scores ? o.scores + (size_t)q0 * k : nullptr;
      p.out_sims = o.sims ? o.sims + (size_t)q0 * k : nullptr;
      p.out_counts = o.counts ? o.counts + q0 : nullptr;
      cudaError_t e = var->fn(p, grid, smem, ix->stream);
      if (e != cudaSuccess) return fail(PCV_ERR_CUDA, "scan launch failed: %s", cudaGetErrorString(e));
      ix->last_launches += 1;
    }
  }
  if (!d_q_list) {
    ix->last_kernel = 1;
    ix->last_scan_bytes = sel_rows * ix->row_bytes * n_groups;
  }
  return PCV_OK;
}

// Everything a single-query scan of this shard will allocate or synchronise on — row ranges for the scan tiling,
// the per-CTA partial lists, the padded-query buffer — done AHEAD of the launch.  A single-process many-GPU handle
// calls this for every shard before it launches the first fused scan (whose last CTA waits for its peers): no
// cudaMalloc or stream synchronisation may then sit between those launches.
int32_t scan_prepare(pcv_index* ix, uint32_t n_queries, uint32_t k, const int64_t* sources, uint32_t n_sources) {
  ScanPlan pl;
  int32_t rc = plan_scan(ix, pl);
  if (rc != PCV_OK) return rc;
  rc = prepare_ranges(ix, sources, n_sources, sources == nullptr, pl.tile_rows, 0);
  if (rc != PCV_OK) return rc;
  const uint32_t groups = std::min<uint32_t>((n_queries + 3) / 4, (uint32_t)pcv::SCAN_MAX_GROUPS);
  cudaError_t ce = ix->partial.reserve((size_t)ix->sm_count * 4 * k * std::max(groups, 1u));
  if (ce == cudaSuccess && (ix->dim_padded != ix->dim || ix->store == PCV_BF16)) ce = ix->q_pad.reserve((size_t)n_queries * ix->dim_padded);
  if (ce != cudaSuccess) return fail(PCV_ERR_OOM, "workspace allocation failed: %s", cudaGetErrorString(ce));
  return PCV_OK;
}

// Enqueue the local (this shard's) search of n_queries padded device queries.
// emit_mode 0: final outputs; 1: (sim,id) candidates into out_sims/out_ids.
int32_t enqueue_local_search(pcv_index* ix, const float* d_q_padded, uint32_t n_queries, uint32_t k,
                             const int64_t* sources, uint32_t n_sources, bool all, uint32_t emit_mode,
                             int64_t* d_out_ids, float* d_out_scores, float* d_out_sims, uint32_t* d_out_counts,
                             const pcv::ExchangeTarget* xchg = nullptr) {
  int32_t rc;
  // rows the source filter selects (search.rs:166)
  uint64_t sel_rows = 0;
  for (const Segment& s : ix->segs) {
    bool sel = all;
    if (!sel)
      for (uint32_t i = 0; i < n_sources; ++i)
        if (sources[i] == s.source_id) { sel = true; break; }
    if (sel) sel_rows += s.end - s.begin;
  }
  const SearchOut out{emit_mode, d_out_ids, d_out_scores, d_out_sims, d_out_counts, xchg};
  ix->last_used_filter = false;

  // K2 / K3: tensor-core path — batches over bf16 rows; batches over split rows go through it as a FILTER
  // over the hi plane, followed by exact fp32 rescoring (pcv_rescore.cuh)
  const bool split = ix->store == PCV_F32_SPLIT;
  const bool cosine = ix->metric == PCV_METRIC_COSINE;
  const uint32_t kk = split ? pcv::split_filter_k(k) : k;  // what the tensor path selects
  const bool gemm_ok = (ix->store == PCV_BF16 || split) && !env_flag("PCV_NO_TENSOR_PATH") &&
                       pcv::gemm_path_applicable(cosine, ix->dim_padded, n_queries, split ? std::max(k, kk) : k, sel_rows, ix->n_rows);
  if (gemm_ok) {
    const uint32_t gemm_tile = pcv::gemm_tile_rows(ix->dim_padded);
    rc = prepare_ranges(ix, sources, n_sources, all, gemm_tile, 1);
    if (rc != PCV_OK) return rc;
    const pcv_index::RangeSet& R = ix->rs[1];
    pcv::GemmCall gc;
    memset(&gc, 0, sizeof gc);
    gc.rows = ix->d_rows; gc.n_rows = ix->n_rows; gc.dim_padded = ix->dim_padded; gc.dim = ix->dim;
    gc.row_bytes = (uint64_t)ix->dim_padded * 2;  // a bf16 matrix, or the hi plane of a split one
    gc.keys_only = split;
    gc.cosine = cosine;
    if (gc.cosine && !ix->d_xinv) {
      const uint64_t n_out = ix->n_rows + 128;
      CU(cudaMalloc((void**)&ix->d_xinv, n_out * sizeof(float)));
      cudaError_t ne = pcv::gemm_row_inv_norms(ix->d_rows, ix->n_rows, ix->dim_padded, ix->d_xinv, n_out, ix->sm_count, ix->stream);
      if (ne != cudaSuccess) return fail(PCV_ERR_CUDA, "row norm kernel failed: %s", cudaGetErrorString(ne));
      ix->last_launches += 1;
    }
    gc.x_inv_norm = ix->d_xinv;
    gc.d_ranges = R.ranges.p; gc.d_range_prefix = R.range_prefix.p;
    gc.n_ranges = (uint32_t)R.h_ranges.size(); gc.total_tiles = R.total_tiles;
    gc.k = kk;
    gc.emit_mode = emit_mode;
    gc.lrank_of_row = ix->d_lrank_of_row; gc.row_of_lrank = ix->d_row_of_lrank;
    gc.ids = ix->d_ids; gc.id_base = ix->id_base;
    gc.sm_count = ix->sm_count; gc.stream = ix->stream;
    // batches beyond 4096 queries run as consecutive chunks (bounds the per-(CTA, query) candidate
    // buffers; the corpus is re-streamed per chunk, which a tensor-bound pass does not notice)
    constexpr uint32_t kChunk = 4096;
    if (split) {
      CU(ix->margin.reserve(std::min(n_queries, kChunk)));
      CU(ix->fb_list.reserve(n_queries));
      CU(cudaMemsetAsync(ix->d_done + CTL_FB_COUNT, 0, sizeof(unsigned int), ix->stream));
    }
    for (uint32_t q0 = 0; q0 < n_queries; q0 += kChunk) {
      const uint32_t nq = std::min(kChunk, n_queries - q0);
      gc.queries = d_q_padded + (size_t)q0 * ix->dim_padded;
      gc.n_queries = nq;
      gc.out_ids = d_out_ids + (size_t)q0 * k;
      gc.out_scores = d_out_scores ? d_out_scores + (size_t)q0 * k : nullptr;
      gc.out_sims = d_out_sims ? d_out_sims + (size_t)q0 * k : nullptr;
      gc.out_counts = d_out_counts ? d_out_counts + q0 : nullptr;
      uint32_t launches = 0;
      cudaError_t e = cudaSuccess;
      const char* what = pcv::gemm_search(ix->gemm, gc, &launches, &e);
      if (what) return fail(e == cudaErrorMemoryAllocation ? PCV_ERR_OOM : PCV_ERR_CUDA, "tcgen05 search: %s failed: %s", what, cudaGetErrorString(e));
      ix->last_launches += launches;
      if (split) {
        // exact fp32 rescoring of the kk candidates per query + proof of completeness
        ScanPlan pl;
        rc = plan_scan(ix, pl);  // lanes per row as K1 picks them: the summation order is K1's
        if (rc != PCV_OK) return rc;
        pcv::split_query_margin_kernel<<<(nq + 7) / 8, 256, 0, ix->stream>>>(gc.queries, nq, ix->dim_padded, ix->d_done + CTL_XMAX2, ix->margin.p);
        CU(cudaGetLastError());
        pcv::RescoreParams rp;
        memset(&rp, 0, sizeof rp);
        rp.cand = ix->gemm.d_topk; rp.kf = kk; rp.k = k; rp.n_queries = nq; rp.q_offset = q0;
        rp.hi = ix->hi_plane(); rp.lo = ix->lo_plane();
        rp.plane_row_bytes = ix->dim_padded * 2; rp.d_chunks = (uint32_t)(ix->row_bytes / 16); rp.lpr_log2 = pl.lpr_log2;
        rp.queries = d_q_padded; rp.q_stride = ix->dim_padded; rp.margin = ix->margin.p;
        rp.row_of_lrank = ix->d_row_of_lrank; rp.ids = ix->d_ids; rp.id_base = ix->id_base;
        rp.emit_mode = emit_mode; rp.dim = ix->dim;
        rp.out_ids = d_out_ids; rp.out_scores = d_out_scores; rp.out_sims = d_out_sims; rp.out_counts = d_out_counts;
        rp.fb_list = ix->fb_list.p; rp.fb_count = ix->d_done + CTL_FB_COUNT;
        if (pl.nj == 6) pcv::rescore_exact_kernel<6><<<nq, 256, 0, ix->stream>>>(rp);
        else pcv::rescore_exact_kernel<12><<<nq, 256, 0, ix->stream>>>(rp);
        CU(cudaGetLastError());
        ix->last_launches += 2;
      }
    }
    ix->last_kernel = 2;
    ix->last_scan_bytes = sel_rows * (split ? ix->row_bytes / 2 : ix->row_bytes) * ((n_queries + kChunk - 1) / kChunk);
    if (split) {
      // queries the proof rejected: the exact scan, one GROUPED launch per 256 queries (each exits at once
      // when the list — whose length only the device knows — is shorter)
      ix->last_used_filter = true;
      rc = enqueue_scan(ix, d_q_padded, n_queries, k, sources, n_sources, all, sel_rows, out, ix->fb_list.p, ix->d_done + CTL_FB_COUNT);
      if (rc != PCV_OK) return rc;
    }
    return PCV_OK;
  }
  return enqueue_scan(ix, d_q_padded, n_queries, k, sources, n_sources, all, sel_rows, out, nullptr, nullptr);
}

// A search on device buffers runs in two phases so that a single-process, many-GPU handle can enqueue
// phase 1 on every shard before any shard starts phase 2 (whose kernel waits for its peers' stores):
//   phase 1  pad the queries, enqueue this shard's local search (world 1: straight into the outputs;
//            sharded: (sim, id) candidates into cand_send)
//   phase 2  sharded only: exchange the candidates and merge them into the outputs
int32_t search_phase_local(pcv_index* ix, const float* d_queries, uint32_t n_queries, uint32_t k, const int64_t* sources,
                           uint32_t n_sources, int64_t* d_out_ids, float* d_out_scores, float* d_out_sims,
                           uint32_t* d_out_counts) {
  const bool all = (sources == nullptr);
  NvtxRange nvtx("pcv:search (enqueue)");
  ix->last_launches = 0;
  ix->last_kernel = 0;
  if (ix->timing) cudaEventRecord(ix->ev0, ix->stream);
  // zero-padded queries
  const float* d_q = d_queries;
  if (ix->dim_padded != ix->dim || ix->store == PCV_BF16) {
    CU(ix->q_pad.reserve((size_t)n_queries * ix->dim_padded));
    pcv::pad_queries_kernel<<<std::min<uint32_t>(n_queries, 1024u), 256, 0, ix->stream>>>(d_queries, ix->q_pad.p, n_queries, ix->dim, ix->dim_padded, ix->store == PCV_BF16 ? 1 : 0);
    CU(cudaGetLastError());
    ix->last_launches += 1;
    d_q = ix->q_pad.p;
  }
  if (ix->world == 1)
    return enqueue_local_search(ix, d_q, n_queries, k, sources, n_sources, all, 0, d_out_ids, d_out_scores, d_out_sims, d_out_counts);
  if (ix->shard_failed)
    return fail(PCV_ERR_STATE, "an earlier collective search failed on this shard: its exchange state is out of step with its peers; rebuild the sharded index");
  const size_t n_pad = (((size_t)n_queries * k) + 1) & ~(size_t)1;
  const bool use_p2p = ix->p2p_attached && (size_t)n_queries * k <= ix->p2p_cap;
  if (!use_p2p && !ix->comm)
    return fail(PCV_ERR_STATE, "sharded search of %u x %u candidates exceeds the peer buffers (%u records) and no NCCL communicator is attached",
                n_queries, k, ix->p2p_cap);
  // One query (the reference's own call, search.rs:157) over peer-mapped buffers: the scan's last CTA delivers
  // this shard's k candidates straight into every shard's buffer, waits for the others' and merges — the
  // whole sharded search is ONE launch per GPU, with no separate exchange kernel.
  ix->fused_exchange = use_p2p && n_queries == 1 && k <= 128 && !env_flag("PCV_NO_FUSED_EXCHANGE");
  if (ix->fused_exchange) {
    pcv::ExchangeTarget x;
    memset(&x, 0, sizeof x);
    for (int r = 0; r < ix->world; ++r) x.peer[r] = ix->p2p_peer[r];
    x.rank = (uint32_t)ix->rank;
    x.world = (uint32_t)ix->world;
    x.cap = ix->p2p_cap;
    x.epoch = ix->p2p_epoch + 1;  // committed in phase 2, once the launch is known to have been accepted
    return enqueue_local_search(ix, d_q, n_queries, k, sources, n_sources, all, 2, d_out_ids, d_out_scores, d_out_sims, d_out_counts, &x);
  }
  CU(ix->cand_send.reserve(n_pad * 12));
  int64_t* s_ids = reinterpret_cast<int64_t*>(ix->cand_send.p);
  float* s_sims = reinterpret_cast<float*>(ix->cand_send.p + n_pad * 8);
  return enqueue_local_search(ix, d_q, n_queries, k, sources, n_sources, all, 1, s_ids, nullptr, s_sims, nullptr);
}

int32_t search_phase_exchange(pcv_index* ix, uint32_t n_queries, uint32_t k, int64_t* d_out_ids, float* d_out_scores,
                              float* d_out_sims, uint32_t* d_out_counts) {
  if (ix->world > 1 && ix->fused_exchange) {
    ix->p2p_epoch += 1;  // the scan launched in phase 1 carried the exchange
  } else if (ix->world > 1) {
    const size_t n_pad = (((size_t)n_queries * k) + 1) & ~(size_t)1;
    const size_t per_rank = n_pad * 12;
    const bool use_p2p = ix->p2p_attached && (size_t)n_queries * k <= ix->p2p_cap;
    int64_t* s_ids = reinterpret_cast<int64_t*>(ix->cand_send.p);
    float* s_sims = reinterpret_cast<float*>(ix->cand_send.p + n_pad * 8);
    NvtxRange nvtx_x(use_p2p ? "pcv:exchange (peer stores + epoch flags + merge)" : "pcv:exchange (ncclAllGather + merge)");
    if (use_p2p) {
      // K5p: stores into peer memory + epoch flags + merge, one launch, no NCCL.  A peer that never
      // publishes its epoch (it failed before this point) makes the kernel trap after a bounded spin:
      // the launch fails instead of hanging the GPU, and this context is lost — documented failure mode.
      pcv::P2PParams pp;
      memset(&pp, 0, sizeof pp);
      pp.s_ids = s_ids;
      pp.s_sims = s_sims;
      pp.n_queries = n_queries;
      pp.k = k;
      pp.dim = ix->dim;
      pp.cosine = ix->metric == PCV_METRIC_COSINE ? 1 : 0;
      pp.rank = (uint32_t)ix->rank;
      pp.world = (uint32_t)ix->world;
      pp.cap = ix->p2p_cap;
      pp.epoch = ix->p2p_epoch + 1;  // committed below, once the launch is known to have been accepted
      for (int r = 0; r < ix->world; ++r) pp.peer[r] = ix->p2p_peer[r];
      pp.done_ctr = ix->d_done + CTL_P2P_DONE;
      pp.out_ids = d_out_ids;
      pp.out_scores = d_out_scores;
      pp.out_sims = d_out_sims;
      pp.out_counts = d_out_counts;
      const uint32_t want = std::max<uint32_t>((n_queries + 7) / 8, (uint32_t)(((size_t)n_queries * k + 2047) / 2048));
      const uint32_t blocks = std::max<uint32_t>(1u, std::min<uint32_t>(want, (uint32_t)ix->sm_count));
      pcv::p2p_exchange_merge_kernel<<<blocks, 256, 0, ix->stream>>>(pp);
      CU(cudaGetLastError());
      ix->p2p_epoch = pp.epoch;
      ix->last_launches += 1;
    } else {
      CU(ix->cand_recv.reserve(per_rank * ix->world));
      NC(nccl_api().AllGather(ix->cand_send.p, ix->cand_recv.p, per_rank, ncclChar, ix->comm, ix->stream));
      const int64_t* r_ids = reinterpret_cast<const int64_t*>(ix->cand_recv.p);
      const float* r_sims = reinterpret_cast<const float*>(ix->cand_recv.p + n_pad * 8);
      const uint32_t warps_per_block = 4;
      const uint32_t blocks = (n_queries + warps_per_block - 1) / warps_per_block;
      pcv::merge_candidates_kernel<<<blocks, warps_per_block * 32, 0, ix->stream>>>(
          r_sims, per_rank / 4, r_ids, per_rank / 8, (uint32_t)ix->world, n_queries, k, ix->dim,
          ix->metric == PCV_METRIC_COSINE ? 1 : 0, d_out_ids, d_out_scores, d_out_sims, d_out_counts);
      CU(cudaGetLastError());
      ix->last_launches += 1;
    }
  }
  if (ix->timing) cudaEventRecord(ix->ev1, ix->stream);
  ix->ev_valid = ix->timing;
  ix->searched = true;
  return PCV_OK;
}

// full search on device buffers (handles padding, shards, merge).  A failure of a COLLECTIVE search leaves
// this shard's epoch / communicator out of step with its peers: the handle refuses further sharded searches.
int32_t search_device_locked(pcv_index* ix, const float* d_queries, uint32_t n_queries, uint32_t k,
                             const int64_t* sources, uint32_t n_sources, int64_t* d_out_ids,
                             float* d_out_scores, float* d_out_sims, uint32_t* d_out_counts) {
  int32_t rc = search_phase_local(ix, d_queries, n_queries, k, sources, n_sources, d_out_ids, d_out_scores, d_out_sims, d_out_counts);
  if (rc == PCV_OK) rc = search_phase_exchange(ix, n_queries, k, d_out_ids, d_out_scores, d_out_sims, d_out_counts);
  if (rc != PCV_OK && ix->world > 1) ix->shard_failed = true;
  return rc;
}

// ===========================================================================
// Single-process, many-GPU handle (pcv_index_create_multi).  The reference Searcher is ONE Send + Sync
// object in ONE process (crates/perceive-tauri/src-tauri/app_state.rs:63-75, crates/perceive-cli/state.rs:28-56);
// this is the same thing over N devices: the front handle owns one ordinary one-device shard per GPU
// (world N, rank r), every shard's receive buffer is reachable from every other device through peer access
// (cudaDeviceEnablePeerAccess — no IPC handles, no second process), and a search enqueues phase 1 on every
// shard's stream, then phase 2 (peer stores + epoch flags + merge, pcv_load.cuh) on every shard's stream.
// Shard 0's device is where device-resident queries live and results are delivered.
// ===========================================================================
// spin for up to `us` microseconds while `keep_waiting()` holds; true if it stopped holding
template <typename F>
bool spin_while(F keep_waiting, int us) {
  const auto t0 = std::chrono::steady_clock::now();
  for (;;) {
    for (int i = 0; i < 64; ++i) {
      if (!keep_waiting()) return true;
#if defined(__x86_64__)
      __builtin_ia32_pause();
#endif
    }
    if (std::chrono::steady_clock::now() - t0 > std::chrono::microseconds(us)) return !keep_waiting();
  }
}

void shard_worker_main(ShardPool* pool, int r, int device) {
  cudaSetDevice(device);  // this thread only ever talks to this device
  uint64_t seen = 0;
  for (;;) {
    // spin for a while first: the next phase, or the next search, usually follows within microseconds, and waking
    // eight sleeping threads through a futex costs more than a 60 us scan
    auto idle = [&] { return pool->ticket.load(std::memory_order_acquire) == seen && !pool->quit.load(std::memory_order_relaxed); };
    if (!spin_while(idle, 300)) {
      std::unique_lock<std::mutex> lk(pool->m);
      pool->cv_work.wait(lk, [&] { return !idle(); });
    }
    if (pool->quit.load(std::memory_order_relaxed)) return;
    seen = pool->ticket.load(std::memory_order_acquire);
    const std::function<int32_t(int)>* task = pool->task;  // published before the ticket moved
    int32_t rc;
    try {
      rc = (*task)(r);
    } catch (const std::bad_alloc&) {
      rc = fail(PCV_ERR_OOM, "out of host memory");
    } catch (...) {
      rc = fail(PCV_ERR_STATE, "unexpected C++ exception in a shard worker");
    }
    pool->rc[r] = rc;
    if (rc != PCV_OK) pool->err[r] = g_err;  // the worker's thread-local message travels with the status
    if (pool->pending.fetch_sub(1, std::memory_order_acq_rel) == 1) {
      std::lock_guard<std::mutex> lk(pool->m);  // the caller may be about to sleep on cv_done: no lost wake-up
      pool->cv_done.notify_one();
    }
  }
}

// Run task(r) for every shard r on that shard's worker, wait for all; the first failure's status and message win.
int32_t run_on_shards(pcv_index* mx, const std::function<int32_t(int)>& task) {
  const int n = (int)mx->shards.size();
  ShardPool* pool = mx->pool;
  if (!pool) {  // one shard: nothing to fan out
    int32_t rc = PCV_OK;
    for (int r = 0; r < n && rc == PCV_OK; ++r) rc = task(r);
    return rc;
  }
  std::fill(pool->rc.begin(), pool->rc.end(), PCV_OK);
  pool->task = &task;
  pool->pending.store(n, std::memory_order_relaxed);
  {
    std::lock_guard<std::mutex> lk(pool->m);  // a worker between its last check and its wait must see the new ticket
    pool->ticket.fetch_add(1, std::memory_order_release);
  }
  pool->cv_work.notify_all();
  auto busy = [&] { return pool->pending.load(std::memory_order_acquire) != 0; };
  if (!spin_while(busy, 300)) {
    std::unique_lock<std::mutex> lk(pool->m);
    pool->cv_done.wait(lk, [&] { return !busy(); });
  }
  pool->task = nullptr;
  for (int r = 0; r < n; ++r)
    if (pool->rc[r] != PCV_OK) {
      g_err = pool->err[r];
      return pool->rc[r];
    }
  return PCV_OK;
}

bool is_multi(const pcv_index* ix) { return ix && !ix->shards.empty(); }

void multi_refresh_layout(pcv_index* mx) {
  mx->rows_epoch += 1;  // called by everything that changes the shards' rows
  mx->shard_row0.assign(mx->shards.size() + 1, 0);
  for (size_t r = 0; r < mx->shards.size(); ++r) mx->shard_row0[r + 1] = mx->shard_row0[r] + mx->shards[r]->n_rows;
  mx->n_rows = mx->shard_row0.back();
}

// (re)allocate every shard's receive buffer for `records` candidates per shard and cross-wire the pointers
int32_t multi_ensure_peer_buffers(pcv_index* mx, size_t records) {
  const int n = (int)mx->shards.size();
  if (n < 2) return PCV_OK;
  const uint32_t cap = (uint32_t)std::max<size_t>((records + 1) & ~(size_t)1, 4096);
  if (mx->shards[0]->p2p_attached && mx->shards[0]->p2p_cap >= cap) return PCV_OK;
  if (records > (1u << 24)) return fail(PCV_ERR_UNSUPPORTED, "%zu candidates per shard exceed the exchange buffers (2^24 records)", records);
  const size_t bytes = 2 * pcv::p2p_half_bytes((uint32_t)n, cap);
  for (pcv_index* sh : mx->shards) {
    CU(cudaSetDevice(sh->device));
    CU(cudaStreamSynchronize(sh->stream));
  }
  for (pcv_index* sh : mx->shards) {
    CU(cudaSetDevice(sh->device));
    if (sh->p2p_local) cudaFree(sh->p2p_local);
    sh->p2p_local = nullptr;
    sh->p2p_attached = false;
    CU(cudaMalloc((void**)&sh->p2p_local, bytes));
    CU(cudaMemset(sh->p2p_local, 0, bytes));  // flags start at epoch 0; searches count from 1
  }
  for (int r = 0; r < n; ++r) {
    pcv_index* sh = mx->shards[r];
    for (int q = 0; q < n; ++q) sh->p2p_peer[q] = mx->shards[q]->p2p_local;
    sh->p2p_world = (uint32_t)n;
    sh->p2p_cap = cap;
    sh->p2p_epoch = 0;
    sh->p2p_attached = true;
    sh->p2p_in_process = true;
  }
  return PCV_OK;
}

int32_t multi_search_device_locked(pcv_index* mx, const float* d_queries, uint32_t n_queries, uint32_t k,
                                   const int64_t* sources, uint32_t n_sources, int64_t* d_out_ids, float* d_out_scores,
                                   float* d_out_sims, uint32_t* d_out_counts) {
  const int n = (int)mx->shards.size();
  pcv_index* root = mx->shards[0];
  if (n == 1) {
    CU(cudaSetDevice(root->device));
    return search_device_locked(root, d_queries, n_queries, k, sources, n_sources, d_out_ids, d_out_scores, d_out_sims, d_out_counts);
  }
  int32_t rc = multi_ensure_peer_buffers(mx, (size_t)n_queries * k);
  if (rc != PCV_OK) return rc;
  const size_t q_bytes = (size_t)n_queries * mx->dim * sizeof(float);
  const size_t nk = (size_t)n_queries * k;
  // the queries travel from shard 0's device to every other shard over NVLink, ordered after whatever
  // produced them on shard 0's stream
  CU(cudaSetDevice(root->device));
  CU(cudaEventRecord(mx->ev0, root->stream));
  auto outs = [&](int r, int64_t*& ids, float*& scores, float*& sims, uint32_t*& counts) {
    if (r == 0) { ids = d_out_ids; scores = d_out_scores; sims = d_out_sims; counts = d_out_counts; return; }
    uint8_t* b = mx->shards[r]->o_pack.p;  // the other shards merge too (all-to-all exchange); their copy is not read back
    ids = reinterpret_cast<int64_t*>(b);
    scores = reinterpret_cast<float*>(b + nk * 8);
    sims = reinterpret_cast<float*>(b + nk * 12);
    counts = reinterpret_cast<uint32_t*>(b + nk * 16);
  };
  // Every shard's share of a phase runs on that shard's worker thread (ShardPool), all shards at once.
  // Phase 0 (single queries only): one query makes every shard's scan carry the exchange in its last CTA
  // (search_phase_local), i.e. phase 1 already launches kernels that wait for one another — so whatever phase 1
  // would allocate or synchronise on is done for ALL shards first.
  const bool fused = n_queries == 1 && k <= 128 && !env_flag("PCV_NO_FUSED_EXCHANGE");
  if (fused) {
    const bool all = sources == nullptr;
    const bool same = mx->prep_epoch == mx->rows_epoch && mx->prep_k == k && mx->prep_all == all &&
                      mx->prep_sources.size() == (all ? 0u : n_sources) && (all || std::equal(mx->prep_sources.begin(), mx->prep_sources.end(), sources));
    if (!same) {
      rc = run_on_shards(mx, [&](int r) { return scan_prepare(mx->shards[r], n_queries, k, sources, n_sources); });
      if (rc != PCV_OK) return rc;
      mx->prep_epoch = mx->rows_epoch;
      mx->prep_k = k;
      mx->prep_all = all;
      mx->prep_sources.assign(sources, sources + (all ? 0u : n_sources));
    }
  }
  // Phase 1 everywhere (may allocate / synchronise a stream), THEN phase 2 everywhere (launch only): a shard's
  // exchange kernel waits for its peers' stores, so no host-side wait may sit between those launches.
  rc = run_on_shards(mx, [&](int r) -> int32_t {
    pcv_index* sh = mx->shards[r];
    if (r > 0) {
      CU(sh->q_in.reserve((size_t)n_queries * mx->dim));
      CU(sh->o_pack.reserve(nk * 16 + (size_t)n_queries * 4 + 64));
      CU(cudaStreamWaitEvent(sh->stream, mx->ev0, 0));
      CU(cudaMemcpyPeerAsync(sh->q_in.p, sh->device, d_queries, root->device, q_bytes, sh->stream));
    }
    int64_t* ids; float* scores; float* sims; uint32_t* counts;
    outs(r, ids, scores, sims, counts);
    return search_phase_local(sh, r == 0 ? d_queries : sh->q_in.p, n_queries, k, sources, n_sources, ids, scores, sims, counts);
  });
  if (rc != PCV_OK) {
    // non-fused: nothing that waits for a peer has been launched yet and the epochs are unchanged.  Fused: the scans
    // of the shards that did launch are waiting for the one that failed — they will trap; the handle is out of step.
    if (fused)
      for (pcv_index* s2 : mx->shards) s2->shard_failed = true;
    return rc;
  }
  if (fused) {
    // the scans launched in phase 1 carried the exchange: what is left per shard is bookkeeping (epoch, timing event)
    for (int r = 0; r < n && rc == PCV_OK; ++r) {
      int64_t* ids; float* scores; float* sims; uint32_t* counts;
      outs(r, ids, scores, sims, counts);
      CU(cudaSetDevice(mx->shards[r]->device));
      rc = search_phase_exchange(mx->shards[r], n_queries, k, ids, scores, sims, counts);
    }
    cudaSetDevice(root->device);
  } else {
    rc = run_on_shards(mx, [&](int r) -> int32_t {
      int64_t* ids; float* scores; float* sims; uint32_t* counts;
      outs(r, ids, scores, sims, counts);
      return search_phase_exchange(mx->shards[r], n_queries, k, ids, scores, sims, counts);
    });
  }
  if (rc != PCV_OK)
    for (pcv_index* s2 : mx->shards) s2->shard_failed = true;  // some shards already wait for the one that failed
  return rc;
}

int32_t multi_set_rows(pcv_index* mx, const float* rows, const int64_t* ids, const int64_t* source_ids, uint64_t n) {
  const uint64_t ns = mx->shards.size();
  for (uint64_t r = 0; r < ns; ++r) {  // contiguous, balanced slices of the caller's rows; every shard orders its own
    const uint64_t b = n * r / ns, e = n * (r + 1) / ns;
    const int32_t rc = pcv_index_set_rows(mx->shards[r], rows ? rows + b * mx->dim : nullptr, ids ? ids + b : nullptr,
                                          source_ids ? source_ids + b : nullptr, e - b);
    if (rc != PCV_OK) {
      for (pcv_index* sh : mx->shards) pcv_index_set_rows(sh, nullptr, nullptr, nullptr, 0);  // never half-loaded
      multi_refresh_layout(mx);
      return rc;
    }
  }
  multi_refresh_layout(mx);
  return PCV_OK;
}

// Searcher::rebuild_source (search.rs:58-79) over shards: the source's old rows leave every shard, its new rows
// are dealt out in balanced id-ordered slices (a shard's segments need no global order: the merge goes by id)
int32_t multi_replace_source(pcv_index* mx, int64_t source_id, const float* rows, const int64_t* ids, uint64_t n) {
  const uint64_t ns = mx->shards.size();
  std::vector<uint64_t> perm(n);
  std::iota(perm.begin(), perm.end(), 0ull);
  std::stable_sort(perm.begin(), perm.end(), [&](uint64_t a, uint64_t b) { return ids[a] < ids[b]; });
  std::vector<float> srows;
  std::vector<int64_t> sids;
  for (uint64_t r = 0; r < ns; ++r) {
    const uint64_t b = n * r / ns, e = n * (r + 1) / ns;
    srows.resize((e - b) * (size_t)mx->dim);
    sids.resize(e - b);
    for (uint64_t i = b; i < e; ++i) {
      memcpy(srows.data() + (i - b) * (size_t)mx->dim, rows + perm[i] * (size_t)mx->dim, (size_t)mx->dim * 4);
      sids[i - b] = ids[perm[i]];
    }
    const int32_t rc = pcv_index_replace_source(mx->shards[r], source_id, srows.data(), sids.data(), e - b);
    if (rc != PCV_OK) { multi_refresh_layout(mx); return rc; }
  }
  multi_refresh_layout(mx);
  return PCV_OK;
}

int32_t multi_get_rows(pcv_index* mx, uint64_t first_row, uint64_t n, float* out_rows, int64_t* out_ids, int64_t* out_source_ids) {
  if (first_row + n > mx->n_rows) return fail(PCV_ERR_INVALID, "rows [%llu,%llu) outside [0,%llu)", (unsigned long long)first_row, (unsigned long long)(first_row + n), (unsigned long long)mx->n_rows);
  for (size_t r = 0; r < mx->shards.size(); ++r) {
    const uint64_t b = std::max(first_row, mx->shard_row0[r]), e = std::min(first_row + n, mx->shard_row0[r + 1]);
    if (b >= e) continue;
    const uint64_t o = b - first_row;
    const int32_t rc = pcv_index_get_rows(mx->shards[r], b - mx->shard_row0[r], e - b, out_rows ? out_rows + o * mx->dim : nullptr,
                                          out_ids ? out_ids + o : nullptr, out_source_ids ? out_source_ids + o : nullptr);
    if (rc != PCV_OK) return rc;
  }
  return PCV_OK;
}

int32_t validate_search(pcv_index* ix, const void* queries, uint32_t n_queries, uint32_t k, const int64_t* sources,
                        uint32_t n_sources, const void* out_ids, const void* out_scores) {
  if (!ix) return fail(PCV_ERR_INVALID, "null index");
  if (n_queries && !queries) return fail(PCV_ERR_INVALID, "null queries");
  if (k == 0 || k > PCV_MAX_K) return fail(PCV_ERR_INVALID, "k=%u outside [1,%u]", k, PCV_MAX_K);
  if (n_queries && (!out_ids || !out_scores)) return fail(PCV_ERR_INVALID, "null output buffer");
  if (sources == nullptr && n_sources != 0) return fail(PCV_ERR_INVALID, "n_sources=%u with null sources", n_sources);
  return PCV_OK;
}

}  // namespace

// ===========================================================================
// No C++ exception may cross the C boundary (a Rust or C caller cannot unwind it): every int32_t
// entry point is a function-try-block that turns one into a status code.
#define PCV_CATCH                                                                                        \
  catch (const std::bad_alloc&) { return fail(PCV_ERR_OOM, "out of host memory"); }                      \
  catch (const std::exception& e) { return fail(PCV_ERR_STATE, "unexpected C++ exception: %s", e.what()); } \
  catch (...) { return fail(PCV_ERR_STATE, "unexpected C++ exception"); }

// error hand-off for the other translation units of the library (not exported)
int32_t pcv_internal_fail(int32_t code, const char* msg) {
  g_err = msg;
  return code;
}

extern "C" {

const char* pcv_last_error(void) { return g_err.c_str(); }
uint32_t pcv_abi_version(void) { return PCV_ABI_VERSION; }

int32_t pcv_device_count(int32_t* out) try {
  if (!out) return fail(PCV_ERR_INVALID, "null out");
  int n = 0;
  cudaError_t e = cudaGetDeviceCount(&n);
  if (e != cudaSuccess) {
    *out = 0;
    return fail(PCV_ERR_CUDA, "cudaGetDeviceCount failed: %s", cudaGetErrorString(e));
  }
  *out = n;
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_create(int32_t device, uint32_t dim, pcv_dtype store, pcv_metric metric, uint32_t flags,
                         pcv_index** out) try {
  if (!out) return fail(PCV_ERR_INVALID, "null out");
  *out = nullptr;
  if (dim == 0 || dim > PCV_MAX_DIM) return fail(PCV_ERR_INVALID, "dim=%u outside [1,%u]", dim, PCV_MAX_DIM);
  if (store != PCV_F32 && store != PCV_BF16 && store != PCV_F32_SPLIT) return fail(PCV_ERR_INVALID, "bad storage type %d", (int)store);
  if (store == PCV_F32_SPLIT && (metric != PCV_METRIC_DOT_REF || dim < 64 || dim > 768))
    return fail(PCV_ERR_UNSUPPORTED, "PCV_F32_SPLIT supports PCV_METRIC_DOT_REF with 64 <= dim <= 768 (got metric %d, dim %u)", (int)metric, dim);
  // the scan kernel holds a row in at most 12 x 32 chunks of 16 bytes: refuse here what no search could serve,
  // before a corpus is uploaded
  if ((store == PCV_F32 && dim > 1536) || (store == PCV_BF16 && dim > 3072))
    return fail(PCV_ERR_UNSUPPORTED, "dim=%u too large for %s rows (max %u): a row must fit 6144 bytes", dim,
                store == PCV_F32 ? "fp32" : "bf16", store == PCV_F32 ? 1536u : 3072u);
  if (metric != PCV_METRIC_DOT_REF && metric != PCV_METRIC_COSINE) return fail(PCV_ERR_INVALID, "bad metric %d", (int)metric);
  if (flags & ~(PCV_FLAG_PRENORMALISE | PCV_FLAG_NO_TIMING)) return fail(PCV_ERR_INVALID, "unknown flags 0x%x", flags);
  int ndev = 0;
  cudaError_t e = cudaGetDeviceCount(&ndev);
  if (e != cudaSuccess || ndev == 0)
    return fail(PCV_ERR_CUDA, "no CUDA device (%s); libperceive_cuda has no CPU fallback", cudaGetErrorString(e));
  if (device < 0 || device >= ndev) return fail(PCV_ERR_INVALID, "device %d outside [0,%d)", device, ndev);
  CU(cudaSetDevice(device));
  cudaDeviceProp prop;
  CU(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(PCV_ERR_UNSUPPORTED, "device %d is sm_%d%d; this library is built for sm_100a (B200) only", device, prop.major, prop.minor);
  pcv_index* ix = new (std::nothrow) pcv_index();
  if (!ix) return fail(PCV_ERR_OOM, "host allocation failed");
  ix->device = device;
  ix->dim = dim;
  const uint32_t epc = store == PCV_F32 ? 4 : 8;  // elements per 16 bytes
  ix->dim_padded = (dim + epc - 1) / epc * epc;
  ix->store = store;
  ix->metric = metric;
  ix->flags = flags;
  ix->timing = !(flags & PCV_FLAG_NO_TIMING) && !env_flag("PCV_NO_TIMING");
  ix->row_bytes = (size_t)ix->dim_padded * elem_size(store);
  ix->sm_count = prop.multiProcessorCount;
  auto bail = [&](cudaError_t ce, const char* what) {
    int32_t rc = fail(PCV_ERR_CUDA, "%s failed: %s", what, cudaGetErrorString(ce));
    pcv_index_destroy(ix);
    return rc;
  };
  if ((e = cudaStreamCreateWithFlags(&ix->own_stream, cudaStreamNonBlocking)) != cudaSuccess) return bail(e, "cudaStreamCreate");
  ix->stream = ix->own_stream;
  if ((e = cudaEventCreate(&ix->ev0)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaEventCreate(&ix->ev1)) != cudaSuccess) return bail(e, "cudaEventCreate");
  if ((e = cudaMalloc((void**)&ix->d_done, CTL_WORDS * sizeof(unsigned int))) != cudaSuccess) return bail(e, "cudaMalloc");
  if ((e = cudaMemset(ix->d_done, 0, CTL_WORDS * sizeof(unsigned int))) != cudaSuccess) return bail(e, "cudaMemset");
  ix->d_flags = ix->d_done + CTL_LOAD_FLAGS;
  *out = ix;
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_create_multi(const int32_t* devices, int32_t n_devices, uint32_t dim, pcv_dtype store, pcv_metric metric,
                               uint32_t flags, pcv_index** out) try {
  if (!out) return fail(PCV_ERR_INVALID, "null out");
  *out = nullptr;
  if (!devices || n_devices < 1 || n_devices > PCV_P2P_MAX_WORLD)
    return fail(PCV_ERR_INVALID, "n_devices=%d outside [1,%d] (or null device list)", n_devices, PCV_P2P_MAX_WORLD);
  for (int a = 0; a < n_devices; ++a)
    for (int b = a + 1; b < n_devices; ++b)
      if (devices[a] == devices[b])
        return fail(PCV_ERR_INVALID, "device %d listed twice: two shards whose exchange kernels wait on one another cannot share a GPU", devices[a]);
  pcv_index* mx = new (std::nothrow) pcv_index();
  if (!mx) return fail(PCV_ERR_OOM, "host allocation failed");
  mx->device = devices[0];
  mx->dim = dim;
  mx->store = store;
  mx->metric = metric;
  mx->flags = flags;
  auto bail = [&](int32_t rc) {
    const std::string keep = g_err;
    for (pcv_index* sh : mx->shards) pcv_index_destroy(sh);
    mx->shards.clear();
    if (mx->ev0) cudaEventDestroy(mx->ev0);
    delete mx;
    g_err = keep;
    return rc;
  };
  for (int r = 0; r < n_devices; ++r) {
    pcv_index* sh = nullptr;
    const int32_t rc = pcv_index_create(devices[r], dim, store, metric, flags, &sh);
    if (rc != PCV_OK) return bail(rc);
    sh->rank = r;
    sh->world = n_devices;
    mx->shards.push_back(sh);
  }
  mx->dim_padded = mx->shards[0]->dim_padded;
  mx->row_bytes = mx->shards[0]->row_bytes;
  mx->sm_count = mx->shards[0]->sm_count;
  mx->world = n_devices;
  // every shard must be able to store into every other shard's receive buffer
  for (int r = 0; r < n_devices; ++r) {
    cudaError_t e = cudaSetDevice(devices[r]);
    for (int q = 0; q < n_devices && e == cudaSuccess; ++q) {
      if (q == r) continue;
      int can = 0;
      e = cudaDeviceCanAccessPeer(&can, devices[r], devices[q]);
      if (e == cudaSuccess && !can)
        return bail(fail(PCV_ERR_UNSUPPORTED, "device %d cannot access device %d's memory (no peer path): a single-process index needs NVLink / PCIe peer access between all of its GPUs", devices[r], devices[q]));
      if (e == cudaSuccess) {
        e = cudaDeviceEnablePeerAccess(devices[q], 0);
        if (e == cudaErrorPeerAccessAlreadyEnabled) { (void)cudaGetLastError(); e = cudaSuccess; }
      }
    }
    if (e != cudaSuccess) return bail(fail(PCV_ERR_CUDA, "enabling peer access from device %d failed: %s", devices[r], cudaGetErrorString(e)));
  }
  cudaSetDevice(devices[0]);
  if (cudaEventCreateWithFlags(&mx->ev0, cudaEventDisableTiming) != cudaSuccess) return bail(fail(PCV_ERR_CUDA, "cudaEventCreate failed"));
  if (n_devices > 1) {  // one worker thread per shard, bound to its device
    mx->pool = new ShardPool();
    mx->pool->rc.assign(n_devices, PCV_OK);
    mx->pool->err.assign(n_devices, std::string());
    for (int r = 0; r < n_devices; ++r) mx->pool->threads.emplace_back(shard_worker_main, mx->pool, r, devices[r]);
  }
  multi_refresh_layout(mx);
  *out = mx;
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_destroy(pcv_index* ix) try {
  if (!ix) return PCV_OK;
  if (is_multi(ix)) {
    if (ix->pool) {
      {
        std::lock_guard<std::mutex> lk(ix->pool->m);
        ix->pool->quit.store(true, std::memory_order_release);
      }
      ix->pool->cv_work.notify_all();
      for (std::thread& t : ix->pool->threads) t.join();
      delete ix->pool;
      ix->pool = nullptr;
    }
    for (pcv_index* sh : ix->shards) {  // nobody may still be storing into a buffer that is about to go
      cudaSetDevice(sh->device);
      if (sh->stream) cudaStreamSynchronize(sh->stream);
    }
    for (pcv_index* sh : ix->shards) pcv_index_destroy(sh);
    ix->shards.clear();
    cudaSetDevice(ix->device);
    if (ix->ev0) cudaEventDestroy(ix->ev0);
    delete ix;
    return PCV_OK;
  }
  cudaSetDevice(ix->device);
  if (ix->own_stream) cudaStreamSynchronize(ix->own_stream);
  if (ix->comm && nccl_api().ok) nccl_api().CommDestroy(ix->comm);
  for (uint32_t r = 0; r < PCV_P2P_MAX_WORLD && !ix->p2p_in_process; ++r)
    if (ix->p2p_peer[r] && ix->p2p_peer[r] != ix->p2p_local) cudaIpcCloseMemHandle(ix->p2p_peer[r]);
  if (ix->p2p_local) cudaFree(ix->p2p_local);
  free_matrix(ix);
  ix->gemm.release();
  ix->partial.release();
  ix->q_pad.release();
  for (auto& r : ix->rs) { r.range_prefix.release(); r.ranges.release(); }
  ix->margin.release();
  ix->fb_list.release();
  ix->o_pack.release();
  ix->q_in.release();
  ix->cand_send.release();
  ix->cand_recv.release();
  ix->pin.release();
  if (ix->d_done) cudaFree(ix->d_done);
  if (ix->ev0) cudaEventDestroy(ix->ev0);
  if (ix->ev1) cudaEventDestroy(ix->ev1);
  if (ix->own_stream) cudaStreamDestroy(ix->own_stream);
  delete ix;
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_set_rows(pcv_index* ix, const float* rows, const int64_t* ids, const int64_t* source_ids, uint64_t n) try {
  if (!ix) return fail(PCV_ERR_INVALID, "null index");
  if (n && (!rows || !ids)) return fail(PCV_ERR_INVALID, "null rows/ids");
  if (n >= 0xfffffff0ull * (is_multi(ix) ? ix->shards.size() : 1)) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-16 rows on one shard");
  if (is_multi(ix)) {
    std::lock_guard<std::mutex> lk(ix->mu);
    return multi_set_rows(ix, rows, ids, source_ids, n);
  }
  NvtxRange nvtx("pcv_index_set_rows");
  std::lock_guard<std::mutex> lk(ix->mu);
  CU(cudaSetDevice(ix->device));
  CU(cudaStreamSynchronize(ix->stream));
  free_matrix(ix);
  if (n == 0) return PCV_OK;
  // order rows by (source_id, id): contiguous per-source segments (search.rs:115-148 groups by source)
  std::vector<uint64_t> perm(n);
  std::iota(perm.begin(), perm.end(), 0ull);
  bool sorted = true;
  for (uint64_t i = 1; i < n && sorted; ++i) {
    const int64_t sa = source_ids ? source_ids[i - 1] : 0, sb = source_ids ? source_ids[i] : 0;
    if (sa > sb || (sa == sb && ids[i - 1] > ids[i])) sorted = false;
  }
  if (!sorted)
    std::stable_sort(perm.begin(), perm.end(), [&](uint64_t a, uint64_t b) {
      const int64_t sa = source_ids ? source_ids[a] : 0, sb = source_ids ? source_ids[b] : 0;
      if (sa != sb) return sa < sb;
      return ids[a] < ids[b];
    });
  // everything is built on the side; the index only changes once every allocation and upload succeeded
  std::vector<int64_t> h_ids(n);
  std::vector<Segment> segs;
  for (uint64_t r = 0; r < n; ++r) {
    const uint64_t s = perm[r];
    h_ids[r] = ids[s];
    const int64_t src = source_ids ? source_ids[s] : 0;
    if (segs.empty() || segs.back().source_id != src) segs.push_back(Segment{src, r, r + 1});
    else segs.back().end = r + 1;
  }
  uint8_t* d_rows = nullptr;
  int64_t* d_ids = nullptr;
  uint32_t *d_lrank = nullptr, *d_rowof = nullptr;
  auto drop = [&]() {
    if (d_rows) cudaFree(d_rows);
    if (d_ids) cudaFree(d_ids);
    if (d_lrank) cudaFree(d_lrank);
    if (d_rowof) cudaFree(d_rowof);
  };
  cudaError_t e = cudaMalloc((void**)&d_rows, n * ix->row_bytes);
  if (e == cudaSuccess) e = cudaMalloc((void**)&d_ids, n * 8);
  if (e != cudaSuccess) {
    drop();
    (void)cudaGetLastError();
    return fail(e == cudaErrorMemoryAllocation ? PCV_ERR_OOM : PCV_ERR_CUDA, "device allocation for %llu rows failed: %s", (unsigned long long)n, cudaGetErrorString(e));
  }
  int32_t rc = upload_rows(ix, rows, sorted ? nullptr : perm.data(), n, d_rows, 0, n);
  if (rc == PCV_OK) rc = check_load_flags(ix);
  if (rc == PCV_OK) {
    e = cudaMemcpy(d_ids, h_ids.data(), n * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = fail(PCV_ERR_CUDA, "id upload failed: %s", cudaGetErrorString(e));
  }
  if (rc == PCV_OK) rc = make_rank_tables(h_ids, &d_lrank, &d_rowof);
  if (rc != PCV_OK) {
    cudaStreamSynchronize(ix->stream);
    drop();
    return rc;  // the index stays empty (free_matrix above), never half-built
  }
  ix->d_rows = d_rows;
  ix->d_ids = d_ids;
  ix->d_lrank_of_row = d_lrank;
  ix->d_row_of_lrank = d_rowof;
  ix->n_rows = n;
  ix->h_ids.swap(h_ids);
  ix->segs.swap(segs);
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_replace_source(pcv_index* ix, int64_t source_id, const float* rows, const int64_t* ids, uint64_t n) try {
  if (!ix) return fail(PCV_ERR_INVALID, "null index");
  if (n && (!rows || !ids)) return fail(PCV_ERR_INVALID, "null rows/ids");
  if (is_multi(ix)) {
    std::lock_guard<std::mutex> lk(ix->mu);
    return multi_replace_source(ix, source_id, rows, ids, n);
  }
  NvtxRange nvtx("pcv_index_replace_source");
  std::lock_guard<std::mutex> lk(ix->mu);
  CU(cudaSetDevice(ix->device));
  CU(cudaStreamSynchronize(ix->stream));
  if (ix->n_rows && ix->h_ids.empty()) return fail(PCV_ERR_STATE, "replace_source on a synthetic index");
  // locate the old segment (may be absent: search.rs:73-76 pushes a new source)
  uint64_t old_b = 0, old_e = 0;
  size_t pos = ix->segs.size();
  for (size_t i = 0; i < ix->segs.size(); ++i) {
    if (ix->segs[i].source_id == source_id) { old_b = ix->segs[i].begin; old_e = ix->segs[i].end; pos = i; break; }
    if (ix->segs[i].source_id > source_id) { old_b = old_e = ix->segs[i].begin; pos = i; break; }
  }
  if (pos == ix->segs.size()) old_b = old_e = ix->n_rows;
  const uint64_t old_n = ix->n_rows;
  const uint64_t new_n = old_n - (old_e - old_b) + n;
  if (new_n >= 0xfffffff0ull) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-16 rows on one shard");
  // new rows ordered by id
  std::vector<uint64_t> perm(n);
  std::iota(perm.begin(), perm.end(), 0ull);
  const bool sorted = std::is_sorted(ids, ids + n);
  if (!sorted) std::stable_sort(perm.begin(), perm.end(), [&](uint64_t a, uint64_t b) { return ids[a] < ids[b]; });
  // host bookkeeping of the new state, built on the side
  std::vector<int64_t> nids;
  nids.reserve(new_n);
  nids.insert(nids.end(), ix->h_ids.begin(), ix->h_ids.begin() + old_b);
  for (uint64_t r = 0; r < n; ++r) nids.push_back(ids[perm[r]]);
  nids.insert(nids.end(), ix->h_ids.begin() + old_e, ix->h_ids.end());
  const int64_t delta = (int64_t)n - (int64_t)(old_e - old_b);
  std::vector<Segment> nsegs;
  bool placed = false;
  for (const Segment& s0 : ix->segs) {
    if (s0.source_id == source_id) continue;  // replaced
    Segment s = s0;
    if (s.source_id > source_id) {
      if (!placed && n) { nsegs.push_back(Segment{source_id, old_b, old_b + n}); placed = true; }
      s.begin = (uint64_t)((int64_t)s.begin + delta);
      s.end = (uint64_t)((int64_t)s.end + delta);
    }
    nsegs.push_back(s);
  }
  if (!placed && n) nsegs.push_back(Segment{source_id, old_b, old_b + n});
  // device side of the new state: rows (kept segments copied device to device), ids, rank tables.  The
  // index is only switched over once ALL of it exists; any failure leaves the old state intact.
  uint8_t* d_new = nullptr;
  int64_t* d_new_ids = nullptr;
  uint32_t *d_lrank = nullptr, *d_rowof = nullptr;
  auto drop = [&]() {
    cudaStreamSynchronize(ix->stream);
    if (d_new) cudaFree(d_new);
    if (d_new_ids) cudaFree(d_new_ids);
    if (d_lrank) cudaFree(d_lrank);
    if (d_rowof) cudaFree(d_rowof);
  };
  cudaError_t e = cudaSuccess;
  if (new_n) {
    e = cudaMalloc((void**)&d_new, new_n * ix->row_bytes);
    if (e == cudaSuccess) e = cudaMalloc((void**)&d_new_ids, new_n * 8);
    if (e != cudaSuccess) {
      drop();
      (void)cudaGetLastError();
      return fail(e == cudaErrorMemoryAllocation ? PCV_ERR_OOM : PCV_ERR_CUDA, "device allocation for %llu rows failed: %s", (unsigned long long)new_n, cudaGetErrorString(e));
    }
  }
  // kept rows: [0, old_b) stay in place, [old_e, old_n) move to old_b + n; a split matrix is two planes
  const int planes = ix->store == PCV_F32_SPLIT ? 2 : 1;
  const size_t prb = ix->row_bytes / planes;  // bytes of one row of one plane
  for (int pl = 0; pl < planes && e == cudaSuccess; ++pl) {
    const uint8_t* src = ix->d_rows + (size_t)pl * old_n * prb;
    uint8_t* dst = d_new + (size_t)pl * new_n * prb;
    if (old_b) e = cudaMemcpyAsync(dst, src, old_b * prb, cudaMemcpyDeviceToDevice, ix->stream);
    if (e == cudaSuccess && old_n > old_e)
      e = cudaMemcpyAsync(dst + (old_b + n) * prb, src + old_e * prb, (old_n - old_e) * prb, cudaMemcpyDeviceToDevice, ix->stream);
  }
  if (e != cudaSuccess) { drop(); return fail(PCV_ERR_CUDA, "segment copy failed: %s", cudaGetErrorString(e)); }
  int32_t rc = upload_rows(ix, rows, sorted ? nullptr : perm.data(), n, d_new, old_b, new_n);
  if (rc == PCV_OK) rc = check_load_flags(ix);
  if (rc == PCV_OK && new_n) {
    e = cudaMemcpy(d_new_ids, nids.data(), new_n * 8, cudaMemcpyHostToDevice);
    if (e != cudaSuccess) rc = fail(PCV_ERR_CUDA, "id upload failed: %s", cudaGetErrorString(e));
  }
  if (rc == PCV_OK) rc = make_rank_tables(nids, &d_lrank, &d_rowof);
  if (rc == PCV_OK) {
    e = cudaStreamSynchronize(ix->stream);
    if (e != cudaSuccess) rc = fail(PCV_ERR_CUDA, "segment copy failed: %s", cudaGetErrorString(e));
  }
  if (rc != PCV_OK) { drop(); return rc; }
  // commit
  if (ix->d_rows) cudaFree(ix->d_rows);
  if (ix->d_ids) cudaFree(ix->d_ids);
  if (ix->d_lrank_of_row) cudaFree(ix->d_lrank_of_row);
  if (ix->d_row_of_lrank) cudaFree(ix->d_row_of_lrank);
  if (ix->d_xinv) { cudaFree(ix->d_xinv); ix->d_xinv = nullptr; }  // row norms follow the rows
  ix->d_rows = d_new;
  ix->d_ids = d_new_ids;
  ix->d_lrank_of_row = d_lrank;
  ix->d_row_of_lrank = d_rowof;
  ix->n_rows = new_n;
  ix->h_ids.swap(nids);
  ix->segs.swap(nsegs);
  ix->hidden_dirty = true;
  for (auto& r : ix->rs) { r.h_ranges.clear(); r.h_range_prefix.clear(); }
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_generate_synthetic(pcv_index* ix, uint64_t n, uint64_t seed, pcv_dist dist, uint64_t first_row) try {
  if (!ix) return fail(PCV_ERR_INVALID, "null index");
  if (dist != PCV_DIST_UNIT_SPHERE && dist != PCV_DIST_SCALED) return fail(PCV_ERR_INVALID, "bad distribution %d", (int)dist);
  if (is_multi(ix)) {  // shard r generates rows [first_row + n*r/N, first_row + n*(r+1)/N) of the same corpus
    std::lock_guard<std::mutex> lk(ix->mu);
    const uint64_t ns = ix->shards.size();
    for (uint64_t r = 0; r < ns; ++r) {
      const uint64_t b = n * r / ns, e = n * (r + 1) / ns;
      const int32_t rc = pcv_index_generate_synthetic(ix->shards[r], e - b, seed, dist, first_row + b);
      if (rc != PCV_OK) return rc;
    }
    multi_refresh_layout(ix);
    return PCV_OK;
  }
  if (n >= 0xfffffff0ull) return fail(PCV_ERR_UNSUPPORTED, "more than 2^32-16 rows on one shard");
  std::lock_guard<std::mutex> lk(ix->mu);
  CU(cudaSetDevice(ix->device));
  CU(cudaStreamSynchronize(ix->stream));
  free_matrix(ix);
  if (n == 0) return PCV_OK;
  {
    cudaError_t me = cudaMalloc((void**)&ix->d_rows, n * ix->row_bytes);
    if (me != cudaSuccess) {
      ix->d_rows = nullptr;
      (void)cudaGetLastError();
      return fail(me == cudaErrorMemoryAllocation ? PCV_ERR_OOM : PCV_ERR_CUDA, "device allocation for %llu rows failed: %s", (unsigned long long)n, cudaGetErrorString(me));
    }
  }
  ix->n_rows = n;
  ix->id_base = (int64_t)first_row + 1;
  ix->segs.push_back(Segment{0, 0, n});
  const int blocks = (int)std::min<uint64_t>((n + 7) / 8, (uint64_t)ix->sm_count * 16);
  if (ix->store == PCV_F32)
    pcv::synth_rows_kernel<float><<<blocks, 256, 0, ix->stream>>>((float*)ix->d_rows, nullptr, n, ix->dim, ix->dim_padded, seed, (int)dist, first_row, nullptr);
  else
    pcv::synth_rows_kernel<uint16_t><<<blocks, 256, 0, ix->stream>>>((uint16_t*)ix->d_rows, ix->store == PCV_F32_SPLIT ? (uint16_t*)ix->lo_plane() : nullptr,
                                                                     n, ix->dim, ix->dim_padded, seed, (int)dist, first_row, ix->d_done + CTL_XMAX2);
  CU(cudaGetLastError());
  CU(cudaStreamSynchronize(ix->stream));
  return PCV_OK;
} PCV_CATCH

int32_t pcv_synthetic_rows_host(uint64_t seed, pcv_dist dist, uint64_t first_row, uint64_t n, uint32_t dim, float* out) try {
  if (n && !out) return fail(PCV_ERR_INVALID, "null out");
  if (dim == 0 || dim > PCV_MAX_DIM) return fail(PCV_ERR_INVALID, "dim=%u outside [1,%u]", dim, PCV_MAX_DIM);
  for (uint64_t r = 0; r < n; ++r) pcv::synth_row_host(seed, (int)dist, first_row + r, dim, out + r * (size_t)dim);
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_set_hidden(pcv_index* ix, const int64_t* ids, uint64_t n) try {
  if (!ix) return fail(PCV_ERR_INVALID, "null index");
  if (n && !ids) return fail(PCV_ERR_INVALID, "null ids");
  if (is_multi(ix)) {  // ids not resident on a shard are remembered there and simply match nothing
    std::lock_guard<std::mutex> lk(ix->mu);
    ix->rows_epoch += 1;
    for (pcv_index* sh : ix->shards) {
      const int32_t rc = pcv_index_set_hidden(sh, ids, n);
      if (rc != PCV_OK) return rc;
    }
    return PCV_OK;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  ix->hidden_ids.assign(ids, ids + n);
  std::sort(ix->hidden_ids.begin(), ix->hidden_ids.end());
  ix->hidden_ids.erase(std::unique(ix->hidden_ids.begin(), ix->hidden_ids.end()), ix->hidden_ids.end());
  ix->hidden_dirty = true;
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_find_id(pcv_index* ix, int64_t id, uint64_t* out_row) try {
  if (!ix || !out_row) return fail(PCV_ERR_INVALID, "null argument");
  if (is_multi(ix)) {  // rows of a many-GPU handle are numbered shard after shard
    std::lock_guard<std::mutex> lk(ix->mu);
    *out_row = ~0ull;
    for (size_t r = 0; r < ix->shards.size(); ++r) {
      uint64_t local = ~0ull;
      const int32_t rc = pcv_index_find_id(ix->shards[r], id, &local);
      if (rc != PCV_OK) return rc;
      if (local != ~0ull) { *out_row = ix->shard_row0[r] + local; break; }
    }
    return PCV_OK;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  const int64_t r = find_row(ix, id);
  *out_row = r < 0 ? ~0ull : (uint64_t)r;
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_get_rows(pcv_index* ix, uint64_t first_row, uint64_t n, float* out_rows, int64_t* out_ids, int64_t* out_source_ids) try {
  if (!ix) return fail(PCV_ERR_INVALID, "null index");
  if (is_multi(ix)) {
    std::lock_guard<std::mutex> lk(ix->mu);
    return multi_get_rows(ix, first_row, n, out_rows, out_ids, out_source_ids);
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  if (first_row + n > ix->n_rows) return fail(PCV_ERR_INVALID, "rows [%llu,%llu) outside [0,%llu)", (unsigned long long)first_row, (unsigned long long)(first_row + n), (unsigned long long)ix->n_rows);
  if (n == 0) return PCV_OK;
  CU(cudaSetDevice(ix->device));
  CU(cudaStreamSynchronize(ix->stream));
  if (out_rows && ix->store == PCV_F32_SPLIT) {
    // the two planes hold the fp32 value exactly: x = (hi << 16) | lo
    const size_t prb = (size_t)ix->dim_padded * 2;
    std::vector<uint16_t> hi(n * ix->dim_padded), lo(n * ix->dim_padded);
    CU(cudaMemcpy(hi.data(), ix->hi_plane() + first_row * prb, n * prb, cudaMemcpyDeviceToHost));
    CU(cudaMemcpy(lo.data(), ix->lo_plane() + first_row * prb, n * prb, cudaMemcpyDeviceToHost));
    for (uint64_t r = 0; r < n; ++r)
      for (uint32_t c = 0; c < ix->dim; ++c) {
        const uint32_t bits = pcv::split_join_bits(hi[r * ix->dim_padded + c], lo[r * ix->dim_padded + c]);
        memcpy(&out_rows[r * ix->dim + c], &bits, 4);
      }
  } else if (out_rows) {
    std::vector<uint8_t> raw(n * ix->row_bytes);
    CU(cudaMemcpy(raw.data(), ix->d_rows + first_row * ix->row_bytes, raw.size(), cudaMemcpyDeviceToHost));
    for (uint64_t r = 0; r < n; ++r)
      for (uint32_t c = 0; c < ix->dim; ++c) {
        const uint16_t* h16 = reinterpret_cast<const uint16_t*>(raw.data() + r * ix->row_bytes);
        if (ix->store == PCV_F32) out_rows[r * ix->dim + c] = reinterpret_cast<const float*>(raw.data() + r * ix->row_bytes)[c];
        else out_rows[r * ix->dim + c] = pcv::bf16_to_f32(h16[c]);
      }
  }
  for (uint64_t r = 0; r < n; ++r) {
    const uint64_t row = first_row + r;
    if (out_ids) out_ids[r] = ix->h_ids.empty() ? ix->id_base + (int64_t)row : ix->h_ids[row];
    if (out_source_ids) {
      int64_t s = 0;
      for (const Segment& sg : ix->segs)
        if (row >= sg.begin && row < sg.end) { s = sg.source_id; break; }
      out_source_ids[r] = s;
    }
  }
  return PCV_OK;
} PCV_CATCH

int32_t pcv_search_device(pcv_index* ix, const float* d_queries, uint32_t n_queries, uint32_t k, const int64_t* sources,
                          uint32_t n_sources, int64_t* d_out_ids, float* d_out_scores, float* d_out_sims, uint32_t* d_out_counts) try {
  int32_t rc = validate_search(ix, d_queries, n_queries, k, sources, n_sources, d_out_ids, d_out_scores);
  if (rc != PCV_OK) return rc;
  if (n_queries == 0) return PCV_OK;
  std::lock_guard<std::mutex> lk(ix->mu);
  CU(cudaSetDevice(ix->device));
  if (is_multi(ix))  // buffers live on the first device of the handle
    return multi_search_device_locked(ix, d_queries, n_queries, k, sources, n_sources, d_out_ids, d_out_scores, d_out_sims, d_out_counts);
  return search_device_locked(ix, d_queries, n_queries, k, sources, n_sources, d_out_ids, d_out_scores, d_out_sims, d_out_counts);
} PCV_CATCH

int32_t pcv_search(pcv_index* ix, const float* queries, uint32_t n_queries, uint32_t k, const int64_t* sources,
                   uint32_t n_sources, int64_t* out_ids, float* out_scores, float* out_sims, uint32_t* out_counts) try {
  int32_t rc = validate_search(ix, queries, n_queries, k, sources, n_sources, out_ids, out_scores);
  if (rc != PCV_OK) return rc;
  if (n_queries == 0) return PCV_OK;
  NvtxRange nvtx("pcv_search (host buffers: H2D, search, D2H)");
  const size_t nq = (size_t)n_queries * ix->dim;
  if (const size_t bad = first_nonfinite(queries, nq); bad < nq)
    return fail(PCV_ERR_NONFINITE, "non-finite value in query %zu", bad / ix->dim);
  if (ix->metric == PCV_METRIC_COSINE)
    for (uint32_t q = 0; q < n_queries; ++q) {
      float ss = 0.0f;  // fp32, as the kernels sum it: a norm that underflows to 0 (or overflows) is as bad as a zero vector
      for (uint32_t c = 0; c < ix->dim; ++c) ss += queries[(size_t)q * ix->dim + c] * queries[(size_t)q * ix->dim + c];
      if (!(ss > 0.0f && std::isfinite(ss)))
        return fail(PCV_ERR_ZERO_NORM, "query %u has a zero (or unrepresentable) norm under the cosine metric", q);
    }
  std::lock_guard<std::mutex> lk(ix->mu);
  CU(cudaSetDevice(ix->device));
  pcv_index* const mx = is_multi(ix) ? ix : nullptr;
  if (mx) ix = mx->shards[0];  // staging, stream and result delivery are the first shard's
  const size_t nk = (size_t)n_queries * k;
  // pinned staging [queries | ids | scores | sims | counts] and a device block with the same output
  // layout, so the results come back in ONE device-to-host copy
  const size_t off_ids = (nq * 4 + 15) & ~(size_t)15;
  const size_t off_scores = off_ids + nk * 8;
  const size_t off_sims = off_scores + nk * 4;
  const size_t off_counts = off_sims + nk * 4;
  const size_t pin_bytes = off_counts + (size_t)n_queries * 4;
  const size_t out_bytes = pin_bytes - off_ids;
  CU(ix->pin.reserve(pin_bytes));
  CU(ix->q_in.reserve(nq));
  // A few hundred bytes of results (the reference's one-query call) are stored by the kernel straight
  // into the pinned block over PCIe — it is device-mapped under unified addressing — which saves the
  // copy-engine hop after the scan; larger result sets go through one device block and one copy.
  uint8_t* d_out = nullptr;
  bool zero_copy = out_bytes <= 4096 && !env_flag("PCV_NO_ZERO_COPY_RESULTS");
  if (zero_copy) {
    void* mapped = nullptr;
    if (cudaHostGetDevicePointer(&mapped, ix->pin.p + off_ids, 0) == cudaSuccess && mapped) d_out = static_cast<uint8_t*>(mapped);
    else { (void)cudaGetLastError(); zero_copy = false; }
  }
  if (!zero_copy) {
    CU(ix->o_pack.reserve(out_bytes));
    d_out = ix->o_pack.p;
  }
  int64_t* d_ids = reinterpret_cast<int64_t*>(d_out);
  float* d_scores = reinterpret_cast<float*>(d_out + (off_scores - off_ids));
  float* d_sims = reinterpret_cast<float*>(d_out + (off_sims - off_ids));
  uint32_t* d_counts = reinterpret_cast<uint32_t*>(d_out + (off_counts - off_ids));
  memcpy(ix->pin.p, queries, nq * 4);
  // Three shortcuts for the single-query call were built, measured and dropped (profiles/README.md, round 2):
  // the query passed through the kernel parameters (no copy, no padding kernel: 249.4 us against 247.6 us on config 2,
  // 42.2 us against 40.7 us on config 1 — a 3.4 KB parameter block and per-lane reads from the constant bank cost more
  // than an asynchronous copy that overlaps the launch anyway),
  // the kernels reading the query straight out of the pinned block over PCIe (281 us per query end to end against
  // 249 us with this copy — 148 CTAs each fetching their slice from host memory cost far more than the copy-engine
  // hop they save), and replaying the whole search as one captured CUDA graph (253 us against 249 us on config 2,
  // 46 us against 43 us on config 1: launching a three-node graph costs more than the copy and the launch it replaces).
  const float* d_q_src = ix->q_in.p;
  CU(cudaMemcpyAsync(ix->q_in.p, ix->pin.p, nq * 4, cudaMemcpyHostToDevice, ix->stream));
  rc = mx ? multi_search_device_locked(mx, d_q_src, n_queries, k, sources, n_sources, d_ids, d_scores, d_sims, d_counts)
          : search_device_locked(ix, d_q_src, n_queries, k, sources, n_sources, d_ids, d_scores, d_sims, d_counts);
  if (rc != PCV_OK) { cudaStreamSynchronize(ix->stream); return rc; }
  if (!zero_copy) CU(cudaMemcpyAsync(ix->pin.p + off_ids, d_out, out_bytes, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  memcpy(out_ids, ix->pin.p + off_ids, nk * 8);
  memcpy(out_scores, ix->pin.p + off_scores, nk * 4);
  if (out_sims) memcpy(out_sims, ix->pin.p + off_sims, nk * 4);
  if (out_counts) memcpy(out_counts, ix->pin.p + off_counts, (size_t)n_queries * 4);
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_best_chunks(pcv_index* ix, const float* query, const float* chunks, uint32_t n_chunks,
                              const uint32_t* doc_chunk_end, uint32_t n_docs, int32_t* out_best_chunk,
                              float* out_best_score, float* out_scores) try {
  if (!ix) return fail(PCV_ERR_INVALID, "null index");
  if (n_docs == 0) return PCV_OK;
  if (!query || !doc_chunk_end || !out_best_chunk) return fail(PCV_ERR_INVALID, "null argument");
  if (n_chunks && !chunks) return fail(PCV_ERR_INVALID, "null chunks");
  if (n_docs > 65535u || n_chunks > (1u << 24)) return fail(PCV_ERR_UNSUPPORTED, "%u documents / %u chunks in one call", n_docs, n_chunks);
  uint32_t prev = 0;
  for (uint32_t d = 0; d < n_docs; ++d) {
    if (doc_chunk_end[d] < prev || doc_chunk_end[d] > n_chunks)
      return fail(PCV_ERR_INVALID, "doc_chunk_end[%u]=%u is not a cumulative count within %u chunks", d, doc_chunk_end[d], n_chunks);
    prev = doc_chunk_end[d];
  }
  const uint32_t dim = ix->dim;
  const size_t nc = (size_t)n_chunks * dim;
  if (first_nonfinite(query, dim) < dim) return fail(PCV_ERR_NONFINITE, "non-finite value in the query");
  // the reference panics on a NaN score (highlight.rs:124 partial_cmp().unwrap())
  if (const size_t bad = first_nonfinite(chunks, nc); bad < nc)
    return fail(PCV_ERR_NONFINITE, "non-finite value in chunk %zu", bad / dim);
  std::lock_guard<std::mutex> lk(ix->mu);
  if (is_multi(ix)) ix = ix->shards[0];  // needs no rows: any shard will do
  CU(cudaSetDevice(ix->device));
  // pinned [query | chunks | ends] -> device; device [scores | best | best_score] -> pinned
  const size_t in_f = dim + nc;
  const size_t in_bytes = in_f * 4 + (size_t)n_docs * 4;
  const size_t out_bytes = (size_t)n_chunks * 4 + (size_t)n_docs * 8;
  CU(ix->pin.reserve(in_bytes + out_bytes));
  CU(ix->q_in.reserve(in_f + n_docs));
  CU(ix->o_pack.reserve(out_bytes));
  memcpy(ix->pin.p, query, (size_t)dim * 4);
  if (nc) memcpy(ix->pin.p + (size_t)dim * 4, chunks, nc * 4);
  memcpy(ix->pin.p + in_f * 4, doc_chunk_end, (size_t)n_docs * 4);
  CU(cudaMemcpyAsync(ix->q_in.p, ix->pin.p, in_bytes, cudaMemcpyHostToDevice, ix->stream));
  float* d_scores = reinterpret_cast<float*>(ix->o_pack.p);
  int32_t* d_best = reinterpret_cast<int32_t*>(d_scores + n_chunks);
  float* d_best_score = reinterpret_cast<float*>(d_best + n_docs);
  pcv::best_chunk_kernel<<<n_docs, 128, 0, ix->stream>>>(ix->q_in.p, ix->q_in.p + dim, dim,
                                                         reinterpret_cast<const uint32_t*>(ix->q_in.p + in_f), d_scores,
                                                         d_best, d_best_score);
  CU(cudaGetLastError());
  uint8_t* h_out = ix->pin.p + in_bytes;
  CU(cudaMemcpyAsync(h_out, ix->o_pack.p, out_bytes, cudaMemcpyDeviceToHost, ix->stream));
  CU(cudaStreamSynchronize(ix->stream));
  ix->last_launches = 1;
  if (out_scores && n_chunks) memcpy(out_scores, h_out, (size_t)n_chunks * 4);
  memcpy(out_best_chunk, h_out + (size_t)n_chunks * 4, (size_t)n_docs * 4);
  if (out_best_score) memcpy(out_best_score, h_out + (size_t)n_chunks * 4 + (size_t)n_docs * 4, (size_t)n_docs * 4);
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_set_stream(pcv_index* ix, void* cuda_stream) try {
  if (!ix) return fail(PCV_ERR_INVALID, "null index");
  if (is_multi(ix)) {  // the caller's stream belongs to the first device; the other shards keep their own
    std::lock_guard<std::mutex> lk(ix->mu);
    return pcv_index_set_stream(ix->shards[0], cuda_stream);
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  CU(cudaSetDevice(ix->device));
  CU(cudaStreamSynchronize(ix->stream));
  ix->stream = cuda_stream ? (cudaStream_t)cuda_stream : ix->own_stream;
  ix->ev_valid = false;
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_synchronize(pcv_index* ix) try {
  if (!ix) return fail(PCV_ERR_INVALID, "null index");
  if (is_multi(ix)) {
    for (pcv_index* sh : ix->shards) {
      const int32_t rc = pcv_index_synchronize(sh);
      if (rc != PCV_OK) return rc;
    }
    return PCV_OK;
  }
  CU(cudaSetDevice(ix->device));
  CU(cudaStreamSynchronize(ix->stream));
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_stats(pcv_index* ix, pcv_stats* out) try {
  if (!ix || !out) return fail(PCV_ERR_INVALID, "null argument");
  if (is_multi(ix)) {  // the first shard's view (it delivers the results), sizes and launches summed over the shards
    std::lock_guard<std::mutex> lk(ix->mu);
    int32_t rc = pcv_index_stats(ix->shards[0], out);
    for (size_t r = 1; r < ix->shards.size() && rc == PCV_OK; ++r) {
      pcv_stats st;
      rc = pcv_index_stats(ix->shards[r], &st);
      out->n_rows += st.n_rows;
      out->matrix_bytes += st.matrix_bytes;
      out->last_scan_bytes += st.last_scan_bytes;
      out->last_launches += st.last_launches;
      out->last_search_ms = std::max(out->last_search_ms, st.last_search_ms);
      out->last_fallback_queries = std::max(out->last_fallback_queries, st.last_fallback_queries);
      out->n_sources = std::max(out->n_sources, st.n_sources);
    }
    out->n_rows_global = out->n_rows;
    return rc;
  }
  std::lock_guard<std::mutex> lk(ix->mu);
  memset(out, 0, sizeof *out);
  out->n_rows = ix->n_rows;
  out->n_rows_global = ix->n_rows;
  out->dim = ix->dim;
  out->dim_padded = ix->dim_padded;
  out->n_sources = (uint32_t)ix->segs.size();
  out->dtype = (uint32_t)ix->store;
  out->matrix_bytes = ix->n_rows * ix->row_bytes;
  out->last_scan_bytes = ix->last_scan_bytes;
  out->last_launches = ix->last_launches;
  out->last_kernel = ix->last_kernel;
  out->sm_count = (uint32_t)ix->sm_count;
  out->world = (uint32_t)ix->world;
  out->rank = (uint32_t)ix->rank;
  if (ix->ev_valid) {
    CU(cudaSetDevice(ix->device));
    CU(cudaEventSynchronize(ix->ev1));
    float ms = 0.f;
    CU(cudaEventElapsedTime(&ms, ix->ev0, ix->ev1));
    out->last_search_ms = ms;
  }
  if (ix->searched && ix->last_used_filter) {  // how many queries the exact fallback scan took, once the search has completed
    unsigned int fb = 0;
    CU(cudaSetDevice(ix->device));
    if (!ix->ev_valid) CU(cudaStreamSynchronize(ix->stream));
    CU(cudaMemcpy(&fb, ix->d_done + CTL_FB_COUNT, sizeof fb, cudaMemcpyDeviceToHost));
    out->last_fallback_queries = fb;
  }
  return PCV_OK;
} PCV_CATCH

int32_t pcv_comm_unique_id(uint8_t out_id[128]) try {
  if (!out_id) return fail(PCV_ERR_INVALID, "null out_id");
  static_assert(sizeof(ncclUniqueId) == 128, "ncclUniqueId size");
  if (!nccl_api().ok) return fail(PCV_ERR_NCCL, "%s", nccl_api().err.c_str());
  ncclUniqueId id;
  NC(nccl_api().GetUniqueId(&id));
  memcpy(out_id, &id, 128);
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_attach_comm(pcv_index* ix, const uint8_t id_bytes[128], int32_t rank, int32_t world) try {
  if (!ix || !id_bytes) return fail(PCV_ERR_INVALID, "null argument");
  if (is_multi(ix) || ix->p2p_in_process) return fail(PCV_ERR_STATE, "a single-process many-GPU handle wires its own exchange; communicators and IPC handles are for one handle per process");
  if (world < 1 || world > 32 || rank < 0 || rank >= world) return fail(PCV_ERR_INVALID, "bad rank %d / world %d (max 32)", rank, world);
  std::lock_guard<std::mutex> lk(ix->mu);
  if (ix->comm) return fail(PCV_ERR_STATE, "communicator already attached");
  CU(cudaSetDevice(ix->device));
  ncclUniqueId id;
  memcpy(&id, id_bytes, 128);
  if (!nccl_api().ok) return fail(PCV_ERR_NCCL, "%s", nccl_api().err.c_str());
  NC(nccl_api().CommInitRank(&ix->comm, world, id, rank));
  ix->rank = rank;
  ix->world = world;
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_p2p_export(pcv_index* ix, int32_t world, uint32_t max_records, uint8_t out_handle[64]) try {
  if (!ix || !out_handle) return fail(PCV_ERR_INVALID, "null argument");
  if (is_multi(ix) || ix->p2p_in_process) return fail(PCV_ERR_STATE, "a single-process many-GPU handle wires its own exchange; communicators and IPC handles are for one handle per process");
  if (world < 2 || world > PCV_P2P_MAX_WORLD) return fail(PCV_ERR_INVALID, "world %d outside [2,%d]", world, PCV_P2P_MAX_WORLD);
  if (max_records == 0 || max_records > (1u << 24)) return fail(PCV_ERR_INVALID, "max_records %u outside [1,2^24]", max_records);
  static_assert(sizeof(cudaIpcMemHandle_t) == 64, "cudaIpcMemHandle_t size");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (ix->p2p_local) return fail(PCV_ERR_STATE, "peer buffer already exported");
  CU(cudaSetDevice(ix->device));
  const uint32_t cap = (max_records + 1u) & ~1u;
  const size_t bytes = 2 * pcv::p2p_half_bytes((uint32_t)world, cap);
  CU(cudaMalloc((void**)&ix->p2p_local, bytes));
  CU(cudaMemset(ix->p2p_local, 0, bytes));  // flags start at epoch 0; searches count from 1
  cudaIpcMemHandle_t h;
  CU(cudaIpcGetMemHandle(&h, ix->p2p_local));
  memcpy(out_handle, &h, 64);
  ix->p2p_world = (uint32_t)world;
  ix->p2p_cap = cap;
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_p2p_attach(pcv_index* ix, const uint8_t* handles, int32_t rank, int32_t world) try {
  if (!ix || !handles) return fail(PCV_ERR_INVALID, "null argument");
  if (is_multi(ix) || ix->p2p_in_process) return fail(PCV_ERR_STATE, "a single-process many-GPU handle wires its own exchange; communicators and IPC handles are for one handle per process");
  std::lock_guard<std::mutex> lk(ix->mu);
  if (!ix->p2p_local) return fail(PCV_ERR_STATE, "pcv_index_p2p_export has not been called");
  if (ix->p2p_attached) return fail(PCV_ERR_STATE, "peer buffers already attached");
  if (world != (int32_t)ix->p2p_world || rank < 0 || rank >= world)
    return fail(PCV_ERR_INVALID, "bad rank %d / world %d (exported for world %u)", rank, world, ix->p2p_world);
  if (ix->comm && (ix->rank != rank || ix->world != world))
    return fail(PCV_ERR_STATE, "rank/world differ from the attached NCCL communicator");
  CU(cudaSetDevice(ix->device));
  for (int r = 0; r < world; ++r) {
    if (r == rank) { ix->p2p_peer[r] = ix->p2p_local; continue; }
    cudaIpcMemHandle_t h;
    memcpy(&h, handles + (size_t)r * 64, 64);
    void* ptr = nullptr;
    cudaError_t e = cudaIpcOpenMemHandle(&ptr, h, cudaIpcMemLazyEnablePeerAccess);
    if (e != cudaSuccess) {
      for (int q = 0; q < r; ++q)
        if (q != rank && ix->p2p_peer[q]) { cudaIpcCloseMemHandle(ix->p2p_peer[q]); ix->p2p_peer[q] = nullptr; }
      ix->p2p_peer[rank] = nullptr;
      cudaGetLastError();  // the failure is reported through the status code; do not leave it pending
      return fail(PCV_ERR_CUDA, "cudaIpcOpenMemHandle(rank %d) failed: %s", r, cudaGetErrorString(e));
    }
    ix->p2p_peer[r] = static_cast<uint8_t*>(ptr);
  }
  ix->rank = rank;
  ix->world = world;
  ix->p2p_attached = true;
  return PCV_OK;
} PCV_CATCH

int32_t pcv_index_p2p_detach(pcv_index* ix) try {
  if (!ix) return fail(PCV_ERR_INVALID, "null index");
  if (is_multi(ix) || ix->p2p_in_process) return fail(PCV_ERR_STATE, "a single-process many-GPU handle wires its own exchange; communicators and IPC handles are for one handle per process");
  std::lock_guard<std::mutex> lk(ix->mu);
  CU(cudaSetDevice(ix->device));
  CU(cudaStreamSynchronize(ix->stream));
  for (uint32_t r = 0; r < PCV_P2P_MAX_WORLD; ++r) {
    if (ix->p2p_peer[r] && ix->p2p_peer[r] != ix->p2p_local) cudaIpcCloseMemHandle(ix->p2p_peer[r]);
    ix->p2p_peer[r] = nullptr;
  }
  ix->p2p_attached = false;  // the exported buffer stays allocated; searches use the NCCL exchange again
  return PCV_OK;
} PCV_CATCH

int32_t pcv_merge_candidates_device(pcv_index* ix, const float* d_sims, const int64_t* d_ids, uint32_t n_lists,
                                    uint32_t n_queries, uint32_t k, int64_t* d_out_ids, float* d_out_scores,
                                    float* d_out_sims, uint32_t* d_out_counts) try {
  if (!ix || !d_sims || !d_ids || !d_out_ids) return fail(PCV_ERR_INVALID, "null argument");
  if (n_lists == 0 || n_lists > 32) return fail(PCV_ERR_INVALID, "n_lists=%u outside [1,32]", n_lists);
  if (k == 0 || k > PCV_MAX_K) return fail(PCV_ERR_INVALID, "k=%u outside [1,%u]", k, PCV_MAX_K);
  if (n_queries == 0) return PCV_OK;
  std::lock_guard<std::mutex> lk(ix->mu);
  if (is_multi(ix)) ix = ix->shards[0];
  CU(cudaSetDevice(ix->device));
  const uint32_t wpb = 4;
  const size_t stride = (size_t)n_queries * k;
  pcv::merge_candidates_kernel<<<(n_queries + wpb - 1) / wpb, wpb * 32, 0, ix->stream>>>(
      d_sims, stride, d_ids, stride, n_lists, n_queries, k, ix->dim, ix->metric == PCV_METRIC_COSINE ? 1 : 0,
      d_out_ids, d_out_scores, d_out_sims, d_out_counts);
  CU(cudaGetLastError());
  return PCV_OK;
} PCV_CATCH

int32_t pcv_decode_embedding(const uint8_t* blob, size_t blob_len, float* out, size_t out_cap, size_t* out_dim) try {
  if (blob_len && !blob) return fail(PCV_ERR_INVALID, "null blob");
  if (blob_len % 4 != 0) return fail(PCV_ERR_INVALID, "embedding blob of %zu bytes is not a multiple of 4", blob_len);
  const size_t d = blob_len / 4;
  if (out_dim) *out_dim = d;
  if (d > out_cap) return fail(PCV_ERR_INVALID, "output capacity %zu < %zu", out_cap, d);
  if (d && !out) return fail(PCV_ERR_INVALID, "null out");
  for (size_t i = 0; i < d; ++i) {
    const uint32_t b = (uint32_t)blob[4 * i] | ((uint32_t)blob[4 * i + 1] << 8) | ((uint32_t)blob[4 * i + 2] << 16) | ((uint32_t)blob[4 * i + 3] << 24);
    memcpy(out + i, &b, 4);
  }
  return PCV_OK;
} PCV_CATCH

int32_t pcv_decode_embeddings_bulk(const uint8_t* blobs, const size_t* lens, size_t n, size_t dim, float* out) try {
  if (n && (!blobs || !lens || !out)) return fail(PCV_ERR_INVALID, "null argument");
  if (dim == 0 || dim > PCV_MAX_DIM) return fail(PCV_ERR_INVALID, "dim=%zu outside [1,%u]", dim, PCV_MAX_DIM);
  size_t off = 0;
  for (size_t i = 0; i < n; ++i) {
    if (lens[i] != dim * 4)
      return fail(PCV_ERR_INVALID, "embedding %zu is %zu bytes, expected %zu (%zu floats)", i, lens[i], dim * 4, dim);
    const uint8_t* b = blobs + off;
    float* o = out + i * dim;
    for (size_t c = 0; c < dim; ++c) {
      const uint32_t w = (uint32_t)b[4 * c] | ((uint32_t)b[4 * c + 1] << 8) | ((uint32_t)b[4 * c + 2] << 16) | ((uint32_t)b[4 * c + 3] << 24);
      memcpy(o + c, &w, 4);
    }
    off += lens[i];
  }
  return PCV_OK;
} PCV_CATCH

int32_t pcv_encode_embedding(const float* v, size_t dim, uint8_t* out, size_t out_cap) try {
  if (dim && (!v || !out)) return fail(PCV_ERR_INVALID, "null argument");
  if (out_cap < dim * 4) return fail(PCV_ERR_INVALID, "output capacity %zu < %zu", out_cap, dim * 4);
  for (size_t i = 0; i < dim; ++i) {
    uint32_t b;
    memcpy(&b, v + i, 4);
    out[4 * i] = (uint8_t)b;
    out[4 * i + 1] = (uint8_t)(b >> 8);
    out[4 * i + 2] = (uint8_t)(b >> 16);
    out[4 * i + 3] = (uint8_t)(b >> 24);
  }
  return PCV_OK;
} PCV_CATCH

float pcv_distance_from_dot(float dot, uint32_t dim) { return pcv::ref_distance(dot, dim); }

}  // extern "C"
