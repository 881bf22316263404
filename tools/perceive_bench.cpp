// perceive_bench.cpp — compiled twin of `perceive bench` (rust/perceive-cli/cmd/bench.rs, unbuilt here: no
// Rust toolchain), over the same C ABI the Rust crate binds.  Same arguments, workloads, seeds, generator
// and JSON keys:
//     perceive_bench [--config c1..c5] [--gpus N] [--steps K] [--warmup W] [--seed S] [--rows R] [--dump-ids]
// One process; --gpus N > 1 shards the corpus over devices 0..N-1 behind ONE handle (pcv_index_create_multi).
// Build: g++ -std=c++17 -O2 -I include tools/perceive_bench.cpp -L perceive_b200 -lperceive_cuda
// (tests/test_bench_twin.py builds and runs it on config 1 and holds its hits to the Python API's).
#include <algorithm>
#include <chrono>
#include <cstdint>
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <stdexcept>
#include <string>
#include <vector>

#include "perceive_cuda.h"

namespace {

struct Workload {
  const char* name;
  uint64_t rows;
  uint32_t dim;
  pcv_dtype store;
  uint32_t batch, k;
  pcv_metric metric;
  pcv_dist dist;
  const char* text;
};

const Workload kWorkloads[] = {
    {"c1", 10000, 384, PCV_F32, 1, 10, PCV_METRIC_DOT_REF, PCV_DIST_UNIT_SPHERE, "1 query vs 10kx384 fp32 docs, top-10 (BASELINE configs[0]; L2-resident)"},
    {"c2", 1000000, 384, PCV_F32, 1, 10, PCV_METRIC_DOT_REF, PCV_DIST_UNIT_SPHERE, "1 query vs 1Mx384 fp32 docs, top-10 (BASELINE configs[1])"},
    {"c3", 10000000, 384, PCV_BF16, 1024, 100, PCV_METRIC_DOT_REF, PCV_DIST_UNIT_SPHERE, "batch 1024 queries vs 10Mx384 bf16 docs, top-100 (BASELINE configs[2])"},
    {"c4", 100000000, 384, PCV_F32_SPLIT, 256, 10, PCV_METRIC_DOT_REF, PCV_DIST_UNIT_SPHERE, "batch 256 queries vs 100Mx384 fp32 docs, top-10, row-sharded (BASELINE configs[3])"},
    {"c5", 50000000, 768, PCV_BF16, 4096, 50, PCV_METRIC_COSINE, PCV_DIST_SCALED, "batch 4096 queries vs 50Mx768 bf16 docs, top-50, cosine (BASELINE configs[4])"},
};

void check(int32_t rc, const char* what) {
  if (rc != PCV_OK) throw std::runtime_error(std::string(what) + ": " + pcv_last_error());
}

}  // namespace

int main(int argc, char** argv) try {
  std::string config = "c2";
  int gpus = 1;
  uint32_t steps = 20, warmup = 5;
  uint64_t seed = 1, rows_override = 0;
  bool dump_ids = false;
  for (int i = 1; i < argc; ++i) {
    const std::string a = argv[i];
    auto val = [&]() -> const char* {
      if (i + 1 >= argc) throw std::runtime_error("missing value after " + a);
      return argv[++i];
    };
    if (a == "--config") config = val();
    else if (a == "--gpus") gpus = std::atoi(val());
    else if (a == "--steps") steps = (uint32_t)std::atoi(val());
    else if (a == "--warmup") warmup = (uint32_t)std::atoi(val());
    else if (a == "--seed") seed = std::strtoull(val(), nullptr, 10);
    else if (a == "--rows") rows_override = std::strtoull(val(), nullptr, 10);
    else if (a == "--dump-ids") dump_ids = true;
    else throw std::runtime_error("unknown argument " + a);
  }
  const Workload* w = nullptr;
  for (const Workload& c : kWorkloads)
    if (config == c.name) w = &c;
  if (!w) throw std::runtime_error("unknown workload " + config + " (c1..c5)");
  const uint64_t rows = rows_override ? rows_override : w->rows;

  std::vector<int32_t> devices;
  for (int d = 0; d < std::max(gpus, 1); ++d) devices.push_back(d);
  pcv_index* ix = nullptr;
  check(pcv_index_create_multi(devices.data(), (int32_t)devices.size(), w->dim, w->store, w->metric, 0, &ix), "pcv_index_create_multi");
  check(pcv_index_generate_synthetic(ix, rows, seed, w->dist, 0), "pcv_index_generate_synthetic");

  // a fresh query batch every step, drawn round-robin from a bounded pool (<= 64 MB)
  const size_t dim = w->dim, B = w->batch, k = w->k;
  const size_t total = (size_t)steps + warmup;
  const size_t pool = std::max<size_t>(1, std::min<size_t>(total, ((size_t)64 << 20) / (B * dim * 4)));
  std::vector<float> queries(pool * B * dim);
  check(pcv_synthetic_rows_host(seed + 1, w->dist, 0, pool * B, w->dim, queries.data()), "pcv_synthetic_rows_host");
  std::vector<int64_t> ids(B * k);
  std::vector<float> scores(B * k), sims(B * k);
  std::vector<uint32_t> counts(B);
  auto search = [&](size_t step) {
    const float* q = queries.data() + (step % pool) * B * dim;
    check(pcv_search(ix, q, (uint32_t)B, (uint32_t)k, nullptr, 0, ids.data(), scores.data(), sims.data(), counts.data()), "pcv_search");
  };

  for (uint32_t s = 0; s < warmup; ++s) search(s);
  double device_ms = 0.0;
  uint32_t launches = 0;
  const auto t0 = std::chrono::steady_clock::now();
  for (uint32_t s = 0; s < steps; ++s) {
    search((size_t)warmup + s);  // host buffers: H2D, search, D2H
    pcv_stats st;
    check(pcv_index_stats(ix, &st), "pcv_index_stats");
    device_ms += st.last_search_ms;
    launches = st.last_launches;
  }
  const double wall_s = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();

  // the last timed step's hits, for comparison with another binding of the same ABI
  std::string last_ids;
  if (dump_ids) {
    last_ids = ", \"last_step_ids\": [";
    for (size_t i = 0; i < B * k; ++i) last_ids += (i ? ", " : "") + std::to_string((long long)ids[i]);
    last_ids += "], \"last_step_score_bits\": [";
    for (size_t i = 0; i < B * k; ++i) {
      uint32_t b;
      std::memcpy(&b, &scores[i], 4);
      last_ids += (i ? ", " : "") + std::to_string(b);
    }
    last_ids += "]";
  }
  // planted check: a corpus row is its own nearest neighbour
  const uint64_t planted_row = 7919 % rows;
  std::vector<float> probe(dim);
  check(pcv_synthetic_rows_host(seed, w->dist, planted_row, 1, w->dim, probe.data()), "pcv_synthetic_rows_host");
  check(pcv_search(ix, probe.data(), 1, (uint32_t)k, nullptr, 0, ids.data(), scores.data(), sims.data(), counts.data()), "pcv_search");
  const bool planted_ok = counts[0] > 0 && ids[0] == (int64_t)planted_row + 1;

  const double done = (double)steps * (double)B;
  std::printf("{\"tool\": \"perceive_bench (C++ twin of `perceive bench`)\", \"metric\": \"queries/sec (exact top-k cosine kNN)\", "
              "\"unit\": \"queries/s\", \"value\": %.6f, \"ms_per_step\": %.6f, "
              "\"e2e\": {\"value\": %.6f, \"unit\": \"queries/s\", \"h2d_bytes_per_step\": %zu, \"d2h_bytes_per_step\": %zu}, "
              "\"n_gpus\": %zu, \"steps\": %u, \"warmup\": %u, \"gpu_launches\": %u, "
              "\"config\": {\"workload\": \"%s\", \"rows\": %llu, \"dim\": %u, \"k\": %u, \"batch\": %u, \"corpus_seed\": %llu, \"query_seed\": %llu}, "
              "\"parity\": {\"planted_top1\": %s}%s}\n",
              done / (device_ms * 1e-3), device_ms / steps, done / wall_s, B * dim * 4, B * k * 16 + B * 4, devices.size(), steps, warmup,
              launches * steps, w->text, (unsigned long long)rows, w->dim, w->k, w->batch, (unsigned long long)seed,
              (unsigned long long)(seed + 1), planted_ok ? "true" : "false", last_ids.c_str());
  pcv_index_destroy(ix);
  return planted_ok ? 0 : 3;
} catch (const std::exception& e) {
  std::fprintf(stderr, "perceive_bench: %s\n", e.what());
  return 1;
}
