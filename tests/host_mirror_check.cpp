// Driven by tests/test_host_mirror.py: exercises include/perceive_search.hpp (the C++ mirror of
// perceive_core::search) the way a caller of the reference would, and prints one JSON line per step.
//   host_mirror_check cpu [db]                      codec + error behaviour without a device
//   host_mirror_check gpu db model_id query.f32     build / search_vector / hide / rebuild_source / --like
#include <cstdio>
#include <cstdlib>
#include <cstring>
#include <fstream>
#include <iterator>
#include <string>

#include "perceive_search.hpp"

using perceive::Error;
using perceive::Searcher;
using perceive::SearchItem;

static void print_items(const char* tag, const std::vector<SearchItem>& items) {
  std::printf("{\"step\": \"%s\", \"ids\": [", tag);
  for (size_t i = 0; i < items.size(); ++i) std::printf("%s%lld", i ? ", " : "", (long long)items[i].id);
  std::printf("], \"score_bits\": [");
  for (size_t i = 0; i < items.size(); ++i) {
    uint32_t b;
    std::memcpy(&b, &items[i].score, 4);
    std::printf("%s%u", i ? ", " : "", b);
  }
  std::printf("]}\n");
}

static int cpu_mode(int argc, char** argv) {
  const std::vector<float> v = {1.0f, -2.5f, 3.25f};
  const std::vector<uint8_t> blob = perceive::serialize_embedding(v);
  if (blob.size() != 12 || blob[3] != 0x3f || perceive::deserialize_embedding(blob.data(), blob.size()) != v) return 1;
  try {
    perceive::deserialize_embedding(blob.data(), 11);  // the reference panics on the trailing partial chunk
    return 2;
  } catch (const Error& e) {
    if (e.code != PCV_ERR_INVALID) return 3;
  }
  try {
    Searcher::build("/nonexistent/perceive.db", 7, 0);
    return 4;
  } catch (const Error& e) {
    if (e.code != PCV_ERR_INVALID) return 5;
  }
  Searcher empty;  // no index: the reference returns no items from a Searcher without sources
  if (!empty.search_vector({1}, 5, std::vector<float>(384, 0.f)).empty() || !empty.embedding_of(1).empty()) return 6;
  if (argc > 2) {
    int32_t n_dev = 0;
    const bool have_gpu = pcv_device_count(&n_dev) == PCV_OK && n_dev > 0;
    if (!have_gpu) {
      try {
        Searcher::build(argv[2], 7, 0);  // rows exist but no device: loud failure, never a CPU path
        return 7;
      } catch (const Error& e) {
        if (e.code != PCV_ERR_CUDA) return 8;
        std::printf("{\"step\": \"no_device\", \"message\": \"%s\"}\n", "build refused without a CUDA device");
      }
    }
  }
  std::printf("{\"step\": \"cpu_ok\"}\n");
  return 0;
}

static int gpu_mode(char** argv) {
  const std::string db = argv[2];
  const uint32_t model_id = (uint32_t)std::stoul(argv[3]);
  std::ifstream qf(argv[4], std::ios::binary);
  const std::string raw((std::istreambuf_iterator<char>(qf)), std::istreambuf_iterator<char>());
  std::vector<float> q(raw.size() / 4);
  std::memcpy(q.data(), raw.data(), q.size() * 4);

  // PCV_MIRROR_DEVICES=0,1,...: ONE Searcher over several GPUs of this process (pcv_index_create_multi)
  perceive::Options opt;
  if (const char* env = std::getenv("PCV_MIRROR_DEVICES")) {
    std::string list(env);
    size_t pos = 0;
    while (pos < list.size()) {
      const size_t comma = list.find(',', pos);
      opt.devices.push_back(std::stoi(list.substr(pos, comma == std::string::npos ? std::string::npos : comma - pos)));
      pos = comma == std::string::npos ? list.size() : comma + 1;
    }
  }
  Searcher s = Searcher::build(db, model_id, 0, opt);
  pcv_stats st;
  perceive::check(pcv_index_stats(s.handle(), &st));
  std::printf("{\"step\": \"shards\", \"world\": %u, \"n_rows\": %llu}\n", st.world, (unsigned long long)st.n_rows);
  std::printf("{\"step\": \"built\", \"dim\": %u, \"n_sources\": %zu}\n", s.dim(), s.sources().size());
  print_items("all", s.search_vector({1, 2, 3}, 10, q));
  print_items("src2", s.search_vector({2}, 10, q));
  print_items("none", s.search_vector({}, 10, q));
  print_items("unknown", s.search_vector({99}, 10, q));
  // two queries at once (the same vector twice): both rows equal the single search
  std::vector<float> two(q);
  two.insert(two.end(), q.begin(), q.end());
  auto both = s.search_vectors({1, 2, 3}, 10, two, 2);
  print_items("batch0", both[0]);
  print_items("batch1", both[1]);
  // hide (perceive-cli/cmd/hide.rs:17 inserts into `hidden`): ignored by default, honoured on request
  auto top = s.search_vector({1, 2, 3}, 5, q);
  s.hidden.insert(top[0].id);
  s.hidden.insert(top[2].id);
  print_items("hidden_ignored", s.search_vector({1, 2, 3}, 5, q));
  s.filter_hidden = true;
  print_items("hidden_filtered", s.search_vector({1, 2, 3}, 5, q));
  s.filter_hidden = false;
  // --like: the stored embedding of the best hit finds itself
  auto like = s.embedding_of(top[0].id);
  print_items("like", s.search_vector({1, 2, 3}, 3, like));
  std::printf("{\"step\": \"like_missing\", \"empty\": %s}\n", s.embedding_of(999999).empty() ? "true" : "false");
  // highlighter: three documents (2, 0 and 3 chunks); chunk 1 of doc 0 and chunk 2 of doc 2 are the query itself
  std::vector<float> chunks;
  for (int c = 0; c < 5; ++c) {
    std::vector<float> row = (c == 1 || c == 4) ? q : s.embedding_of(top[(size_t)c % top.size()].id);
    chunks.insert(chunks.end(), row.begin(), row.end());
  }
  auto best = s.best_chunks(q, chunks, {2, 2, 5});
  std::printf("{\"step\": \"best_chunks\", \"best\": [%d, %d, %d]}\n", best[0], best[1], best[2]);
  // rebuild_source re-reads the database (argv[5], when given, is the source to rebuild)
  if (argv[5]) {
    s.rebuild_source(db, std::stoll(argv[5]), model_id, 0);
    print_items("after_rebuild", s.search_vector({1, 2, 3}, 5, q));
  }
  Searcher moved = std::move(s);  // handles move, never copy
  print_items("moved", moved.search_vector({1, 2, 3}, 3, q));
  return 0;
}

int main(int argc, char** argv) {
  try {
    if (argc >= 2 && std::string(argv[1]) == "cpu") return cpu_mode(argc, argv);
    if (argc >= 5 && std::string(argv[1]) == "gpu") return gpu_mode(argv);
  } catch (const std::exception& e) {
    std::fprintf(stderr, "host_mirror_check: %s\n", e.what());
    return 100;
  }
  std::fprintf(stderr, "usage: host_mirror_check cpu [db] | gpu db model_id query.f32 [source_to_rebuild]\n");
  return 64;
}
