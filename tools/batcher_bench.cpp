// batcher_bench.cpp — queries/s of the reference-shaped call (one query per call, many concurrent
// callers: Collection.Search under RLock, collection.go:193-204) through scn_batcher_search,
// against the same calls issued one launch each (scn_search_flat with nq = 1).
//
//   g++ -O2 -std=c++17 -pthread tools/batcher_bench.cpp -Iinclude -Lscintirete_b200 -lscn_gpu \
//       -Wl,-rpath,$PWD/scintirete_b200 -o gpurun_out/batcher_bench
//   gpurun_out/batcher_bench [rows=200000] [dim=768] [threads=256] [calls_per_thread=200] [window_us=200]
#include <atomic>
#include <chrono>
#include <cstdio>
#include <cstdlib>
#include <random>
#include <thread>
#include <vector>

#include "scn_gpu.h"

#define CHECK(x)                                                                   \
  do {                                                                             \
    int32_t rc__ = (x);                                                            \
    if (rc__ != 0) {                                                               \
      std::fprintf(stderr, "%s -> %d: %s\n", #x, rc__, scn_last_error());          \
      std::exit(1);                                                                \
    }                                                                              \
  } while (0)

int main(int argc, char** argv) {
  const uint64_t rows = argc > 1 ? std::strtoull(argv[1], nullptr, 10) : 200000;
  const uint32_t dim = argc > 2 ? (uint32_t)std::atoi(argv[2]) : 768;
  const int threads = argc > 3 ? std::atoi(argv[3]) : 256;
  const int calls = argc > 4 ? std::atoi(argv[4]) : 200;
  const uint32_t window_us = argc > 5 ? (uint32_t)std::atoi(argv[5]) : 200;
  const uint32_t k = 10;
  scn_store* s = nullptr;
  CHECK(scn_store_create(0, dim, SCN_METRIC_COSINE, &s));
  CHECK(scn_store_reserve(s, rows));
  std::mt19937 rng(1234);
  std::normal_distribution<float> nd;
  std::vector<float> blk((size_t)16384 * dim);
  for (uint64_t r = 0; r < rows; r += 16384) {
    const uint64_t n = std::min<uint64_t>(16384, rows - r);
    for (size_t i = 0; i < n * dim; ++i) blk[i] = nd(rng);
    CHECK(scn_store_append(s, blk.data(), nullptr, n));
  }
  std::vector<float> q((size_t)threads * dim);
  for (auto& v : q) v = nd(rng);

  auto run = [&](scn_batcher* b, int n_threads, int n_calls) {
    std::vector<std::thread> ts;
    std::atomic<uint64_t> checksum{0};
    auto t0 = std::chrono::steady_clock::now();
    for (int t = 0; t < n_threads; ++t)
      ts.emplace_back([&, t] {
        uint64_t ids[k];
        float dist[k];
        uint32_t cnt = 0;
        uint64_t acc = 0;
        for (int c = 0; c < n_calls; ++c) {
          if (b) CHECK(scn_batcher_search(b, q.data() + (size_t)t * dim, k, 0, ids, dist, &cnt));
          else CHECK(scn_search_flat(s, q.data() + (size_t)t * dim, 1, k, ids, dist, &cnt));
          acc += ids[0];
        }
        checksum += acc;
      });
    for (auto& th : ts) th.join();
    const double sec = std::chrono::duration<double>(std::chrono::steady_clock::now() - t0).count();
    return std::make_pair((double)n_threads * n_calls / sec, checksum.load());
  };

  // warm-up
  run(nullptr, 4, 5);
  auto solo = run(nullptr, threads, std::max(1, calls / 10));
  scn_batcher* b = nullptr;
  CHECK(scn_batcher_create(s, 0, 4096, window_us, &b));
  run(b, threads, 5);
  auto co = run(b, threads, calls);
  uint64_t st[4] = {0, 0, 0, 0};
  CHECK(scn_batcher_stats(b, st, 4));
  std::printf(
      "{\"workload\": \"%llux%u cosine flat k=10, %d concurrent callers, one query per call\", \"uncoalesced_qps\": %.1f, "
      "\"coalesced_qps\": %.1f, \"speedup\": %.2f, \"calls\": %llu, \"batches\": %llu, \"mean_batch\": %.1f, \"max_batch\": %llu, "
      "\"window_us\": %u, \"checksums_equal\": %s}\n",
      (unsigned long long)rows, dim, threads, solo.first, co.first, co.first / solo.first, (unsigned long long)st[0],
      (unsigned long long)st[1], (double)st[0] / (double)std::max<uint64_t>(1, st[1]), (unsigned long long)st[3], window_us,
      (solo.second / std::max(1, calls / 10) == co.second / calls) ? "true" : "false");
  CHECK(scn_batcher_destroy(b));
  scn_store_destroy(s);
  return 0;
}
