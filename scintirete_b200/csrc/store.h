// store.h — device-memory mirror of the reference's node store (hnsw.go:17-26, 115) on one GPU.
//
// HBM layout (rows in insertion order; row == id-1 for auto-assigned ids, collection.go:115-116):
//   vec      float   [cap][pitch]     fp32 rows, pitch = dim rounded up to 8 floats, zero padded
//   norm     float   [cap]            ||x|| in the reference's sequential fp32 order (cosine normB)
//   mirror   bf16    [cap][kpad]      tensor-core operand: bf16(x) (L2, IP) or bf16(x/||x||) (cosine),
//                                     kpad = dim rounded up to 64, zero padded
//   aux      float   [cap]            filter-side per-row term: sum(mirror^2) for L2, else unused (0)
//   ids      u64     [cap]            external id of each row
//   deleted  u32     [cap/32]         soft-delete bitmap (HNSWNode.Deleted)
//   adj0     u32     [n][2M]          layer-0 neighbour rows, ROW_NONE padded (one 128 B line at M=16)
//   levels   u8      [n]              len(Connections)-1
//   up_off   u32     [n]              first upper-layer list of the node (in lists), for level >= 1
//   adj_up   u32     [lists][M]       upper-layer neighbour rows, ROW_NONE padded
#pragma once

#include <cuda_bf16.h>

#include <mutex>
#include <string>
#include <unordered_map>
#include <vector>

#include "common.cuh"

namespace scn {
struct BuildState;                       // hnsw_build.cu: host mirror of the graph under construction
}

struct scn_store {
  int32_t device = 0;
  uint32_t dim = 0;
  uint32_t pitch = 0;  // floats per fp32 row
  uint32_t kpad = 0;   // bf16 elements per mirror row
  int32_t metric = 0;

  uint64_t rows = 0, cap = 0, live = 0;
  float* d_vec = nullptr;
  float* d_norm = nullptr;
  __nv_bfloat16* d_mirror = nullptr;
  float* d_aux = nullptr;
  uint64_t* d_ids = nullptr;
  uint32_t* d_deleted = nullptr;
  // running maxima over rows, for the tensor-path certificate (device float[4]):
  //   [0] max ||mirror row||   [1] max ||x_true - mirror row||   [2] max ||x||   [3] reserved
  float* d_bounds = nullptr;
  // device counters of the last search: flat: [0] queries through the tensor filter, [1] queries
  // re-scanned exactly (certificate failed); hnsw: [0] distance evaluations, [1] expansions
  unsigned long long* d_counters = nullptr;

  bool auto_ids = true;  // every id so far was auto_base+row+1 -> no host map needed
  uint64_t auto_base = 0;  // first auto id - 1 (a row shard of a larger collection starts at its global row)
  std::unordered_map<uint64_t, uint32_t> row_of;

  // graph
  bool has_graph = false;
  int32_t m = 0, max_layer = -1;
  uint64_t entry_id = 0;
  uint32_t entry_row = scn::ROW_NONE;
  uint64_t graph_nodes = 0, graph_edges = 0, upper_lists = 0;
  // host: highest layer on which a row has at least one edge (getNodeLayer, hnsw.go:472-484), kept
  // so that a deleted entry point can be replaced like findNewEntrypoint does (hnsw.go:617-634)
  std::vector<uint8_t> h_node_layer;
  uint32_t* d_adj0 = nullptr;
  uint8_t* d_levels = nullptr;
  uint32_t* d_up_off = nullptr;
  uint32_t* d_adj_up = nullptr;
  scn::BuildState* build = nullptr;   // present once the graph was built / extended on this device (scn_hnsw_insert)

  // options
  int64_t opt_flat_path = 0;
  int64_t opt_tensor_min_batch = 1;  // the bf16 filter streams half the bytes of the fp32 scan: it wins at every batch size
  int64_t opt_overfetch = 0;  // 0 = auto
  int64_t opt_hnsw_gather = -1;   // hnsw_search row gather: -1 auto; 0 registers (LDG.256); 1 <512,1>, 2 <512,2>, 3 <256,1> shared-memory stages
  int64_t opt_hnsw_gather_long = 1;  // the auto choice for rows longer than 512 bytes (C1-768: <512,1> 21.9 ms / 2.9 ms at nq = 10 000 / 1 000; <512,2> 22.5 / 3.1; <256,1> 21.2 / 3.6)
  int64_t opt_hnsw_global = 1;    // hnsw_search: 1 = visited tables of the first pass in global memory (L2) instead of shared memory
  int64_t opt_hnsw_per_sm = 0;    // hnsw_search, global tables: cap on resident queries per SM; 0 = whatever fits
  int64_t opt_hnsw_early = 1;     // hnsw_search: rows requested before the visited test (copies overlap the probes)
  int64_t opt_hnsw_exact_ties = 0;  // hnsw_search: 1 = walks that end with a distance tie at the edge of W are redone by the exact walk kernel
  int64_t opt_hnsw_hash = 0;      // hnsw_search: entries of the visited table of the first pass (shared or global memory); 0 = auto
  int64_t opt_tensor_hint = 1;    // lists of a query seed their threshold from the finished ones
  int64_t opt_tensor_hint_target = 0;  // rows of the shard that should beat a published threshold; 0 = 3 k''
  int64_t opt_tensor_bn = 0;      // 128 forces 128-row tiles in the tensor filter (0 = auto)
  int64_t opt_tensor_chunks = 0;  // row chunks per query block in the tensor filter; 0 = auto
  int64_t opt_tensor_pair = 1;    // 1 = CTA-pair filter kernel (cta_group::2, M=256 x N=128, queries stationary in TMEM) at kpad 512 / 640 / 768, batches >= 256
  int64_t opt_tensor_fused = -1;  // merge -> exact rerank -> certificate behind the filter in ONE launch (finish_queries_kernel): 1 always, 0 never, -1 auto (batches of up to 512 queries)
  int64_t opt_tensor_pair_ew = 0; // pair kernel: epilogue warps per TMEM lane quarter; 0 = auto (2), 1, 2
  int64_t opt_tensor_share = 1;   // lists of a query exchange bounds while they are built: 1 = where it pays (flat_tensor.cu), 2 = whenever a query has >= 16 lists, 0 = never
  int64_t opt_pdl = 1;            // the short kernels behind the tensor filter are launched chained (programmatic dependent launch, common.cuh)
  int64_t opt_build_window = 0;   // scn_hnsw_insert: inserts searched speculatively per round; 0 = adaptive, 1 = none (serial)
  int64_t opt_profile = 0;

  // introspection (protected by mu). Profiled kernels leave (name, start, stop) event triples
  // here; scn_last_timings synchronises on them, sums per name and clears the list.
  std::mutex mu;
  struct TimedEvent {
    std::string name;
    cudaEvent_t a, b;
  };
  std::vector<TimedEvent> pending;

  uint64_t device_bytes() const;
  bool lookup(uint64_t id, uint32_t* row) const {
    if (auto_ids) {
      if (id <= auto_base || id - auto_base > rows) return false;
      *row = (uint32_t)(id - auto_base - 1);
      return true;
    }
    auto it = row_of.find(id);
    if (it == row_of.end()) return false;
    *row = it->second;
    return true;
  }
};

namespace scn {

// RAII device guard
struct DeviceGuard {
  int prev = -1;
  explicit DeviceGuard(int dev) {
    cudaGetDevice(&prev);
    if (prev != dev) cudaSetDevice(dev);
    else prev = -1;
  }
  ~DeviceGuard() {
    if (prev >= 0) cudaSetDevice(prev);
  }
};

// Stream-ordered scratch allocation tied to a call.
struct Scratch {
  cudaStream_t stream;
  std::vector<void*> ptrs;
  char* pool = nullptr;        // one block reserved up front (reserve()): later allocs are carved out of it
  size_t pool_bytes = 0, pool_used = 0;
  explicit Scratch(cudaStream_t s) : stream(s) {}
  static size_t padded(size_t bytes) { return (bytes + 16 + 255) & ~(size_t)255; }
  // One stream-ordered allocation for a whole call instead of one per buffer (a call of the tensor path
  // needs about fifteen): `bytes` = sum of padded(size) over the buffers that will follow.
  int32_t reserve(size_t bytes) {
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, bytes, stream);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return scn::fail(SCN_ERR_RESOURCE, "device scratch allocation of %zu bytes failed: %s", bytes, cudaGetErrorString(e));
    }
    ptrs.push_back(p);
    pool = static_cast<char*>(p);
    pool_bytes = bytes;
    pool_used = 0;
    return SCN_OK;
  }
  template <class T>
  int32_t alloc(T** out, size_t count) {
    const size_t need = padded(count * sizeof(T));
    if (pool && pool_used + need <= pool_bytes) {
      *out = reinterpret_cast<T*>(pool + pool_used);
      pool_used += need;
      return SCN_OK;
    }
    void* p = nullptr;
    cudaError_t e = cudaMallocAsync(&p, count * sizeof(T) + 16, stream);
    if (e != cudaSuccess) {
      cudaGetLastError();
      return scn::fail(SCN_ERR_RESOURCE, "device scratch allocation of %zu bytes failed: %s", count * sizeof(T),
                       cudaGetErrorString(e));
    }
    ptrs.push_back(p);
    *out = reinterpret_cast<T*>(p);
    return SCN_OK;
  }
  ~Scratch() {
    for (void* p : ptrs) cudaFreeAsync(p, stream);
  }
};

// Optional per-kernel CUDA-event timing on the launching stream.
struct Profiler {
  scn_store* s;
  cudaStream_t stream;
  bool on;
  std::vector<std::pair<std::string, std::pair<cudaEvent_t, cudaEvent_t>>> ev;
  Profiler(scn_store* st, cudaStream_t str) : s(st), stream(str), on(st && st->opt_profile != 0) {}
  void begin(const char* name) {
    if (!on) return;
    cudaEvent_t a, b;
    cudaEventCreate(&a);
    cudaEventCreate(&b);
    cudaEventRecord(a, stream);
    ev.push_back({name, {a, b}});
  }
  void end() {
    if (!on) return;
    cudaEventRecord(ev.back().second.second, stream);
  }
  // hands the recorded events to the store (no synchronisation here)
  void collect();
  ~Profiler();
};

void free_build_state(scn_store* s);

// One rank's host-buffer call of the fused shard exchange in three steps (exchange.cu); shards.cu puts a
// host barrier between them when several shards share a device.
struct HostExchangeCall;
HostExchangeCall* host_exchange_begin(scn_store* s, scn_exchange* ex, const float* q_slice, uint64_t nq, uint32_t k, uint64_t row_base,
                                      int32_t* rc);
int32_t host_exchange_search(HostExchangeCall* h);
int32_t host_exchange_finish(HostExchangeCall* h, uint64_t* out_ids, float* out_dist, uint32_t* out_counts);   // consumes h
int32_t host_exchange_sync(HostExchangeCall* h);
void host_exchange_abort(HostExchangeCall* h);
cudaStream_t thread_stream(int device);
// Host -> device copy of a caller's buffer, enqueued on `stream`. Pinned / registered memory is
// copied directly; pageable memory (a Go slice, a numpy array) goes through two pinned staging
// chunks of this thread, so that the host-side memcpy of one chunk overlaps the DMA of the previous
// one. Either way the caller's buffer has been read completely when the function returns.
int32_t copy_to_device(void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream);

// kernels / launchers implemented in the other translation units
int32_t launch_prepare_rows(scn_store* s, uint64_t first_row, uint64_t n, cudaStream_t stream);

int32_t flat_search_exact(scn_store* s, const float* d_q, const uint32_t* d_qlist, const uint32_t* d_nq_dev,
                          uint64_t nq, uint32_t k, uint64_t row_base, uint64_t* d_out_keys, cudaStream_t stream,
                          Profiler* prof);
int32_t keys_to_results(scn_store* s, const uint64_t* d_keys, uint64_t n, uint64_t row_base, uint64_t* d_out_ids,
                        float* d_out_dist, uint32_t* d_out_counts, uint32_t k, cudaStream_t stream, bool chained = false);
int32_t flat_search_tensor(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint64_t row_base,
                           uint64_t* d_out_keys, cudaStream_t stream, Profiler* prof);
bool tensor_path_supported(const scn_store* s, uint32_t k);
int32_t rerank_rows(scn_store* s, const float* d_q, uint64_t nq, const uint32_t* d_cand_rows, uint32_t ncand,
                    uint32_t k, uint64_t row_base, uint64_t* d_out_keys, cudaStream_t stream,
                    const uint32_t* d_qlist = nullptr, const uint32_t* d_nq_dev = nullptr);
int32_t tensor_debug_scores(scn_store* s, const float* d_q, uint64_t nq, float* d_scores, cudaStream_t stream);
int32_t mark_aux_deleted(scn_store* s, const uint32_t* h_rows, uint32_t n, cudaStream_t stream);
int32_t hnsw_search(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                    float* d_out_dist, uint32_t* d_out_counts, cudaStream_t stream, Profiler* prof);
int32_t hnsw_search_exact(scn_store* s, const float* d_q, const uint32_t* d_qlist, const uint32_t* d_nq_dev, uint64_t nq, uint32_t k,
                          uint32_t ef, uint32_t hash_size, uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts,
                          unsigned long long* d_failed, cudaStream_t st, Scratch& scratch);
int32_t vector_ops(int32_t op, const float* d_a, const float* d_b, uint64_t n, uint32_t dim, float* d_out, cudaStream_t stream);
int32_t distance_batch(int32_t metric, const float* d_q, uint64_t nq, const float* d_x, uint64_t nx, uint32_t dim,
                       float* d_out, cudaStream_t stream);
int32_t merge_topk(const uint64_t* d_keys, const uint64_t* d_ids, uint32_t n_shards, uint64_t nq, uint32_t k,
                   uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts, cudaStream_t stream,
                   uint64_t shard_stride = 0 /* elements between shard lists; 0 = nq*k */,
                   const uint32_t* d_status = nullptr /* non-zero word: the lists are incomplete -> empty results */);
int32_t flat_keys(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint64_t row_base, uint64_t* d_keys,
                  cudaStream_t stream, Profiler* prof);

}  // namespace scn
