// exchange.cu — row-sharded exact search with both of its exchanges fused over NVLink peer memory
// ("optional fusion").
//
// The NCCL formulation of one sharded batch is: H2D of the whole query batch on every rank ->
// shard-local search -> keys_to_results -> all_gather(keys) -> all_gather(ids) -> merge_topk on
// every rank. Here every rank owns one exchange buffer, mapped into every other rank (cudaIpc
// between processes, cudaDeviceEnablePeerAccess inside one process — the reference server is ONE
// process, so that is the drop-in's natural form):
//
//   flags      u32 [2 kinds][2 parities][16 peers]          arrival flags (value = call epoch, never 0)
//   lists      u64 keys [2 parities][world][slice_cap * k]  + ids, same shape
//   queries    f32 [2 parities][max_nq * dim]               the assembled query batch (dim > 0 only)
//
// and a batch of nq queries is cut into `world` contiguous slices (slice r = queries
// [r*per, min(nq, (r+1)*per)), per = ceil(nq / world)); rank r answers slice r:
//
//   1. query gather (optional): each rank copies only ITS slice from the host (1/world of the PCIe
//      traffic) and a kernel stores it into every rank's query buffer with 16-byte P2P stores, then
//      raises the query flags; a one-warp kernel waits for the world flags of the call.
//   2. shard-local search of the WHOLE batch over the local rows -> sorted (key, row) lists.
//   3. push: every list element goes to exactly one peer — the owner of its query's slice — as a
//      (key, id) pair, 16-byte P2P stores; the last block raises the list flags (st.release.sys).
//   4. wait for the world list flags (time-out reported through a status word, never a hang), then
//      merge_topk over the rank's own slice from LOCAL memory. A timed-out call yields empty results.
//
// Keys carry global rows, so the merged slices are bit-identical to the single-GPU search. Parity
// double buffering is enough: a rank can start call e+2 (which overwrites parity(e) buffers of its
// peers) only after its own wait of call e+1 saw every peer's list flag e+1, which a peer raises
// after its search e+1, i.e. after it finished reading parity(e) (its merge and its search of call
// e precede it in stream order). Ranks driven inside one process must use different streams (or
// devices): a rank's wait spins until the other ranks' pushes have run.
#include <cstring>

#include "store.h"

using namespace scn;

struct scn_exchange {
  int32_t device = 0;
  uint32_t rank = 0, world = 1, k = 0, dim = 0;
  uint64_t max_nq = 0;
  uint64_t slice_cap = 0;                    // queries per slice the list buffers can hold
  uint64_t epoch = 0;
  unsigned char* local = nullptr;            // this rank's buffer
  unsigned char* peer[16] = {};              // peer[r] = rank r's buffer as seen from here (peer[rank] = local)
  bool opened_ipc[16] = {};
  bool connected = false;
  uint32_t* d_done = nullptr;                // [0] push blocks done, [1] scatter blocks done, [2] status
  uint32_t* d_status = nullptr;              // 0 ok, 1 = a wait timed out since the last status read
  size_t bytes = 0;
};

namespace {

constexpr size_t FLAG_BYTES = 4096;  // [2 kinds][2 parities][16] u32 flags, padded
enum : uint32_t { FLAG_LISTS = 0, FLAG_QUERIES = 1 };

struct Layout {
  uint32_t world;
  size_t le;        // list elements per (parity, shard): slice_cap * k
  size_t q_floats;  // floats per query buffer: max_nq * dim
};

__host__ __device__ inline uint32_t* flag_of(unsigned char* base, uint32_t kind, uint32_t parity, uint32_t r) {
  return reinterpret_cast<uint32_t*>(base) + (kind * 2 + parity) * 16 + r;
}
__host__ __device__ inline uint64_t* keys_of(unsigned char* base, const Layout& L, uint32_t parity, uint32_t shard) {
  return reinterpret_cast<uint64_t*>(base + FLAG_BYTES) + ((size_t)parity * L.world + shard) * L.le;
}
__host__ __device__ inline uint64_t* ids_of(unsigned char* base, const Layout& L, uint32_t parity, uint32_t shard) {
  return reinterpret_cast<uint64_t*>(base + FLAG_BYTES) + ((size_t)2 * L.world + (size_t)parity * L.world + shard) * L.le;
}
__host__ __device__ inline float* queries_of(unsigned char* base, const Layout& L, uint32_t parity) {
  return reinterpret_cast<float*>(base + FLAG_BYTES + (size_t)4 * L.world * L.le * sizeof(uint64_t)) + (size_t)parity * L.q_floats;
}

struct PeerArgs {
  unsigned char* peer[16];
  Layout L;
  uint32_t rank, parity, epoch;
  uint32_t* done;
};

// the last block of a grid raises this rank's flag of `kind` on every peer
__device__ __forceinline__ void publish(const PeerArgs& a, uint32_t kind) {
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(a.done, 1u);
    if (prev == gridDim.x - 1) {
      *a.done = 0;
      __threadfence_system();
      for (uint32_t p = 0; p < a.L.world; ++p) {
        uint32_t* f = flag_of(a.peer[p], kind, a.parity, a.rank);
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(a.epoch) : "memory");
      }
    }
  }
}

// Step 3. keys -> ids, then every (key, id) pair to the owner of its query's slice. With an even k
// a thread moves two consecutive elements of one query with one 16-byte store per list.
__global__ void __launch_bounds__(256) push_results_kernel(PeerArgs a, const uint64_t* __restrict__ keys,
                                                           const uint64_t* __restrict__ row_ids, uint64_t n /* nq*k */,
                                                           uint32_t k, uint32_t per /* queries per slice */, uint32_t row_base) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  if ((k & 1u) == 0) {
    for (uint64_t i = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) * 2; i < n; i += stride * 2) {
      const ulonglong2 kk = *reinterpret_cast<const ulonglong2*>(keys + i);
      ulonglong2 id;
      id.x = (kk.x == KEY_NONE) ? 0ull : row_ids[(uint32_t)kk.x - row_base];
      id.y = (kk.y == KEY_NONE) ? 0ull : row_ids[(uint32_t)kk.y - row_base];
      const uint32_t q = (uint32_t)(i / k);
      const uint32_t p = q / per;
      const size_t dst = i - (size_t)p * per * k;  // even: k is
      *reinterpret_cast<ulonglong2*>(keys_of(a.peer[p], a.L, a.parity, a.rank) + dst) = kk;
      *reinterpret_cast<ulonglong2*>(ids_of(a.peer[p], a.L, a.parity, a.rank) + dst) = id;
    }
  } else {
    for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < n; i += stride) {
      const uint64_t key = keys[i];
      const uint64_t id = (key == KEY_NONE) ? 0ull : row_ids[(uint32_t)key - row_base];
      const uint32_t q = (uint32_t)(i / k);
      const uint32_t p = q / per;
      const size_t dst = i - (size_t)p * per * k;
      keys_of(a.peer[p], a.L, a.parity, a.rank)[dst] = key;
      ids_of(a.peer[p], a.L, a.parity, a.rank)[dst] = id;
    }
  }
  publish(a, FLAG_LISTS);
}

// Step 1. this rank's query slice -> every rank's query buffer (16-byte stores when aligned)
__global__ void __launch_bounds__(256) scatter_queries_kernel(PeerArgs a, const float* __restrict__ src, uint64_t n_floats,
                                                              uint64_t dst_off /* floats */) {
  const uint64_t stride = (uint64_t)gridDim.x * blockDim.x;
  const uint64_t t0 = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  const bool vec = ((n_floats | dst_off) & 3ull) == 0 && (reinterpret_cast<uintptr_t>(src) & 15u) == 0;
  if (vec) {
    for (uint64_t i = t0; i < n_floats / 4; i += stride) {
      const float4 v = __ldg(reinterpret_cast<const float4*>(src) + i);
      for (uint32_t p = 0; p < a.L.world; ++p) reinterpret_cast<float4*>(queries_of(a.peer[p], a.L, a.parity) + dst_off)[i] = v;
    }
  } else {
    for (uint64_t i = t0; i < n_floats; i += stride) {
      const float v = __ldg(src + i);
      for (uint32_t p = 0; p < a.L.world; ++p) queries_of(a.peer[p], a.L, a.parity)[dst_off + i] = v;
    }
  }
  publish(a, FLAG_QUERIES);
}

// one warp: lane r waits for rank r's flag of this call
__global__ void wait_flags_kernel(const uint32_t* flags /* [16] of the kind and parity */, uint32_t world, uint32_t epoch,
                                  uint32_t* status, unsigned long long timeout_ns) {
  const uint32_t r = threadIdx.x;
  if (r >= world) return;
  const uint32_t* f = flags + r;
  unsigned long long t0;
  asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t0));
  for (;;) {
    uint32_t v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(f) : "memory");
    if (v == epoch) return;
    unsigned long long t;
    asm volatile("mov.u64 %0, %%globaltimer;" : "=l"(t));
    if (t - t0 > timeout_ns) {
      atomicExch(status, 1u);
      return;
    }
    __nanosleep(200);
  }
}

constexpr unsigned long long WAIT_TIMEOUT_NS = 5000000000ull;

uint32_t epoch_value(uint64_t epoch) { return (uint32_t)(epoch % 0xFFFFFFFFull) + 1u; }  // never 0, the flags' initial value

}  // namespace

extern "C" {

int32_t scn_exchange_create(int32_t device, uint32_t rank, uint32_t world, uint64_t max_nq, uint32_t k, uint32_t dim,
                            scn_exchange** out) {
  if (!out) return fail(SCN_ERR_INVALID_PARAMETERS, "out is NULL");
  *out = nullptr;
  if (world == 0 || world > 16 || rank >= world) return fail(SCN_ERR_INVALID_PARAMETERS, "world must be in [1, 16] and rank < world");
  if (k == 0 || k > 1024 || max_nq == 0) return fail(SCN_ERR_INVALID_PARAMETERS, "max_nq and k must be positive (k <= 1024)");
  DeviceGuard g(device);
  scn_exchange* ex = new scn_exchange();
  ex->device = device;
  ex->rank = rank;
  ex->world = world;
  ex->k = k;
  ex->dim = dim;
  ex->max_nq = max_nq;
  ex->slice_cap = (max_nq + world - 1) / world;
  const size_t le = (size_t)ex->slice_cap * k;
  ex->bytes = FLAG_BYTES + (size_t)4 * world * le * sizeof(uint64_t) + (size_t)2 * max_nq * dim * sizeof(float);
  cudaError_t e = cudaMalloc(&ex->local, ex->bytes);  // cudaMalloc (not the async pool): IPC-exportable
  if (e == cudaSuccess) e = cudaMemset(ex->local, 0, FLAG_BYTES);
  if (e == cudaSuccess) e = cudaMalloc(&ex->d_done, 4 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(ex->d_done, 0, 4 * sizeof(uint32_t));
  if (e != cudaSuccess) {
    cudaFree(ex->local);
    cudaFree(ex->d_done);
    delete ex;
    return cuda_fail(e, "exchange buffer allocation", __FILE__, __LINE__);
  }
  ex->d_status = ex->d_done + 2;
  ex->peer[rank] = ex->local;
  ex->connected = (world == 1);
  *out = ex;
  return SCN_OK;
}

int32_t scn_exchange_local_handle(scn_exchange* ex, void* out_handle) {
  if (!ex || !out_handle) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  static_assert(sizeof(cudaIpcMemHandle_t) == SCN_IPC_HANDLE_BYTES, "IPC handle size");
  DeviceGuard g(ex->device);
  cudaIpcMemHandle_t h;
  SCN_CUDA(cudaIpcGetMemHandle(&h, ex->local));
  std::memcpy(out_handle, &h, sizeof h);
  return SCN_OK;
}

// between processes: handles[r] = rank r's scn_exchange_local_handle
int32_t scn_exchange_connect(scn_exchange* ex, const void* handles) {
  if (!ex || !handles) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  DeviceGuard g(ex->device);
  for (uint32_t r = 0; r < ex->world; ++r) {
    if (r == ex->rank) continue;
    cudaIpcMemHandle_t h;
    std::memcpy(&h, static_cast<const unsigned char*>(handles) + (size_t)r * SCN_IPC_HANDLE_BYTES, sizeof h);
    void* p = nullptr;
    SCN_CUDA(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    ex->peer[r] = static_cast<unsigned char*>(p);
    ex->opened_ipc[r] = true;
  }
  ex->connected = true;
  return SCN_OK;
}

// inside one process (the reference server's shape): peers[r] = rank r's exchange object
int32_t scn_exchange_connect_local(scn_exchange* ex, scn_exchange* const* peers) {
  if (!ex || !peers) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  DeviceGuard g(ex->device);
  for (uint32_t r = 0; r < ex->world; ++r) {
    if (r == ex->rank) continue;
    if (!peers[r] || peers[r]->world != ex->world || peers[r]->rank != r || peers[r]->max_nq != ex->max_nq || peers[r]->k != ex->k ||
        peers[r]->dim != ex->dim)
      return fail(SCN_ERR_INVALID_PARAMETERS, "peer %u does not match this exchange", r);
    if (peers[r]->device != ex->device) {
      int can = 0;
      SCN_CUDA(cudaDeviceCanAccessPeer(&can, ex->device, peers[r]->device));
      if (!can) return fail(SCN_ERR_INTERNAL, "device %d cannot access device %d", ex->device, peers[r]->device);
      cudaError_t e = cudaDeviceEnablePeerAccess(peers[r]->device, 0);
      if (e != cudaSuccess && e != cudaErrorPeerAccessAlreadyEnabled) return cuda_fail(e, "cudaDeviceEnablePeerAccess", __FILE__, __LINE__);
      cudaGetLastError();
    }
    ex->peer[r] = peers[r]->local;
  }
  ex->connected = true;
  return SCN_OK;
}

int32_t scn_exchange_destroy(scn_exchange* ex) {
  if (!ex) return SCN_OK;
  DeviceGuard g(ex->device);
  cudaDeviceSynchronize();
  for (uint32_t r = 0; r < ex->world; ++r)
    if (ex->opened_ipc[r]) cudaIpcCloseMemHandle(ex->peer[r]);
  cudaFree(ex->local);
  cudaFree(ex->d_done);
  delete ex;
  return SCN_OK;
}

int32_t scn_exchange_slice(const scn_exchange* ex, uint64_t nq, uint32_t rank, uint64_t* out_first, uint64_t* out_count) {
  if (!ex || rank >= ex->world) return fail(SCN_ERR_INVALID_PARAMETERS, "bad exchange or rank");
  const uint64_t per = (nq + ex->world - 1) / ex->world;
  const uint64_t lo = std::min<uint64_t>(nq, (uint64_t)rank * per);
  if (out_first) *out_first = lo;
  if (out_count) *out_count = std::min<uint64_t>(nq, lo + per) - lo;
  return SCN_OK;
}

}  // extern "C"

// One rank's call, cut at the points where it starts to wait for its peers. Ranks that share a DEVICE
// (tests, smoke: several shards of one process on one GPU) must never have a kernel spinning on a flag
// that another launch on the same GPU is to raise — nothing guarantees that two launches run at the
// same time — so a driver of such ranks runs the three steps with a host barrier AND a stream
// synchronisation between them (shards.cu): every wait kernel then finds its flags already raised.
// With one rank per device the steps simply follow each other.
namespace scn {

struct ExchangeCall {
  scn_store* s;
  scn_exchange* ex;
  cudaStream_t st;
  Profiler prof;
  Scratch scratch;
  PeerArgs a;
  const float* d_q_full = nullptr;
  uint64_t* d_keys = nullptr;
  uint64_t nq = 0, q_lo = 0, q_n = 0, row_base = 0;
  uint32_t k = 0, per = 0;
  bool q_is_slice = false;
  ExchangeCall(scn_store* s_, scn_exchange* ex_, cudaStream_t st_) : s(s_), ex(ex_), st(st_), prof(s_, st_), scratch(st_) {}
};

// step 1: argument checks, this rank's query slice -> every rank's query buffer
int32_t exchange_begin(ExchangeCall& c, const float* d_q, int32_t q_is_slice, uint64_t nq, uint32_t k, uint64_t row_base) {
  scn_store* s = c.s;
  scn_exchange* ex = c.ex;
  if (!s || !ex) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  if (!ex->connected) return fail(SCN_ERR_INVALID_PARAMETERS, "exchange is not connected to its peers");
  if (s->device != ex->device) return fail(SCN_ERR_INVALID_PARAMETERS, "store and exchange live on different devices");
  if (k != ex->k || nq == 0 || nq > ex->max_nq) return fail(SCN_ERR_INVALID_PARAMETERS, "nq must be in [1, %llu] and k == %u", (unsigned long long)ex->max_nq, ex->k);
  if (q_is_slice && ex->dim != s->dim) return fail(SCN_ERR_INVALID_PARAMETERS, "the exchange was created for dim %u, the store has %u", ex->dim, s->dim);
  if (row_base + s->rows >= (uint64_t)ROW_NONE) return fail(SCN_ERR_INVALID_PARAMETERS, "global row index exceeds 32 bits");
  scn_exchange_slice(ex, nq, ex->rank, &c.q_lo, &c.q_n);
  if ((c.q_n && !d_q && q_is_slice) || (!q_is_slice && !d_q)) return fail(SCN_ERR_INVALID_PARAMETERS, "query pointer is NULL");
  c.nq = nq;
  c.k = k;
  c.row_base = row_base;
  c.q_is_slice = q_is_slice != 0;
  const uint64_t epoch = ++ex->epoch;
  PeerArgs& a = c.a;
  for (int r = 0; r < 16; ++r) a.peer[r] = ex->peer[r];
  a.L.world = ex->world;
  a.L.le = (size_t)ex->slice_cap * k;
  a.L.q_floats = (size_t)ex->max_nq * ex->dim;
  a.rank = ex->rank;
  a.parity = (uint32_t)(epoch & 1);
  a.epoch = epoch_value(epoch);
  c.per = (uint32_t)((nq + ex->world - 1) / ex->world);
  c.d_q_full = d_q;
  SCN_TRY(c.scratch.alloc(&c.d_keys, nq * k));   // (allocated here: nothing is allocated once a rank may be waiting)
  if (c.q_is_slice) {
    c.prof.begin("scatter_queries");
    a.done = ex->d_done + 1;
    const uint64_t n_floats = c.q_n * s->dim;
    const unsigned blocks = (unsigned)std::max<uint64_t>(1, std::min<uint64_t>(148, (n_floats / 4 + 255) / 256));
    scatter_queries_kernel<<<blocks, 256, 0, c.st>>>(a, d_q, n_floats, c.q_lo * s->dim);
    SCN_LAUNCHED();
    c.prof.end();
    c.d_q_full = queries_of(ex->local, a.L, a.parity);
  }
  return SCN_OK;
}

// step 2: (wait for the query slices,) shard-local search of the whole batch, push each list to the
// owner of its query's slice
int32_t exchange_search(ExchangeCall& c) {
  scn_store* s = c.s;
  scn_exchange* ex = c.ex;
  PeerArgs& a = c.a;
  if (c.q_is_slice) {
    c.prof.begin("wait_queries");
    wait_flags_kernel<<<1, 32, 0, c.st>>>(flag_of(ex->local, FLAG_QUERIES, a.parity, 0), ex->world, a.epoch, ex->d_status, WAIT_TIMEOUT_NS);
    SCN_LAUNCHED();
    c.prof.end();
  }
  SCN_TRY(flat_keys(s, c.d_q_full, c.nq, c.k, c.row_base, c.d_keys, c.st, &c.prof));
  a.done = ex->d_done;
  c.prof.begin("push_results");
  const uint64_t n = c.nq * c.k;
  const unsigned blocks = (unsigned)std::min<uint64_t>(148, (n / 2 + 255) / 256 + 1);
  push_results_kernel<<<blocks, 256, 0, c.st>>>(a, c.d_keys, s->d_ids, n, c.k, c.per, (uint32_t)c.row_base);
  SCN_LAUNCHED();
  c.prof.end();
  return SCN_OK;
}

// step 3: wait for the peers' lists, merge this rank's slice from local memory
int32_t exchange_finish(ExchangeCall& c, uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts) {
  scn_exchange* ex = c.ex;
  PeerArgs& a = c.a;
  if (c.q_n && (!d_out_ids || !d_out_dist)) return fail(SCN_ERR_INVALID_PARAMETERS, "output pointer is NULL");
  c.prof.begin("wait_peers");
  wait_flags_kernel<<<1, 32, 0, c.st>>>(flag_of(ex->local, FLAG_LISTS, a.parity, 0), ex->world, a.epoch, ex->d_status, WAIT_TIMEOUT_NS);
  SCN_LAUNCHED();
  c.prof.end();
  if (c.q_n) {
    c.prof.begin("merge_topk");
    SCN_TRY(merge_topk(keys_of(ex->local, a.L, a.parity, 0), ids_of(ex->local, a.L, a.parity, 0), ex->world, c.q_n, c.k, d_out_ids,
                       d_out_dist, d_out_counts, c.st, a.L.le, ex->d_status));
    c.prof.end();
  }
  c.prof.collect();
  return SCN_OK;
}

// Host-buffer form in the same three steps (shards.cu drives them): the per-call device buffers live in `h`.
struct HostExchangeCall {
  ExchangeCall call;
  float* d_q = nullptr;
  uint64_t* d_ids = nullptr;
  float* d_dist = nullptr;
  uint32_t* d_counts = nullptr;
  HostExchangeCall(scn_store* s, scn_exchange* ex, cudaStream_t st) : call(s, ex, st) {}
};

HostExchangeCall* host_exchange_begin(scn_store* s, scn_exchange* ex, const float* q_slice, uint64_t nq, uint32_t k, uint64_t row_base,
                                      int32_t* rc) {
  *rc = SCN_OK;
  if (!s || !ex) {
    *rc = fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
    return nullptr;
  }
  if (ex->dim != s->dim) {
    *rc = fail(SCN_ERR_INVALID_PARAMETERS, "the exchange was created for dim %u, the store has %u", ex->dim, s->dim);
    return nullptr;
  }
  DeviceGuard g(s->device);
  HostExchangeCall* h = new HostExchangeCall(s, ex, thread_stream(s->device));
  uint64_t q_lo = 0, q_n = 0;
  scn_exchange_slice(ex, nq, ex->rank, &q_lo, &q_n);
  auto bail = [&](int32_t r) -> HostExchangeCall* {
    *rc = r;
    delete h;
    return nullptr;
  };
  if (q_n && !q_slice) return bail(fail(SCN_ERR_INVALID_PARAMETERS, "query pointer is NULL"));
  const uint64_t m = std::max<uint64_t>(q_n, 1);
  int32_t r = h->call.scratch.alloc(&h->d_q, m * s->dim);
  if (r == SCN_OK) r = h->call.scratch.alloc(&h->d_ids, m * k);
  if (r == SCN_OK) r = h->call.scratch.alloc(&h->d_dist, m * k);
  if (r == SCN_OK) r = h->call.scratch.alloc(&h->d_counts, m);
  if (r == SCN_OK && q_n) r = copy_to_device(h->d_q, q_slice, q_n * s->dim * sizeof(float), h->call.st);
  if (r == SCN_OK) r = exchange_begin(h->call, h->d_q, 1, nq, k, row_base);
  if (r != SCN_OK) return bail(r);
  return h;
}

int32_t host_exchange_search(HostExchangeCall* h) {
  DeviceGuard g(h->call.s->device);
  return exchange_search(h->call);
}

// consumes h
int32_t host_exchange_finish(HostExchangeCall* h, uint64_t* out_ids, float* out_dist, uint32_t* out_counts) {
  DeviceGuard g(h->call.s->device);
  const uint64_t q_n = h->call.q_n;
  const uint32_t k = h->call.k;
  cudaStream_t st = h->call.st;
  scn_exchange* ex = h->call.ex;
  int32_t rc = (q_n && (!out_ids || !out_dist)) ? fail(SCN_ERR_INVALID_PARAMETERS, "output pointer is NULL") : SCN_OK;
  if (rc == SCN_OK) rc = exchange_finish(h->call, h->d_ids, h->d_dist, h->d_counts);
  if (rc == SCN_OK && q_n) {
    cudaError_t e = cudaMemcpyAsync(out_ids, h->d_ids, q_n * k * sizeof(uint64_t), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess) e = cudaMemcpyAsync(out_dist, h->d_dist, q_n * k * sizeof(float), cudaMemcpyDeviceToHost, st);
    if (e == cudaSuccess && out_counts) e = cudaMemcpyAsync(out_counts, h->d_counts, q_n * sizeof(uint32_t), cudaMemcpyDeviceToHost, st);
    if (e != cudaSuccess) rc = cuda_fail(e, "result copy", __FILE__, __LINE__);
  }
  const int32_t st_rc = scn_exchange_status(ex, st);   // synchronises the stream
  delete h;
  return rc != SCN_OK ? rc : st_rc;
}

// everything this rank has enqueued so far has run (shards sharing a device: see shards.cu)
int32_t host_exchange_sync(HostExchangeCall* h) {
  DeviceGuard g(h->call.s->device);
  SCN_CUDA(cudaStreamSynchronize(h->call.st));
  return SCN_OK;
}

void host_exchange_abort(HostExchangeCall* h) {
  if (!h) return;
  cudaStreamSynchronize(h->call.st);
  delete h;
}

}  // namespace scn

extern "C" {

int32_t scn_search_flat_exchange_dev(scn_store* s, scn_exchange* ex, const float* d_q, int32_t q_is_slice, uint64_t nq, uint32_t k,
                                     uint64_t row_base, uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts,
                                     void* stream) {
  if (!s || !ex) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  DeviceGuard g(s->device);
  ExchangeCall c(s, ex, (cudaStream_t)stream);
  SCN_TRY(exchange_begin(c, d_q, q_is_slice, nq, k, row_base));
  SCN_TRY(exchange_search(c));
  return exchange_finish(c, d_out_ids, d_out_dist, d_out_counts);
}

// Synchronises `stream`; SCN_ERR_SEARCH_FAILED if a peer missed the 5 s arrival time-out since the
// previous call of this function (the results of such a call are empty). Clears the status.
int32_t scn_exchange_status(scn_exchange* ex, void* stream) {
  if (!ex) return fail(SCN_ERR_INVALID_PARAMETERS, "exchange is NULL");
  DeviceGuard g(ex->device);
  uint32_t v = 0;
  SCN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  SCN_CUDA(cudaMemcpy(&v, ex->d_status, sizeof v, cudaMemcpyDeviceToHost));
  if (v) {
    SCN_CUDA(cudaMemset(ex->d_status, 0, sizeof v));
    return fail(SCN_ERR_SEARCH_FAILED, "a peer shard did not deliver its part of the exchange within 5 s");
  }
  return SCN_OK;
}

// Host-buffer form of one rank's call: q_slice is this rank's slice of the batch (host memory,
// scn_exchange_slice(ex, nq, rank)), the outputs receive the results of that slice. Blocking.
int32_t scn_search_flat_exchange(scn_store* s, scn_exchange* ex, const float* q_slice, uint64_t nq, uint32_t k, uint64_t row_base,
                                 uint64_t* out_ids, float* out_dist, uint32_t* out_counts) {
  int32_t rc = SCN_OK;
  HostExchangeCall* h = host_exchange_begin(s, ex, q_slice, nq, k, row_base, &rc);
  if (!h) return rc;
  rc = host_exchange_search(h);
  if (rc != SCN_OK) {
    host_exchange_abort(h);
    return rc;
  }
  return host_exchange_finish(h, out_ids, out_dist, out_counts);
}

}  // extern "C"
