// exchange.cu — row-sharded exact search with the per-shard top-k exchange fused into the search's
// own epilogue over NVLink peer memory ("optional fusion").
//
// The NCCL formulation is: shard-local search -> keys_to_results -> all_gather(keys) ->
// all_gather(ids) -> merge_topk: three extra launches and two collectives for 16 B x nq x k per
// rank (1.6 MB at 10 000 x 10) — pure latency. Here every rank owns one exchange buffer
// [2 parities][world][max_nq*k] of (key, id) lists plus one arrival flag per (parity, peer), mapped
// into every other rank (cudaIpc between processes, cudaDeviceEnablePeerAccess inside one process —
// the reference server is ONE process, so that is the drop-in's natural form). The epilogue kernel
// of the shard-local search converts its keys to ids and stores both lists straight into all
// world buffers (plain 16-byte NVLink P2P stores), then the last block publishes the flags with a
// system-scope release. A one-warp wait kernel acquires the world flags (with a time-out that is
// reported, never a hang) and the ordinary merge kernel reads only local memory. Parity double
// buffering is enough: a rank can start call e+2 only after its merge of call e+1 saw every peer's
// flag e+1, which each peer sets after its own merge of call e.
#include <cstring>

#include "store.h"

using namespace scn;

struct scn_exchange {
  int32_t device = 0;
  uint32_t rank = 0, world = 1, k = 0;
  uint64_t max_nq = 0;
  uint64_t epoch = 0;
  unsigned char* local = nullptr;            // this rank's buffer
  unsigned char* peer[16] = {};              // peer[r] = rank r's buffer as seen from here (peer[rank] = local)
  bool opened_ipc[16] = {};
  bool connected = false;
  uint32_t* d_done = nullptr;                // block completion counter of the push kernel
  uint32_t* d_status = nullptr;              // 0 ok, 1 = wait timed out
  size_t bytes = 0;
};

namespace {

constexpr size_t FLAG_BYTES = 4096;  // [2][16] u32 flags, padded

__host__ __device__ inline size_t list_elems(uint64_t max_nq, uint32_t k) { return (size_t)max_nq * k; }
// layout: flags | keys[2][world][max_nq*k] | ids[2][world][max_nq*k]
__host__ __device__ inline uint32_t* flags_of(unsigned char* base) { return reinterpret_cast<uint32_t*>(base); }
__host__ __device__ inline uint64_t* keys_of(unsigned char* base, uint32_t parity, uint32_t world, uint32_t shard, size_t le) {
  return reinterpret_cast<uint64_t*>(base + FLAG_BYTES) + ((size_t)parity * world + shard) * le;
}
__host__ __device__ inline uint64_t* ids_of(unsigned char* base, uint32_t parity, uint32_t world, uint32_t shard, size_t le) {
  return reinterpret_cast<uint64_t*>(base + FLAG_BYTES) + ((size_t)2 * world + (size_t)parity * world + shard) * le;
}

struct PushArgs {
  unsigned char* peer[16];
  const uint64_t* keys;   // [nq*k] shard-local sorted keys (global rows)
  const uint64_t* row_ids;
  uint32_t world, rank, parity, epoch;
  uint64_t n;             // nq*k
  uint64_t row_base;
  size_t le;
  uint32_t* done;
};

// keys -> ids, then both lists into every rank's buffer; the last block raises the flags
__global__ void __launch_bounds__(256) push_results_kernel(PushArgs a) {
  for (uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x; i < a.n; i += (uint64_t)gridDim.x * blockDim.x) {
    const uint64_t key = a.keys[i];
    const uint64_t id = (key == KEY_NONE) ? 0ull : a.row_ids[(uint32_t)key - (uint32_t)a.row_base];
    for (uint32_t p = 0; p < a.world; ++p) {
      keys_of(a.peer[p], a.parity, a.world, a.rank, a.le)[i] = key;
      ids_of(a.peer[p], a.parity, a.world, a.rank, a.le)[i] = id;
    }
  }
  __threadfence_system();
  __syncthreads();
  if (threadIdx.x == 0) {
    const uint32_t prev = atomicAdd(a.done, 1u);
    if (prev == gridDim.x - 1) {
      *a.done = 0;
      __threadfence_system();
      for (uint32_t p = 0; p < a.world; ++p) {
        uint32_t* f = flags_of(a.peer[p]) + a.parity * 16 + a.rank;
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(f), "r"(a.epoch) : "memory");
      }
    }
  }
}

// one warp: lane r waits for rank r's flag of this call
__global__ void wait_flags_kernel(const uint32_t* flags, uint32_t world, uint32_t parity, uint32_t epoch, uint32_t* status,
                                  unsigned long long timeout_ns) {
  const uint32_t r = threadIdx.x;
  if (r >= world) return;
  const uint32_t* f = flags + parity * 16 + r;
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

}  // namespace

extern "C" {

int32_t scn_exchange_create(int32_t device, uint32_t rank, uint32_t world, uint64_t max_nq, uint32_t k, scn_exchange** out) {
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
  ex->max_nq = max_nq;
  ex->bytes = FLAG_BYTES + (size_t)4 * world * list_elems(max_nq, k) * sizeof(uint64_t);
  cudaError_t e = cudaMalloc(&ex->local, ex->bytes);  // cudaMalloc (not the async pool): IPC-exportable
  if (e == cudaSuccess) e = cudaMemset(ex->local, 0, FLAG_BYTES);
  if (e == cudaSuccess) e = cudaMalloc(&ex->d_done, 2 * sizeof(uint32_t));
  if (e == cudaSuccess) e = cudaMemset(ex->d_done, 0, 2 * sizeof(uint32_t));
  if (e != cudaSuccess) {
    cudaFree(ex->local);
    cudaFree(ex->d_done);
    delete ex;
    return cuda_fail(e, "exchange buffer allocation", __FILE__, __LINE__);
  }
  ex->d_status = ex->d_done + 1;
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
    if (!peers[r] || peers[r]->world != ex->world || peers[r]->rank != r || peers[r]->max_nq != ex->max_nq || peers[r]->k != ex->k)
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

int32_t scn_search_flat_exchange_dev(scn_store* s, scn_exchange* ex, const float* d_q, uint64_t nq, uint32_t k, uint64_t row_base,
                                     uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts, void* stream) {
  if (!s || !ex) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  if (!ex->connected) return fail(SCN_ERR_INVALID_PARAMETERS, "exchange is not connected to its peers");
  if (s->device != ex->device) return fail(SCN_ERR_INVALID_PARAMETERS, "store and exchange live on different devices");
  if (k != ex->k || nq == 0 || nq > ex->max_nq) return fail(SCN_ERR_INVALID_PARAMETERS, "nq must be in [1, %llu] and k == %u", (unsigned long long)ex->max_nq, ex->k);
  if (!d_q || !d_out_ids || !d_out_dist) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  DeviceGuard g(s->device);
  cudaStream_t st = (cudaStream_t)stream;
  Profiler prof(s, st);
  Scratch scratch(st);
  uint64_t* d_keys = nullptr;
  SCN_TRY(scratch.alloc(&d_keys, nq * k));
  SCN_TRY(flat_keys(s, d_q, nq, k, row_base, d_keys, st, &prof));
  const uint64_t epoch = ++ex->epoch;
  PushArgs a;
  for (int r = 0; r < 16; ++r) a.peer[r] = ex->peer[r];
  a.keys = d_keys;
  a.row_ids = s->d_ids;
  a.world = ex->world;
  a.rank = ex->rank;
  a.parity = (uint32_t)(epoch & 1);
  a.epoch = (uint32_t)epoch;
  a.n = nq * k;
  a.row_base = row_base;
  a.le = list_elems(ex->max_nq, k);
  a.done = ex->d_done;
  prof.begin("push_results");
  const unsigned blocks = (unsigned)std::min<uint64_t>(148, (a.n + 255) / 256);
  push_results_kernel<<<blocks, 256, 0, st>>>(a);
  SCN_LAUNCHED();
  prof.end();
  prof.begin("wait_peers");
  wait_flags_kernel<<<1, 32, 0, st>>>(flags_of(ex->local), ex->world, a.parity, a.epoch, ex->d_status, 5000000000ull);
  SCN_LAUNCHED();
  prof.end();
  prof.begin("merge_topk");
  SCN_TRY(merge_topk(keys_of(ex->local, a.parity, ex->world, 0, a.le), ids_of(ex->local, a.parity, ex->world, 0, a.le), ex->world, nq, k,
                     d_out_ids, d_out_dist, d_out_counts, st, a.le));
  prof.end();
  prof.collect();
  return SCN_OK;
}

// 0 = every wait so far was satisfied; 1 = a peer did not arrive within the time-out (results of
// that call are invalid). Synchronises the stream first.
int32_t scn_exchange_status(scn_exchange* ex, void* stream) {
  if (!ex) return fail(SCN_ERR_INVALID_PARAMETERS, "exchange is NULL");
  DeviceGuard g(ex->device);
  uint32_t v = 0;
  SCN_CUDA(cudaStreamSynchronize((cudaStream_t)stream));
  SCN_CUDA(cudaMemcpy(&v, ex->d_status, sizeof v, cudaMemcpyDeviceToHost));
  if (v) return fail(SCN_ERR_SEARCH_FAILED, "a peer shard did not deliver its top-k lists within 5 s");
  return SCN_OK;
}

}  // extern "C"
