// store.cu — device vector store + graph mirror + C-ABI plumbing (see include/scn_gpu.h).
#include "store.h"

#include <cstdarg>
#include <cstdio>
#include <cstring>

namespace scn {

static thread_local std::string g_last_error;
static std::atomic<uint64_t> g_launches{0};

int32_t fail(int32_t code, const char* fmt, ...) {
  char buf[512];
  va_list ap;
  va_start(ap, fmt);
  vsnprintf(buf, sizeof buf, fmt, ap);
  va_end(ap);
  g_last_error = buf;
  return code;
}

int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line) {
  cudaGetLastError();
  int32_t code = (e == cudaErrorMemoryAllocation) ? SCN_ERR_RESOURCE : SCN_ERR_INTERNAL;
  return fail(code, "CUDA error %d (%s) at %s:%d: %s", (int)e, cudaGetErrorString(e), file, line, what);
}

void count_launch(uint64_t n) { g_launches.fetch_add(n, std::memory_order_relaxed); }

int max_optin_smem() {
  // all devices of a box are the same part (the library refuses anything but sm_100)
  static const int v = [] {
    int dev = 0, val = 0;
    cudaGetDevice(&dev);
    cudaDeviceGetAttribute(&val, cudaDevAttrMaxSharedMemoryPerBlockOptin, dev);
    return val;
  }();
  return v;
}

int static_smem_of(const void* kernel) {
  cudaFuncAttributes fa{};
  if (cudaFuncGetAttributes(&fa, kernel) != cudaSuccess) {
    cudaGetLastError();
    return 1024;
  }
  return (int)fa.sharedSizeBytes;
}

// Per-thread CUDA resources: one non-blocking stream per device (concurrent Search calls from
// different goroutines/threads overlap on the GPU; the reference allows concurrent readers,
// hnsw.go:293) and two pinned staging chunks for pageable callers. Released when the thread ends
// (cgo calls run on whatever OS thread the Go scheduler picks, and those threads come and go).
struct ThreadResources {
  cudaStream_t streams[64] = {};
  void* stage[2] = {nullptr, nullptr};
  cudaEvent_t stage_ev[2] = {nullptr, nullptr};
  bool stage_busy[2] = {false, false};
  ~ThreadResources() {
    // best effort: at process exit the driver may already be gone, errors are ignored
    for (int d = 0; d < 64; ++d)
      if (streams[d]) cudaStreamDestroy(streams[d]);
    for (int b = 0; b < 2; ++b) {
      if (stage_ev[b]) cudaEventDestroy(stage_ev[b]);
      if (stage[b]) cudaFreeHost(stage[b]);
    }
    cudaGetLastError();
  }
};
static thread_local ThreadResources g_thread;

cudaStream_t thread_stream(int device) {
  if (device < 0 || device >= 64) return nullptr;
  if (!g_thread.streams[device]) {
    DeviceGuard g(device);
    cudaStreamCreateWithFlags(&g_thread.streams[device], cudaStreamNonBlocking);
  }
  return g_thread.streams[device];
}

constexpr size_t STAGE_CHUNK = (size_t)4 << 20;

int32_t copy_to_device(void* d_dst, const void* h_src, size_t bytes, cudaStream_t stream) {
  if (bytes == 0) return SCN_OK;
  cudaPointerAttributes at{};
  const bool known = cudaPointerGetAttributes(&at, h_src) == cudaSuccess;
  if (!known) cudaGetLastError();
  if ((known && at.type != cudaMemoryTypeUnregistered) || bytes <= (64u << 10)) {
    SCN_CUDA(cudaMemcpyAsync(d_dst, h_src, bytes, cudaMemcpyDefault, stream));
    return SCN_OK;
  }
  ThreadResources& t = g_thread;
  for (int b = 0; b < 2; ++b) {
    if (!t.stage[b]) {
      SCN_CUDA(cudaHostAlloc(&t.stage[b], STAGE_CHUNK, cudaHostAllocPortable));
      SCN_CUDA(cudaEventCreateWithFlags(&t.stage_ev[b], cudaEventDisableTiming));
    }
  }
  size_t off = 0;
  for (int b = 0; off < bytes; b ^= 1) {
    const size_t n = std::min(STAGE_CHUNK, bytes - off);
    if (t.stage_busy[b]) SCN_CUDA(cudaEventSynchronize(t.stage_ev[b]));  // the DMA that last read this chunk
    std::memcpy(t.stage[b], static_cast<const unsigned char*>(h_src) + off, n);
    SCN_CUDA(cudaMemcpyAsync(static_cast<unsigned char*>(d_dst) + off, t.stage[b], n, cudaMemcpyHostToDevice, stream));
    SCN_CUDA(cudaEventRecord(t.stage_ev[b], stream));
    t.stage_busy[b] = true;
    off += n;
  }
  return SCN_OK;
}

void Profiler::collect() {
  if (!on || ev.empty()) return;
  std::lock_guard<std::mutex> lk(s->mu);
  for (auto& e : ev) s->pending.push_back({e.first, e.second.first, e.second.second});
  ev.clear();
}

Profiler::~Profiler() {
  for (auto& e : ev) {  // only reached when collect() was not called (error paths)
    cudaEventDestroy(e.second.first);
    cudaEventDestroy(e.second.second);
  }
}

// ---- row preparation (K6): norms, bf16 mirror, certificate bounds -----------------------------
// One thread per row, strictly sequential over the row so that `norm` is bit-identical to the
// reference's normB (distance.go:58-66) / VectorMagnitude (distance.go:175-181).
template <int METRIC>
__global__ void __launch_bounds__(128) prepare_rows_kernel(const float* __restrict__ vec, uint32_t pitch, uint32_t dim,
                                                           uint32_t kpad, uint64_t first, uint64_t n,
                                                           float* __restrict__ norm, __nv_bfloat16* __restrict__ mirror,
                                                           float* __restrict__ aux, float* __restrict__ bounds) {
  uint64_t r = first + (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  float mir_norm = 0.f, err_norm = 0.f, x_norm = 0.f;
  if (r < first + n) {
    const float* x = vec + r * pitch;
    float ss = 0.0f;
    for (uint32_t i = 0; i < dim; ++i) ss = __fadd_rn(ss, __fmul_rn(x[i], x[i]));
    float nrm = __fsqrt_rn(ss);
    norm[r] = nrm;
    x_norm = nrm;
    float inv = (METRIC == M_COS && nrm > 0.0f) ? (1.0f / nrm) : 1.0f;
    float mm = 0.f, ee = 0.f;
    __nv_bfloat16* mrow = mirror + r * kpad;
    for (uint32_t i = 0; i < kpad; i += 8) {
      __align__(16) __nv_bfloat16 out[8];
#pragma unroll
      for (int j = 0; j < 8; ++j) {
        float v = (i + j < dim) ? x[i + j] * inv : 0.0f;
        __nv_bfloat16 b = __float2bfloat16_rn(v);
        float bv = __bfloat162float(b);
        mm = fmaf(bv, bv, mm);
        ee = fmaf(v - bv, v - bv, ee);
        out[j] = b;
      }
      *reinterpret_cast<uint4*>(mrow + i) = *reinterpret_cast<const uint4*>(out);
    }
    aux[r] = (METRIC == M_L2) ? mm : 0.0f;
    mir_norm = sqrtf(mm) * 1.0000005f;
    err_norm = sqrtf(ee) * 1.0000005f;
  }
  // block max -> global max (floats are non-negative: integer compare on the bits is monotone)
  for (int o = 16; o > 0; o >>= 1) {
    mir_norm = fmaxf(mir_norm, __shfl_xor_sync(0xffffffffu, mir_norm, o));
    err_norm = fmaxf(err_norm, __shfl_xor_sync(0xffffffffu, err_norm, o));
    x_norm = fmaxf(x_norm, __shfl_xor_sync(0xffffffffu, x_norm, o));
  }
  if ((threadIdx.x & 31) == 0) {
    atomicMax(reinterpret_cast<unsigned int*>(bounds + 0), __float_as_uint(mir_norm));
    atomicMax(reinterpret_cast<unsigned int*>(bounds + 1), __float_as_uint(err_norm));
    if (x_norm == x_norm) atomicMax(reinterpret_cast<unsigned int*>(bounds + 2), __float_as_uint(x_norm));
  }
}

int32_t launch_prepare_rows(scn_store* s, uint64_t first_row, uint64_t n, cudaStream_t stream) {
  if (n == 0) return SCN_OK;
  dim3 grid((unsigned)((n + 127) / 128)), block(128);
#define PREP(MT)                                                                                              \
  prepare_rows_kernel<MT><<<grid, block, 0, stream>>>(s->d_vec, s->pitch, s->dim, s->kpad, first_row, n, s->d_norm, \
                                                      s->d_mirror, s->d_aux, s->d_bounds)
  switch (s->metric) {
    case M_L2: PREP(M_L2); break;
    case M_COS: PREP(M_COS); break;
    default: PREP(M_IP); break;
  }
#undef PREP
  SCN_LAUNCHED();
  return SCN_OK;
}

// ---- compaction gather: dst[new_row] = src[src_row[new_row]], rows of `bytes` bytes (multiple of 4) --
__global__ void __launch_bounds__(256) gather_rows_kernel(const unsigned char* __restrict__ src, unsigned char* __restrict__ dst,
                                                          const uint32_t* __restrict__ src_row, uint64_t n_new, uint32_t bytes) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= n_new) return;
  const unsigned char* s = src + (uint64_t)src_row[warp] * bytes;
  unsigned char* d = dst + warp * bytes;
  if ((bytes & 15u) == 0) {
    for (uint32_t i = lane; i < bytes / 16; i += 32) reinterpret_cast<uint4*>(d)[i] = reinterpret_cast<const uint4*>(s)[i];
  } else {
    for (uint32_t i = lane; i < bytes / 4; i += 32) reinterpret_cast<uint32_t*>(d)[i] = reinterpret_cast<const uint32_t*>(s)[i];
  }
}

template <class T>
static int32_t gather_array(T** arr, const uint32_t* d_src_row, uint64_t n_new, uint64_t new_cap, uint32_t elems_per_row,
                            cudaStream_t st) {
  T* np = nullptr;
  SCN_CUDA(cudaMalloc(&np, std::max<uint64_t>(new_cap * elems_per_row, 1) * sizeof(T)));
  SCN_CUDA(cudaMemsetAsync(np, 0, std::max<uint64_t>(new_cap * elems_per_row, 1) * sizeof(T), st));
  if (n_new) {
    const uint32_t bytes = elems_per_row * (uint32_t)sizeof(T);
    const uint64_t threads = n_new * 32;
    gather_rows_kernel<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(reinterpret_cast<const unsigned char*>(*arr),
                                                                          reinterpret_cast<unsigned char*>(np), d_src_row, n_new,
                                                                          bytes);
    SCN_LAUNCHED();
  }
  SCN_CUDA(cudaStreamSynchronize(st));
  cudaFree(*arr);
  *arr = np;
  return SCN_OK;
}

}  // namespace scn

using namespace scn;

uint64_t scn_store::device_bytes() const {
  uint64_t b = cap * ((uint64_t)pitch * 4 + (uint64_t)kpad * 2 + 4 + 4 + 8) + (cap / 32 + 1) * 4 + 16;
  if (has_graph) b += graph_nodes * ((uint64_t)2 * m * 4 + 1 + 4) + upper_lists * (uint64_t)m * 4;
  return b;
}

template <class T>
static int32_t grow(T** p, uint64_t old_count, uint64_t new_count, cudaStream_t st) {
  T* np = nullptr;
  SCN_CUDA(cudaMalloc(&np, std::max<uint64_t>(new_count, 1) * sizeof(T)));
  if (*p && old_count) SCN_CUDA(cudaMemcpyAsync(np, *p, old_count * sizeof(T), cudaMemcpyDeviceToDevice, st));
  if (new_count > old_count)
    SCN_CUDA(cudaMemsetAsync(np + old_count, 0, (new_count - old_count) * sizeof(T), st));
  SCN_CUDA(cudaStreamSynchronize(st));
  if (*p) cudaFree(*p);
  *p = np;
  return SCN_OK;
}

static int32_t ensure_capacity(scn_store* s, uint64_t need) {
  if (need <= s->cap) return SCN_OK;
  if (need >= (uint64_t)ROW_NONE) return fail(SCN_ERR_RESOURCE, "store limited to 2^32-2 rows per device");
  uint64_t nc = std::max<uint64_t>(need, s->cap + s->cap / 2);
  nc = (nc + 255) / 256 * 256;
  cudaStream_t st = thread_stream(s->device);
  SCN_TRY(grow(&s->d_vec, s->rows * s->pitch, nc * s->pitch, st));
  SCN_TRY(grow(&s->d_norm, s->rows, nc, st));
  SCN_TRY(grow(&s->d_mirror, s->rows * s->kpad, nc * s->kpad, st));
  SCN_TRY(grow(&s->d_aux, s->rows, nc, st));
  SCN_TRY(grow(&s->d_ids, s->rows, nc, st));
  SCN_TRY(grow(&s->d_deleted, (s->cap + 31) / 32, (nc + 31) / 32, st));
  s->cap = nc;
  return SCN_OK;
}

static void free_graph(scn_store* s) {
  free_build_state(s);
  cudaFree(s->d_adj0);
  cudaFree(s->d_levels);
  cudaFree(s->d_up_off);
  cudaFree(s->d_adj_up);
  s->d_adj0 = s->d_up_off = s->d_adj_up = nullptr;
  s->d_levels = nullptr;
  s->has_graph = false;
  s->max_layer = -1;
  s->entry_id = 0;
  s->entry_row = ROW_NONE;
  s->graph_nodes = s->graph_edges = s->upper_lists = 0;
  s->h_node_layer.clear();
}

extern "C" {

const char* scn_last_error(void) { return g_last_error.c_str(); }
uint64_t scn_launch_count(void) { return g_launches.load(); }

int32_t scn_host_alloc(uint64_t bytes, void** out) {
  if (!out) return fail(SCN_ERR_INVALID_PARAMETERS, "out is NULL");
  *out = nullptr;
  cudaError_t e = cudaHostAlloc(out, std::max<uint64_t>(bytes, 1), cudaHostAllocPortable);
  if (e != cudaSuccess) {
    cudaGetLastError();
    return fail(SCN_ERR_RESOURCE, "pinned host allocation of %llu bytes failed: %s", (unsigned long long)bytes, cudaGetErrorString(e));
  }
  return SCN_OK;
}

int32_t scn_host_free(void* p) {
  if (p) SCN_CUDA(cudaFreeHost(p));
  return SCN_OK;
}

int32_t scn_store_create(int32_t device, uint32_t dim, int32_t metric, scn_store** out) {
  if (!out) return fail(SCN_ERR_INVALID_PARAMETERS, "out is NULL");
  *out = nullptr;
  if (metric != M_L2 && metric != M_COS && metric != M_IP)
    return fail(SCN_ERR_INVALID_PARAMETERS, "unsupported distance metric");  // distance.go:138
  if (dim == 0 || dim > 65536) return fail(SCN_ERR_INVALID_PARAMETERS, "dimension must be in [1, 65536]");
  int ndev = 0;
  SCN_CUDA(cudaGetDeviceCount(&ndev));
  if (device < 0 || device >= ndev) return fail(SCN_ERR_INVALID_PARAMETERS, "no CUDA device %d (have %d)", device, ndev);
  DeviceGuard g(device);
  cudaDeviceProp prop;
  SCN_CUDA(cudaGetDeviceProperties(&prop, device));
  if (prop.major != 10)
    return fail(SCN_ERR_INTERNAL, "device %d is sm_%d%d; this library is built for sm_100a only", device, prop.major,
                prop.minor);
  {
    // Per-call scratch comes from the stream-ordered pool; keep freed blocks cached across the
    // synchronisations of the blocking entry points instead of returning them to the driver.
    cudaMemPool_t pool;
    if (cudaDeviceGetDefaultMemPool(&pool, device) == cudaSuccess) {
      uint64_t keep = ~0ull;
      cudaMemPoolSetAttribute(pool, cudaMemPoolAttrReleaseThreshold, &keep);
      // Never let an allocation on one stream wait for a free that is still pending on another
      // stream: two ranks of the fused shard exchange driven on ONE device would deadlock (rank B's
      // allocation would wait for rank A's free, which is ordered after A's wait for B's push).
      int off = 0;
      cudaMemPoolSetAttribute(pool, cudaMemPoolReuseAllowInternalDependencies, &off);
    }
    cudaGetLastError();
  }
  scn_store* s = new scn_store();
  s->device = device;
  s->dim = dim;
  s->pitch = round_up(dim, 8);  // rows 32-byte aligned (256-bit loads in hnsw_search)
  s->kpad = round_up(dim, 64);
  s->metric = metric;
  if (cudaMalloc(&s->d_bounds, 4 * sizeof(float)) != cudaSuccess) {
    delete s;
    return cuda_fail(cudaGetLastError(), "cudaMalloc(bounds)", __FILE__, __LINE__);
  }
  cudaMemset(s->d_bounds, 0, 4 * sizeof(float));
  if (cudaMalloc(&s->d_counters, 4 * sizeof(unsigned long long)) != cudaSuccess) {
    cudaFree(s->d_bounds);
    delete s;
    return cuda_fail(cudaGetLastError(), "cudaMalloc(counters)", __FILE__, __LINE__);
  }
  cudaMemset(s->d_counters, 0, 4 * sizeof(unsigned long long));
  *out = s;
  return SCN_OK;
}

int32_t scn_store_destroy(scn_store* s) {
  if (!s) return SCN_OK;
  DeviceGuard g(s->device);
  cudaDeviceSynchronize();
  for (auto& e : s->pending) {
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
  }
  free_graph(s);
  cudaFree(s->d_vec);
  cudaFree(s->d_norm);
  cudaFree(s->d_mirror);
  cudaFree(s->d_aux);
  cudaFree(s->d_ids);
  cudaFree(s->d_deleted);
  cudaFree(s->d_bounds);
  cudaFree(s->d_counters);
  delete s;
  return SCN_OK;
}

int32_t scn_store_reserve(scn_store* s, uint64_t rows) {
  if (!s) return fail(SCN_ERR_INVALID_PARAMETERS, "store is NULL");
  DeviceGuard g(s->device);
  return ensure_capacity(s, rows);
}

int32_t scn_store_clear(scn_store* s) {
  if (!s) return fail(SCN_ERR_INVALID_PARAMETERS, "store is NULL");
  DeviceGuard g(s->device);
  cudaDeviceSynchronize();
  free_graph(s);
  s->rows = s->live = 0;
  s->auto_ids = true;
  s->row_of.clear();
  if (s->d_deleted) SCN_CUDA(cudaMemset(s->d_deleted, 0, ((s->cap + 31) / 32) * 4));
  SCN_CUDA(cudaMemset(s->d_bounds, 0, 4 * sizeof(float)));
  return SCN_OK;
}

static int32_t append_common(scn_store* s, const float* src, bool src_on_device, const uint64_t* ids, uint64_t n) {
  if (!s) return fail(SCN_ERR_INVALID_PARAMETERS, "store is NULL");
  if (n == 0) return SCN_OK;
  if (!src) return fail(SCN_ERR_INVALID_PARAMETERS, "vectors pointer is NULL");
  DeviceGuard g(s->device);
  // validate ids before touching anything (hnsw.go:192-194: duplicate -> error; id 0 is the
  // reference's "no entrypoint" sentinel, hnsw.go:210/296)
  std::vector<uint64_t> auto_buf;
  if (ids) {
    bool still_auto = s->auto_ids;
    for (uint64_t i = 0; i < n && still_auto; ++i) still_auto = (ids[i] == s->auto_base + s->rows + i + 1);
    if (!still_auto) {
      if (s->auto_ids) {  // materialise the implicit map
        s->row_of.reserve((size_t)(s->rows + n) * 2);
        for (uint64_t r = 0; r < s->rows; ++r) s->row_of[s->auto_base + r + 1] = (uint32_t)r;
        s->auto_ids = false;
      }
      for (uint64_t i = 0; i < n; ++i) {
        if (ids[i] == 0) return fail(SCN_ERR_INVALID_PARAMETERS, "vector id 0 is reserved");
        if (s->row_of.count(ids[i]))
          return fail(SCN_ERR_INVALID_PARAMETERS, "vector with ID %llu already exists", (unsigned long long)ids[i]);
      }
      // duplicates inside the batch
      for (uint64_t i = 0; i < n; ++i) {
        auto ins = s->row_of.emplace(ids[i], (uint32_t)(s->rows + i));
        if (!ins.second) {
          for (uint64_t j = 0; j < i; ++j) s->row_of.erase(ids[j]);
          return fail(SCN_ERR_INVALID_PARAMETERS, "vector with ID %llu already exists", (unsigned long long)ids[i]);
        }
      }
    }
  } else {
    if (!s->auto_ids) {
      // explicit ids were used before: auto ids continue from row+1 only if free
      for (uint64_t i = 0; i < n; ++i)
        if (s->row_of.count(s->auto_base + s->rows + i + 1))
          return fail(SCN_ERR_INVALID_PARAMETERS, "auto id %llu collides with an explicit id",
                      (unsigned long long)(s->auto_base + s->rows + i + 1));
      for (uint64_t i = 0; i < n; ++i) s->row_of[s->auto_base + s->rows + i + 1] = (uint32_t)(s->rows + i);
    }
    auto_buf.resize(n);
    for (uint64_t i = 0; i < n; ++i) auto_buf[i] = s->auto_base + s->rows + i + 1;
    ids = auto_buf.data();
  }
  int32_t rc = ensure_capacity(s, s->rows + n);
  if (rc != SCN_OK) {
    if (!s->auto_ids)
      for (uint64_t i = 0; i < n; ++i) s->row_of.erase(ids[i]);
    return rc;
  }
  cudaStream_t st = thread_stream(s->device);
  cudaMemcpyKind kind = src_on_device ? cudaMemcpyDeviceToDevice : cudaMemcpyHostToDevice;
  float* dst = s->d_vec + s->rows * s->pitch;
  if (s->pitch == s->dim) {
    SCN_CUDA(cudaMemcpyAsync(dst, src, n * s->dim * sizeof(float), kind, st));
  } else {
    SCN_CUDA(cudaMemsetAsync(dst, 0, n * s->pitch * sizeof(float), st));
    SCN_CUDA(cudaMemcpy2DAsync(dst, s->pitch * sizeof(float), src, s->dim * sizeof(float), s->dim * sizeof(float), n,
                               kind, st));
  }
  SCN_CUDA(cudaMemcpyAsync(s->d_ids + s->rows, ids, n * sizeof(uint64_t), cudaMemcpyHostToDevice, st));
  SCN_TRY(launch_prepare_rows(s, s->rows, n, st));
  SCN_CUDA(cudaStreamSynchronize(st));
  s->rows += n;
  s->live += n;
  return SCN_OK;
}

int32_t scn_store_append(scn_store* s, const float* vecs, const uint64_t* ids, uint64_t n) {
  return append_common(s, vecs, false, ids, n);
}
int32_t scn_store_append_dev(scn_store* s, const float* d_vecs, const uint64_t* ids, uint64_t n) {
  return append_common(s, d_vecs, true, ids, n);
}

// keep_entry: restore semantics — ImportGraphState takes entrypoint / maxLayer verbatim (hnsw.go:791-793),
// so a snapshot whose entry point is flagged deleted keeps it (and Search then returns nothing, like
// the reference: searchLayer drops a deleted entry point, hnsw.go:492-502).
static int32_t mark_deleted_impl(scn_store* s, const uint64_t* ids, uint64_t n, bool keep_entry) {
  if (!s) return fail(SCN_ERR_INVALID_PARAMETERS, "store is NULL");
  if (n == 0) return SCN_OK;
  DeviceGuard g(s->device);
  std::vector<uint32_t> rows(n);
  for (uint64_t i = 0; i < n; ++i)
    if (!s->lookup(ids[i], &rows[i]))
      return fail(SCN_ERR_VECTOR_NOT_FOUND, "vector %llu not found", (unsigned long long)ids[i]);  // hnsw.go:271-273
  // read-modify-write of the bitmap words on the host side of the mirror (mutators are exclusive)
  size_t words = (s->rows + 31) / 32;
  std::vector<uint32_t> bits(words);
  cudaStream_t st = thread_stream(s->device);
  SCN_CUDA(cudaMemcpyAsync(bits.data(), s->d_deleted, words * 4, cudaMemcpyDeviceToHost, st));
  SCN_CUDA(cudaStreamSynchronize(st));
  for (uint32_t r : rows) {
    uint32_t& w = bits[r >> 5];
    uint32_t b = 1u << (r & 31);
    if (!(w & b)) {  // already deleted -> no-op (hnsw.go:275-277)
      w |= b;
      s->live--;
    }
  }
  SCN_CUDA(cudaMemcpyAsync(s->d_deleted, bits.data(), words * 4, cudaMemcpyHostToDevice, st));
  SCN_CUDA(cudaStreamSynchronize(st));
  // hnsw.go:280-283: the entry point was deleted -> findNewEntrypoint (617-634): the live node with
  // the highest getNodeLayer becomes the entry point and its layer the new maxLayer. (The reference
  // walks a Go map, i.e. ties fall in random order; here, as in the oracle, in insertion order.)
  if (!keep_entry && s->has_graph && s->entry_row != ROW_NONE && ((bits[s->entry_row >> 5] >> (s->entry_row & 31)) & 1u)) {
    int best = -1;
    uint32_t best_row = ROW_NONE;
    const uint64_t n_graph = std::min<uint64_t>(s->graph_nodes, s->h_node_layer.size());
    for (uint64_t r = 0; r < n_graph; ++r) {
      if ((bits[r >> 5] >> (r & 31)) & 1u) continue;
      if ((int)s->h_node_layer[r] > best) {
        best = s->h_node_layer[r];
        best_row = (uint32_t)r;
      }
    }
    s->entry_row = best_row;
    s->max_layer = best;
    s->entry_id = 0;
    if (best_row != ROW_NONE) SCN_CUDA(cudaMemcpy(&s->entry_id, s->d_ids + best_row, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  }
  // the tensor filter learns about deletions through its per-row additive term (+Inf)
  return mark_aux_deleted(s, rows.data(), (uint32_t)rows.size(), st);
}

int32_t scn_store_mark_deleted(scn_store* s, const uint64_t* ids, uint64_t n) { return mark_deleted_impl(s, ids, n, false); }
int32_t scn_store_restore_deleted(scn_store* s, const uint64_t* ids, uint64_t n) { return mark_deleted_impl(s, ids, n, true); }

// Collection.Compact (collection.go:283-313): drop the soft-deleted vectors for good. Surviving
// rows keep their insertion order (and their ids); the graph is dropped, because the reference
// rebuilds the index from the surviving vectors (index.Build) and hands the new graph over again.
int32_t scn_store_compact(scn_store* s, uint64_t* out_removed) {
  if (!s) return fail(SCN_ERR_INVALID_PARAMETERS, "store is NULL");
  DeviceGuard g(s->device);
  cudaDeviceSynchronize();
  if (out_removed) *out_removed = s->rows - s->live;
  free_graph(s);
  if (s->live == s->rows) return SCN_OK;
  cudaStream_t st = thread_stream(s->device);
  const size_t words = (s->rows + 31) / 32;
  std::vector<uint32_t> bits(words);
  SCN_CUDA(cudaMemcpy(bits.data(), s->d_deleted, words * 4, cudaMemcpyDeviceToHost));
  std::vector<uint32_t> src_row;
  src_row.reserve(s->live);
  for (uint64_t r = 0; r < s->rows; ++r)
    if (!((bits[r >> 5] >> (r & 31)) & 1u)) src_row.push_back((uint32_t)r);
  const uint64_t n_new = src_row.size();
  // ids of the survivors: the implicit id = row + 1 rule no longer holds once a row is gone
  std::vector<uint64_t> old_ids;
  if (!s->auto_ids) {
    old_ids.resize(s->rows);
    SCN_CUDA(cudaMemcpy(old_ids.data(), s->d_ids, s->rows * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  }
  std::unordered_map<uint64_t, uint32_t> row_of;
  row_of.reserve(n_new * 2);
  for (uint64_t i = 0; i < n_new; ++i) row_of[s->auto_ids ? s->auto_base + (uint64_t)src_row[i] + 1 : old_ids[src_row[i]]] = (uint32_t)i;
  uint32_t* d_src = nullptr;
  SCN_CUDA(cudaMalloc(&d_src, std::max<uint64_t>(n_new, 1) * sizeof(uint32_t)));
  SCN_CUDA(cudaMemcpy(d_src, src_row.data(), n_new * sizeof(uint32_t), cudaMemcpyHostToDevice));
  const uint64_t new_cap = std::max<uint64_t>((n_new + 255) / 256 * 256, 256);
  int32_t rc = gather_array(&s->d_vec, d_src, n_new, new_cap, s->pitch, st);
  if (rc == SCN_OK) rc = gather_array(&s->d_norm, d_src, n_new, new_cap, 1, st);
  if (rc == SCN_OK) rc = gather_array(&s->d_mirror, d_src, n_new, new_cap, s->kpad, st);
  if (rc == SCN_OK) rc = gather_array(&s->d_aux, d_src, n_new, new_cap, 1, st);
  if (rc == SCN_OK) rc = gather_array(&s->d_ids, d_src, n_new, new_cap, 1, st);
  cudaFree(d_src);
  SCN_TRY(rc);
  cudaFree(s->d_deleted);
  s->d_deleted = nullptr;
  SCN_CUDA(cudaMalloc(&s->d_deleted, ((new_cap + 31) / 32) * 4));
  SCN_CUDA(cudaMemset(s->d_deleted, 0, ((new_cap + 31) / 32) * 4));
  s->cap = new_cap;
  s->rows = s->live = n_new;
  s->auto_ids = false;
  s->row_of.swap(row_of);
  return SCN_OK;
}

int32_t scn_store_stats(scn_store* s, scn_stats* out) {
  if (!s || !out) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  std::memset(out, 0, sizeof *out);
  out->rows = s->rows;
  out->live_rows = s->live;
  out->capacity_rows = s->cap;
  out->device_bytes = s->device_bytes();
  out->dim = s->dim;
  out->metric = s->metric;
  out->device = s->device;
  out->has_graph = s->has_graph ? 1 : 0;
  out->max_layer = s->max_layer;
  out->m = s->m;
  out->entry_id = s->entry_id;
  out->graph_edges = s->graph_edges;
  return SCN_OK;
}

// rows by index -> dense [n][dim] block (one warp per row), for scn_store_get
__global__ void __launch_bounds__(256) get_rows_kernel(const float* __restrict__ vec, uint32_t pitch, uint32_t dim,
                                                       const uint32_t* __restrict__ rows, uint64_t n, float* __restrict__ out) {
  const uint64_t warp = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const uint32_t lane = threadIdx.x & 31;
  if (warp >= n) return;
  const float* src = vec + (uint64_t)rows[warp] * pitch;
  float* dst = out + warp * dim;
  for (uint32_t i = lane; i < dim; i += 32) dst[i] = __ldg(src + i);
}

int32_t scn_store_get(scn_store* s, const uint64_t* ids, uint64_t n, float* out) {
  if (!s || (!ids && n) || (!out && n)) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  DeviceGuard g(s->device);
  cudaStream_t st = thread_stream(s->device);
  std::vector<uint32_t> rows(n);
  for (uint64_t i = 0; i < n; ++i) {
    if (!s->lookup(ids[i], &rows[i])) return fail(SCN_ERR_VECTOR_NOT_FOUND, "vector %llu not found", (unsigned long long)ids[i]);
  }
  if (s->live != s->rows) {  // HNSW.Get: a soft-deleted node is "not found" (hnsw.go:364-366)
    const size_t words = (s->rows + 31) / 32;
    std::vector<uint32_t> bits(words);
    SCN_CUDA(cudaMemcpyAsync(bits.data(), s->d_deleted, words * 4, cudaMemcpyDeviceToHost, st));
    SCN_CUDA(cudaStreamSynchronize(st));
    for (uint64_t i = 0; i < n; ++i)
      if ((bits[rows[i] >> 5] >> (rows[i] & 31)) & 1u)
        return fail(SCN_ERR_VECTOR_NOT_FOUND, "vector %llu not found", (unsigned long long)ids[i]);
  }
  if (n <= 4) {  // HNSW.Get of one vector: straight copies
    for (uint64_t i = 0; i < n; ++i)
      SCN_CUDA(cudaMemcpyAsync(out + i * s->dim, s->d_vec + (uint64_t)rows[i] * s->pitch, s->dim * sizeof(float),
                               cudaMemcpyDeviceToHost, st));
    SCN_CUDA(cudaStreamSynchronize(st));
    return SCN_OK;
  }
  // result decoration for whole batches (include_vector): gather on the device, one copy per block of rows
  const uint64_t block_rows = std::max<uint64_t>(1, ((uint64_t)64 << 20) / ((uint64_t)s->dim * sizeof(float)));
  Scratch scratch(st);
  uint32_t* d_rows = nullptr;
  float* d_out = nullptr;
  SCN_TRY(scratch.alloc(&d_rows, std::min(n, block_rows)));
  SCN_TRY(scratch.alloc(&d_out, std::min(n, block_rows) * s->dim));
  for (uint64_t i0 = 0; i0 < n; i0 += block_rows) {
    const uint64_t m = std::min(block_rows, n - i0);
    SCN_CUDA(cudaMemcpyAsync(d_rows, rows.data() + i0, m * sizeof(uint32_t), cudaMemcpyHostToDevice, st));
    get_rows_kernel<<<(unsigned)((m * 32 + 255) / 256), 256, 0, st>>>(s->d_vec, s->pitch, s->dim, d_rows, m, d_out);
    SCN_LAUNCHED();
    SCN_CUDA(cudaMemcpyAsync(out + i0 * s->dim, d_out, m * s->dim * sizeof(float), cudaMemcpyDeviceToHost, st));
    SCN_CUDA(cudaStreamSynchronize(st));
  }
  return SCN_OK;
}

int32_t scn_graph_upload(scn_store* s, int32_t m, int32_t max_layer, uint64_t entry_id, uint64_t n_nodes,
                         const uint64_t* node_ids, const int32_t* list_counts, const uint32_t* edge_counts,
                         const uint64_t* edges) {
  if (!s) return fail(SCN_ERR_INVALID_PARAMETERS, "store is NULL");
  if (m <= 0 || m > 512) return fail(SCN_ERR_INVALID_PARAMETERS, "M must be in [1, 512]");
  if (n_nodes && (!node_ids || !list_counts || !edge_counts)) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  DeviceGuard g(s->device);
  cudaDeviceSynchronize();
  const uint32_t s0 = 2 * (uint32_t)m, su = (uint32_t)m;
  uint64_t n = s->rows;
  std::vector<uint32_t> adj0((size_t)n * s0, ROW_NONE);
  std::vector<uint8_t> levels(n, 0);
  std::vector<uint32_t> up_off(n, 0);
  // first pass: rows, levels, upper offsets
  std::vector<uint32_t> node_row(n_nodes);
  uint64_t upper = 0;
  for (uint64_t i = 0; i < n_nodes; ++i) {
    uint32_t r;
    if (!s->lookup(node_ids[i], &r))
      return fail(SCN_ERR_INDEX_BUILD_FAILED, "graph node %llu is not in the store", (unsigned long long)node_ids[i]);
    node_row[i] = r;
    int lc = list_counts[i];
    if (lc < 0 || lc > 255) return fail(SCN_ERR_INVALID_PARAMETERS, "node %llu has %d layers", (unsigned long long)node_ids[i], lc);
    levels[r] = (uint8_t)(lc > 0 ? lc - 1 : 0);
    up_off[r] = (uint32_t)upper;
    if (lc > 1) upper += (uint64_t)(lc - 1);
  }
  if (upper >= 0xFFFFFFFFull) return fail(SCN_ERR_RESOURCE, "too many upper-layer lists");
  std::vector<uint32_t> adj_up((size_t)upper * su, ROW_NONE);
  std::vector<uint8_t> node_layer(n, 0);
  uint64_t li = 0, ei = 0, total_edges = 0;
  for (uint64_t i = 0; i < n_nodes; ++i) {
    uint32_t r = node_row[i];
    for (int l = 0; l < list_counts[i]; ++l) {
      uint32_t c = edge_counts[li++];
      if (c > 0) node_layer[r] = (uint8_t)l;  // highest non-empty layer wins (lists come in layer order)
      uint32_t cap_l = (l == 0) ? s0 : su;
      if (c > cap_l)
        return fail(SCN_ERR_INVALID_PARAMETERS, "node %llu layer %d has %u neighbours (max %u)",
                    (unsigned long long)node_ids[i], l, c, cap_l);
      uint32_t* dst = (l == 0) ? &adj0[(size_t)r * s0] : &adj_up[((size_t)up_off[r] + (l - 1)) * su];
      // A neighbour named twice is kept once, at its first position: searchLayer skips the second
      // occurrence as visited (hnsw.go:523-525; as deleted, 527-530, if the first one was), so the
      // walk is the same, and the kernels may rely on a list holding distinct rows.
      uint32_t kept = 0;
      for (uint32_t e = 0; e < c; ++e) {
        uint32_t nr;
        if (!s->lookup(edges[ei + e], &nr))
          return fail(SCN_ERR_INDEX_BUILD_FAILED, "neighbour %llu of node %llu is not in the store",
                      (unsigned long long)edges[ei + e], (unsigned long long)node_ids[i]);
        bool seen = false;
        for (uint32_t f = 0; f < kept && !seen; ++f) seen = dst[f] == nr;
        if (!seen) dst[kept++] = nr;
      }
      ei += c;
      total_edges += kept;
    }
  }
  uint32_t entry_row = ROW_NONE;
  if (entry_id != 0 && !s->lookup(entry_id, &entry_row))
    return fail(SCN_ERR_INDEX_BUILD_FAILED, "entrypoint %llu is not in the store", (unsigned long long)entry_id);
  free_graph(s);
  SCN_CUDA(cudaMalloc(&s->d_adj0, std::max<size_t>(adj0.size(), 1) * 4));
  SCN_CUDA(cudaMalloc(&s->d_levels, std::max<size_t>(levels.size(), 1)));
  SCN_CUDA(cudaMalloc(&s->d_up_off, std::max<size_t>(up_off.size(), 1) * 4));
  SCN_CUDA(cudaMalloc(&s->d_adj_up, std::max<size_t>(adj_up.size(), 1) * 4));
  SCN_CUDA(cudaMemcpy(s->d_adj0, adj0.data(), adj0.size() * 4, cudaMemcpyHostToDevice));
  SCN_CUDA(cudaMemcpy(s->d_levels, levels.data(), levels.size(), cudaMemcpyHostToDevice));
  SCN_CUDA(cudaMemcpy(s->d_up_off, up_off.data(), up_off.size() * 4, cudaMemcpyHostToDevice));
  SCN_CUDA(cudaMemcpy(s->d_adj_up, adj_up.data(), adj_up.size() * 4, cudaMemcpyHostToDevice));
  s->has_graph = true;
  s->h_node_layer.swap(node_layer);
  s->m = m;
  s->max_layer = max_layer;
  s->entry_id = entry_id;
  s->entry_row = entry_row;
  s->graph_nodes = n;
  s->graph_edges = total_edges;
  s->upper_lists = upper;
  return SCN_OK;
}

int32_t scn_set_option(scn_store* s, const char* name, int64_t value) {
  if (!s || !name) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  std::string n(name);
  if (n == "flat_path") {
    if (value < 0 || value > 2) return fail(SCN_ERR_INVALID_PARAMETERS, "flat_path must be 0, 1 or 2");
    s->opt_flat_path = value;
  } else if (n == "tensor_min_batch") {
    s->opt_tensor_min_batch = value;
  } else if (n == "overfetch") {
    s->opt_overfetch = value;
  } else if (n == "hnsw_gather") {
    s->opt_hnsw_gather = value;
  } else if (n == "hnsw_gather_long") {
    s->opt_hnsw_gather_long = value;
  } else if (n == "hnsw_global") {
    s->opt_hnsw_global = value;
  } else if (n == "hnsw_per_sm") {
    s->opt_hnsw_per_sm = value;
  } else if (n == "hnsw_early") {
    s->opt_hnsw_early = value;
  } else if (n == "hnsw_exact_ties") {
    s->opt_hnsw_exact_ties = value;
  } else if (n == "hnsw_hash") {
    s->opt_hnsw_hash = value;
  } else if (n == "tensor_hint_target") {
    s->opt_tensor_hint_target = value;
  } else if (n == "tensor_hint") {
    s->opt_tensor_hint = value;
  } else if (n == "tensor_bn") {
    s->opt_tensor_bn = value;
  } else if (n == "tensor_pair") {
    s->opt_tensor_pair = value;
  } else if (n == "tensor_fused") {
    s->opt_tensor_fused = value;
  } else if (n == "tensor_pair_ew") {
    s->opt_tensor_pair_ew = value;
  } else if (n == "tensor_share") {
    s->opt_tensor_share = value;
  } else if (n == "pdl") {
    s->opt_pdl = value;
  } else if (n == "tensor_chunks") {
    s->opt_tensor_chunks = value;
  } else if (n == "build_window") {
    s->opt_build_window = value;
  } else if (n == "auto_id_base") {
    // a row shard of a larger collection: auto-assigned ids are value + row + 1 (global row + 1)
    if (s->rows != 0 || value < 0) return fail(SCN_ERR_INVALID_PARAMETERS, "auto_id_base can only be set on an empty store");
    s->auto_base = (uint64_t)value;
  } else if (n == "profile") {
    s->opt_profile = value;
  } else {
    return fail(SCN_ERR_INVALID_PARAMETERS, "unknown option '%s'", name);
  }
  return SCN_OK;
}

int32_t scn_last_timings(scn_store* s, const char** names, float* ms, uint32_t* counts, int32_t max_entries) {
  if (!s) return 0;
  DeviceGuard g(s->device);
  std::lock_guard<std::mutex> lk(s->mu);
  static thread_local std::vector<std::string> name_store;
  name_store.clear();
  std::vector<float> total;
  std::vector<uint32_t> cnt;
  for (auto& e : s->pending) {
    cudaEventSynchronize(e.b);
    float t = 0;
    cudaEventElapsedTime(&t, e.a, e.b);
    cudaEventDestroy(e.a);
    cudaEventDestroy(e.b);
    size_t i = 0;
    for (; i < name_store.size(); ++i)
      if (name_store[i] == e.name) break;
    if (i == name_store.size()) {
      name_store.push_back(e.name);
      total.push_back(0.f);
      cnt.push_back(0);
    }
    total[i] += t;
    cnt[i] += 1;
  }
  s->pending.clear();
  int32_t n = (int32_t)std::min<size_t>(name_store.size(), (size_t)std::max(0, max_entries));
  for (int32_t i = 0; i < n; ++i) {
    if (names) names[i] = name_store[i].c_str();
    if (ms) ms[i] = total[i];
    if (counts) counts[i] = cnt[i];
  }
  return n;
}

int32_t scn_last_counters(scn_store* s, uint64_t* out, int32_t n) {
  if (!s || !out) return 0;
  DeviceGuard g(s->device);
  unsigned long long h[4] = {0, 0, 0, 0};
  if (cudaMemcpy(h, s->d_counters, sizeof h, cudaMemcpyDeviceToHost) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  int32_t m = std::min(n, 4);
  for (int32_t i = 0; i < m; ++i) out[i] = h[i];
  return m;
}

}  // extern "C"
