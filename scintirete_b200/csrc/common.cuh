// common.cuh — shared host/device helpers for libscn_gpu (B200 / sm_100a).
#pragma once

#include <cuda_runtime.h>
#include <stdint.h>

#include <atomic>
#include <string>

#include "../../include/scn_gpu.h"

namespace scn {

// ---- error plumbing (status codes = utils.ErrorCode, internal/utils/errors.go:11-49) --------
int32_t fail(int32_t code, const char* fmt, ...);
int32_t cuda_fail(cudaError_t e, const char* what, const char* file, int line);
void count_launch(uint64_t n = 1);

#define SCN_CUDA(expr)                                                        \
  do {                                                                        \
    cudaError_t e__ = (expr);                                                 \
    if (e__ != cudaSuccess) return scn::cuda_fail(e__, #expr, __FILE__, __LINE__); \
  } while (0)

#define SCN_TRY(expr)            \
  do {                           \
    int32_t rc__ = (expr);       \
    if (rc__ != SCN_OK) return rc__; \
  } while (0)

// Opt-in dynamic shared memory. Search entry points run concurrently from many host threads, so the
// per-function limit must never be lowered between another thread's set and its launch: every call
// sets the SAME value, the device's opt-in maximum (idempotent, hence race-free).
int max_optin_smem();
int static_smem_of(const void* kernel);
#define SCN_ALLOW_SMEM(kernel, bytes)                                                                     \
  do {                                                                                                    \
    static const int static__ = scn::static_smem_of((const void*)(kernel)); /* statically allocated part */ \
    const int limit__ = scn::max_optin_smem() - static__;                                                 \
    if ((size_t)(bytes) > (size_t)limit__)                                                                \
      return scn::fail(SCN_ERR_INVALID_PARAMETERS, "%zu bytes of shared memory exceed the device limit",   \
                       (size_t)(bytes));                                                                  \
    /* once per kernel and device: the value never changes, so racing first calls are harmless */         \
    static std::atomic<uint64_t> done__{0};                                                               \
    int dev__ = 0;                                                                                        \
    cudaGetDevice(&dev__);                                                                                \
    const uint64_t bit__ = 1ull << (dev__ & 63);                                                          \
    if (!(done__.load(std::memory_order_acquire) & bit__)) {                                              \
      SCN_CUDA(cudaFuncSetAttribute(kernel, cudaFuncAttributeMaxDynamicSharedMemorySize, limit__));       \
      done__.fetch_or(bit__, std::memory_order_release);                                                  \
    }                                                                                                     \
  } while (0)

// ---- programmatic dependent launch (PDL) -------------------------------------------------------
// The kernels behind the filter of a small batch run for a few microseconds each, less than the gap
// between two dependent launches. A kernel launched with launch_chained(..., pdl = true) may become
// resident as soon as every block of the kernel before it has executed pdl_trigger() (or exited); it
// must execute pdl_wait() before it touches anything an earlier kernel of the stream wrote — the wait
// returns when the preceding grid has completed and its writes are visible. Every thread of such a
// kernel waits at its top (a grid whose threads all left without waiting would complete early and
// release ITS dependents before the grids further up the chain are done). Both instructions are
// no-ops in a kernel that was launched the ordinary way.
#ifdef __CUDACC__
__device__ __forceinline__ void pdl_wait() { asm volatile("griddepcontrol.wait;" ::: "memory"); }
__device__ __forceinline__ void pdl_trigger() { asm volatile("griddepcontrol.launch_dependents;" ::: "memory"); }

template <class... KArgs, class... Args>
inline cudaError_t launch_chained(void (*kernel)(KArgs...), dim3 grid, dim3 block, size_t smem, cudaStream_t stream, bool pdl,
                                  Args&&... args) {
  cudaLaunchConfig_t cfg{};
  cfg.gridDim = grid;
  cfg.blockDim = block;
  cfg.dynamicSmemBytes = smem;
  cfg.stream = stream;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
  attr[0].val.programmaticStreamSerializationAllowed = 1;
  cfg.attrs = attr;
  cfg.numAttrs = pdl ? 1 : 0;
  return cudaLaunchKernelEx(&cfg, kernel, static_cast<KArgs>(args)...);
}
#endif

// after a kernel launch
#define SCN_LAUNCHED()                                                         \
  do {                                                                         \
    scn::count_launch();                                                       \
    cudaError_t e__ = cudaGetLastError();                                      \
    if (e__ != cudaSuccess) return scn::cuda_fail(e__, "kernel launch", __FILE__, __LINE__); \
  } while (0)

// ---- order-preserving distance keys ---------------------------------------------------------
// key = ord(dist) << 32 | row. Ascending u64 order == (distance ascending, row ascending), the
// flat oracle's stable order. -0.0 is folded into +0.0 (they compare equal in the reference's
// float comparisons); any NaN sorts after +Inf (a NaN is never "<" anything in the reference).
__host__ __device__ __forceinline__ uint32_t f32_ord(float d) {
  if (d != d) return 0xFFFFFFFFu;
  d = d + 0.0f;
#ifdef __CUDA_ARCH__
  uint32_t u = __float_as_uint(d);
#else
  union { float f; uint32_t u; } c;
  c.f = d;
  uint32_t u = c.u;
#endif
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

__host__ __device__ __forceinline__ float ord_f32(uint32_t o) {
  uint32_t u = (o & 0x80000000u) ? (o & 0x7FFFFFFFu) : ~o;
#ifdef __CUDA_ARCH__
  return __uint_as_float(u);
#else
  union { float f; uint32_t u; } c;
  c.u = u;
  return c.f;
#endif
}

__host__ __device__ __forceinline__ uint64_t make_key(float d, uint32_t row) {
  return ((uint64_t)f32_ord(d) << 32) | row;
}

constexpr uint64_t KEY_NONE = ~0ull;
constexpr uint32_t ROW_NONE = 0xFFFFFFFFu;

// ---- the reference's arithmetic, bit for bit -------------------------------------------------
// Go on amd64 rounds every float32 multiply and add separately (no FMA) and accumulates in
// source order (distance.go:26-30, 58-63, 109-112). __fmul_rn/__fadd_rn/__fsub_rn are never
// contracted by nvcc, so these reproduce the reference's bits exactly.
enum : int { M_L2 = SCN_METRIC_L2, M_COS = SCN_METRIC_COSINE, M_IP = SCN_METRIC_INNER_PRODUCT };

template <int METRIC>
__device__ __forceinline__ float acc_step(float acc, float q, float x) {
  if (METRIC == M_L2) {
    float diff = __fsub_rn(q, x);
    return __fadd_rn(acc, __fmul_rn(diff, diff));
  } else {
    return __fadd_rn(acc, __fmul_rn(q, x));
  }
}

// Four steps at once: the element-wise part (difference, product) on the packed fp32x2 pipe
// (FADD2 / FMUL2: two IEEE round-to-nearest results per instruction, bit-identical to the scalar
// ops), the four additions into the accumulator one after the other in source order.
__device__ __forceinline__ unsigned long long pack_f32x2(float lo, float hi) {
  unsigned long long r;
  asm("mov.b64 %0, {%1, %2};" : "=l"(r) : "f"(lo), "f"(hi));
  return r;
}
template <int METRIC>
__device__ __forceinline__ float acc_step4(float acc, const float4& q, const float4& x) {
  unsigned long long p01, p23;
  const unsigned long long q01 = pack_f32x2(q.x, q.y), q23 = pack_f32x2(q.z, q.w);
  const unsigned long long x01 = pack_f32x2(x.x, x.y), x23 = pack_f32x2(x.z, x.w);
  if (METRIC == M_L2) {
    unsigned long long d01, d23;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d01) : "l"(q01), "l"(x01));
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d23) : "l"(q23), "l"(x23));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(p01) : "l"(d01));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(p23) : "l"(d23));
  } else {
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p01) : "l"(q01), "l"(x01));
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(p23) : "l"(q23), "l"(x23));
  }
  float t0, t1, t2, t3;
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t0), "=f"(t1) : "l"(p01));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(t2), "=f"(t3) : "l"(p23));
  acc = __fadd_rn(acc, t0);
  acc = __fadd_rn(acc, t1);
  acc = __fadd_rn(acc, t2);
  return __fadd_rn(acc, t3);
}

// acc = the sequential sum; qnorm / xnorm = sqrt of the sequential sums of squares (cosine only).
template <int METRIC>
__device__ __forceinline__ float finish_distance(float acc, float qnorm, float xnorm) {
  if (METRIC == M_L2) return __fsqrt_rn(acc);                  // distance.go:31
  if (METRIC == M_IP) return -acc;                             // distance.go:115
  if (qnorm == 0.0f || xnorm == 0.0f) return 1.0f;             // distance.go:68-70
  float cs = __fdiv_rn(acc, __fmul_rn(qnorm, xnorm));          // distance.go:72
  if (cs > 1.0f) cs = 1.0f;
  else if (cs < -1.0f) cs = -1.0f;                              // distance.go:74-78
  return __fsub_rn(1.0f, cs);                                  // distance.go:81
}

// One thread, one (query,row) pair, reference order. q and x must be 16-byte aligned with
// `pitch4` float4s readable (zero padded past dim: adding +0.0 terms never changes the sum).
template <int METRIC>
__device__ __forceinline__ float exact_acc_thread(const float* __restrict__ q, const float* __restrict__ x,
                                                  uint32_t pitch4) {
  const float4* q4 = reinterpret_cast<const float4*>(q);
  const float4* x4 = reinterpret_cast<const float4*>(x);
  float acc = 0.0f;
#pragma unroll 4
  for (uint32_t i = 0; i < pitch4; ++i) {
    float4 a = q4[i];
    float4 b = __ldg(x4 + i);
    acc = acc_step4<METRIC>(acc, a, b);
  }
  return acc;
}

// The reference's Distance(query, row), bit for bit, for gather-style callers (HNSW traversal,
// rerank): one thread walks one row in the reference's sequential fp32 order; q is in shared memory. The row is fetched in chunks of 16 float4 that are double
// buffered in registers, so 16-32 independent 128-bit loads per lane (up to 16 KB per warp) are in
// flight before the first dependent add: a 512-byte row costs one memory latency instead of the
// eight that a 4-wide unrolled loop serialises.
// DCH = float4 per register chunk: 16 (two chunks = 128 data registers, one memory latency per
// 512-byte row; HNSW traversal) or 8 (half the registers, more resident warps; rerank).

// 256-bit read-only load (rows are 32-byte aligned: pitch is a multiple of 8 floats). One lane
// reads one row, so a warp-wide load touches up to 32 different lines and L1 spends a tag cycle on
// each: the wider load halves the requests and touches every 32-byte sector once instead of twice.
// (ptxas 12.9 crashes on LDG.256 inside a __noinline__ function, hence the single inlined call site.)
struct __align__(32) F8 {
  float4 a, b;
};
__device__ __forceinline__ void ldg256(const float4* p, float4& a, float4& b) {
  const F8 v = *reinterpret_cast<const F8*>(p);
  a = v.a;
  b = v.b;
}

// FULL: the chunk lies entirely inside the row -> straight-line code without predicates, so the
// scheduler can hoist the query LDS ahead of the dependent add chain.
template <bool FULL, int DCH>
__device__ __forceinline__ void load_chunk(float4 (&b)[DCH], const float4* __restrict__ x4, uint32_t c, uint32_t pitch4) {
#pragma unroll
  for (int i = 0; i < DCH; i += 2) {
    const uint32_t j = c * DCH + i;
    if (FULL || j < pitch4) ldg256(x4 + j, b[i], b[i + 1]);  // pitch4 is even
    else b[i] = b[i + 1] = make_float4(0.f, 0.f, 0.f, 0.f);  // +0 terms never change the sum
  }
}

template <int METRIC, bool FULL, int DCH>
__device__ __forceinline__ float acc_chunk(float acc, const float4 (&b)[DCH], const float4* __restrict__ q4, uint32_t c, uint32_t pitch4) {
  float4 qa[DCH];
#pragma unroll
  for (int i = 0; i < DCH; ++i) {
    const uint32_t j = c * DCH + i;
    qa[i] = (FULL || j < pitch4) ? q4[j] : make_float4(0.f, 0.f, 0.f, 0.f);
  }
#pragma unroll
  for (int i = 0; i < DCH; ++i) {
    acc = acc_step4<METRIC>(acc, qa[i], b[i]);
  }
  return acc;
}

template <int METRIC, int DCH = 16>
__device__ __forceinline__ float row_distance(const float* __restrict__ vec, const float* __restrict__ norm, uint32_t pitch,
                                              const float* sq, float qn, uint32_t row) {
  const float4* x4 = reinterpret_cast<const float4*>(vec + (size_t)row * pitch);
  const float4* q4 = reinterpret_cast<const float4*>(sq);
  const uint32_t pitch4 = pitch >> 2;
  const uint32_t nfull = pitch4 / DCH;          // chunks that lie entirely inside the row
  float4 b0[DCH], b1[DCH];
  float xn = 0.0f;
  float acc = 0.0f;
  if (nfull > 0) load_chunk<true, DCH>(b0, x4, 0, pitch4);
  if (nfull > 1) load_chunk<true, DCH>(b1, x4, 1, pitch4);
  if (METRIC == M_COS) xn = __ldg(norm + row);
  uint32_t c = 0;
  for (; c + 1 < nfull; c += 2) {
    acc = acc_chunk<METRIC, true, DCH>(acc, b0, q4, c, pitch4);
    if (c + 2 < nfull) load_chunk<true, DCH>(b0, x4, c + 2, pitch4);
    acc = acc_chunk<METRIC, true, DCH>(acc, b1, q4, c + 1, pitch4);
    if (c + 3 < nfull) load_chunk<true, DCH>(b1, x4, c + 3, pitch4);
  }
  if (c < nfull) {  // odd number of full chunks: the last one sits in b0
    acc = acc_chunk<METRIC, true, DCH>(acc, b0, q4, c, pitch4);
    ++c;
  }
  if (c * DCH < pitch4) {  // ragged tail (< 16 float4)
    load_chunk<false, DCH>(b1, x4, c, pitch4);
    acc = acc_chunk<METRIC, false, DCH>(acc, b1, q4, c, pitch4);
  }
  return finish_distance<METRIC>(acc, qn, xn);
}

// sequential sum of squares -> sqrt, i.e. VectorMagnitude (distance.go:175-181) and the normA /
// normB terms of CosineDistance (distance.go:58-66)
__device__ __forceinline__ float exact_norm_thread(const float* __restrict__ v, uint32_t n) {
  float s = 0.0f;
  for (uint32_t i = 0; i < n; ++i) s = __fadd_rn(s, __fmul_rn(v[i], v[i]));
  return __fsqrt_rn(s);
}

// Same value from a zero-padded, 16-byte-aligned copy (n4 float4s): the +0 terms of the padding do
// not change the sum, 128-bit loads run ahead of the dependent add chain.
__device__ __forceinline__ float exact_norm_padded(const float* __restrict__ v, uint32_t n4) {
  const float4* v4 = reinterpret_cast<const float4*>(v);
  float s = 0.0f;
#pragma unroll 8
  for (uint32_t i = 0; i < n4; ++i) {
    const float4 a = v4[i];
    s = __fadd_rn(s, __fmul_rn(a.x, a.x));
    s = __fadd_rn(s, __fmul_rn(a.y, a.y));
    s = __fadd_rn(s, __fmul_rn(a.z, a.z));
    s = __fadd_rn(s, __fmul_rn(a.w, a.w));
  }
  return __fsqrt_rn(s);
}

// Block-cooperative copy of a query row into a zero-padded shared-memory buffer: eight independent
// loads per thread are in flight before the first store (a plain load->store loop serialises one
// memory latency per element and thread).
__device__ __forceinline__ void stage_query(float* __restrict__ dst, const float* __restrict__ src, uint32_t dim, uint32_t pitch,
                                            uint32_t tid, uint32_t nthreads) {
  for (uint32_t i0 = tid; i0 < pitch; i0 += 8 * nthreads) {
    float t[8];
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t i = i0 + u * nthreads;
      t[u] = (i < dim) ? __ldg(src + i) : 0.0f;
    }
#pragma unroll
    for (int u = 0; u < 8; ++u) {
      const uint32_t i = i0 + u * nthreads;
      if (i < pitch) dst[i] = t[u];
    }
  }
}

__device__ __forceinline__ bool bit_test(const uint32_t* __restrict__ bits, uint32_t i) {
  return (__ldg(bits + (i >> 5)) >> (i & 31)) & 1u;
}

// ---- cp.async helpers -------------------------------------------------------------------------
__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gmem_src, bool valid) {
  uint32_t s = (uint32_t)__cvta_generic_to_shared(smem_dst);
  int sz = valid ? 16 : 0;  // src-size 0 -> zero fill
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16, %2;\n" ::"r"(s), "l"(gmem_src), "r"(sz));
}
__device__ __forceinline__ void cp_async_commit() { asm volatile("cp.async.commit_group;\n" ::); }
template <int N>
__device__ __forceinline__ void cp_async_wait() {
  asm volatile("cp.async.wait_group %0;\n" ::"n"(N));
}

__device__ __forceinline__ uint32_t cvta_smem(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

// ---- warp gather: the reference's Distance(query, row) for up to 32 rows, one per lane ------------
// The rows of a batch are copied into shared memory with warp-wide cp.async: one instruction moves
// one whole 128-byte line of each of four rows (16 bytes per lane), so DRAM sees whole lines and L1
// spends four tag cycles per instruction instead of 32. Each lane then walks its own row (stage rows GA_ROW bytes apart:
// conflict-free LDS.128) in the reference's sequential fp32 order.
// History (1 M x 128, ef = 128, 10 k queries): lane-per-row LDG.256 into registers 9.1 ms (32 lines
// per instruction: L1-tag-bound, 253 registers); one 1-D bulk copy (UBLKCP) per lane and row 5.9 ms —
// exactly the copy engine's request rate (one request per ~46 cycles and SM: 39.9 M requests / 148
// SMs); warp-wide cp.async: see profiles/.
// CH = bytes of a row per stage (512, or 256: two rows per instruction), NBUF = stage buffers:
//   <512, 2>  rows longer than 512 bytes: two stages in flight            (32 x 528 x 2 bytes)
//   <512, 1>  rows of <= 512 bytes, one stage                             (32 x 528 bytes)
//   <256, 1>  half the shared memory, stages strictly one after the other (32 x 272 bytes): the
//             smallest footprint — for walks whose parallelism is bounded by shared memory
__host__ __device__ constexpr uint32_t ga_row(uint32_t ch) { return ch + 16; }  // stage rows: conflict-free LDS.128
__host__ __device__ constexpr uint32_t ga_stage_bytes(uint32_t ch, uint32_t nbuf) { return nbuf * 32 * ga_row(ch); }

// stage `ch` of the rows of all lanes in `mask` -> buffer ch % NBUF, one commit group.
// A quarter-warp copies one 128-byte line of one row per instruction (four rows per instruction,
// the lines of a row by consecutive instructions with immediate offsets), so one shuffle and one
// 64-bit address feed CH/128 copies.
template <uint32_t CH, uint32_t NBUF>
__device__ __forceinline__ void gather_issue(const float* __restrict__ vec, uint32_t pitch, uint32_t row, uint32_t mask, uint32_t ch,
                                             unsigned char* stage, uint32_t lane) {
  const uint32_t row_bytes = pitch * 4;  // multiple of 32
  const uint32_t bytes = min(CH, row_bytes - ch * CH);
  const uint32_t sub = lane >> 3, piece = lane & 7;
  const uint32_t dst0 = cvta_smem(stage + ((ch % NBUF) * 32 + sub) * ga_row(CH) + piece * 16);
  const unsigned char* src0 = reinterpret_cast<const unsigned char*>(vec) + (size_t)ch * CH + piece * 16;
  // the row index doubles as the predicate: lanes outside `mask` hand out 0xFFFFFFFF
  const uint32_t rowv = ((mask >> lane) & 1u) ? row : 0xFFFFFFFFu;
  if (bytes == CH) {  // full stage: straight-line, predicated copies (no branch per row)
#pragma unroll
    for (uint32_t i = 0; i < 32; i += 4) {
      const uint32_t r = __shfl_sync(0xffffffffu, rowv, i + sub);
      const unsigned char* src = src0 + (size_t)r * row_bytes;  // not dereferenced when r is the marker
      if (CH == 256)
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0xffffffff;\n\t"
                     "@p cp.async.cg.shared.global [%0], [%1], 16;\n\t"
                     "@p cp.async.cg.shared.global [%0+128], [%1+128], 16;\n\t}" ::"r"(dst0 + i * ga_row(CH)), "l"(src), "r"(r) : "memory");
      else
        asm volatile("{\n\t.reg .pred p;\n\tsetp.ne.u32 p, %2, 0xffffffff;\n\t"
                     "@p cp.async.cg.shared.global [%0], [%1], 16;\n\t"
                     "@p cp.async.cg.shared.global [%0+128], [%1+128], 16;\n\t"
                     "@p cp.async.cg.shared.global [%0+256], [%1+256], 16;\n\t"
                     "@p cp.async.cg.shared.global [%0+384], [%1+384], 16;\n\t}" ::"r"(dst0 + i * ga_row(CH)), "l"(src), "r"(r) : "memory");
    }
  } else {  // ragged last stage
#pragma unroll
    for (uint32_t i = 0; i < 32; i += 4) {
      const uint32_t r = __shfl_sync(0xffffffffu, rowv, i + sub);
      if (r != 0xFFFFFFFFu) {
        const unsigned char* src = src0 + (size_t)r * row_bytes;
#pragma unroll
        for (uint32_t j = 0; j < CH / 128; ++j) {
          if (piece * 16 + j * 128 < bytes)
            asm volatile("cp.async.cg.shared.global [%0], [%1], 16;\n" ::"r"(dst0 + i * ga_row(CH) + j * 128), "l"(src + j * 128) : "memory");
        }
      }
    }
  }
  cp_async_commit();
}
// the first stage(s) of a gather (what fits the buffers); gather_finish issues the rest
template <uint32_t CH, uint32_t NBUF>
__device__ __forceinline__ void gather_begin(const float* __restrict__ vec, uint32_t pitch, uint32_t row, uint32_t mask,
                                             unsigned char* stage, uint32_t lane) {
  gather_issue<CH, NBUF>(vec, pitch, row, mask, 0, stage, lane);
  if (NBUF > 1 && pitch * 4 > CH) gather_issue<CH, NBUF>(vec, pitch, row, mask, 1, stage, lane);
}
// Completes a gather begun for a superset of the lanes in `mask` (lanes may have been dropped
// since, e.g. rows found in the visited set while their copies were in flight: their data is
// ignored, later stages are requested for the lanes of `mask` only). Returns +Inf on other lanes.
template <int METRIC, uint32_t CH, uint32_t NBUF>
__device__ __forceinline__ float gather_finish(const float* __restrict__ vec, const float* __restrict__ norm, uint32_t pitch,
                                               const float* sq, float qn, uint32_t row, uint32_t mask, unsigned char* stage,
                                               uint32_t lane) {
  const uint32_t row_bytes = pitch * 4;
  const uint32_t n_chunks = (row_bytes + CH - 1) / CH;
  const bool valid = (mask >> lane) & 1u;
  if (!mask) {  // nobody left: drain what was begun
    cp_async_wait<0>();
    __syncwarp();
    return __int_as_float(0x7f800000);
  }
  float acc = 0.0f;
  const float xn = (METRIC == M_COS && valid) ? __ldg(norm + row) : 0.0f;
  for (uint32_t ch = 0; ch < n_chunks; ++ch) {
    if (NBUF > 1 && ch + 1 < n_chunks) cp_async_wait<1>();  // stage ch + 1 may still be in flight
    else cp_async_wait<0>();
    __syncwarp();  // every lane's pieces of this stage have landed
    const uint32_t n4 = min(CH, row_bytes - ch * CH) / 16;
    if (valid) {
      const float4* x4 = reinterpret_cast<const float4*>(stage + ((ch % NBUF) * 32 + lane) * ga_row(CH));
      const float4* q4 = reinterpret_cast<const float4*>(sq) + ch * (CH / 16);
      if (n4 == CH / 16) {
#pragma unroll
        for (uint32_t i = 0; i < CH / 16; ++i) {
          const float4 xa = x4[i], qa = q4[i];
          acc = acc_step4<METRIC>(acc, qa, xa);
        }
      } else {
        for (uint32_t i = 0; i < n4; ++i) {
          const float4 xa = x4[i], qa = q4[i];
          acc = acc_step4<METRIC>(acc, qa, xa);
        }
      }
    }
    __syncwarp();  // this buffer is free again
    if (ch + NBUF < n_chunks) gather_issue<CH, NBUF>(vec, pitch, row, mask, ch + NBUF, stage, lane);
  }
  return valid ? finish_distance<METRIC>(acc, qn, xn) : __int_as_float(0x7f800000);
}

template <int METRIC, uint32_t CH, uint32_t NBUF>
__device__ __forceinline__ float gather_distance(const float* __restrict__ vec, const float* __restrict__ norm, uint32_t pitch,
                                                 const float* sq, float qn, uint32_t row, uint32_t mask, unsigned char* stage,
                                                 uint32_t lane) {
  gather_begin<CH, NBUF>(vec, pitch, row, mask, stage, lane);
  return gather_finish<METRIC, CH, NBUF>(vec, norm, pitch, sq, qn, row, mask, stage, lane);
}

inline uint32_t round_up(uint32_t v, uint32_t m) { return (v + m - 1) / m * m; }
inline uint32_t next_pow2(uint32_t v) {
  uint32_t p = 1;
  while (p < v) p <<= 1;
  return p;
}

}  // namespace scn
