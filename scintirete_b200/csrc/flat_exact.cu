// flat_exact.cu — exact (bit-faithful) flat scan, exact rerank, batched distances, top-k merges.
//
// K1e  flat_exact_scan_kernel : HBM-streaming scan of the fp32 rows. Each thread owns one row of
//      a 256-row tile and accumulates, in the reference's sequential order and rounding
//      (distance.go:26-30, 58-63, 109-112), the distance to up to QB queries at once, two queries per
//      packed fp32x2 instruction. Tiles are staged through shared memory with 16-byte cp.async
//      (coalesced 128-byte row segments, three stages in flight), read back conflict-free as float4
//      (row pitch 36 floats). Roofline: HBM. Algorithmic bytes per pass = N*pitch*4 (+ N*4 norms for
//      cosine). Measured at 1 M x 768, cosine: 0.62 ms for one query (0.76 of HBM), 0.92 ms for eight
//      (0.51; round 1: 1.05-1.17 ms). Two rows per thread on 512-row tiles with 64-byte stage segments —
//      half the shared-memory reads of query values per row — measured 0.97 ms and was dropped.
// K4   block top-k: per-query sorted key lists in shared memory, threshold-gated queues.
// K3   rerank_kernel : exact distances for gathered candidate rows + block bitonic top-k.
// K7   merge_topk_kernel : per-query merge of G sorted shard lists.
#include "store.h"

namespace scn {

constexpr int SCAN_THREADS = 256;  // rows per tile
constexpr int SCAN_KC = 32;        // floats of each row per stage
constexpr int SCAN_PITCH = 36;     // smem row pitch in floats (16B aligned, conflict-free LDS.128)
constexpr int SCAN_STAGES = 3;

// ---- warp-cooperative sorted-list insertion ---------------------------------------------------
// list[0..k) ascending, KEY_NONE padded. Inserts x if x < list[k-1]. All 32 lanes participate.
__device__ __forceinline__ void warp_list_insert(uint64_t* list, uint32_t k, uint64_t x, int lane) {
  if (x >= list[k - 1]) return;
  // position = number of entries < x
  uint32_t pos = 0;
  for (uint32_t base = 0; base < k; base += 32) {
    uint32_t i = base + lane;
    bool lt = (i < k) && (list[i] < x);
    pos += __popc(__ballot_sync(0xffffffffu, lt));
  }
  // shift [pos, k-2] right by one, highest chunk first
  for (int base = (int)((k - 1) / 32) * 32; base >= 0; base -= 32) {
    uint32_t i = base + lane;  // destination index
    uint64_t v = 0;
    bool mv = (i < k) && (i > pos);
    if (mv) v = list[i - 1];
    __syncwarp();
    if (mv) list[i] = v;
    __syncwarp();
  }
  if (lane == 0) list[pos] = x;
  __syncwarp();
}

// block-wide bitonic sort of n (power of two) u64 keys in shared memory
__device__ __forceinline__ void block_bitonic_sort(uint64_t* a, uint32_t n) {
  for (uint32_t size = 2; size <= n; size <<= 1) {
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      __syncthreads();
      for (uint32_t t = threadIdx.x; t < n / 2; t += blockDim.x) {
        uint32_t lo = 2 * t - (t & (stride - 1));
        uint32_t hi = lo + stride;
        bool up = ((lo & size) == 0);
        uint64_t x = a[lo], y = a[hi];
        if ((x > y) == up) {
          a[lo] = y;
          a[hi] = x;
        }
      }
    }
  }
  __syncthreads();
}

// ---- K1e ------------------------------------------------------------------------------------
struct ScanParams {
  const float* vec;
  const float* norm;
  const uint32_t* deleted;
  uint32_t pitch;     // floats
  uint32_t n_rows;
  uint32_t dim_pad;   // dim rounded up to SCAN_KC
  const float* q;     // [.][pitch_q] device queries
  uint32_t q_pitch;   // floats between queries in q
  uint32_t q_dim;     // valid floats per query
  const uint32_t* qlist;   // optional: indices into q
  const uint32_t* nq_dev;  // optional: device-resident number of queries (<= nq)
  uint32_t nq;
  uint32_t k;
  uint32_t row_base;  // added to the row field of the emitted keys
  uint64_t* partial;  // [gridDim.x][nq][k]
  unsigned long long one2;  // (1.0f, 1.0f): see pair_step
};

// One accumulation step for a PAIR of queries on the packed fp32x2 pipe: acc2 += q2 * x2 (or (q2 - x2)^2), every half
// an IEEE round-to-nearest operation of its own, i.e. bit for bit the scalar __fsub_rn / __fmul_rn / __fadd_rn.
template <int METRIC>
__device__ __forceinline__ void pair_step(unsigned long long& acc2, unsigned long long q2, unsigned long long x2, unsigned long long one2) {
  // The sum is taken as prod * one + acc with `one` = (1.0f, 1.0f) handed in as a KERNEL PARAMETER: one rounding of the exact
  // value prod + acc, i.e. the add. Spelled add.rn.f32x2, ptxas 12.9 contracts the mul.rn.f32x2 in front of it into ONE FFMA2
  // — a single rounding of q * x + acc instead of the reference's two — with or without -fmad=false, and also when the one is a
  // literal or the product is written fma(a, b, -0). A multiplier it cannot see through keeps FMUL2 and FFMA2 apart
  // (cuobjdump: 128 + 128 per stage). tests/test_gpu_flat.py compares the bits with the oracle's.
  unsigned long long prod;
  if (METRIC == M_L2) {
    unsigned long long d;
    asm("sub.rn.f32x2 %0, %1, %2;" : "=l"(d) : "l"(q2), "l"(x2));
    asm("mul.rn.f32x2 %0, %1, %1;" : "=l"(prod) : "l"(d));
  } else {
    asm("mul.rn.f32x2 %0, %1, %2;" : "=l"(prod) : "l"(q2), "l"(x2));
  }
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(acc2) : "l"(prod), "l"(one2), "l"(acc2));
}

// The queries of a pass sit in shared memory element-major, s_q[e][QP] with QP = QB rounded up to 2: the values all
// queries need for element e are adjacent (one or two LDS.128 at an immediate offset), and queries (2p, 2p+1) form the
// pairs whose sums advance together: per element and pair one FMUL2 and one FADD2 (round 1 spent one FADD per element
// and QUERY plus half an FMUL2 — 38.7 % + 19.2 % of all issued instructions, and another 22 % on the addresses of
// query-major rows of run-time length; profiles/r02_ncu_flat_exact_scan_*).
template <int METRIC, int QB>
__global__ void __launch_bounds__(SCAN_THREADS, 1) flat_exact_scan_kernel(ScanParams p) {
  constexpr int NP = (QB + 1) / 2;  // query pairs
  constexpr int QP = 2 * NP;        // floats per element in s_q
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const uint32_t k = p.k;
  float* s_q = reinterpret_cast<float*>(smem_raw);                    // [dim_pad][QP]
  float* s_tile = s_q + (size_t)QP * p.dim_pad;                       // [STAGES][256][PITCH]
  uint64_t* s_queue = reinterpret_cast<uint64_t*>(s_tile + (size_t)SCAN_STAGES * SCAN_THREADS * SCAN_PITCH);  // [QB][256]
  uint64_t* s_list = s_queue + (size_t)QB * SCAN_THREADS;             // [QB][k]
  __shared__ uint64_t s_tau[QB];
  __shared__ uint32_t s_qcount[QB];
  __shared__ float s_qnorm[QB];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  pdl_trigger();
  pdl_wait();
  uint32_t nq_total = p.nq_dev ? min(*p.nq_dev, p.nq) : p.nq;
  const uint32_t n_tiles = (p.n_rows + SCAN_THREADS - 1) / SCAN_THREADS;
  const uint32_t n_chunks = p.dim_pad / SCAN_KC;
  // this thread's share of every stage copy: 16-byte segment `seg` of rows r0, r0 + 32, ... of the tile
  const uint32_t r0 = tid >> 3, seg = tid & 7;

  for (uint32_t q0 = 0; q0 < nq_total; q0 += QB) {
    const uint32_t nq_here = min((uint32_t)QB, nq_total - q0);
    __syncthreads();
    // queries -> smem (element-major, zero padded), lists -> KEY_NONE
    for (uint32_t i = tid; i < QP * p.dim_pad; i += SCAN_THREADS) {
      const uint32_t qi = i / p.dim_pad, e = i - qi * p.dim_pad;   // (consecutive threads read consecutive elements of a query)
      float v = 0.0f;
      if (qi < nq_here && e < p.q_dim) {
        uint32_t src = p.qlist ? p.qlist[q0 + qi] : (q0 + qi);
        v = p.q[(size_t)src * p.q_pitch + e];
      }
      s_q[(size_t)e * QP + qi] = v;
    }
    for (uint32_t i = tid; i < QB * k; i += SCAN_THREADS) s_list[i] = KEY_NONE;
    if (tid < QB) {
      s_tau[tid] = KEY_NONE;
      s_qcount[tid] = 0;
    }
    __syncthreads();
    if (METRIC == M_COS && tid < QB) {   // ||q|| as the reference sums it (distance.go:58-66)
      float sum = 0.0f;
      for (uint32_t e = 0; e < p.q_dim; ++e) {
        const float v = s_q[(size_t)e * QP + tid];
        sum = __fadd_rn(sum, __fmul_rn(v, v));
      }
      s_qnorm[tid] = __fsqrt_rn(sum);
    }
    __syncthreads();

    // flattened (tile, chunk) pipeline over this block's tiles
    const uint32_t my_tiles = (n_tiles > blockIdx.x) ? (n_tiles - blockIdx.x + gridDim.x - 1) / gridDim.x : 0;
    const uint32_t total = my_tiles * n_chunks;
    auto issue = [&](uint32_t it) {
      if (it < total) {
        const uint32_t tile = blockIdx.x + (it / n_chunks) * gridDim.x;
        const uint32_t chunk = it % n_chunks;
        float* dst = s_tile + (size_t)(it % SCAN_STAGES) * SCAN_THREADS * SCAN_PITCH + r0 * SCAN_PITCH + seg * 4;
        // 256 rows x 8 segments of 16 B; consecutive threads take consecutive segments of a row
        const uint32_t col = chunk * SCAN_KC + seg * 4;
        const uint32_t row0 = tile * SCAN_THREADS + r0;
        const bool col_ok = col < p.pitch;
        const float* src = p.vec + (size_t)row0 * p.pitch + col;
        const size_t step = (size_t)32 * p.pitch;
#pragma unroll
        for (int j = 0; j < SCAN_THREADS / 32; ++j) {
          const bool valid = col_ok && (row0 + j * 32 < p.n_rows);
          cp_async16(dst + j * 32 * SCAN_PITCH, valid ? src + j * step : p.vec, valid);
        }
      }
      cp_async_commit();
    };
    for (int s = 0; s < SCAN_STAGES - 1; ++s) issue(s);

    unsigned long long acc2[NP];
    for (uint32_t it = 0; it < total; ++it) {
      const uint32_t chunk = it % n_chunks;
      if (chunk == 0) {
#pragma unroll
        for (int pr = 0; pr < NP; ++pr) acc2[pr] = 0ull;   // (+0.0, +0.0)
      }
      cp_async_wait<SCAN_STAGES - 2>();
      __syncthreads();                 // stage `it` landed for everyone; stage it-1 is free
      issue(it + SCAN_STAGES - 1);
      const float4* x4 = reinterpret_cast<const float4*>(s_tile + (size_t)(it % SCAN_STAGES) * SCAN_THREADS * SCAN_PITCH +
                                                          tid * SCAN_PITCH);
      const float* qe = s_q + (size_t)chunk * SCAN_KC * QP;   // the QP query values of element e at qe[e * QP]
#pragma unroll
      for (int j = 0; j < SCAN_KC / 4; ++j) {
        const float4 x = x4[j];
        const float xs[4] = {x.x, x.y, x.z, x.w};
#pragma unroll
        for (int u = 0; u < 4; ++u) {
          const unsigned long long x2 = pack_f32x2(xs[u], xs[u]);
          const float* qv = qe + (j * 4 + u) * QP;
          if (NP == 1) {
            const float2 a = *reinterpret_cast<const float2*>(qv);
            pair_step<METRIC>(acc2[0], pack_f32x2(a.x, a.y), x2, p.one2);
          } else {
#pragma unroll
            for (int h = 0; h < NP / 2; ++h) {
              const float4 a = *reinterpret_cast<const float4*>(qv + 4 * h);
              pair_step<METRIC>(acc2[2 * h], pack_f32x2(a.x, a.y), x2, p.one2);
              pair_step<METRIC>(acc2[2 * h + 1], pack_f32x2(a.z, a.w), x2, p.one2);
            }
          }
        }
      }
      if (chunk == n_chunks - 1) {
        // tile finished: gate against the block threshold, queue, merge
        uint32_t tile = blockIdx.x + (it / n_chunks) * gridDim.x;
        uint32_t row = tile * SCAN_THREADS + tid;
        bool live = (row < p.n_rows) && !bit_test(p.deleted, row);
        float xn = (METRIC == M_COS && live) ? __ldg(p.norm + row) : 0.0f;
#pragma unroll
        for (int qi = 0; qi < QB; ++qi) {
          if (qi < (int)nq_here && live) {
            float lo, hi;
            asm("mov.b64 {%0, %1}, %2;" : "=f"(lo), "=f"(hi) : "l"(acc2[qi >> 1]));
            float d = finish_distance<METRIC>((qi & 1) ? hi : lo, s_qnorm[qi], xn);
            uint64_t key = make_key(d, row + p.row_base);
            if (key < s_tau[qi]) {
              uint32_t slot = atomicAdd(&s_qcount[qi], 1u);
              s_queue[(size_t)qi * SCAN_THREADS + slot] = key;
            }
          }
        }
        __syncthreads();
        for (uint32_t qi = warp; qi < nq_here; qi += SCAN_THREADS / 32) {
          uint32_t cnt = s_qcount[qi];
          if (cnt) {
            uint64_t* list = s_list + (size_t)qi * k;
            for (uint32_t c = 0; c < cnt; ++c) warp_list_insert(list, k, s_queue[(size_t)qi * SCAN_THREADS + c], lane);
            if (lane == 0) {
              s_tau[qi] = list[k - 1];
              s_qcount[qi] = 0;
            }
          }
        }
        // the next __syncthreads (top of the loop) orders these writes before the next gate
      }
    }
    cp_async_wait<0>();
    __syncthreads();
    for (uint32_t i = tid; i < nq_here * k; i += SCAN_THREADS) {
      uint32_t qi = i / k, j = i - qi * k;
      p.partial[((size_t)blockIdx.x * p.nq + (q0 + qi)) * k + j] = s_list[(size_t)qi * k + j];
    }
  }
}

// ---- partial-list merge: one block per query, P sorted lists of k -> k -------------------------
// buf holds `cap` keys (power of two >= max(2*kp, 2048)); each round appends cap-kp fresh keys
// behind the current best kp, sorts, and keeps the first kp.
__global__ void __launch_bounds__(256) merge_partials_kernel(const uint64_t* __restrict__ partial, uint32_t n_parts,
                                                             uint32_t nq_stride, const uint32_t* __restrict__ qlist,
                                                             const uint32_t* __restrict__ nq_dev, uint32_t k, uint32_t kp,
                                                             uint32_t cap, uint64_t* __restrict__ out_keys) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  uint64_t* buf = reinterpret_cast<uint64_t*>(smem_raw);  // [cap]
  uint32_t q = blockIdx.x;
  pdl_trigger();
  pdl_wait();
  if (nq_dev && q >= *nq_dev) return;
  for (uint32_t i = threadIdx.x; i < kp; i += blockDim.x) buf[i] = KEY_NONE;
  const uint64_t total = (uint64_t)n_parts * k;
  const uint32_t fresh = cap - kp;
  for (uint64_t base = 0; base < total; base += fresh) {
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < fresh; i += blockDim.x) {
      uint64_t g = base + i;
      uint64_t v = KEY_NONE;
      if (g < total) {
        uint32_t part = (uint32_t)(g / k), j = (uint32_t)(g % k);
        v = partial[((size_t)part * nq_stride + q) * k + j];
      }
      buf[kp + i] = v;
    }
    block_bitonic_sort(buf, cap);  // best kp stay in buf[0..kp)
  }
  __syncthreads();
  uint32_t dst = qlist ? qlist[q] : q;
  for (uint32_t i = threadIdx.x; i < k; i += blockDim.x) out_keys[(size_t)dst * k + i] = buf[i];
}

static size_t scan_smem_bytes(int qb, uint32_t dim_pad, uint32_t k) {
  return (size_t)((qb + 1) / 2 * 2) * dim_pad * 4 + (size_t)SCAN_STAGES * SCAN_THREADS * SCAN_PITCH * 4 + (size_t)qb * SCAN_THREADS * 8 +
         (size_t)qb * k * 8;
}

template <int METRIC, int QB>
static int32_t launch_scan(const ScanParams& p, int grid, size_t smem, cudaStream_t stream, bool pdl) {
  SCN_ALLOW_SMEM((flat_exact_scan_kernel<METRIC, QB>), smem);
  SCN_CUDA(launch_chained(flat_exact_scan_kernel<METRIC, QB>, dim3(grid), dim3(SCAN_THREADS), smem, stream, pdl, p));
  SCN_LAUNCHED();
  return SCN_OK;
}

template <int METRIC>
static int32_t launch_scan_qb(int qb, const ScanParams& p, int grid, size_t smem, cudaStream_t stream, bool pdl) {
  switch (qb) {
    case 1: return launch_scan<METRIC, 1>(p, grid, smem, stream, pdl);
    case 2: return launch_scan<METRIC, 2>(p, grid, smem, stream, pdl);
    case 4: return launch_scan<METRIC, 4>(p, grid, smem, stream, pdl);
    default: return launch_scan<METRIC, 8>(p, grid, smem, stream, pdl);
  }
}

// Exact scan of the whole shard for nq queries (or for the device-side list qlist[0..*nq_dev)).
// Writes sorted keys [nq][k] (ord(dist)<<32 | row_base+row), KEY_NONE padded.
int32_t flat_search_exact(scn_store* s, const float* d_q, const uint32_t* d_qlist, const uint32_t* d_nq_dev,
                          uint64_t nq, uint32_t k, uint64_t row_base, uint64_t* d_out_keys, cudaStream_t stream,
                          Profiler* prof) {
  if (nq == 0) return SCN_OK;
  if (s->rows == 0) {
    SCN_CUDA(cudaMemsetAsync(d_out_keys, 0xFF, nq * k * sizeof(uint64_t), stream));
    return SCN_OK;
  }
  int sms = 0;
  SCN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
  const uint32_t dim_pad = round_up(s->dim, SCAN_KC);
  // queries per pass: as many as shared memory allows, at most 8
  int qb = nq >= 8 ? 8 : (nq >= 4 ? 4 : (nq >= 2 ? 2 : 1));
  while (qb > 1 && scan_smem_bytes(qb, dim_pad, k) > 200 * 1024) qb >>= 1;
  size_t smem = scan_smem_bytes(qb, dim_pad, k);
  if (smem > 220 * 1024) return fail(SCN_ERR_INVALID_PARAMETERS, "k=%u / dim=%u exceed the exact scan's shared memory", k, s->dim);
  const uint32_t n_tiles = (uint32_t)((s->rows + SCAN_THREADS - 1) / SCAN_THREADS);
  int grid = (int)std::min<uint32_t>((uint32_t)sms, n_tiles);
  Scratch scratch(stream);
  uint64_t* d_partial = nullptr;
  SCN_TRY(scratch.alloc(&d_partial, (size_t)grid * nq * k));
  ScanParams p;
  p.vec = s->d_vec;
  p.norm = s->d_norm;
  p.deleted = s->d_deleted;
  p.pitch = s->pitch;
  p.n_rows = (uint32_t)s->rows;
  p.dim_pad = dim_pad;
  p.q = d_q;
  p.q_pitch = s->dim;
  p.q_dim = s->dim;
  p.qlist = d_qlist;
  p.nq_dev = d_nq_dev;
  p.nq = (uint32_t)nq;
  p.k = k;
  p.row_base = (uint32_t)row_base;
  p.partial = d_partial;
  p.one2 = 0x3f8000003f800000ull;
  // behind the tensor path this is the last resort and normally has nothing to do: chained launches (see
  // common.cuh) let its start-up overlap the kernels before it
  const bool pdl = s->opt_pdl != 0 && d_nq_dev != nullptr && !(prof && prof->on);
  if (prof) prof->begin("flat_exact_scan");
  int32_t rc;
  switch (s->metric) {
    case M_L2: rc = launch_scan_qb<M_L2>(qb, p, grid, smem, stream, pdl); break;
    case M_COS: rc = launch_scan_qb<M_COS>(qb, p, grid, smem, stream, pdl); break;
    default: rc = launch_scan_qb<M_IP>(qb, p, grid, smem, stream, pdl); break;
  }
  if (prof) prof->end();
  SCN_TRY(rc);
  const uint32_t kp = std::max(32u, next_pow2(k));
  if (prof) prof->begin("merge_partials");
  const uint32_t cap = std::max(2 * kp, 2048u);
  SCN_CUDA(launch_chained(merge_partials_kernel, dim3((unsigned)nq), dim3(256), cap * sizeof(uint64_t), stream, pdl, d_partial, (uint32_t)grid,
                          (uint32_t)nq, d_qlist, d_nq_dev, k, kp, cap, d_out_keys));
  SCN_LAUNCHED();
  if (prof) prof->end();
  return SCN_OK;
}

// ---- keys -> (ids, distances, counts) ---------------------------------------------------------
__global__ void keys_to_results_kernel(const uint64_t* __restrict__ keys, uint64_t n, uint32_t k, uint32_t row_base,
                                       const uint64_t* __restrict__ ids, uint64_t* __restrict__ out_ids,
                                       float* __restrict__ out_dist, uint32_t* __restrict__ out_counts) {
  uint64_t q = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  pdl_trigger();
  pdl_wait();
  if (q >= n) return;
  uint32_t cnt = 0;
  for (uint32_t j = 0; j < k; ++j) {
    uint64_t key = keys[q * k + j];
    if (key == KEY_NONE) {
      if (out_ids) out_ids[q * k + j] = 0;
      if (out_dist) out_dist[q * k + j] = __int_as_float(0x7f800000);
    } else {
      if (out_ids) out_ids[q * k + j] = ids[(uint32_t)key - row_base];
      if (out_dist) out_dist[q * k + j] = ord_f32((uint32_t)(key >> 32));
      ++cnt;
    }
  }
  if (out_counts) out_counts[q] = cnt;
}

// chained: the launch directly follows the kernels of flat_keys on the same stream (see launch_chained, common.cuh)
int32_t keys_to_results(scn_store* s, const uint64_t* d_keys, uint64_t n, uint64_t row_base, uint64_t* d_out_ids,
                        float* d_out_dist, uint32_t* d_out_counts, uint32_t k, cudaStream_t stream, bool chained) {
  if (n == 0) return SCN_OK;
  SCN_CUDA(launch_chained(keys_to_results_kernel, dim3((unsigned)((n + 127) / 128)), dim3(128), 0, stream, chained && s->opt_pdl != 0 && s->opt_profile == 0, d_keys,
                          n, k, (uint32_t)row_base, s->d_ids, d_out_ids, d_out_dist, d_out_counts));
  SCN_LAUNCHED();
  return SCN_OK;
}

// ---- K3: exact rerank of gathered candidate rows ------------------------------------------------
// One warp per query. Candidates are taken 32 at a time, one row per lane. Rows arrive in
// shared memory through warp-wide cp.async (one instruction = 512 contiguous bytes of one row, two
// stages in flight; common.cuh), then every lane walks its own row in the reference's sequential
// fp32 order (conflict-free LDS.128: rows are 528 bytes apart). The warp sorts the keys and emits
// the first k distinct ones.

__host__ __device__ inline size_t rerank_smem_bytes(uint32_t pitch, uint32_t ncand_pad) {
  return (size_t)pitch * 4 + (size_t)ncand_pad * 8 + ga_stage_bytes(512, 2);
}

template <int METRIC>
__global__ void __launch_bounds__(32) rerank_kernel(const float* __restrict__ vec, const float* __restrict__ norm,
                                                    const uint32_t* __restrict__ deleted, uint32_t pitch, uint32_t dim,
                                                    uint32_t n_rows, const float* __restrict__ q,
                                                    const uint32_t* __restrict__ cand, uint32_t ncand, uint32_t ncand_pad,
                                                    uint32_t k, uint32_t row_base, const uint32_t* __restrict__ qlist,
                                                    const uint32_t* __restrict__ nq_dev, uint32_t nq,
                                                    uint64_t* __restrict__ out_keys) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  float* s_q = reinterpret_cast<float*>(smem_raw);                                   // [pitch]
  uint64_t* s_keys = reinterpret_cast<uint64_t*>(s_q + pitch);                       // [ncand_pad]
  unsigned char* s_stage = reinterpret_cast<unsigned char*>(s_keys + ncand_pad);     // [2][32][528]
  const uint32_t lane = threadIdx.x;
  pdl_trigger();
  pdl_wait();
  const uint32_t n_slots = nq_dev ? min(*nq_dev, nq) : nq;
  for (uint32_t slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
    const uint32_t qi = qlist ? qlist[slot] : slot;
    __syncwarp();
    stage_query(s_q, q + (size_t)qi * dim, dim, pitch, lane, 32);
    __syncwarp();
    float qnorm = 0.0f;
    if (METRIC == M_COS) {
      if (lane == 0) qnorm = exact_norm_padded(s_q, pitch / 4);
      qnorm = __shfl_sync(0xffffffffu, qnorm, 0);
    }
    for (uint32_t c0 = 0; c0 < ncand_pad; c0 += 32) {
      const uint32_t c = c0 + lane;
      uint32_t row = ROW_NONE;
      if (c < ncand) {
        row = cand[(size_t)qi * ncand + c];
        if (row >= n_rows || bit_test(deleted, row)) row = ROW_NONE;
      }
      const bool valid = row != ROW_NONE;
      const uint32_t mask = __ballot_sync(0xffffffffu, valid);
      uint64_t key = KEY_NONE;
      if (mask) {
        const float d = gather_distance<METRIC, 512, 2>(vec, norm, pitch, s_q, qnorm, row, mask, s_stage, lane);
        if (valid) key = make_key(d, row + row_base);
      }
      s_keys[c] = key;
    }
    __syncwarp();
    block_bitonic_sort(s_keys, ncand_pad);
    // emit the first k distinct keys (a candidate list may name a row twice: identical keys are adjacent)
    for (uint32_t i = lane; i < k; i += 32) out_keys[(size_t)qi * k + i] = KEY_NONE;
    __syncwarp();
    if (lane == 0) {
      uint32_t w = 0;
      uint64_t prev = KEY_NONE;
      for (uint32_t i = 0; i < ncand_pad && w < k; ++i) {
        uint64_t v = s_keys[i];
        if (v == KEY_NONE) break;
        if (v != prev) out_keys[(size_t)qi * k + w++] = v;
        prev = v;
      }
    }
  }
}

// Exact rerank of candidate rows cand[q][ncand] for all nq queries, or — with qlist / nq_dev — for
// the listed queries only (device-resident count; the launch exits at once when it is zero).
int32_t rerank_rows(scn_store* s, const float* d_q, uint64_t nq, const uint32_t* d_cand_rows, uint32_t ncand, uint32_t k,
                    uint64_t row_base, uint64_t* d_out_keys, cudaStream_t stream, const uint32_t* d_qlist,
                    const uint32_t* d_nq_dev) {
  if (nq == 0) return SCN_OK;
  uint32_t ncand_pad = std::max(32u, next_pow2(ncand));
  size_t smem = rerank_smem_bytes(s->pitch, ncand_pad);
  if (smem > 200 * 1024) return fail(SCN_ERR_INVALID_PARAMETERS, "too many rerank candidates (%u)", ncand);
  const unsigned grid = d_qlist ? (unsigned)std::min<uint64_t>(nq, 148 * 6) : (unsigned)nq;
  const bool pdl = s->opt_pdl != 0 && s->opt_profile == 0 && d_nq_dev != nullptr;   // the second-chance rerank of the tensor path
#define RR(MT)                                                                                                    \
  do {                                                                                                            \
    SCN_ALLOW_SMEM((rerank_kernel<MT>), smem);                                                                    \
    SCN_CUDA(launch_chained(rerank_kernel<MT>, dim3(grid), dim3(32), smem, stream, pdl, s->d_vec, s->d_norm, s->d_deleted, s->pitch, s->dim, \
                            (uint32_t)s->rows, d_q, d_cand_rows, ncand, ncand_pad, k, (uint32_t)row_base, d_qlist, d_nq_dev, (uint32_t)nq,   \
                            d_out_keys));                                                                                                 \
  } while (0)
  switch (s->metric) {
    case M_L2: RR(M_L2); break;
    case M_COS: RR(M_COS); break;
    default: RR(M_IP); break;
  }
#undef RR
  SCN_LAUNCHED();
  return SCN_OK;
}

// ---- scn_distance_batch: the scalar DistanceCalculator, one thread per (query, target) ---------
template <int METRIC>
__global__ void distance_batch_kernel(const float* __restrict__ q, uint64_t nq, const float* __restrict__ x, uint64_t nx,
                                      uint32_t dim, float* __restrict__ out) {
  uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= nq * nx) return;
  const float* a = q + (idx / nx) * dim;
  const float* b = x + (idx % nx) * dim;
  float acc = 0.0f, na = 0.0f, nb = 0.0f;
  for (uint32_t i = 0; i < dim; ++i) {
    acc = acc_step<METRIC>(acc, a[i], b[i]);
    if (METRIC == M_COS) {  // distance.go:58-63: three accumulators in one loop
      na = __fadd_rn(na, __fmul_rn(a[i], a[i]));
      nb = __fadd_rn(nb, __fmul_rn(b[i], b[i]));
    }
  }
  if (METRIC == M_COS) {
    na = __fsqrt_rn(na);
    nb = __fsqrt_rn(nb);
  }
  out[idx] = finish_distance<METRIC>(acc, na, nb);
}

int32_t distance_batch(int32_t metric, const float* d_q, uint64_t nq, const float* d_x, uint64_t nx, uint32_t dim,
                       float* d_out, cudaStream_t stream) {
  uint64_t total = nq * nx;
  if (total == 0) return SCN_OK;
  unsigned grid = (unsigned)((total + 127) / 128);
  switch (metric) {
    case M_L2: distance_batch_kernel<M_L2><<<grid, 128, 0, stream>>>(d_q, nq, d_x, nx, dim, d_out); break;
    case M_COS: distance_batch_kernel<M_COS><<<grid, 128, 0, stream>>>(d_q, nq, d_x, nx, dim, d_out); break;
    default: distance_batch_kernel<M_IP><<<grid, 128, 0, stream>>>(d_q, nq, d_x, nx, dim, d_out); break;
  }
  SCN_LAUNCHED();
  return SCN_OK;
}

// ---- scn_vector_ops: VectorMagnitude / NormalizeVector / DotProduct (distance.go:152-192) --------
// One thread per vector, the reference's sequential fp32 order (a warp reads 32 different vectors:
// these are API helpers for small inputs, not a streaming path).
__global__ void vector_ops_kernel(int op, const float* __restrict__ a, const float* __restrict__ b, uint64_t n, uint32_t dim,
                                  float* __restrict__ out) {
  const uint64_t i = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n) return;
  const float* v = a + i * dim;
  if (op == SCN_VEC_DOT) {
    const float* w = b + i * dim;
    float p = 0.0f;
    for (uint32_t j = 0; j < dim; ++j) p = __fadd_rn(p, __fmul_rn(v[j], w[j]));   // distance.go:189-191
    out[i] = p;
    return;
  }
  const float norm = exact_norm_thread(v, dim);                                     // distance.go:155-159, 176-180
  if (op == SCN_VEC_MAGNITUDE) {
    out[i] = norm;
    return;
  }
  float* o = out + i * dim;
  for (uint32_t j = 0; j < dim; ++j) o[j] = (norm == 0.0f) ? v[j] : __fdiv_rn(v[j], norm);  // distance.go:161-170
}

int32_t vector_ops(int32_t op, const float* d_a, const float* d_b, uint64_t n, uint32_t dim, float* d_out, cudaStream_t stream) {
  if (n == 0) return SCN_OK;
  vector_ops_kernel<<<(unsigned)((n + 127) / 128), 128, 0, stream>>>(op, d_a, d_b, n, dim, d_out);
  SCN_LAUNCHED();
  return SCN_OK;
}

// ---- K7: merge of G per-shard lists [G][nq][k] ---------------------------------------------------
// One warp per query. Keys carry the global row, so ascending key order is exactly the flat
// oracle's (distance, row) order over the whole database. ids ride along.
__global__ void __launch_bounds__(128) merge_topk_kernel(const uint64_t* __restrict__ keys, const uint64_t* __restrict__ ids,
                                                         uint32_t n_shards, uint64_t nq, uint32_t k, uint64_t shard_stride,
                                                         uint64_t* __restrict__ out_ids, float* __restrict__ out_dist,
                                                         uint32_t* __restrict__ out_counts, const uint32_t* __restrict__ status) {
  uint64_t q = (uint64_t)blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  if (q >= nq) return;
  if (status && *status != 0) {  // a peer never delivered its lists: no result is better than a partial one
    for (uint32_t j = lane; j < k; j += 32) {
      out_ids[q * k + j] = 0;
      out_dist[q * k + j] = __int_as_float(0x7f800000);
    }
    if (lane == 0 && out_counts) out_counts[q] = 0;
    return;
  }
  // each lane owns shards lane, lane+32, ...; heads advance as winners are emitted
  uint32_t head[8];  // supports up to 256 shards
#pragma unroll
  for (int i = 0; i < 8; ++i) head[i] = 0;
  uint32_t cnt = 0;
  for (uint32_t j = 0; j < k; ++j) {
    uint64_t best = KEY_NONE;
    uint32_t best_src = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) {
      uint32_t sh = lane + 32 * i;
      if (sh < n_shards && head[i] < k) {
        uint64_t v = keys[(size_t)sh * shard_stride + q * k + head[i]];
        if (v < best) {
          best = v;
          best_src = sh;
        }
      }
    }
    uint64_t wbest = best;
    for (int o = 16; o > 0; o >>= 1) {
      uint64_t other = __shfl_xor_sync(0xffffffffu, wbest, o);
      wbest = other < wbest ? other : wbest;
    }
    if (wbest == KEY_NONE) {
      if (lane == 0) {
        out_ids[q * k + j] = 0;
        out_dist[q * k + j] = __int_as_float(0x7f800000);
      }
      continue;
    }
    // the lane holding the winner (keys are unique: global rows differ) emits and advances
    if (best == wbest) {
      uint32_t i = best_src / 32;
      out_ids[q * k + j] = ids[(size_t)best_src * shard_stride + q * k + head[i]];
      out_dist[q * k + j] = ord_f32((uint32_t)(wbest >> 32));
#pragma unroll
      for (int t = 0; t < 8; ++t)
        if (t == (int)i) head[t]++;
    }
    ++cnt;
  }
  if (lane == 0 && out_counts) out_counts[q] = cnt;
}

int32_t merge_topk(const uint64_t* d_keys, const uint64_t* d_ids, uint32_t n_shards, uint64_t nq, uint32_t k,
                   uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts, cudaStream_t stream, uint64_t shard_stride,
                   const uint32_t* d_status) {
  if (nq == 0) return SCN_OK;
  if (n_shards == 0 || n_shards > 256) return fail(SCN_ERR_INVALID_PARAMETERS, "n_shards must be in [1, 256]");
  merge_topk_kernel<<<(unsigned)((nq + 3) / 4), 128, 0, stream>>>(d_keys, d_ids, n_shards, nq, k, shard_stride ? shard_stride : nq * k,
                                                                  d_out_ids, d_out_dist, d_out_counts, d_status);
  SCN_LAUNCHED();
  return SCN_OK;
}

}  // namespace scn
