// hnsw_search.cu — K5: batched HNSW.Search (hnsw.go:292-350) — greedy descent through the upper
// layers, layer-0 beam search (searchLayer, hnsw.go:487-557) and the top-k rerank, one warp per
// query.
//
// Mapping of the reference's state onto the warp:
//   candidates (W)  -> one sorted list of ef (key, row) pairs in shared memory;
//                      key = ord(dist) << 32 | admission_seq << 1 | expanded  (stable: on equal
//                      distance the earlier-admitted entry stays ahead, hnsw.go:675-686)
//   dynamic (C)     -> the `expanded` bit: every live element of C is an un-expanded element of W
//                      (an element evicted from W has dist >= W[ef-1].dist and can only ever
//                      trigger the `break` at hnsw.go:516-518), so "pop the closest of C" is "take
//                      the first un-expanded entry of W" and the loop ends when there is none.
//                      (Only exact float ties with W[ef-1] can make the two differ.)
//   visited         -> open-addressing hash of row indices in shared memory (global-memory table
//                      as the overflow path)
//   Distance()      -> traversal: 128-bit coalesced row loads, FMA, warp-shuffle reduction
//                      (ordering-only; L2 is compared squared); returned distances: recomputed in
//                      the reference's exact sequential fp32 arithmetic (hnsw.go:330).
// Per neighbour the reference's order of tests is kept: visited? -> deleted? (not marked visited,
// not traversed) -> mark visited -> distance -> admit if |W| < ef or d < W[ef-1].d (strict).
// Roofline: HBM random gather; algorithmic bytes per query = evals*dim*4 + hops*2M*4.
#include "store.h"

namespace scn {

constexpr int HNSW_WARPS = 2;
constexpr uint32_t HASH_EMPTY = 0u;

struct HnswArgs {
  const float* vec;
  const float* norm;
  const uint32_t* deleted;
  const uint64_t* ids;
  const uint32_t* adj0;
  const uint32_t* adj_up;
  const uint8_t* levels;
  const uint32_t* up_off;
  uint32_t pitch, dim, n_rows;
  uint32_t s0, su;
  uint32_t entry_row;
  int32_t max_layer;
  const float* q;
  const uint32_t* qlist;   // optional list of query indices (overflow pass)
  uint32_t* nq_dev;        // optional device count for qlist
  uint32_t nq, k, ef, ef_pad;
  uint32_t hash_size;      // entries per warp
  uint32_t* ghash;         // global tables [warps][hash_size] when USE_GLOBAL
  uint32_t* overflow_list; // queries whose visited table overflowed (smem pass)
  uint32_t* overflow_count;
  uint64_t* out_ids;
  float* out_dist;
  uint32_t* out_counts;
  unsigned long long* stats;  // [0] distance evaluations, [1] expansions (optional)
};

template <int METRIC>
__device__ __forceinline__ float traversal_finish(float acc, float qn, float xn) {
  if (METRIC == M_L2) return acc;  // squared: same ordering as the reference's sqrt
  if (METRIC == M_IP) return -acc;
  if (qn == 0.0f || xn == 0.0f) return 1.0f;
  float cs = acc / (qn * xn);
  cs = fminf(1.0f, fmaxf(-1.0f, cs));
  return 1.0f - cs;
}

// Distances from the query (shared memory) to up to 4 rows at once; every lane returns all four
// reduced sums. Rows equal to ROW_NONE are skipped (their result is unspecified).
template <int METRIC>
__device__ __forceinline__ void warp_dist4(const float* __restrict__ sq, const float* __restrict__ vec, uint32_t pitch,
                                           const uint32_t r[4], int lane, float out[4]) {
  const uint32_t p4 = pitch >> 2;
  const float4* q4 = reinterpret_cast<const float4*>(sq);
  const float4* x0 = reinterpret_cast<const float4*>(vec + (size_t)(r[0] == ROW_NONE ? 0 : r[0]) * pitch);
  const float4* x1 = reinterpret_cast<const float4*>(vec + (size_t)(r[1] == ROW_NONE ? 0 : r[1]) * pitch);
  const float4* x2 = reinterpret_cast<const float4*>(vec + (size_t)(r[2] == ROW_NONE ? 0 : r[2]) * pitch);
  const float4* x3 = reinterpret_cast<const float4*>(vec + (size_t)(r[3] == ROW_NONE ? 0 : r[3]) * pitch);
  float a0 = 0.f, a1 = 0.f, a2 = 0.f, a3 = 0.f;
#pragma unroll 2
  for (uint32_t i = lane; i < p4; i += 32) {
    float4 qv = q4[i];
    float4 v0 = __ldg(x0 + i), v1 = __ldg(x1 + i), v2 = __ldg(x2 + i), v3 = __ldg(x3 + i);
    if (METRIC == M_L2) {
      float d;
      d = qv.x - v0.x; a0 = fmaf(d, d, a0); d = qv.y - v0.y; a0 = fmaf(d, d, a0);
      d = qv.z - v0.z; a0 = fmaf(d, d, a0); d = qv.w - v0.w; a0 = fmaf(d, d, a0);
      d = qv.x - v1.x; a1 = fmaf(d, d, a1); d = qv.y - v1.y; a1 = fmaf(d, d, a1);
      d = qv.z - v1.z; a1 = fmaf(d, d, a1); d = qv.w - v1.w; a1 = fmaf(d, d, a1);
      d = qv.x - v2.x; a2 = fmaf(d, d, a2); d = qv.y - v2.y; a2 = fmaf(d, d, a2);
      d = qv.z - v2.z; a2 = fmaf(d, d, a2); d = qv.w - v2.w; a2 = fmaf(d, d, a2);
      d = qv.x - v3.x; a3 = fmaf(d, d, a3); d = qv.y - v3.y; a3 = fmaf(d, d, a3);
      d = qv.z - v3.z; a3 = fmaf(d, d, a3); d = qv.w - v3.w; a3 = fmaf(d, d, a3);
    } else {
      a0 = fmaf(qv.x, v0.x, a0); a0 = fmaf(qv.y, v0.y, a0); a0 = fmaf(qv.z, v0.z, a0); a0 = fmaf(qv.w, v0.w, a0);
      a1 = fmaf(qv.x, v1.x, a1); a1 = fmaf(qv.y, v1.y, a1); a1 = fmaf(qv.z, v1.z, a1); a1 = fmaf(qv.w, v1.w, a1);
      a2 = fmaf(qv.x, v2.x, a2); a2 = fmaf(qv.y, v2.y, a2); a2 = fmaf(qv.z, v2.z, a2); a2 = fmaf(qv.w, v2.w, a2);
      a3 = fmaf(qv.x, v3.x, a3); a3 = fmaf(qv.y, v3.y, a3); a3 = fmaf(qv.z, v3.z, a3); a3 = fmaf(qv.w, v3.w, a3);
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a0 += __shfl_xor_sync(0xffffffffu, a0, o);
    a1 += __shfl_xor_sync(0xffffffffu, a1, o);
    a2 += __shfl_xor_sync(0xffffffffu, a2, o);
    a3 += __shfl_xor_sync(0xffffffffu, a3, o);
  }
  out[0] = a0; out[1] = a1; out[2] = a2; out[3] = a3;
}

// Evaluates the traversal distance of the `ns` rows held by lanes [0, ns) (value `nb`); lane j
// receives the distance of its own row.
template <int METRIC>
__device__ __forceinline__ float eval_rows(const HnswArgs& a, const float* sq, float qn, uint32_t nb, uint32_t ns,
                                           int lane) {
  float mine = 0.0f;
  for (uint32_t t = 0; t < ns; t += 4) {
    uint32_t r[4];
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      uint32_t v = __shfl_sync(0xffffffffu, nb, (t + u) & 31);
      r[u] = (t + u < ns) ? v : ROW_NONE;
    }
    float d[4];
    warp_dist4<METRIC>(sq, a.vec, a.pitch, r, lane, d);
#pragma unroll
    for (int u = 0; u < 4; ++u) {
      if ((uint32_t)lane == t + u && t + u < ns) {
        float xn = (METRIC == M_COS) ? __ldg(a.norm + r[u]) : 0.0f;
        mine = traversal_finish<METRIC>(d[u], qn, xn);
      }
    }
  }
  return mine;
}

template <bool USE_GLOBAL>
__device__ __forceinline__ bool visited_insert(uint32_t* tab, uint32_t hash_size, uint32_t row) {
  uint32_t h = __umulhi(row * 2654435761u, hash_size);
  const uint32_t key = row + 1;
  for (uint32_t probes = 0; probes < hash_size; ++probes) {
    uint32_t old = atomicCAS(tab + h, HASH_EMPTY, key);
    if (old == HASH_EMPTY) return true;
    if (old == key) return false;
    if (++h == hash_size) h = 0;
  }
  return false;  // table full (guarded against by the overflow check)
}

template <int METRIC, bool USE_GLOBAL>
__global__ void __launch_bounds__(HNSW_WARPS * 32) hnsw_search_kernel(HnswArgs a) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t warp_global = blockIdx.x * HNSW_WARPS + warp;
  const uint32_t warps_total = gridDim.x * HNSW_WARPS;
  // per-warp carve-up: q[pitch] | wkey[ef_pad] | wrow[ef_pad] | hash[hash_size] (smem mode)
  const size_t per_warp = (size_t)a.pitch * 4 + (size_t)a.ef_pad * 12 + (USE_GLOBAL ? 0 : (size_t)a.hash_size * 4);
  unsigned char* base = smem_raw + per_warp * warp;
  uint64_t* wkey = reinterpret_cast<uint64_t*>(base);
  float* sq = reinterpret_cast<float*>(wkey + a.ef_pad);
  uint32_t* wrow = reinterpret_cast<uint32_t*>(sq + a.pitch);
  uint32_t* hash = USE_GLOBAL ? (a.ghash + (size_t)warp_global * a.hash_size) : (wrow + a.ef_pad);
  const uint32_t ef = a.ef;
  const uint32_t nq = a.qlist ? min(*a.nq_dev, a.nq) : a.nq;
  unsigned long long evals = 0, hops = 0;

  for (uint32_t qslot = warp_global; qslot < nq; qslot += warps_total) {
    const uint32_t qi = a.qlist ? a.qlist[qslot] : qslot;
    __syncwarp();
    for (uint32_t i = lane; i < a.pitch; i += 32) sq[i] = (i < a.dim) ? a.q[(size_t)qi * a.dim + i] : 0.0f;
    if (USE_GLOBAL) {
      for (uint32_t i = lane; i < a.hash_size; i += 32) hash[i] = HASH_EMPTY;
    } else {
      for (uint32_t i = lane; i < a.hash_size; i += 32) hash[i] = HASH_EMPTY;
    }
    __syncwarp();
    float qn = 0.0f;
    if (METRIC == M_COS) {
      if (lane == 0) qn = exact_norm_thread(sq, a.dim);
      qn = __shfl_sync(0xffffffffu, qn, 0);
    }

    uint32_t n_out = 0;
    uint32_t cnt = 0;
    bool overflow = false;
    uint32_t cur = a.entry_row;
    const bool have_entry = (cur != ROW_NONE) && (cur < a.n_rows) && !bit_test(a.deleted, cur) && a.max_layer >= 0;
    if (have_entry) {
      // ---- entry distance -------------------------------------------------------------------
      float dcur;
      {
        uint32_t r[4] = {cur, ROW_NONE, ROW_NONE, ROW_NONE};
        float d[4];
        warp_dist4<METRIC>(sq, a.vec, a.pitch, r, lane, d);
        dcur = traversal_finish<METRIC>(d[0], qn, METRIC == M_COS ? __ldg(a.norm + cur) : 0.0f);
        ++evals;
      }
      // ---- greedy descent, layers maxLayer..1 with numClosest = 1 (hnsw.go:309-311) -----------
      for (int layer = a.max_layer; layer >= 1; --layer) {
        for (;;) {
          if ((int)a.levels[cur] < layer) break;  // GetConnections(layer) is empty (hnsw.go:45-50)
          ++hops;
          const uint32_t* list = a.adj_up + ((size_t)a.up_off[cur] + (layer - 1)) * a.su;
          float best = dcur;
          uint32_t best_row = cur;
          for (uint32_t c0 = 0; c0 < a.su; c0 += 32) {
            uint32_t nb = (c0 + lane < a.su) ? __ldg(list + c0 + lane) : ROW_NONE;
            bool ok = (nb != ROW_NONE) && !bit_test(a.deleted, nb);
            uint32_t mask = __ballot_sync(0xffffffffu, ok);
            uint32_t ns = __popc(mask);
            if (!ns) continue;
            uint32_t src = __fns(mask, 0, lane + 1);
            uint32_t nbj = __shfl_sync(0xffffffffu, nb, src & 31);
            float dj = eval_rows<METRIC>(a, sq, qn, nbj, ns, lane);
            evals += ns;
            // first strict minimum in list order (sequential `d < W[0].d` updates)
            float dm = ((uint32_t)lane < ns) ? dj : __int_as_float(0x7f800000);
            uint32_t im = lane;
#pragma unroll
            for (int o = 16; o > 0; o >>= 1) {
              float od = __shfl_xor_sync(0xffffffffu, dm, o);
              uint32_t oi = __shfl_xor_sync(0xffffffffu, im, o);
              if (od < dm || (od == dm && oi < im)) {
                dm = od;
                im = oi;
              }
            }
            if (dm < best) {
              best = dm;
              best_row = __shfl_sync(0xffffffffu, nbj, im);
            }
          }
          if (best_row == cur) break;
          cur = best_row;
          dcur = best;
        }
      }
      // ---- layer 0 beam (hnsw.go:314, 487-557) ---------------------------------------------
      uint32_t seq = 0, visited = 1;
      if (lane == 0) {
        wkey[0] = ((uint64_t)f32_ord(dcur) << 32) | ((uint64_t)(seq) << 1);
        wrow[0] = cur;
      }
      seq = 1;
      cnt = 1;
      visited_insert<USE_GLOBAL>(hash, a.hash_size, cur);
      __syncwarp();
      for (;;) {
        // closest un-expanded entry of W
        int p = -1;
        for (uint32_t b0 = 0; b0 < cnt; b0 += 32) {
          uint32_t i = b0 + lane;
          bool un = (i < cnt) && !(wkey[i] & 1ull);
          uint32_t m = __ballot_sync(0xffffffffu, un);
          if (m) {
            p = (int)b0 + __ffs(m) - 1;
            break;
          }
        }
        if (p < 0) break;
        cur = wrow[p];
        __syncwarp();
        if (lane == 0) wkey[p] |= 1ull;
        __syncwarp();
        ++hops;
        if (visited + a.s0 > a.hash_size - (a.hash_size >> 3)) {  // keep the table below 7/8 full
          overflow = true;
          break;
        }
        const uint32_t* list = a.adj0 + (size_t)cur * a.s0;
        for (uint32_t c0 = 0; c0 < a.s0; c0 += 32) {
          uint32_t nb = (c0 + lane < a.s0) ? __ldg(list + c0 + lane) : ROW_NONE;
          bool ok = (nb != ROW_NONE);
          // reference order: visited? -> deleted? -> mark visited. A deleted row is never
          // inserted, so testing `deleted` first and inserting only live rows is equivalent.
          if (ok) ok = !bit_test(a.deleted, nb);
          if (ok) ok = visited_insert<USE_GLOBAL>(hash, a.hash_size, nb);
          uint32_t mask = __ballot_sync(0xffffffffu, ok);
          uint32_t ns = __popc(mask);
          if (!ns) continue;
          visited += ns;
          uint32_t src = __fns(mask, 0, lane + 1);
          uint32_t nbj = __shfl_sync(0xffffffffu, nb, src & 31);
          float dj = eval_rows<METRIC>(a, sq, qn, nbj, ns, lane);
          evals += ns;
          // sequential admission in adjacency order (hnsw.go:536-546)
          for (uint32_t j = 0; j < ns; ++j) {
            float d = __shfl_sync(0xffffffffu, dj, j);
            uint32_t row = __shfl_sync(0xffffffffu, nbj, j);
            uint32_t od = f32_ord(d);
            if (cnt >= ef && !(od < (uint32_t)(wkey[ef - 1] >> 32))) continue;  // strict <, NaN never admitted
            uint64_t key = ((uint64_t)od << 32) | ((uint64_t)seq << 1);
            ++seq;
            // position: after every entry with distance <= d (stable tail insertion)
            uint32_t pos = 0;
            for (uint32_t b0 = 0; b0 < cnt; b0 += 32) {
              uint32_t i = b0 + lane;
              bool lt = (i < cnt) && ((wkey[i] >> 32) <= od);
              pos += __popc(__ballot_sync(0xffffffffu, lt));
            }
            const uint32_t last = (cnt < ef) ? cnt : ef - 1;  // destination of the last shifted entry
            for (int b0 = (int)(last / 32) * 32; b0 >= 0; b0 -= 32) {
              uint32_t i = b0 + lane;
              bool mv = (i <= last) && (i > pos);
              uint64_t kv = 0;
              uint32_t rv = 0;
              if (mv) {
                kv = wkey[i - 1];
                rv = wrow[i - 1];
              }
              __syncwarp();
              if (mv) {
                wkey[i] = kv;
                wrow[i] = rv;
              }
              __syncwarp();
            }
            if (lane == 0) {
              wkey[pos] = key;
              wrow[pos] = row;
            }
            if (cnt < ef) ++cnt;
            __syncwarp();
          }
        }
      }
      n_out = min(a.k, cnt);
    }

    if (overflow && !USE_GLOBAL) {
      if (lane == 0) {
        uint32_t slot = atomicAdd(a.overflow_count, 1u);
        a.overflow_list[slot] = qi;
      }
      continue;  // the overflow pass will produce this query's results
    }

    // ---- rerank (hnsw.go:317-347): exact distances of the first n_out candidates, stable sort ----
    const uint32_t np = max(32u, 1u << (32 - __clz(max(n_out, 1u) - 1)));
    __syncwarp();
    for (uint32_t i0 = 0; i0 < np; i0 += 32) {
      uint32_t i = i0 + lane;
      uint64_t key = KEY_NONE;
      if (i < n_out) {
        uint32_t row = wrow[i];
        float acc = exact_acc_thread<METRIC>(sq, a.vec + (size_t)row * a.pitch, a.pitch >> 2);
        float d = finish_distance<METRIC>(acc, qn, METRIC == M_COS ? __ldg(a.norm + row) : 0.0f);
        key = ((uint64_t)f32_ord(d) << 32) | i;
      }
      __syncwarp();
      if (i < a.ef_pad) wkey[i] = key;
    }
    __syncwarp();
    for (uint32_t size = 2; size <= np; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t t = lane; t < np / 2; t += 32) {
          uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
          bool up = ((lo & size) == 0);
          uint64_t x = wkey[lo], y = wkey[hi];
          if ((x > y) == up) {
            wkey[lo] = y;
            wkey[hi] = x;
          }
        }
        __syncwarp();
      }
    }
    for (uint32_t i = lane; i < a.k; i += 32) {
      uint64_t id = 0;
      float d = __int_as_float(0x7f800000);
      if (i < n_out) {
        uint64_t key = wkey[i];
        id = a.ids[wrow[(uint32_t)key]];
        d = ord_f32((uint32_t)(key >> 32));
      }
      a.out_ids[(size_t)qi * a.k + i] = id;
      a.out_dist[(size_t)qi * a.k + i] = d;
    }
    if (lane == 0 && a.out_counts) a.out_counts[qi] = n_out;
  }
  if (a.stats && lane == 0) {
    atomicAdd(a.stats + 0, evals);
    atomicAdd(a.stats + 1, hops);
  }
}

template <int METRIC>
static int32_t launch_hnsw(HnswArgs& a, int sms, cudaStream_t stream, Scratch& scratch, Profiler* prof) {
  // pass 1: shared-memory visited table
  const size_t per_warp = (size_t)a.pitch * 4 + (size_t)a.ef_pad * 12 + (size_t)a.hash_size * 4;
  const size_t smem = per_warp * HNSW_WARPS;
  SCN_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<METRIC, false>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem));
  int per_sm = 0;
  SCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hnsw_search_kernel<METRIC, false>, HNSW_WARPS * 32, smem));
  if (per_sm < 1) return fail(SCN_ERR_INVALID_PARAMETERS, "ef=%u / dim=%u need more shared memory than one SM has", a.ef, a.dim);
  uint32_t blocks_needed = (a.nq + HNSW_WARPS - 1) / HNSW_WARPS;
  int grid = (int)std::min<uint32_t>(blocks_needed, (uint32_t)(sms * per_sm));
  SCN_TRY(scratch.alloc(&a.overflow_list, (size_t)a.nq));
  SCN_TRY(scratch.alloc(&a.overflow_count, 1));
  SCN_CUDA(cudaMemsetAsync(a.overflow_count, 0, sizeof(uint32_t), stream));
  if (prof) prof->begin("hnsw_search");
  hnsw_search_kernel<METRIC, false><<<grid, HNSW_WARPS * 32, smem, stream>>>(a);
  SCN_LAUNCHED();
  if (prof) prof->end();
  // pass 2 (always enqueued, exits at once when nothing overflowed): global-memory visited table
  HnswArgs b = a;
  b.qlist = a.overflow_list;
  b.nq_dev = a.overflow_count;
  // 8x the shared-memory table, never more than 2x the row count (a table that cannot fill up)
  b.hash_size = (uint32_t)std::min<uint64_t>(std::min<uint64_t>((uint64_t)1 << 20, next_pow2(a.n_rows) * 2ull),
                                             std::max<uint64_t>(next_pow2(a.hash_size) * 8ull, 1024));
  b.hash_size = std::max<uint32_t>(b.hash_size, 1024u);
  const int grid2 = std::min<int>(sms, (int)blocks_needed);
  SCN_TRY(scratch.alloc(&b.ghash, (size_t)grid2 * HNSW_WARPS * b.hash_size));
  const size_t smem2 = ((size_t)a.pitch * 4 + (size_t)a.ef_pad * 12) * HNSW_WARPS;
  SCN_CUDA(cudaFuncSetAttribute(hnsw_search_kernel<METRIC, true>, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem2));
  if (prof) prof->begin("hnsw_search_overflow");
  hnsw_search_kernel<METRIC, true><<<grid2, HNSW_WARPS * 32, smem2, stream>>>(b);
  SCN_LAUNCHED();
  if (prof) prof->end();
  return SCN_OK;
}

int32_t hnsw_search(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint32_t ef, uint64_t* d_out_ids,
                    float* d_out_dist, uint32_t* d_out_counts, cudaStream_t stream, Profiler* prof) {
  if (nq == 0) return SCN_OK;
  int sms = 0;
  SCN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
  Scratch scratch(stream);
  HnswArgs a{};
  a.vec = s->d_vec;
  a.norm = s->d_norm;
  a.deleted = s->d_deleted;
  a.ids = s->d_ids;
  a.adj0 = s->d_adj0;
  a.adj_up = s->d_adj_up;
  a.levels = s->d_levels;
  a.up_off = s->d_up_off;
  a.pitch = s->pitch;
  a.dim = s->dim;
  a.n_rows = (uint32_t)s->rows;
  a.s0 = 2 * (uint32_t)s->m;
  a.su = (uint32_t)s->m;
  a.entry_row = s->entry_row;
  a.max_layer = s->max_layer;
  a.q = d_q;
  a.qlist = nullptr;
  a.nq_dev = nullptr;
  a.nq = (uint32_t)nq;
  a.k = k;
  a.ef = ef;
  a.ef_pad = std::max(32u, next_pow2(ef));
  // visited table: ~2M*1.25 rows per expansion, ~ef expansions; keep it under 7/8 full
  a.hash_size = round_up(std::max<uint32_t>(1024u, (uint32_t)std::min<uint64_t>((uint64_t)ef * (uint32_t)s->m * 5 / 2, 1u << 16)), 512);
  a.out_ids = d_out_ids;
  a.out_dist = d_out_dist;
  a.out_counts = d_out_counts;
  SCN_CUDA(cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(unsigned long long), stream));
  a.stats = s->opt_profile ? s->d_counters : nullptr;
  // shrink the table until at least one block fits
  while ((((size_t)a.pitch * 4 + (size_t)a.ef_pad * 12 + (size_t)a.hash_size * 4) * HNSW_WARPS) > 200 * 1024 &&
         a.hash_size > 1024)
    a.hash_size = round_up(a.hash_size / 2, 512);
  int32_t rc;
  switch (s->metric) {
    case M_L2: rc = launch_hnsw<M_L2>(a, sms, stream, scratch, prof); break;
    case M_COS: rc = launch_hnsw<M_COS>(a, sms, stream, scratch, prof); break;
    default: rc = launch_hnsw<M_IP>(a, sms, stream, scratch, prof); break;
  }
  SCN_TRY(rc);
  return SCN_OK;
}

}  // namespace scn
