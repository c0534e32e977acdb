// hnsw_build.cu — GPU-assisted HNSW construction with the reference's SERIAL semantics
// (hnsw.go:190-257 insertVector, 487-557 searchLayer, 560-583 selectNeighbors,
// 586-614 pruneConnections).
//
// The reference inserts one vector at a time: every insert searches the graph the previous insert
// left behind. What an insert costs is its searches (greedy descent, then searchLayer with
// efConstruction on every layer of the new node: thousands of distance evaluations); what it
// changes is tiny (its own lists and one slot of <= 2M neighbour lists). So the searches of a WINDOW
// of upcoming inserts are run speculatively on the device against one snapshot of the graph, one
// warp per insert, and the inserts are then committed on the host strictly in order. An insert's
// speculative search is used only if it is provably the search the serial algorithm would have run
// on the graph as it is at commit time:
//
//   * the kernel logs every expansion (node, layer, W[ef-1] at that moment — or "W not full" — and
//     which neighbours of the list were admitted);
//   * a commit records every change of a neighbour SET (row, layer, added node, removed node);
//   * a search is valid iff for every list it expanded that has changed since the snapshot: no added
//     node would have been admitted (its distance to the new vector — taken from a device-computed
//     window x window matrix — is not below the logged threshold, and W was full), and no removed
//     node had been admitted by that expansion (logged bit). A non-admitted evaluation has no effect
//     on a walk, so such a search is step for step the serial one. Entry point / maxLayer changes
//     invalidate the rest of the window.
//
// The first invalid insert ends the round; the window is searched again on the updated graph
// (a round costs one wave of walks, whatever the window size). All distance arithmetic is done on
// the device in the reference's order and rounding; the host only compares and sorts the returned
// values: neighbour selection is the first maxConn entries of the (distance, admission)-sorted W,
// and pruneConnections needs no new distances because every stored edge keeps the distance it was
// created with (Distance(a, b) and Distance(b, a) are the same bits for all three metrics).
// Level draws (selectLayer, hnsw.go:458-469) are the caller's: the Go shim draws them from its own
// math/rand stream, so the graph is the one the CPU index would have built. As everywhere in this
// library, only exact float ties with W[ef-1] could make a walk differ from the reference's.
#include <algorithm>
#include <chrono>
#include <cstring>
#include <unordered_map>

#include "store.h"
#include "visited.cuh"

namespace scn {

constexpr uint32_t BUILD_LOGCAP = 2048;        // expansion records kept per search (a level-L node logs ~250 (L + 1))
constexpr uint32_t LOG_NOT_FULL = 0xFFFFFFFFu;  // "W held fewer than ef entries": everything is admitted
constexpr uint32_t LOG_OVERFLOW = 0xFFFFFFFFu;  // log_cnt: the visited table filled up, search unusable

// host mirror of the graph under construction (device layout, plus the distance of every edge)
struct BuildState {
  uint32_t m = 0, s0 = 0, su = 0;
  uint64_t n_nodes = 0;                 // rows 0..n_nodes-1 are in the graph
  std::vector<uint8_t> level;           // len(Connections) - 1 per row
  std::vector<uint32_t> up_off;         // first upper list per row
  uint64_t upper_lists = 0;
  std::vector<uint32_t> adj0, ord0;     // [n][s0] neighbour rows / ord(distance) of the edge
  std::vector<uint8_t> cnt0;            // [n]
  std::vector<uint32_t> adju, ordu;     // [lists][su]
  std::vector<uint8_t> cntu;            // [lists]
  std::vector<uint8_t> deleted;         // [n] byte per row (host copy of the bitmap)
  uint64_t cap_rows = 0, cap_lists = 0; // device capacity of the adjacency arrays
};

void free_build_state(scn_store* s) {
  delete s->build;
  s->build = nullptr;
}

struct BuildArgs {
  const float* vec;
  const float* norm;
  const uint32_t* deleted;
  const uint32_t* adj0;
  const uint32_t* adj_up;
  const uint8_t* levels;
  const uint32_t* up_off;
  uint32_t pitch, n_rows, s0, su, has_deleted;
  uint32_t entry_row;
  int32_t max_layer;
  const uint32_t* q_rows;    // [n_slots] row of the node to insert
  const uint8_t* q_levels;   // [n_slots]
  const uint32_t* out_off;   // [n_slots] first output list of the slot
  uint32_t n_slots, efc, ef_pad;
  uint32_t* ghash;           // [grid][hash_size]
  uint32_t hash_size, row_bits, tag_max;
  uint32_t* w_rows;          // [lists][efc]
  uint32_t* w_ord;           // [lists][efc]
  uint32_t* w_cnt;           // [lists]
  uint4* log;                // [n_slots][BUILD_LOGCAP] {row, W[ef-1] ord or LOG_NOT_FULL, admitted mask, layer | chunk << 8}
  uint32_t* log_cnt;         // [n_slots]
  unsigned long long* stats; // [0] distance evaluations, [1] expansions
  // ---- search mode (ext_q != nullptr): the walk of HNSW.Search for external queries — the exact second pass
  // of hnsw_search.cu for queries whose walk met a distance tie at the edge of W or overflowed its table
  const float* ext_q;        // [.][dim]
  const uint32_t* qlist;     // query indices
  const uint32_t* nq_dev;    // device-resident number of listed queries
  const uint64_t* ids;
  uint32_t dim, k;
  uint64_t* out_ids;
  float* out_dist;
  uint32_t* out_counts;
  unsigned long long* failed;
};

__host__ __device__ inline size_t build_warp_bytes(uint32_t pitch, uint32_t ef_pad, uint32_t stage_bytes) {
  // wkey[2][ef_pad] u64 | snk[32] u64 | q[pitch] f32 | wrow[2][ef_pad] u32 | snr[32] u32 | eps[64] u32 | stage
  // (ef_pad = capacity of a candidate array: 2 * efConstruction, W plus ghosts)
  return (size_t)ef_pad * 16 + 256 + (size_t)pitch * 4 + (size_t)ef_pad * 8 + 128 + 256 + stage_bytes;
}

// ---- the candidate lists of searchLayer, exactly (ties included) -----------------------------------
// `candidates` (W) and `dynamic` (C) of hnsw.go:487-557 live in ONE sorted array of keys
// ord(dist) << 32 | admission_seq << 1 | expanded (ascending key order = the reference's stable
// insertion order):
//   positions [0, cnt)           W, cnt <= ef
//   positions [cnt, cnt + gcnt)  "ghosts": elements that were admitted, have since been pushed out of W,
//                                and whose distance EQUALS W[ef-1]'s. They are the only evicted members
//                                of C the reference can still expand (it pops C in (distance, admission)
//                                order and stops at the first one with dist > W[ef-1].dist, hnsw.go:516-518;
//                                an evicted element never has dist < W[ef-1].dist). All of them were in W
//                                when its last distance took its current value, so there are at most ef.
// C's un-popped part = the un-expanded entries of [0, cnt + gcnt), in array order.
// Admission is the reference's SEQUENTIAL rule (hnsw.go:536-542): the neighbours of a list are judged
// one after the other against the W the previous ones left, i.e. neighbour j enters iff fewer than ef
// elements of (W before the list) u (earlier neighbours of the list) have distance <= its own.
struct WState {
  uint64_t *key, *okey;
  uint32_t *row, *orow;
  uint32_t cnt, gcnt, p_lo;
};

// number of keys in k[0, n) that are < x  (k ascending)
__device__ __forceinline__ uint32_t lower_bound_keys(const uint64_t* k, uint32_t n, uint64_t x) {
  uint32_t lo = 0, hi = n;
  while (lo < hi) {
    const uint32_t mid = (lo + hi) >> 1;
    if (k[mid] < x) lo = mid + 1;
    else hi = mid;
  }
  return lo;
}

// Judges the candidate lanes (`cand`: evaluated for the first time and, if W is full, closer than
// W[ef-1] was when the list part was read) in list order, merges the admitted ones into the other
// buffer. Returns the mask of admitted lanes.
__device__ __forceinline__ uint32_t merge_admitted(WState& w, uint64_t* snk, uint32_t* snr, bool cand, uint64_t key, uint32_t nb,
                                                   uint32_t ef, uint32_t cap, uint32_t lane) {
  const uint32_t cand_mask = __ballot_sync(0xffffffffu, cand);
  const uint32_t nc = __popc(cand_mask);
  if (!nc) return 0u;
  const uint32_t tot = w.cnt + w.gcnt;
  const uint32_t slot = __popc(cand_mask & ((1u << lane) - 1u));
  if (cand) {
    snk[slot] = key;
    snr[slot] = nb;
  }
  __syncwarp();
  // lane j < nc owns candidate j (list order)
  const bool mine = lane < nc;
  const uint64_t nkey = mine ? snk[lane] : KEY_NONE;
  const uint32_t nord = (uint32_t)(nkey >> 32);
  uint32_t lb = 0, le_old = 0;
  if (mine) {
    lb = lower_bound_keys(w.key, tot, nkey);                                           // old entries ahead of it
    le_old = lower_bound_keys(w.key, w.cnt, ((uint64_t)nord << 32) | 0xFFFFFFFFull);   // old W entries with dist <= its own
  }
  // earlier candidates of the list with dist <= its own (admitted or not: one that was refused had
  // dist >= W[ef-1] then, and W[ef-1] has not grown since)
  uint32_t le_new = 0;
  for (uint32_t k = 0; k < nc; ++k) {
    const uint32_t ok_ = __shfl_sync(0xffffffffu, nord, k);
    le_new += (k < lane && ok_ <= nord) ? 1u : 0u;
  }
  const bool adm = mine && (le_old + le_new < ef);
  const uint32_t adm_mask = __ballot_sync(0xffffffffu, adm);   // over candidate slots
  const uint32_t na = __popc(adm_mask);
  // admitted lanes in list positions (for the log)
  const uint32_t lanes_admitted = __ballot_sync(0xffffffffu, cand && ((adm_mask >> slot) & 1u));
  if (!na) return 0u;
  // position of an admitted new key: old entries ahead + admitted new keys ahead
  uint32_t npos = lb;
  for (uint32_t k = 0; k < nc; ++k) {
    const uint64_t kk = snk[k];   // broadcast
    npos += (((adm_mask >> k) & 1u) && kk < nkey) ? 1u : 0u;
  }
  // old entries move back by the number of admitted new keys ahead of them
  for (uint32_t i0 = 0; i0 < tot; i0 += 32) {
    const uint32_t i = i0 + lane;
    const uint64_t kv = (i < tot) ? w.key[i] : KEY_NONE;
    const uint32_t rw = (i < tot) ? w.row[i] : ROW_NONE;
    uint32_t below = 0;
    for (uint32_t k = 0; k < nc; ++k) {
      const uint64_t kk = snk[k];
      below += (((adm_mask >> k) & 1u) && kk < kv) ? 1u : 0u;
    }
    const uint32_t pos = i + below;
    if (i < tot && pos < cap) {
      w.okey[pos] = kv;
      w.orow[pos] = rw;
    }
  }
  if (adm && npos < cap) {
    w.okey[npos] = nkey;
    w.orow[npos] = snr[lane];
  }
  w.p_lo = min(w.p_lo, __reduce_min_sync(0xffffffffu, adm ? npos : 0xFFFFFFFFu));
  __syncwarp();
  uint64_t* tk = w.key;
  w.key = w.okey;
  w.okey = tk;
  uint32_t* tr = w.row;
  w.row = w.orow;
  w.orow = tr;
  const uint32_t ntot = min(cap, tot + na);
  w.cnt = min(ef, ntot);
  // ghosts: the run of entries right behind W[ef-1] that share its distance
  uint32_t g = 0;
  if (ntot > ef) {
    const uint32_t wo = reinterpret_cast<const uint32_t*>(w.key)[2 * (ef - 1) + 1];
    g = lower_bound_keys(w.key + ef, ntot - ef, ((uint64_t)wo << 32) | 0xFFFFFFFFull);   // every lane computes the same value
  }
  w.gcnt = g;
  return lanes_admitted;
}

// ---- row gather of the build walk: ALL stages of a batch in flight at once -------------------------------
// The build runs few walks (one warp per pending insert, a window of tens) and each walk is a chain of
// dependent expansions, so shared memory is spent on latency instead of occupancy: the rows of a batch
// are requested in full (NB stages of 512 bytes per row, up to 6 = 3 KB rows) before anything is waited
// for — one memory latency per expansion instead of one per stage. Longer rows go block by block.
template <uint32_t NB>
__device__ __forceinline__ void gather_all_begin(const float* __restrict__ vec, uint32_t pitch, uint32_t row, uint32_t mask,
                                                 unsigned char* stage, uint32_t lane) {
  const uint32_t n_chunks = (pitch * 4 + 511) / 512;
  for (uint32_t ch = 0; ch < min(NB, n_chunks); ++ch) gather_issue<512, NB>(vec, pitch, row, mask, ch, stage, lane);
}
// completes a gather begun for a superset of `mask`; +Inf on the other lanes
template <int METRIC, uint32_t NB>
__device__ __forceinline__ float gather_all_finish(const float* __restrict__ vec, const float* __restrict__ norm, uint32_t pitch,
                                                   const float* sq, float qn, uint32_t row, uint32_t mask, unsigned char* stage,
                                                   uint32_t lane) {
  const uint32_t row_bytes = pitch * 4;
  const uint32_t n_chunks = (row_bytes + 511) / 512;
  const bool valid = (mask >> lane) & 1u;
  float acc = 0.0f;
  const float xn = (METRIC == M_COS && valid) ? __ldg(norm + row) : 0.0f;
  for (uint32_t c0 = 0; c0 < n_chunks; c0 += NB) {
    cp_async_wait<0>();
    __syncwarp();   // every lane's pieces of these stages have landed
    if (valid) {
      for (uint32_t ch = c0; ch < min(c0 + NB, n_chunks); ++ch) {
        const uint32_t n4 = min(512u, row_bytes - ch * 512) / 16;
        const float4* x4 = reinterpret_cast<const float4*>(stage + ((ch % NB) * 32 + lane) * ga_row(512));
        const float4* q4 = reinterpret_cast<const float4*>(sq) + ch * 32;
        if (n4 == 32) {
#pragma unroll
          for (uint32_t i = 0; i < 32; ++i) acc = acc_step4<METRIC>(acc, q4[i], x4[i]);
        } else {
          for (uint32_t i = 0; i < n4; ++i) acc = acc_step4<METRIC>(acc, q4[i], x4[i]);
        }
      }
    }
    __syncwarp();   // the buffers are free again
    if (mask)
      for (uint32_t ch = c0 + NB; ch < min(c0 + 2 * NB, n_chunks); ++ch) gather_issue<512, NB>(vec, pitch, row, mask, ch, stage, lane);
  }
  return valid ? finish_distance<METRIC>(acc, qn, xn) : __int_as_float(0x7f800000);
}

// One warp per pending insert: all of insertVector's searches (hnsw.go:215-226) against the current
// device graph. Output: for every layer lc <= level the sorted result list W of
// searchLayer(vec, eps, efConstruction, lc), and the expansion log of the whole walk.
template <int METRIC, uint32_t NB>
__global__ void __launch_bounds__(32) hnsw_build_search_kernel(BuildArgs a) {
  extern __shared__ __align__(16) unsigned char smem_build[];
  const uint32_t lane = threadIdx.x;
  uint64_t* wkey0 = reinterpret_cast<uint64_t*>(smem_build);
  uint64_t* wkey1 = wkey0 + a.ef_pad;
  uint64_t* snk = wkey1 + a.ef_pad;
  float* sq = reinterpret_cast<float*>(snk + 32);
  uint32_t* wrow0 = reinterpret_cast<uint32_t*>(sq + a.pitch);
  uint32_t* wrow1 = wrow0 + a.ef_pad;
  uint32_t* snr = wrow1 + a.ef_pad;
  uint32_t* eps = snr + 32;
  unsigned char* stage = reinterpret_cast<unsigned char*>(eps + 64);
  uint32_t* hash = a.ghash + (size_t)blockIdx.x * a.hash_size;
  const uint32_t n_groups = a.hash_size >> 2;
  const uint32_t row_bits = a.row_bits;
  const float INF = __int_as_float(0x7f800000);
  uint32_t tag = 0;
  unsigned long long evals = 0, hops = 0;

  const bool search_mode = a.ext_q != nullptr;
  const uint32_t n_slots = (search_mode && a.nq_dev) ? min(*a.nq_dev, a.n_slots) : a.n_slots;
  for (uint32_t slot = blockIdx.x; slot < n_slots; slot += gridDim.x) {
    const uint32_t qi = search_mode ? (a.qlist ? a.qlist[slot] : slot) : slot;
    const uint32_t x = search_mode ? 0u : a.q_rows[slot];
    const int L = search_mode ? 0 : (int)a.q_levels[slot];
    __syncwarp();
    float qn = 0.0f;
    if (search_mode) {
      stage_query(sq, a.ext_q + (size_t)qi * a.dim, a.dim, a.pitch, lane, 32);
      __syncwarp();
      if (METRIC == M_COS) {
        if (lane == 0) qn = exact_norm_padded(sq, a.pitch / 4);
        qn = __shfl_sync(0xffffffffu, qn, 0);
      }
    } else {
      for (uint32_t i = lane; i < a.pitch / 4; i += 32)
        reinterpret_cast<float4*>(sq)[i] = __ldg(reinterpret_cast<const float4*>(a.vec + (size_t)x * a.pitch) + i);
      if (METRIC == M_COS) qn = __ldg(a.norm + x);   // == the sequential norm of the query (store.cu prepare_rows)
    }
    uint4* log = a.log ? a.log + (size_t)slot * BUILD_LOGCAP : nullptr;
    uint32_t n_log = 0;
    bool overflow = false;
    uint32_t n_eps = 0;
    if (a.entry_row != ROW_NONE) {
      if (lane == 0) eps[0] = a.entry_row;
      n_eps = 1;
    }
    __syncwarp();
    const int top = max(a.max_layer, L);   // hnsw.go:205-207: maxLayer is raised before the descent
    for (int lc = top; lc >= 0; --lc) {
      const uint32_t ef = (lc > L) ? 1u : a.efc;              // hnsw.go:219-221 / 224-225
      const uint32_t stride = (lc == 0) ? a.s0 : a.su;
      // ---- a fresh visited set per searchLayer call (hnsw.go:488)
      if (tag == 0 || tag >= a.tag_max) {
        for (uint32_t i = lane; i < n_groups; i += 32) reinterpret_cast<uint4*>(hash)[i] = make_uint4(HASH_EMPTY, HASH_EMPTY, HASH_EMPTY, HASH_EMPTY);
        tag = 1;
      } else {
        ++tag;
      }
      __syncwarp();
      WState w{wkey0, wkey1, wrow0, wrow1, 0u, 0u, 0u};
      const uint32_t cap = 2u * ef;   // W + ghosts (<= a.ef_pad)
      uint32_t seq = 1, visited = 0;
      // ---- entry points (hnsw.go:492-508): evaluate, mark visited, W = C = sorted(entries)
      for (uint32_t e0 = 0; e0 < n_eps && !overflow; e0 += 32) {
        const uint32_t nb = (e0 + lane < n_eps) ? eps[e0 + lane] : ROW_NONE;
        bool ok = (nb != ROW_NONE) && (nb < a.n_rows) && !(a.has_deleted && bit_test(a.deleted, nb));
        ok = visited_insert_warp(hash, n_groups, nb, ok, false, make_uint4(0, 0, 0, 0), lane, tag, row_bits);
        const uint32_t mask = __ballot_sync(0xffffffffu, ok);
        if (!mask) continue;
        visited += __popc(mask);
        evals += __popc(mask);
        gather_all_begin<NB>(a.vec, a.pitch, nb, mask, stage, lane);
        const float d = gather_all_finish<METRIC, NB>(a.vec, a.norm, a.pitch, sq, qn, nb, mask, stage, lane);
        const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
        const uint64_t key = ((uint64_t)f32_ord(d) << 32) | ((uint64_t)(seq + rank) << 1);
        seq += __popc(mask);
        merge_admitted(w, snk, snr, ok, key, nb, max(ef, n_eps), max(cap, n_eps), lane);   // entries all stay (n_eps <= ef, checked on the host)
      }
      w.p_lo = 0;
      uint32_t pre_row = ROW_NONE, pre_nb = ROW_NONE;   // layer 0: adjacency fetched ahead for the runner-up
      // ---- beam (hnsw.go:510-548): expand the closest un-expanded entry of W until there is none
      while (!overflow) {
        const uint32_t tot = w.cnt + w.gcnt;   // un-expanded ghosts are expanded after all of W (they tie with W[ef-1])
        uint32_t mm = 0, b0 = w.p_lo & ~31u;
        for (; b0 < tot; b0 += 32) {
          const uint32_t i = b0 + lane;
          const bool un = (i < tot) && !(reinterpret_cast<const uint32_t*>(w.key)[2 * i] & 1u);
          mm = __ballot_sync(0xffffffffu, un);
          if (mm) break;
        }
        if (!mm) break;
        const uint32_t p = b0 + __ffs(mm) - 1;
        w.p_lo = p + 1;
        const uint32_t cur = w.row[p];
        // the runner-up is the most likely next expansion: its adjacency line is fetched now (layer 0), so
        // that the next expansion does not start with a dependent memory round trip
        const uint32_t m2 = mm & (mm - 1);
        const uint32_t row2 = (lc == 0 && m2) ? w.row[b0 + __ffs(m2) - 1] : ROW_NONE;
        __syncwarp();
        if (lane == 0) reinterpret_cast<uint32_t*>(w.key)[2 * p] |= 1u;
        __syncwarp();
        ++hops;
        if (lc > 0 && (int)a.levels[cur] < lc) continue;   // GetConnections(layer) is empty for good (hnsw.go:45-50)
        if (visited + stride > a.hash_size - (a.hash_size >> 3)) {
          overflow = true;
          break;
        }
        const uint32_t* list = (lc == 0) ? a.adj0 + (size_t)cur * a.s0 : a.adj_up + ((size_t)a.up_off[cur] + (lc - 1)) * a.su;
        uint32_t nb_first = ROW_NONE;
        if (lc == 0) {
          nb_first = (cur == pre_row) ? pre_nb : ((lane < stride) ? __ldg(list + lane) : ROW_NONE);
          pre_row = row2;
          if (row2 != ROW_NONE) pre_nb = (lane < stride) ? __ldg(a.adj0 + (size_t)row2 * a.s0 + lane) : ROW_NONE;
        }
        for (uint32_t c0 = 0; c0 < stride; c0 += 32) {
          const uint32_t nb = (lc == 0 && c0 == 0) ? nb_first : ((c0 + lane < stride) ? __ldg(list + c0 + lane) : ROW_NONE);
          bool ok = (nb != ROW_NONE);
          const uint32_t listed = __ballot_sync(0xffffffffu, ok);
          if (listed) gather_all_begin<NB>(a.vec, a.pitch, nb, listed, stage, lane);   // rows fly while the table is probed
          if (ok && a.has_deleted) ok = !bit_test(a.deleted, nb);   // deleted: not marked visited, not traversed (hnsw.go:527-530)
          ok = visited_insert_warp(hash, n_groups, nb, ok, false, make_uint4(0, 0, 0, 0), lane, tag, row_bits);
          const uint32_t mask = __ballot_sync(0xffffffffu, ok);
          // threshold in force when this part of the list is examined (hnsw.go:536-542)
          const uint32_t worst = (w.cnt >= ef) ? reinterpret_cast<const uint32_t*>(w.key)[2 * (ef - 1) + 1] : LOG_NOT_FULL;
          float d = INF;
          if (listed) d = gather_all_finish<METRIC, NB>(a.vec, a.norm, a.pitch, sq, qn, nb, mask, stage, lane);
          visited += __popc(mask);
          evals += __popc(mask);
          const uint32_t od = f32_ord(d);
          const bool cand = ok && (worst == LOG_NOT_FULL || od < worst);   // the rest cannot be admitted: W[ef-1] only shrinks
          const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
          const uint64_t key = ((uint64_t)od << 32) | ((uint64_t)(seq + rank) << 1);
          seq += __popc(mask);
          const uint32_t mask_in = merge_admitted(w, snk, snr, cand, key, nb, ef, cap, lane);
          if (log && lane == 0 && n_log < BUILD_LOGCAP) log[n_log] = make_uint4(cur, worst, mask_in, (uint32_t)lc | ((c0 >> 5) << 8));
          ++n_log;
        }
      }
      if (overflow) break;
      // ---- result of this layer (hnsw.go:551-556), entry points of the next (hnsw.go:220, 248)
      __syncwarp();
      if (search_mode) {
        if (lc == 0) {   // hnsw.go:317-347: the first min(TopK, |W|) candidates (none is deleted, all are sorted)
          const uint32_t n_out = min(a.k, w.cnt);
          for (uint32_t i = lane; i < a.k; i += 32) {
            a.out_ids[(size_t)qi * a.k + i] = (i < n_out) ? a.ids[w.row[i]] : 0ull;
            a.out_dist[(size_t)qi * a.k + i] = (i < n_out) ? ord_f32((uint32_t)(w.key[i] >> 32)) : INF;
          }
          if (lane == 0 && a.out_counts) a.out_counts[qi] = n_out;
        }
      } else if (lc <= L) {
        const size_t o = ((size_t)a.out_off[slot] + lc) * a.efc;
        for (uint32_t i = lane; i < w.cnt; i += 32) {
          a.w_rows[o + i] = w.row[i];
          a.w_ord[o + i] = (uint32_t)(w.key[i] >> 32);
        }
        if (lane == 0) a.w_cnt[a.out_off[slot] + lc] = w.cnt;
      }
      n_eps = min(w.cnt, (lc > L) ? 1u : stride);
      for (uint32_t i = lane; i < n_eps; i += 32) eps[i] = w.row[i];
      __syncwarp();
    }
    if (search_mode) {
      if (lane == 0 && a.stats) atomicAdd(a.stats + 2, 1ull);   // walks redone exactly
      if (overflow) {   // the largest table overflowed too: no result rather than a truncated beam
        for (uint32_t i = lane; i < a.k; i += 32) {
          a.out_ids[(size_t)qi * a.k + i] = 0ull;
          a.out_dist[(size_t)qi * a.k + i] = INF;
        }
        if (lane == 0) {
          if (a.out_counts) a.out_counts[qi] = 0;
          atomicAdd(a.failed, 1ull);
        }
      }
    } else if (lane == 0) {
      a.log_cnt[slot] = overflow ? LOG_OVERFLOW : n_log;
    }
  }
  if (a.stats && lane == 0 && !search_mode) {
    atomicAdd(a.stats + 0, evals);
    atomicAdd(a.stats + 1, hops);
  }
}

// out[i * n + j] = ord(Distance(row_i, row_j)) for j < i: the distances between the nodes of a window,
// in the reference's order and rounding (one thread per pair)
template <int METRIC>
__global__ void window_pairs_kernel(const float* __restrict__ vec, const float* __restrict__ norm, uint32_t pitch,
                                    const uint32_t* __restrict__ rows, uint32_t n, uint32_t* __restrict__ out) {
  const uint32_t idx = blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n * n) return;
  const uint32_t i = idx / n, j = idx - i * n;
  if (j >= i) return;
  const float4* a4 = reinterpret_cast<const float4*>(vec + (size_t)rows[i] * pitch);
  const float4* b4 = reinterpret_cast<const float4*>(vec + (size_t)rows[j] * pitch);
  float acc = 0.0f;
  for (uint32_t t = 0; t < pitch / 4; ++t) acc = acc_step4<METRIC>(acc, __ldg(a4 + t), __ldg(b4 + t));
  const float d = finish_distance<METRIC>(acc, METRIC == M_COS ? __ldg(norm + rows[i]) : 0.0f, METRIC == M_COS ? __ldg(norm + rows[j]) : 0.0f);
  out[idx] = f32_ord(d);
}

// ord(Distance(row, neighbour)) for every edge of an uploaded graph (one thread per slot)
template <int METRIC>
__global__ void edge_distances_kernel(const float* __restrict__ vec, const float* __restrict__ norm, uint32_t pitch,
                                      const uint32_t* __restrict__ adj, const uint32_t* __restrict__ owner /* optional: row per list */,
                                      uint32_t stride, uint64_t n_slots, uint32_t* __restrict__ out) {
  const uint64_t idx = (uint64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (idx >= n_slots) return;
  const uint32_t nb = adj[idx];
  if (nb == ROW_NONE) {
    out[idx] = 0xFFFFFFFFu;
    return;
  }
  const uint32_t r = owner ? owner[idx / stride] : (uint32_t)(idx / stride);
  const float4* a4 = reinterpret_cast<const float4*>(vec + (size_t)r * pitch);
  const float4* b4 = reinterpret_cast<const float4*>(vec + (size_t)nb * pitch);
  float acc = 0.0f;
  for (uint32_t t = 0; t < pitch / 4; ++t) acc = acc_step4<METRIC>(acc, __ldg(a4 + t), __ldg(b4 + t));
  out[idx] = f32_ord(finish_distance<METRIC>(acc, METRIC == M_COS ? __ldg(norm + r) : 0.0f, METRIC == M_COS ? __ldg(norm + nb) : 0.0f));
}

// dst[off[i] .. off[i] + len) = src[i * 32 .. ): the adjacency lists a round of commits changed
__global__ void apply_lists_kernel(uint32_t* __restrict__ adj0, uint32_t* __restrict__ adj_up, const uint64_t* __restrict__ dst /* bit 63: upper */,
                                   const uint32_t* __restrict__ src, uint32_t n, uint32_t s0, uint32_t su, uint32_t src_stride) {
  const uint32_t i = blockIdx.x * (blockDim.x / 32) + (threadIdx.x >> 5);
  const uint32_t lane = threadIdx.x & 31;
  if (i >= n) return;
  const uint64_t d = dst[i];
  const bool upper = (d >> 63) != 0;
  uint32_t* out = (upper ? adj_up : adj0) + (d & ~(1ull << 63));
  const uint32_t len = upper ? su : s0;
  for (uint32_t t = lane; t < len; t += 32) out[t] = src[(size_t)i * src_stride + t];
}

namespace {

template <class T>
int32_t grow_dev(T** p, uint64_t old_count, uint64_t new_count, int fill, cudaStream_t st) {
  T* np = nullptr;
  SCN_CUDA(cudaMalloc(&np, std::max<uint64_t>(new_count, 1) * sizeof(T)));
  if (*p && old_count) SCN_CUDA(cudaMemcpyAsync(np, *p, old_count * sizeof(T), cudaMemcpyDeviceToDevice, st));
  if (new_count > old_count) SCN_CUDA(cudaMemsetAsync(np + old_count, fill, (new_count - old_count) * sizeof(T), st));
  SCN_CUDA(cudaStreamSynchronize(st));
  if (*p) cudaFree(*p);
  *p = np;
  return SCN_OK;
}

// pinned host buffers of one scn_hnsw_insert call
struct PinnedArena {
  std::vector<void*> p;
  ~PinnedArena() {
    for (void* x : p) cudaFreeHost(x);
  }
  template <class T>
  int32_t alloc(T** out, size_t count) {
    void* x = nullptr;
    if (cudaHostAlloc(&x, std::max<size_t>(count, 1) * sizeof(T), cudaHostAllocDefault) != cudaSuccess) {
      cudaGetLastError();
      return fail(SCN_ERR_RESOURCE, "pinned host allocation failed");
    }
    p.push_back(x);
    *out = static_cast<T*>(x);
    return SCN_OK;
  }
};

struct Change {       // one change of a neighbour list during this round
  uint32_t added;     // row that joined the list, or ROW_NONE
  uint32_t removed;   // row that left the list, or ROW_NONE
  bool reordered;     // the surviving neighbours changed their relative order (first prune of a list)
};

struct RoundLog {
  // key = row * 256 + layer
  std::unordered_map<uint64_t, std::vector<Change>> changes;
  std::unordered_map<uint64_t, std::vector<uint32_t>> snapshot;  // the list as the round's searches saw it
  bool global_changed = false;                                     // entry point or maxLayer moved
  void clear() {
    changes.clear();
    snapshot.clear();
    global_changed = false;
  }
};

inline uint64_t list_key(uint32_t row, uint32_t layer) { return (uint64_t)row * 256 + layer; }

struct Builder {
  scn_store* s;
  BuildState* b;
  uint32_t efc;
  cudaStream_t st;
  RoundLog round;
  std::vector<uint64_t> dirty;   // list keys whose device copy is stale

  uint32_t* adj(uint32_t row, uint32_t layer) {
    return layer == 0 ? &b->adj0[(size_t)row * b->s0] : &b->adju[((size_t)b->up_off[row] + layer - 1) * b->su];
  }
  uint32_t* ord(uint32_t row, uint32_t layer) {
    return layer == 0 ? &b->ord0[(size_t)row * b->s0] : &b->ordu[((size_t)b->up_off[row] + layer - 1) * b->su];
  }
  uint8_t& cnt(uint32_t row, uint32_t layer) { return layer == 0 ? b->cnt0[row] : b->cntu[(size_t)b->up_off[row] + layer - 1]; }
  uint32_t cap(uint32_t layer) const { return layer == 0 ? b->s0 : b->su; }

  int node_layer(uint32_t row) {  // getNodeLayer, hnsw.go:472-484
    for (int l = (int)b->level[row]; l >= 0; --l)
      if (cnt(row, (uint32_t)l) > 0) return l;
    return 0;
  }

  void touch(uint32_t row, uint32_t layer, bool record_snapshot) {
    const uint64_t k = list_key(row, layer);
    if (record_snapshot && !round.snapshot.count(k)) {
      const uint32_t* a = adj(row, layer);
      round.snapshot[k].assign(a, a + cnt(row, layer));
    }
    dirty.push_back(k);
  }

  // AddConnection(layer, nb) on `row` + pruneConnections (hnsw.go:78-89, 586-614). `is_new_node`: the
  // list belongs to the node being inserted (nobody can have expanded it: no change record needed).
  void add_edge(uint32_t row, uint32_t layer, uint32_t nb, uint32_t d_ord, bool is_new_node) {
    if (layer > b->level[row]) return;   // the node has no such layer: silent no-op (hnsw.go:79)
    uint32_t* a = adj(row, layer);
    uint32_t* o = ord(row, layer);
    uint8_t& c = cnt(row, layer);
    for (uint32_t i = 0; i < c; ++i)
      if (a[i] == nb) return;            // already connected (hnsw.go:81-85)
    touch(row, layer, !is_new_node);
    const uint32_t mc = cap(layer);
    if (c < mc) {
      a[c] = nb;
      o[c] = d_ord;
      ++c;
      if (!is_new_node) round.changes[list_key(row, layer)].push_back({nb, ROW_NONE, false});
      return;
    }
    // the list would hold maxConn + 1 entries: keep the closest maxConn live ones, stable by distance
    // (hnsw.go:594-613; ties keep list order, the new edge is last)
    struct E {
      uint32_t row, ord;
    };
    E all[65];
    uint32_t n_all = 0;
    for (uint32_t i = 0; i < c; ++i)
      if (!b->deleted[a[i]]) all[n_all++] = {a[i], o[i]};
    if (!b->deleted[nb]) all[n_all++] = {nb, d_ord};
    std::stable_sort(all, all + n_all, [](const E& x, const E& y) { return x.ord < y.ord; });
    const uint32_t keep = std::min<uint32_t>(mc, n_all);
    if (!is_new_node) {
      // set difference old -> new
      std::vector<Change>& ch = round.changes[list_key(row, layer)];
      bool nb_kept = false;
      for (uint32_t i = 0; i < keep; ++i) nb_kept |= (all[i].row == nb);
      for (uint32_t i = 0; i < c; ++i) {
        bool kept = false;
        for (uint32_t j = 0; j < keep && !kept; ++j) kept = all[j].row == a[i];
        if (!kept) ch.push_back({ROW_NONE, a[i], false});
      }
      if (nb_kept) ch.push_back({nb, ROW_NONE, false});
      // did the survivors keep their relative order? (the order in which a walk meets the neighbours
      // decides ties between equal distances)
      uint32_t last = 0;
      bool reordered = false;
      for (uint32_t j = 0; j < keep && !reordered; ++j) {
        if (all[j].row == nb) continue;
        uint32_t pos = 0;
        while (pos < c && a[pos] != all[j].row) ++pos;
        reordered = pos < last;
        last = pos;
      }
      if (reordered) ch.push_back({ROW_NONE, ROW_NONE, true});
    }
    for (uint32_t i = 0; i < mc; ++i) {
      a[i] = (i < keep) ? all[i].row : ROW_NONE;
      o[i] = (i < keep) ? all[i].ord : 0xFFFFFFFFu;
    }
    c = (uint8_t)keep;
  }
};

}  // namespace

// Ensures the store carries a host mirror of its graph (edge distances computed on the device) and
// device adjacency arrays with room for `rows` rows / `lists` upper lists.
static int32_t prepare_build_state(scn_store* s, int32_t m, cudaStream_t st) {
  if (s->build && s->build->n_nodes == 0 && !s->has_graph) free_build_state(s);   // nothing built yet: M is still free
  if (s->build) {
    if ((int32_t)s->build->m != m) return fail(SCN_ERR_INVALID_PARAMETERS, "the graph was built with M=%u, not %d", s->build->m, m);
    return SCN_OK;
  }
  BuildState* b = new BuildState();
  b->m = (uint32_t)m;
  b->s0 = 2 * (uint32_t)m;
  b->su = (uint32_t)m;
  if (s->has_graph) {
    if (s->m != m) {
      delete b;
      return fail(SCN_ERR_INVALID_PARAMETERS, "the uploaded graph has M=%d, not %d", s->m, m);
    }
    // mirror the uploaded graph; every edge gets its distance from the device
    const uint64_t n = s->graph_nodes;
    b->n_nodes = n;
    b->level.resize(n);
    b->up_off.resize(n);
    b->adj0.resize(n * b->s0);
    b->ord0.resize(n * b->s0);
    b->cnt0.assign(n, 0);
    b->upper_lists = s->upper_lists;
    b->adju.resize(b->upper_lists * b->su);
    b->ordu.resize(b->upper_lists * b->su);
    b->cntu.assign(b->upper_lists, 0);
    b->cap_rows = n;
    b->cap_lists = b->upper_lists;
    auto bail = [&](int32_t rc) {
      delete b;
      return rc;
    };
    if (n) {
      if (cudaMemcpy(b->level.data(), s->d_levels, n, cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(b->up_off.data(), s->d_up_off, n * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
          cudaMemcpy(b->adj0.data(), s->d_adj0, n * b->s0 * 4, cudaMemcpyDeviceToHost) != cudaSuccess ||
          (b->upper_lists && cudaMemcpy(b->adju.data(), s->d_adj_up, b->upper_lists * b->su * 4, cudaMemcpyDeviceToHost) != cudaSuccess))
        return bail(cuda_fail(cudaGetLastError(), "graph download", __FILE__, __LINE__));
      Scratch scratch(st);
      uint32_t *d_o0 = nullptr, *d_ou = nullptr, *d_owner = nullptr;
      if (scratch.alloc(&d_o0, n * b->s0) != SCN_OK || scratch.alloc(&d_ou, std::max<uint64_t>(b->upper_lists * b->su, 1)) != SCN_OK ||
          scratch.alloc(&d_owner, std::max<uint64_t>(b->upper_lists, 1)) != SCN_OK)
        return bail(SCN_ERR_RESOURCE);
      std::vector<uint32_t> owner(b->upper_lists);
      for (uint64_t r = 0; r < n; ++r)
        for (uint32_t l = 1; l <= b->level[r]; ++l) owner[(size_t)b->up_off[r] + l - 1] = (uint32_t)r;
      if (b->upper_lists) cudaMemcpyAsync(d_owner, owner.data(), b->upper_lists * 4, cudaMemcpyHostToDevice, st);
#define EDGE(MT)                                                                                                                   \
  do {                                                                                                                             \
    edge_distances_kernel<MT><<<(unsigned)((n * b->s0 + 255) / 256), 256, 0, st>>>(s->d_vec, s->d_norm, s->pitch, s->d_adj0, nullptr, \
                                                                                 b->s0, n * b->s0, d_o0);                          \
    if (b->upper_lists)                                                                                                            \
      edge_distances_kernel<MT><<<(unsigned)((b->upper_lists * b->su + 255) / 256), 256, 0, st>>>(                                  \
          s->d_vec, s->d_norm, s->pitch, s->d_adj_up, d_owner, b->su, b->upper_lists * b->su, d_ou);                                \
  } while (0)
      switch (s->metric) {
        case M_L2: EDGE(M_L2); break;
        case M_COS: EDGE(M_COS); break;
        default: EDGE(M_IP); break;
      }
#undef EDGE
      count_launch(2);
      cudaMemcpyAsync(b->ord0.data(), d_o0, n * b->s0 * 4, cudaMemcpyDeviceToHost, st);
      if (b->upper_lists) cudaMemcpyAsync(b->ordu.data(), d_ou, b->upper_lists * b->su * 4, cudaMemcpyDeviceToHost, st);
      if (cudaStreamSynchronize(st) != cudaSuccess) return bail(cuda_fail(cudaGetLastError(), "edge distances", __FILE__, __LINE__));
      for (uint64_t r = 0; r < n; ++r) {
        uint32_t c = 0;
        while (c < b->s0 && b->adj0[r * b->s0 + c] != ROW_NONE) ++c;
        b->cnt0[r] = (uint8_t)c;
      }
      for (uint64_t l = 0; l < b->upper_lists; ++l) {
        uint32_t c = 0;
        while (c < b->su && b->adju[l * b->su + c] != ROW_NONE) ++c;
        b->cntu[l] = (uint8_t)c;
      }
    }
  }
  s->build = b;
  return SCN_OK;
}

static int32_t ensure_graph_capacity(scn_store* s, BuildState* b, uint64_t rows, uint64_t lists, cudaStream_t st) {
  if (rows > b->cap_rows) {
    const uint64_t nc = std::max<uint64_t>(rows, b->cap_rows + b->cap_rows / 2);
    SCN_TRY(grow_dev(&s->d_adj0, b->cap_rows * b->s0, nc * b->s0, 0xFF, st));
    SCN_TRY(grow_dev(&s->d_levels, b->cap_rows, nc, 0, st));
    SCN_TRY(grow_dev(&s->d_up_off, b->cap_rows, nc, 0, st));
    b->cap_rows = nc;
  }
  if (lists > b->cap_lists || !s->d_adj_up) {
    const uint64_t nc = std::max<uint64_t>(std::max<uint64_t>(lists, 1), b->cap_lists + b->cap_lists / 2);
    SCN_TRY(grow_dev(&s->d_adj_up, b->cap_lists * b->su, nc * b->su, 0xFF, st));
    b->cap_lists = nc;
  }
  return SCN_OK;
}

// nb = resident 512-byte stages per row (1, 2, 4 or 6). `launch` = false: only sets the shared-memory
// limit and reports how many walks fit an SM.
template <int METRIC, uint32_t NB>
static int32_t build_search_nb(const BuildArgs& a, int grid, size_t smem, cudaStream_t st, bool launch, int* per_sm) {
  SCN_ALLOW_SMEM((hnsw_build_search_kernel<METRIC, NB>), smem);
  if (!launch) {
    SCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(per_sm, hnsw_build_search_kernel<METRIC, NB>, 32, smem));
    return SCN_OK;
  }
  hnsw_build_search_kernel<METRIC, NB><<<grid, 32, smem, st>>>(a);
  SCN_LAUNCHED();
  return SCN_OK;
}
template <int METRIC>
static int32_t build_search_metric(const BuildArgs& a, uint32_t nb, int grid, size_t smem, cudaStream_t st, bool launch, int* per_sm) {
  switch (nb) {
    case 1: return build_search_nb<METRIC, 1>(a, grid, smem, st, launch, per_sm);
    case 2: return build_search_nb<METRIC, 2>(a, grid, smem, st, launch, per_sm);
    case 4: return build_search_nb<METRIC, 4>(a, grid, smem, st, launch, per_sm);
    default: return build_search_nb<METRIC, 6>(a, grid, smem, st, launch, per_sm);
  }
}
static int32_t build_search(int metric, const BuildArgs& a, uint32_t nb, int grid, size_t smem, cudaStream_t st, bool launch, int* per_sm) {
  switch (metric) {
    case M_L2: return build_search_metric<M_L2>(a, nb, grid, smem, st, launch, per_sm);
    case M_COS: return build_search_metric<M_COS>(a, nb, grid, smem, st, launch, per_sm);
    default: return build_search_metric<M_IP>(a, nb, grid, smem, st, launch, per_sm);
  }
}

// The exact walk for the queries listed in d_qlist[0 .. *d_nq_dev): HNSW.Search with the reference's tie rules
// (see WState). Always enqueued behind hnsw_search_kernel; exits at once when the list is empty.
int32_t hnsw_search_exact(scn_store* s, const float* d_q, const uint32_t* d_qlist, const uint32_t* d_nq_dev, uint64_t nq, uint32_t k,
                          uint32_t ef, uint32_t hash_size, uint64_t* d_out_ids, float* d_out_dist, uint32_t* d_out_counts,
                          unsigned long long* d_failed, cudaStream_t st, Scratch& scratch) {
  int sms = 0;
  SCN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
  const uint32_t ef_pad = std::max(64u, round_up(2 * ef, 32));
  const uint32_t n_chunks = (s->pitch * 4 + 511) / 512;
  const uint32_t nb = n_chunks <= 1 ? 1u : n_chunks <= 2 ? 2u : n_chunks <= 4 ? 4u : 6u;
  const size_t smem = build_warp_bytes(s->pitch, ef_pad, ga_stage_bytes(512, nb));
  if (smem > 200 * 1024) return fail(SCN_ERR_INVALID_PARAMETERS, "ef=%u / dim=%u need more shared memory than one SM has", ef, s->dim);
  int per_sm = 0;
  SCN_TRY(build_search(s->metric, BuildArgs{}, nb, 0, smem, st, false, &per_sm));
  if (per_sm < 1) return fail(SCN_ERR_INVALID_PARAMETERS, "the exact walk kernel does not fit an SM");
  // the tables live in global memory: bound their total (flagged queries are few)
  const uint64_t max_blocks = std::max<uint64_t>(1, ((uint64_t)512 << 20) / ((uint64_t)hash_size * 4));
  const int grid = (int)std::min<uint64_t>(std::min<uint64_t>((uint64_t)sms * per_sm, max_blocks), nq);
  BuildArgs a{};
  a.vec = s->d_vec;
  a.norm = s->d_norm;
  a.deleted = s->d_deleted;
  a.adj0 = s->d_adj0;
  a.adj_up = s->d_adj_up;
  a.levels = s->d_levels;
  a.up_off = s->d_up_off;
  a.pitch = s->pitch;
  a.n_rows = (uint32_t)s->rows;
  a.s0 = 2 * (uint32_t)s->m;
  a.su = (uint32_t)s->m;
  a.has_deleted = (s->live != s->rows) ? 1u : 0u;
  a.entry_row = s->entry_row;
  a.max_layer = s->max_layer;
  a.n_slots = (uint32_t)nq;
  a.efc = ef;
  a.ef_pad = ef_pad;
  SCN_TRY(scratch.alloc(&a.ghash, (size_t)grid * hash_size));
  a.hash_size = hash_size;
  a.row_bits = 1;
  while ((1ull << a.row_bits) <= (uint64_t)s->rows) ++a.row_bits;
  a.tag_max = (uint32_t)((1ull << (32 - a.row_bits)) - 1);
  a.ext_q = d_q;
  a.qlist = d_qlist;
  a.nq_dev = d_nq_dev;
  a.ids = s->d_ids;
  a.dim = s->dim;
  a.k = k;
  a.out_ids = d_out_ids;
  a.out_dist = d_out_dist;
  a.out_counts = d_out_counts;
  a.failed = d_failed;
  a.stats = s->opt_profile ? s->d_counters : nullptr;
  return build_search(s->metric, a, nb, grid, smem, st, true, nullptr);
}

}  // namespace scn

using namespace scn;

extern "C" {

int32_t scn_hnsw_insert(scn_store* s, uint64_t n, const int32_t* levels, int32_t m, int32_t ef_construction, scn_build_stats* stats) {
  if (!s) return fail(SCN_ERR_INVALID_PARAMETERS, "store is NULL");
  if (stats) std::memset(stats, 0, sizeof *stats);
  if (n == 0) return SCN_OK;
  if (!levels) return fail(SCN_ERR_INVALID_PARAMETERS, "levels pointer is NULL");
  if (m < 1 || m > 32) return fail(SCN_ERR_INVALID_PARAMETERS, "the GPU-assisted build supports M in [1, 32]");
  if (ef_construction < 2 * m || ef_construction > 1024)
    return fail(SCN_ERR_INVALID_PARAMETERS, "the GPU-assisted build needs 2*M <= efConstruction <= 1024");
  DeviceGuard g(s->device);
  cudaStream_t st = thread_stream(s->device);
  SCN_CUDA(cudaDeviceSynchronize());
  const auto t_start = std::chrono::steady_clock::now();
  SCN_TRY(prepare_build_state(s, m, st));
  BuildState* b = s->build;
  const uint64_t first = b->n_nodes;
  if (first + n > s->rows)
    return fail(SCN_ERR_INVALID_PARAMETERS, "%llu rows are to be inserted but only %llu rows of the store are not in the graph yet",
                (unsigned long long)n, (unsigned long long)(s->rows - first));
  if (s->rows >= (1ull << 31)) return fail(SCN_ERR_INVALID_PARAMETERS, "HNSW supports at most 2^31 - 1 rows per device");
  for (uint64_t i = 0; i < n; ++i)
    if (levels[i] < 0 || levels[i] > 254) return fail(SCN_ERR_INVALID_PARAMETERS, "level %d of node %llu is out of range", levels[i], (unsigned long long)i);

  // ---- host mirror and device arrays for the new nodes (empty lists) --------------------------------
  const uint64_t total = first + n;
  b->level.resize(total);
  b->up_off.resize(total);
  uint64_t lists = b->upper_lists;
  for (uint64_t i = 0; i < n; ++i) {
    b->level[first + i] = (uint8_t)levels[i];
    b->up_off[first + i] = (uint32_t)lists;
    lists += (uint64_t)levels[i];
  }
  if (lists >= 0xFFFFFFFFull) return fail(SCN_ERR_RESOURCE, "too many upper-layer lists");
  b->adj0.resize(total * b->s0, ROW_NONE);
  b->ord0.resize(total * b->s0, 0xFFFFFFFFu);
  b->cnt0.resize(total, 0);
  b->adju.resize(lists * b->su, ROW_NONE);
  b->ordu.resize(lists * b->su, 0xFFFFFFFFu);
  b->cntu.resize(lists, 0);
  b->upper_lists = lists;
  b->deleted.assign(total, 0);
  if (s->live != s->rows) {
    const size_t words = (s->rows + 31) / 32;
    std::vector<uint32_t> bits(words);
    SCN_CUDA(cudaMemcpy(bits.data(), s->d_deleted, words * 4, cudaMemcpyDeviceToHost));
    for (uint64_t r = 0; r < total; ++r) b->deleted[r] = (bits[r >> 5] >> (r & 31)) & 1u;
  }
  SCN_TRY(ensure_graph_capacity(s, b, total, lists, st));
  SCN_CUDA(cudaMemcpyAsync(s->d_levels + first, b->level.data() + first, n, cudaMemcpyHostToDevice, st));
  SCN_CUDA(cudaMemcpyAsync(s->d_up_off + first, b->up_off.data() + first, n * 4, cudaMemcpyHostToDevice, st));
  SCN_CUDA(cudaStreamSynchronize(st));
  uint32_t entry_row = s->has_graph ? s->entry_row : ROW_NONE;
  int32_t max_layer = s->has_graph ? s->max_layer : -1;

  // ---- per-round device buffers --------------------------------------------------------------------
  int sms = 0;
  SCN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
  const uint32_t efc = (uint32_t)ef_construction;
  const uint32_t ef_pad = std::max(64u, round_up(2 * efc, 32));   // W + ghosts (see WState)
  const uint32_t n_chunks = (s->pitch * 4 + 511) / 512;
  const uint32_t nb = n_chunks <= 1 ? 1u : n_chunks <= 2 ? 2u : n_chunks <= 4 ? 4u : 6u;   // resident stages per row
  const size_t smem = build_warp_bytes(s->pitch, ef_pad, ga_stage_bytes(512, nb));
  if (smem > 200 * 1024) return fail(SCN_ERR_INVALID_PARAMETERS, "efConstruction=%u / dim=%u need more shared memory than one SM has", efc, s->dim);
  int per_sm = 0;
  SCN_TRY(build_search(s->metric, BuildArgs{}, nb, 0, smem, st, false, &per_sm));
  if (per_sm < 1) return fail(SCN_ERR_INVALID_PARAMETERS, "the build search kernel does not fit an SM");
  const uint32_t max_window = (uint32_t)std::min<uint64_t>((uint64_t)sms * per_sm, 2048);   // one wave of walks
  // visited table per resident walk: ~2M*1.25 rows per expansion, ~1.5 efc expansions, under 7/8 full
  uint32_t hash_size = (uint32_t)std::min<uint64_t>(next_pow2((uint32_t)std::min<uint64_t>((uint64_t)efc * b->s0 * 4, 1u << 22)), 1u << 22);
  hash_size = std::max(hash_size, 4096u);
  const uint32_t max_out_lists = max_window * 2 + 64;   // lists of a window: sum(level + 1), bounded per round below
  Scratch scratch(st);
  uint32_t *d_qrows = nullptr, *d_outoff = nullptr, *d_ghash = nullptr, *d_wrows = nullptr, *d_word = nullptr, *d_wcnt = nullptr,
           *d_logcnt = nullptr, *d_pairs = nullptr, *d_upd_src = nullptr;
  uint8_t* d_qlevels = nullptr;
  uint4* d_log = nullptr;
  uint64_t* d_upd_dst = nullptr;
  const uint32_t pair_cap = 256;                          // pair matrix for the first pair_cap slots of a window
  const uint32_t upd_cap = 1u << 16;                      // changed lists pushed per apply call
  SCN_TRY(scratch.alloc(&d_qrows, max_window));
  SCN_TRY(scratch.alloc(&d_qlevels, max_window));
  SCN_TRY(scratch.alloc(&d_outoff, max_window));
  SCN_TRY(scratch.alloc(&d_ghash, (size_t)max_window * hash_size));
  SCN_TRY(scratch.alloc(&d_wrows, (size_t)max_out_lists * efc));
  SCN_TRY(scratch.alloc(&d_word, (size_t)max_out_lists * efc));
  SCN_TRY(scratch.alloc(&d_wcnt, max_out_lists));
  SCN_TRY(scratch.alloc(&d_log, (size_t)max_window * BUILD_LOGCAP));
  SCN_TRY(scratch.alloc(&d_logcnt, max_window));
  SCN_TRY(scratch.alloc(&d_pairs, (size_t)pair_cap * pair_cap));
  SCN_TRY(scratch.alloc(&d_upd_src, (size_t)upd_cap * 64));
  SCN_TRY(scratch.alloc(&d_upd_dst, upd_cap));
  // pinned host mirrors of the round outputs
  uint32_t *h_wrows = nullptr, *h_word = nullptr, *h_wcnt = nullptr, *h_logcnt = nullptr, *h_pairs = nullptr, *h_upd_src = nullptr;
  uint4* h_log = nullptr;
  uint64_t* h_upd_dst = nullptr;
  PinnedArena pinned;
  SCN_TRY(pinned.alloc(&h_wrows, (size_t)max_out_lists * efc));
  SCN_TRY(pinned.alloc(&h_word, (size_t)max_out_lists * efc));
  SCN_TRY(pinned.alloc(&h_wcnt, max_out_lists));
  SCN_TRY(pinned.alloc(&h_log, (size_t)max_window * BUILD_LOGCAP));
  SCN_TRY(pinned.alloc(&h_logcnt, max_window));
  SCN_TRY(pinned.alloc(&h_pairs, (size_t)pair_cap * pair_cap));
  SCN_TRY(pinned.alloc(&h_upd_src, (size_t)upd_cap * 64));
  SCN_TRY(pinned.alloc(&h_upd_dst, upd_cap));
  std::vector<uint32_t> h_qrows(max_window), h_outoff(max_window);
  std::vector<uint8_t> h_qlevels(max_window);

  Builder B{s, b, efc, st, {}, {}};
  uint32_t row_bits = 1;
  while ((1ull << row_bits) <= (uint64_t)s->rows) ++row_bits;
  const uint32_t tag_max = (uint32_t)((1ull << (32 - row_bits)) - 1);
  SCN_CUDA(cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(unsigned long long), st));

  uint64_t done = 0, rounds = 0, searched = 0, conflicts = 0, overflowed = 0;
  uint64_t why[6] = {0, 0, 0, 0, 0, 0};   // what ended the rounds (scn_build_stats.conflict_kind)
  double t_device = 0.0, t_commit = 0.0;
  double avg_commits = 4.0;
  uint32_t big_hash_rounds = 0;

  auto push_dirty = [&]() -> int32_t {
    // device copies of the lists the commits changed (one record per list, last state wins)
    std::sort(B.dirty.begin(), B.dirty.end());
    B.dirty.erase(std::unique(B.dirty.begin(), B.dirty.end()), B.dirty.end());
    size_t i = 0;
    while (i < B.dirty.size()) {
      const uint32_t cnt = (uint32_t)std::min<size_t>(upd_cap, B.dirty.size() - i);
      for (uint32_t t = 0; t < cnt; ++t) {
        const uint64_t k = B.dirty[i + t];
        const uint32_t row = (uint32_t)(k >> 8), layer = (uint32_t)(k & 255);
        const uint32_t* a = B.adj(row, layer);
        const uint32_t len = B.cap(layer);
        std::memcpy(h_upd_src + (size_t)t * 64, a, len * 4);
        h_upd_dst[t] = layer == 0 ? (uint64_t)row * b->s0 : ((1ull << 63) | (((uint64_t)b->up_off[row] + layer - 1) * b->su));
      }
      SCN_CUDA(cudaMemcpyAsync(d_upd_src, h_upd_src, (size_t)cnt * 64 * 4, cudaMemcpyHostToDevice, st));
      SCN_CUDA(cudaMemcpyAsync(d_upd_dst, h_upd_dst, (size_t)cnt * 8, cudaMemcpyHostToDevice, st));
      apply_lists_kernel<<<(cnt + 3) / 4, 128, 0, st>>>(s->d_adj0, s->d_adj_up, d_upd_dst, d_upd_src, cnt, b->s0, b->su, 64);
      SCN_LAUNCHED();
      SCN_CUDA(cudaStreamSynchronize(st));   // the pinned staging is reused
      i += cnt;
    }
    B.dirty.clear();
    return SCN_OK;
  };

  while (done < n) {
    // ---- the first node of an empty index becomes the entry point (hnsw.go:210-213): no search -----
    if (entry_row == ROW_NONE) {
      const uint32_t x = (uint32_t)(first + done);
      if ((int)b->level[x] > max_layer) max_layer = b->level[x];
      entry_row = x;
      ++done;
      b->n_nodes = first + done;
      continue;
    }
    // ---- window of the next inserts, all searched against the graph as it is now ---------------------
    // a round lasts as long as its slowest walk (a level-L node runs L + 1 searches): speculate about as far
    // as the rounds have been getting, not much further
    uint32_t want = (uint32_t)std::min<double>(max_window, 1.5 * avg_commits + 4.0);
    if (s->opt_build_window > 0) want = (uint32_t)std::min<int64_t>(s->opt_build_window, max_window);   // 1 = no speculation at all
    uint32_t W = (uint32_t)std::min<uint64_t>(want, n - done);
    uint32_t n_lists = 0;
    for (uint32_t i = 0; i < W; ++i) {
      const uint32_t x = (uint32_t)(first + done + i);
      if (n_lists + b->level[x] + 1 > max_out_lists) {
        W = i;
        break;
      }
      h_qrows[i] = x;
      h_qlevels[i] = b->level[x];
      h_outoff[i] = n_lists;
      n_lists += b->level[x] + 1u;
    }
    if (W == 0) return fail(SCN_ERR_INTERNAL, "a node has more layers than the build buffers hold");
    const bool big = big_hash_rounds > 0;
    if (big) {
      W = 1;   // the walk that overflowed its table, alone, with the largest table there is room for
      n_lists = b->level[h_qrows[0]] + 1u;
      --big_hash_rounds;
    }
    const auto t_round = std::chrono::steady_clock::now();
    SCN_CUDA(cudaMemcpyAsync(d_qrows, h_qrows.data(), W * 4, cudaMemcpyHostToDevice, st));
    SCN_CUDA(cudaMemcpyAsync(d_qlevels, h_qlevels.data(), W, cudaMemcpyHostToDevice, st));
    SCN_CUDA(cudaMemcpyAsync(d_outoff, h_outoff.data(), W * 4, cudaMemcpyHostToDevice, st));
    BuildArgs a{};
    a.vec = s->d_vec;
    a.norm = s->d_norm;
    a.deleted = s->d_deleted;
    a.adj0 = s->d_adj0;
    a.adj_up = s->d_adj_up;
    a.levels = s->d_levels;
    a.up_off = s->d_up_off;
    a.pitch = s->pitch;
    a.n_rows = (uint32_t)s->rows;
    a.s0 = b->s0;
    a.su = b->su;
    a.has_deleted = (s->live != s->rows) ? 1u : 0u;
    a.entry_row = entry_row;
    a.max_layer = max_layer;
    a.q_rows = d_qrows;
    a.q_levels = d_qlevels;
    a.out_off = d_outoff;
    a.n_slots = W;
    a.efc = efc;
    a.ef_pad = ef_pad;
    a.ghash = d_ghash;
    a.hash_size = big ? (uint32_t)std::min<uint64_t>(std::max<uint64_t>(next_pow2((uint32_t)s->rows) * 2ull, hash_size),
                                                     std::min<uint64_t>((uint64_t)max_window * hash_size, 1u << 30))
                      : hash_size;
    a.row_bits = row_bits;
    a.tag_max = tag_max;
    a.w_rows = d_wrows;
    a.w_ord = d_word;
    a.w_cnt = d_wcnt;
    a.log = d_log;
    a.log_cnt = d_logcnt;
    a.stats = s->d_counters;
    SCN_TRY(build_search(s->metric, a, nb, (int)W, smem, st, true, nullptr));
    const uint32_t P = std::min(W, pair_cap);
    if (P > 1) {
      switch (s->metric) {
        case M_L2: window_pairs_kernel<M_L2><<<(P * P + 127) / 128, 128, 0, st>>>(s->d_vec, s->d_norm, s->pitch, d_qrows, P, d_pairs); break;
        case M_COS: window_pairs_kernel<M_COS><<<(P * P + 127) / 128, 128, 0, st>>>(s->d_vec, s->d_norm, s->pitch, d_qrows, P, d_pairs); break;
        default: window_pairs_kernel<M_IP><<<(P * P + 127) / 128, 128, 0, st>>>(s->d_vec, s->d_norm, s->pitch, d_qrows, P, d_pairs); break;
      }
      SCN_LAUNCHED();
      SCN_CUDA(cudaMemcpyAsync(h_pairs, d_pairs, (size_t)P * P * 4, cudaMemcpyDeviceToHost, st));
    }
    SCN_CUDA(cudaMemcpyAsync(h_wcnt, d_wcnt, n_lists * 4, cudaMemcpyDeviceToHost, st));
    SCN_CUDA(cudaMemcpyAsync(h_wrows, d_wrows, (size_t)n_lists * efc * 4, cudaMemcpyDeviceToHost, st));
    SCN_CUDA(cudaMemcpyAsync(h_word, d_word, (size_t)n_lists * efc * 4, cudaMemcpyDeviceToHost, st));
    SCN_CUDA(cudaMemcpyAsync(h_logcnt, d_logcnt, W * 4, cudaMemcpyDeviceToHost, st));
    SCN_CUDA(cudaMemcpyAsync(h_log, d_log, (size_t)W * BUILD_LOGCAP * sizeof(uint4), cudaMemcpyDeviceToHost, st));
    SCN_CUDA(cudaStreamSynchronize(st));
    ++rounds;
    searched += W;
    const auto t_searched = std::chrono::steady_clock::now();
    t_device += std::chrono::duration<double>(t_searched - t_round).count();

    // ---- commit in order while the speculative searches are provably the serial ones -----------------
    B.round.clear();
    uint32_t committed = 0;
    for (uint32_t i = 0; i < W; ++i) {
      const uint32_t x = h_qrows[i];
      const uint32_t L = b->level[x];
      if (h_logcnt[i] == LOG_OVERFLOW) {
        if (i == 0) {
          if (big) return fail(SCN_ERR_INDEX_BUILD_FAILED, "the walk of row %u visits more rows than the largest visited table holds", x);
          ++overflowed;
          big_hash_rounds = 1;
        }
        break;
      }
      // -- validation against the changes of this round
      bool valid = !B.round.global_changed;
      int reason = 3;
      if (valid && !B.round.changes.empty()) {
        if (h_logcnt[i] > BUILD_LOGCAP || i >= P) {
          valid = false;   // the log is incomplete / no pair distances: usable only on an unchanged graph
          reason = (i >= P) ? 5 : 4;
        } else {
          const uint4* lg = h_log + (size_t)i * BUILD_LOGCAP;
          for (uint32_t r = 0; r < h_logcnt[i] && valid; ++r) {
            const uint32_t row = lg[r].x, worst = lg[r].y, amask = lg[r].z, layer = lg[r].w & 255u, chunk = lg[r].w >> 8;
            auto it = B.round.changes.find(list_key(row, layer));
            if (it == B.round.changes.end()) continue;
            const std::vector<uint32_t>& snap = B.round.snapshot[list_key(row, layer)];
            for (const Change& c : it->second) {
              if (c.reordered && amask != 0) {
                valid = false;   // the walk admitted neighbours of a list whose order has changed since
                reason = 2;
                break;
              }
              if (c.added != ROW_NONE) {
                // a node committed earlier in this round: slot index = its row - first row of the window
                const uint32_t j = c.added - h_qrows[0];
                if (j >= i || worst == LOG_NOT_FULL || h_pairs[(size_t)i * P + j] < worst) {
                  valid = false;
                  reason = (worst == LOG_NOT_FULL) ? 0 : 1;
                  break;
                }
              }
              if (c.removed != ROW_NONE) {
                uint32_t pos = ROW_NONE;
                for (uint32_t t = 0; t < snap.size(); ++t)
                  if (snap[t] == c.removed) pos = t;
                if (pos == ROW_NONE) {
                  // added and removed again within this round: it matters only if it would have been admitted
                  const uint32_t j = c.removed - h_qrows[0];
                  if (j >= i || worst == LOG_NOT_FULL || h_pairs[(size_t)i * P + j] < worst) {
                    valid = false;
                    reason = (worst == LOG_NOT_FULL) ? 0 : 1;
                    break;
                  }
                } else if ((pos >> 5) == chunk && ((amask >> (pos & 31)) & 1u)) {
                  valid = false;   // the walk admitted a neighbour that is no longer in this list
                  reason = 2;
                  break;
                }
              }
            }
          }
        }
      }
      if (!valid) {
        ++conflicts;
        ++why[reason];
        break;
      }
      // -- insertVector's link phase (hnsw.go:224-254) with the lists the device search returned
      if ((int)L > max_layer) {
        max_layer = (int)L;
        B.round.global_changed = true;
      }
      for (int lc = (int)L; lc >= 0; --lc) {
        const uint32_t li = h_outoff[i] + (uint32_t)lc;
        const uint32_t n_cand = h_wcnt[li];
        const uint32_t mc = B.cap((uint32_t)lc);
        const uint32_t n_sel = std::min(n_cand, mc);   // selectNeighbors: W is sorted by (distance, admission) already
        const uint32_t* cr = h_wrows + (size_t)li * efc;
        const uint32_t* co = h_word + (size_t)li * efc;
        for (uint32_t t = 0; t < n_sel; ++t) {
          B.add_edge(x, (uint32_t)lc, cr[t], co[t], true);
          B.add_edge(cr[t], (uint32_t)lc, x, co[t], false);
        }
      }
      if ((int)L > B.node_layer(entry_row)) {   // hnsw.go:252-254
        entry_row = x;
        B.round.global_changed = true;
      }
      ++committed;
    }
    done += committed;
    b->n_nodes = first + done;
    avg_commits = 0.75 * avg_commits + 0.25 * (double)committed;
    SCN_TRY(push_dirty());
    t_commit += std::chrono::duration<double>(std::chrono::steady_clock::now() - t_searched).count();
  }

  // ---- the store's graph fields --------------------------------------------------------------------
  s->has_graph = true;
  s->m = m;
  s->max_layer = max_layer;
  s->entry_row = entry_row;
  s->entry_id = 0;
  if (entry_row != ROW_NONE) SCN_CUDA(cudaMemcpy(&s->entry_id, s->d_ids + entry_row, sizeof(uint64_t), cudaMemcpyDeviceToHost));
  s->graph_nodes = b->n_nodes;
  s->upper_lists = b->upper_lists;
  uint64_t edges = 0;
  s->h_node_layer.assign(b->n_nodes, 0);
  for (uint64_t r = 0; r < b->n_nodes; ++r) {
    edges += b->cnt0[r];
    for (uint32_t l = 1; l <= b->level[r]; ++l) edges += b->cntu[(size_t)b->up_off[r] + l - 1];
    s->h_node_layer[r] = (uint8_t)B.node_layer((uint32_t)r);
  }
  s->graph_edges = edges;
  if (stats) {
    unsigned long long h[4] = {0, 0, 0, 0};
    SCN_CUDA(cudaMemcpy(h, s->d_counters, sizeof h, cudaMemcpyDeviceToHost));
    stats->inserted = n;
    stats->rounds = rounds;
    stats->searches = searched;
    stats->conflicts = conflicts;
    stats->table_overflows = overflowed;
    stats->distance_evals = h[0];
    stats->expansions = h[1];
    for (int i = 0; i < 6; ++i) stats->conflict_kind[i] = why[i];
    stats->seconds = std::chrono::duration<double>(std::chrono::steady_clock::now() - t_start).count();
    stats->device_seconds = t_device;
    stats->commit_seconds = t_commit;
  }
  return SCN_OK;
}

int32_t scn_graph_export_sizes(scn_store* s, uint64_t* n_nodes, uint64_t* n_lists, uint64_t* n_edges) {
  if (!s || !n_nodes || !n_lists || !n_edges) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  *n_nodes = *n_lists = *n_edges = 0;
  if (!s->has_graph) return SCN_OK;
  DeviceGuard g(s->device);
  SCN_TRY(prepare_build_state(s, s->m, thread_stream(s->device)));
  const BuildState* b = s->build;
  *n_nodes = b->n_nodes;
  for (uint64_t r = 0; r < b->n_nodes; ++r) {
    *n_lists += b->level[r] + 1u;
    *n_edges += b->cnt0[r];
    for (uint32_t l = 1; l <= b->level[r]; ++l) *n_edges += b->cntu[(size_t)b->up_off[r] + l - 1];
  }
  return SCN_OK;
}

int32_t scn_graph_export(scn_store* s, uint64_t* node_ids, int32_t* list_counts, uint32_t* edge_counts, uint64_t* edges,
                         uint64_t* entry_id, int32_t* max_layer) {
  if (!s) return fail(SCN_ERR_INVALID_PARAMETERS, "store is NULL");
  if (entry_id) *entry_id = s->has_graph ? s->entry_id : 0;
  if (max_layer) *max_layer = s->has_graph ? s->max_layer : -1;
  if (!s->has_graph) return SCN_OK;
  if (!node_ids || !list_counts || !edge_counts || !edges) return fail(SCN_ERR_INVALID_PARAMETERS, "NULL argument");
  DeviceGuard g(s->device);
  SCN_TRY(prepare_build_state(s, s->m, thread_stream(s->device)));
  const BuildState* b = s->build;
  std::vector<uint64_t> ids(b->n_nodes);
  if (b->n_nodes) SCN_CUDA(cudaMemcpy(ids.data(), s->d_ids, b->n_nodes * sizeof(uint64_t), cudaMemcpyDeviceToHost));
  uint64_t li = 0, ei = 0;
  for (uint64_t r = 0; r < b->n_nodes; ++r) {
    node_ids[r] = ids[r];
    list_counts[r] = (int32_t)b->level[r] + 1;
    for (uint32_t l = 0; l <= b->level[r]; ++l) {
      const uint32_t* a = l == 0 ? &b->adj0[r * b->s0] : &b->adju[((size_t)b->up_off[r] + l - 1) * b->su];
      const uint32_t c = l == 0 ? b->cnt0[r] : b->cntu[(size_t)b->up_off[r] + l - 1];
      edge_counts[li++] = c;
      for (uint32_t t = 0; t < c; ++t) edges[ei++] = ids[a[t]];
    }
  }
  return SCN_OK;
}

}  // extern "C"
