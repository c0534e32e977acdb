// hnsw_search.cu — K5: batched HNSW.Search (hnsw.go:292-350) — greedy descent through the upper
// layers, layer-0 beam search (searchLayer, hnsw.go:487-557) and the top-k cut, one warp per query.
//
// Mapping of the reference's state onto the warp:
//   candidates (W)  -> one sorted list of ef (key, row) pairs in shared memory (double buffered);
//                      key = ord(dist) << 32 | admission_seq << 1 | expanded  (stable: on equal
//                      distance the earlier-admitted entry stays ahead, hnsw.go:675-686)
//   dynamic (C)     -> the `expanded` bit: every live element of C is an un-expanded element of W
//                      (an element evicted from W has dist >= W[ef-1].dist and can only ever
//                      trigger the `break` at hnsw.go:516-518), so "pop the closest of C" is "take
//                      the first un-expanded entry of W" and the loop ends when there is none.
//                      (Only exact float ties with W[ef-1] can make the two differ.)
//   visited         -> open-addressing table of row indices, groups of four 32-bit slots, private
//                      to the warp and EPOCH-TAGGED: an entry is (tag << row_bits) | (row + 1) and only
//                      entries carrying the current query's tag count, so the table is cleared once
//                      per warp and launch (and when the tag wraps) instead of once per query — at
//                      ef = 128 that was a 64 KB store burst per query, a fifth of the kernel's DRAM
//                      traffic. The table lives in GLOBAL memory (L2-resident) by default, because it
//                      was what limited the number of resident queries, with a warp-collective,
//                      atomic-free insert and group snapshots fetched ahead for the predicted next
//                      expansion (visited_insert_warp); in shared memory with a CAS insert as the
//                      alternative (option hnsw_global=0). A table that fills up sends the query to
//                      an overflow pass with an 8x table.
//   Distance()      -> one lane per neighbour: the lane walks its neighbour's row in the reference's
//                      exact sequential fp32 order, so all (up to 32) neighbours of an expansion are
//                      evaluated concurrently and every distance — hence the whole walk — is
//                      bit-identical to the reference's. The rows reach the lanes through shared
//                      memory (warp-wide cp.async, one 128-byte line of four rows per instruction,
//                      requested BEFORE the visited test so that the copies fly while the table is
//                      probed; common.cuh gather_*), or through registers (LDG.256, option
//                      hnsw_gather=0). (The first version evaluated 4 rows at a time
//                      warp-cooperatively with shuffle reductions and serialised ~6 memory latencies
//                      per expansion: 25 us per expansion at C1.)
//   admission       -> the reference admits neighbours one by one against the current W[ef-1]
//                      (strict <) and re-sorts; the result is the best ef of W u new under the
//                      stable (dist, admission order) order, which is computed here in one step:
//                      warp bitonic sort of the new keys + rank-based merge into the other buffer.
// Per neighbour the reference's order of tests is kept: visited? -> deleted? (not marked visited,
// not traversed) -> mark visited -> distance -> admit if |W| < ef or d < W[ef-1].d (strict).
// The rerank of hnsw.go:317-347 recomputes the same distances and stably re-sorts an already
// sorted list, so the first min(k, |W|) entries of W are the result.
// Roofline: HBM random gather; algorithmic bytes per query = evals*dim*4 + hops*2M*4. In practice
// the walk is a chain of dependent round trips: throughput = resident queries / per-expansion
// latency (DESIGN.md 4.2 has the measurements behind every choice above).
#include "store.h"
#include "visited.cuh"

namespace scn {

// GM, how the rows of a batch reach the lanes: 0 = LDG.256 into registers; else warp-wide cp.async
// into shared memory (common.cuh): 1 = <512, 1>, 2 = <512, 2> (rows > 512 bytes), 3 = <256, 1>
__host__ __device__ constexpr uint32_t gm_ch(int gm) { return gm == 3 ? 256u : 512u; }
__host__ __device__ constexpr uint32_t gm_nbuf(int gm) { return gm == 2 ? 2u : 1u; }
__host__ __device__ constexpr uint32_t gm_bytes(int gm) { return gm == 0 ? 0u : ga_stage_bytes(gm_ch(gm), gm_nbuf(gm)); }
constexpr int HNSW_MAX_WARPS = 2;  // warps (= queries) per CTA: 2 for the register gather, 1 for the shared-memory gather

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
  uint32_t has_deleted;    // 0: no row is soft-deleted, the bitmap need not be read
  uint32_t global_first;   // 1: the first pass keeps its visited tables in global memory too (L2-resident; more warps per SM)
  uint32_t early_issue;    // shared-memory gather: request the rows of a list before the visited test
  uint32_t exact_ties;     // 1: a walk whose end state still holds a distance tie at the edge of W is redone exactly
  uint32_t entry_row;
  int32_t max_layer;
  const float* q;
  const uint32_t* qlist;   // optional list of query indices (overflow pass)
  uint32_t* nq_dev;        // optional device count for qlist
  uint32_t nq, k, ef, ef_pad;
  uint32_t max_per_sm;     // global_first: cap on resident CTAs per SM (0 = whatever fits)
  uint32_t hash_size;      // entries per warp
  uint32_t row_bits;       // bits of an entry that hold row + 1; the bits above hold the query tag
  uint32_t tag_max;        // largest tag (>= 1); the table is cleared when the tag would exceed it
  uint32_t* ghash;         // global tables [warps][hash_size] when USE_GLOBAL
  uint32_t* overflow_list; // first pass: queries whose visited table overflowed (nullptr in the overflow pass)
  uint32_t* overflow_count;
  uint64_t* out_ids;
  float* out_dist;
  uint32_t* out_counts;
  unsigned long long* stats;  // [0] distance evaluations, [1] expansions (optional)
  unsigned long long* failed; // counts queries whose visited table overflowed in the overflow pass too (no result)
};

__host__ __device__ inline size_t hnsw_warp_bytes(uint32_t pitch, uint32_t ef_pad, uint32_t hash_size, bool global_hash,
                                                  uint32_t stage_bytes) {
  // wkey[2][ef_pad] u64 | snk[32] u64 | q[pitch] f32 | wrow[2][ef_pad] u32 | snr[32] u32 |
  // stage (shared-memory gather only) | hash
  return (size_t)ef_pad * 16 + 256 + (size_t)pitch * 4 + (size_t)ef_pad * 8 + 128 + stage_bytes +
         (global_hash ? 0 : (size_t)hash_size * 4);
}

template <int METRIC, bool USE_GLOBAL, int GM>
__global__ void __launch_bounds__(HNSW_MAX_WARPS * 32) hnsw_search_kernel(HnswArgs a) {
  constexpr bool GATHER = GM != 0;
  constexpr uint32_t GCH = gm_ch(GM), GNB = gm_nbuf(GM);
  extern __shared__ __align__(16) unsigned char smem_hnsw[];
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const uint32_t block_warps = blockDim.x >> 5;
  const uint32_t warp_global = blockIdx.x * block_warps + warp;
  const uint32_t warps_total = gridDim.x * block_warps;
  unsigned char* base = smem_hnsw + hnsw_warp_bytes(a.pitch, a.ef_pad, a.hash_size, USE_GLOBAL, gm_bytes(GM)) * warp;
  uint64_t* wkey0 = reinterpret_cast<uint64_t*>(base);
  uint64_t* wkey1 = wkey0 + a.ef_pad;
  uint64_t* snk = wkey1 + a.ef_pad;                              // sorted new keys
  float* sq = reinterpret_cast<float*>(snk + 32);
  uint32_t* wrow0 = reinterpret_cast<uint32_t*>(sq + a.pitch);
  uint32_t* wrow1 = wrow0 + a.ef_pad;
  uint32_t* snr = wrow1 + a.ef_pad;                              // rows of the sorted new keys
  unsigned char* stage = reinterpret_cast<unsigned char*>(snr + 32);
  uint32_t* smem_hash = reinterpret_cast<uint32_t*>(stage + gm_bytes(GM));
  uint32_t* hash = USE_GLOBAL ? (a.ghash + (size_t)warp_global * a.hash_size) : smem_hash;
  const uint32_t ef = a.ef;
  const uint32_t n_groups = a.hash_size >> 2;  // hash_size is a multiple of 4
  const uint32_t nq = a.qlist ? min(*a.nq_dev, a.nq) : a.nq;
  const float INF = __int_as_float(0x7f800000);
  unsigned long long evals = 0, hops = 0;
  const uint32_t row_bits = a.row_bits;
  uint32_t tag = 0;  // 0 = the table has not been cleared yet (scratch memory holds garbage)

  for (uint32_t qslot = warp_global; qslot < nq; qslot += warps_total) {
    const uint32_t qi = a.qlist ? a.qlist[qslot] : qslot;
    __syncwarp();
    stage_query(sq, a.q + (size_t)qi * a.dim, a.dim, a.pitch, lane, 32);
    if (tag == 0 || tag >= a.tag_max) {  // first query of this warp, or the tag wrapped: start from an empty table
      for (uint32_t i = lane; i < n_groups; i += 32) reinterpret_cast<uint4*>(hash)[i] = make_uint4(HASH_EMPTY, HASH_EMPTY, HASH_EMPTY, HASH_EMPTY);
      tag = 1;
    } else {
      ++tag;  // every entry of the previous queries is stale from here on
    }
    __syncwarp();
    float qn = 0.0f;
    if (METRIC == M_COS) {
      if (lane == 0) qn = exact_norm_padded(sq, a.pitch / 4);
      qn = __shfl_sync(0xffffffffu, qn, 0);
    }

    uint64_t* wkey = wkey0;   // current list
    uint32_t* wrow = wrow0;
    uint64_t* okey = wkey1;   // merge target
    uint32_t* orow = wrow1;
    uint32_t cnt = 0;
    bool overflow = false;
    bool tie = false;            // the reference's walk may go on where this one ends: redo it exactly (second pass)
    uint32_t ghost_ord = 0xFFFFFFFFu;   // distance of an un-expanded element that left W with the very distance W[ef-1] had then
    uint32_t cur = a.entry_row;
    const bool have_entry = (cur != ROW_NONE) && (cur < a.n_rows) && !bit_test(a.deleted, cur) && a.max_layer >= 0;
    if (have_entry) {
      // The three phases of HNSW.Search (entry distance, greedy descent, layer-0 beam) run through
      // ONE loop: produce a batch of up to 32 rows (one per lane) -> evaluate it -> consume the
      // distances. The distance code (128 registers of row data in flight) is thus instantiated
      // once. All state below is warp-uniform.
      enum { PH_ENTRY = 0, PH_DESCENT = 1, PH_BEAM = 2 };
      int phase = PH_ENTRY;
      int layer = a.max_layer;
      bool in_list = false;            // an adjacency list is being walked (chunk c0 of list_len)
      const uint32_t* list = nullptr;
      uint32_t list_len = 0, c0 = 0;
      float dcur = 0.0f;
      float best = 0.0f;               // descent: closest neighbour so far of the expanded node
      uint32_t best_row = cur;
      uint32_t seq = 1, visited = 1;   // beam: admission sequence, size of the visited set
      uint32_t p_lo = 0;               // beam: every entry of W before p_lo has been expanded
      uint32_t pre_row = ROW_NONE;     // beam: adjacency prefetched for this row (the runner-up)
      uint32_t pre_nb = ROW_NONE;
      uint32_t pre_grp_row = ROW_NONE; // global table: home groups of pre_nb fetched for this row, after the last insert
      uint4 pre_grp = make_uint4(0, 0, 0, 0);
      for (;;) {
        // ================= produce: the next non-empty batch of rows, or the end =================
        uint32_t nb = ROW_NONE;
        bool ok = false;
        uint32_t mask = 0;
        bool finished = false;
        bool begun = false;  // shared-memory gather: the copies of this batch are already in flight
        for (;;) {
          if (phase == PH_ENTRY) {
            nb = cur;
            ok = (lane == 0);
            mask = 1u;
            break;
          }
          if (phase == PH_DESCENT) {
            // greedy descent, layers maxLayer..1 with numClosest = 1 (hnsw.go:309-311)
            if (!in_list) {
              if (layer < 1) {
                // ---- layer 0 beam starts from the descent's result (hnsw.go:314) ----
                phase = PH_BEAM;
                if (lane == 0) {
                  wkey[0] = ((uint64_t)f32_ord(dcur) << 32);
                  wrow[0] = cur;
                }
                cnt = 1;
                if (lane == 0) visited_insert<USE_GLOBAL>(hash, n_groups, cur, (tag << row_bits) | (cur + 1), tag, row_bits);
                __syncwarp();
                continue;
              }
              ++hops;
              if ((int)a.levels[cur] < layer) {  // GetConnections(layer) is empty (hnsw.go:45-50)
                --layer;
                continue;
              }
              list = a.adj_up + ((size_t)a.up_off[cur] + (layer - 1)) * a.su;
              list_len = a.su;
              c0 = 0;
              in_list = true;
              best = dcur;
              best_row = cur;
            }
            nb = (c0 + lane < list_len) ? __ldg(list + c0 + lane) : ROW_NONE;
            ok = (nb != ROW_NONE) && !(a.has_deleted && bit_test(a.deleted, nb));
            mask = __ballot_sync(0xffffffffu, ok);
            evals += __popc(mask);
            break;
          }
          // ---- PH_BEAM (searchLayer on layer 0, hnsw.go:487-557) ----
          if (!in_list) {
            // closest un-expanded entry of W (== pop of the reference's `dynamic` list)
            uint32_t m = 0, b0 = p_lo & ~31u;
            for (; b0 < cnt; b0 += 32) {
              const uint32_t i = b0 + lane;
              const bool un = (i < cnt) && !(reinterpret_cast<const uint32_t*>(wkey)[2 * i] & 1u);
              m = __ballot_sync(0xffffffffu, un);
              if (m) break;
            }
            if (!m) {
              // every entry of W has been expanded; the reference would go on with a ghost that still ties with W[ef-1]
              if (a.exact_ties && cnt >= ef && ghost_ord == reinterpret_cast<const uint32_t*>(wkey)[2 * (ef - 1) + 1]) tie = true;
              finished = true;
              break;
            }
            const uint32_t p = b0 + __ffs(m) - 1;
            p_lo = p + 1;
            cur = wrow[p];
            // the runner-up is the most likely next expansion: fetch its adjacency line now, so
            // that the adjacency -> rows dependency of the next hop starts from registers
            const uint32_t m2 = m & (m - 1);
            const uint32_t row2 = m2 ? wrow[b0 + __ffs(m2) - 1] : ROW_NONE;
            __syncwarp();
            if (lane == 0) reinterpret_cast<uint32_t*>(wkey)[2 * p] |= 1u;
            __syncwarp();
            ++hops;
            if (visited + a.s0 > a.hash_size - (a.hash_size >> 3)) {  // keep the table below 7/8 full
              overflow = true;
              finished = true;
              break;
            }
            list = a.adj0 + (size_t)cur * a.s0;
            list_len = a.s0;
            c0 = 0;
            in_list = true;
            if (cur == pre_row) nb = pre_nb;
            else nb = ((uint32_t)lane < list_len) ? __ldg(list + lane) : ROW_NONE;
            pre_row = row2;
            if (row2 != ROW_NONE && (uint32_t)lane < list_len) pre_nb = __ldg(a.adj0 + (size_t)row2 * a.s0 + lane);
          } else {
            nb = (c0 + lane < list_len) ? __ldg(list + c0 + lane) : ROW_NONE;
          }
          ok = (nb != ROW_NONE);
          if (GATHER) {
            // Ask for the rows of ALL listed neighbours before the visited test: the copies then fly
            // while the table is probed (two dependent round trips when it lives in global
            // memory). Rows that turn out to be visited or deleted (about one in five) were
            // fetched for nothing; bandwidth is not what bounds the walk.
            const uint32_t listed = __ballot_sync(0xffffffffu, ok);
            if (listed && a.early_issue) {
              gather_begin<GCH, GNB>(a.vec, a.pitch, nb, listed, stage, lane);
              begun = true;
            }
          }
          // reference order: visited? -> deleted? -> mark visited. A deleted row is never
          // inserted, so testing `deleted` first and inserting only live rows is equivalent.
          if (ok && a.has_deleted) ok = !bit_test(a.deleted, nb);
          if (USE_GLOBAL) ok = visited_insert_warp(hash, n_groups, nb, ok, c0 == 0 && cur == pre_grp_row, pre_grp, lane, tag, row_bits);
          else if (ok) ok = visited_insert<false>(hash, n_groups, nb, (tag << row_bits) | (nb + 1), tag, row_bits);
          mask = __ballot_sync(0xffffffffu, ok);
          const uint32_t ns = __popc(mask);
          if (USE_GLOBAL && c0 + 32 >= list_len) {
            // nothing is inserted any more before the next expansion: if that is the runner-up, its
            // probes can start from these group snapshots (off the critical path)
            pre_grp_row = pre_row;
            if (pre_row != ROW_NONE && pre_nb != ROW_NONE) pre_grp = ld_group<true>(hash + home_group(pre_nb, n_groups) * 4);
          }
          if (!ns) {  // nothing new in this chunk
            if (GATHER && begun) {  // drain the copies
              gather_finish<METRIC, GCH, GNB>(a.vec, a.norm, a.pitch, sq, qn, nb, 0u, stage, lane);
              begun = false;
            }
            c0 += 32;
            if (c0 >= list_len) in_list = false;
            continue;
          }
          visited += ns;
          evals += ns;
          break;
        }
        if (finished) break;

        // ================= evaluate: the reference's Distance(), one row per lane ================
        // (register gather: 2 x 8 LDG.256 per lane in flight; shared-memory gather: one 512-byte request
        // per row and stage to the copy engine, rows land in shared memory)
        float d = INF;
        if (GATHER) {
          if (!begun) gather_begin<GCH, GNB>(a.vec, a.pitch, nb, mask, stage, lane);
          d = gather_finish<METRIC, GCH, GNB>(a.vec, a.norm, a.pitch, sq, qn, nb, mask, stage, lane);
        }
        else if (ok) d = row_distance<METRIC>(a.vec, a.norm, a.pitch, sq, qn, nb);

        // ================= consume ==============================================================
        if (phase == PH_ENTRY) {
          dcur = __shfl_sync(0xffffffffu, d, 0);
          ++evals;
          phase = PH_DESCENT;
          continue;
        }
        if (phase == PH_DESCENT) {
          // first strict minimum in list order (the reference's sequential `d < W[0].d` updates)
          float dm = (ok && d == d) ? d : INF;
          uint32_t im = lane;
#pragma unroll
          for (int o = 16; o > 0; o >>= 1) {
            const float od = __shfl_xor_sync(0xffffffffu, dm, o);
            const uint32_t oi = __shfl_xor_sync(0xffffffffu, im, o);
            if (od < dm || (od == dm && oi < im)) {
              dm = od;
              im = oi;
            }
          }
          if (dm < best) {
            best = dm;
            best_row = __shfl_sync(0xffffffffu, nb, im);
          }
          c0 += 32;
          if (c0 >= list_len) {
            in_list = false;
            if (best_row == cur) --layer;  // no closer neighbour: this layer is done
            else {
              cur = best_row;
              dcur = best;
            }
          }
          continue;
        }
        // ---- PH_BEAM: admission (hnsw.go:536-542): while |W| < ef everything enters; once full only
        // d < W[ef-1].d (strict; a NaN never passes). Entering while the list is not full and
        // being pushed out later is the same as losing the final cut of the merge below.
        c0 += 32;
        if (c0 >= list_len) in_list = false;
        const uint32_t od = f32_ord(d);
        bool in = ok;
        if (in && cnt >= ef) in = od < reinterpret_cast<const uint32_t*>(wkey)[2 * (ef - 1) + 1];
        // stable admission order = adjacency order
        const uint32_t rank = __popc(mask & ((1u << lane) - 1u));
        const uint64_t key = ((uint64_t)od << 32) | ((uint64_t)(seq + rank) << 1);
        seq += __popc(mask);
        const uint32_t mask_in = __ballot_sync(0xffffffffu, in);
        const uint32_t nn = __popc(mask_in);
        if (!nn) continue;
        // ---- merge of W[0..cnt) with the nn admitted keys into the other buffer, by rank ----------
        // (keys are unique: the admission sequence number breaks distance ties in the reference's
        // stable order.) Once W is full only a handful of neighbours pass the W[ef-1] test, so the
        // admitted keys are compacted in adjacency order (lane j < nn then owns new entry j) and
        // every entry counts the keys of the other list below it: old entry i moves to
        // i + #{new < it}; new entry j moves to #{old < it} + #{new < it}.
        const uint32_t slot = __popc(mask_in & ((1u << lane) - 1u));
        if (in) {
          snk[slot] = key;
          snr[slot] = nb;
        }
        __syncwarp();
        const bool mine = (uint32_t)lane < nn;
        const uint64_t nkey = mine ? snk[lane] : KEY_NONE;
        uint32_t npos = 0;  // #{old < nkey} + #{new < nkey}
        uint32_t dropped = 0xFFFFFFFFu;   // smallest ord(distance) among the entries that do not make the cut
        for (uint32_t i0 = 0; i0 < cnt; i0 += 128) {
          uint64_t kv[4];
          uint32_t rw[4], below[4];
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * 32 + lane;
            kv[u] = (i < cnt) ? wkey[i] : KEY_NONE;
            rw[u] = (i < cnt) ? wrow[i] : ROW_NONE;
            below[u] = 0;
          }
          for (uint32_t j = 0; j < nn; ++j) {
            const uint64_t nk = snk[j];  // broadcast
            uint32_t c = 0;
#pragma unroll
            for (int u = 0; u < 4; ++u) {
              const bool lt = nk < kv[u];
              below[u] += lt ? 1u : 0u;
              c += (!lt && kv[u] != KEY_NONE) ? 1u : 0u;  // old key below new key j (keys are distinct)
            }
            c = __reduce_add_sync(0xffffffffu, c);
            if ((uint32_t)lane == j) npos += c;
          }
#pragma unroll
          for (int u = 0; u < 4; ++u) {
            const uint32_t i = i0 + u * 32 + lane;
            const uint32_t pos = i + below[u];
            if (i < cnt && pos < ef) {
              okey[pos] = kv[u];
              orow[pos] = rw[u];
            } else if (i < cnt && !(kv[u] & 1ull)) {   // (an expanded entry that leaves W is gone for good)
              dropped = min(dropped, (uint32_t)(kv[u] >> 32));
            }
          }
        }
        for (uint32_t j = 0; j < nn; ++j) npos += (snk[j] < nkey) ? 1u : 0u;
        if (mine && npos < ef) {
          okey[npos] = nkey;
          orow[npos] = snr[lane];
        } else if (mine) {
          dropped = min(dropped, (uint32_t)(nkey >> 32));
        }
        // a new (un-expanded) entry may have landed in front of p_lo
        p_lo = min(p_lo, __reduce_min_sync(0xffffffffu, mine ? npos : 0xFFFFFFFFu));
        cnt = min(ef, cnt + nn);
        __syncwarp();
        // Exact-tie rule of the reference: an admitted, un-expanded element that is pushed out of W while its
        // distance EQUALS the new W[ef-1] stays in `dynamic` and is still expanded if the tie holds when W has
        // been worked off (hnsw.go:516-518 stops only at dist > W[ef-1].dist). Remember the tie; if it still
        // holds at the end of the walk, the walk is redone by the exact walk kernel (hnsw_build.cu, WState) in
        // the second pass. (A NEW key that misses the cut with that distance is treated alike: whether the
        // reference's one-by-one admission, 536-542, had let it in is decided there.)
        dropped = __reduce_min_sync(0xffffffffu, dropped);
        if (cnt == ef && dropped == (uint32_t)(okey[ef - 1] >> 32)) ghost_ord = dropped;
        uint64_t* tk = wkey;
        wkey = okey;
        okey = tk;
        uint32_t* tr = wrow;
        wrow = orow;
        orow = tr;
      }
    }

    if (overflow || tie) {
      if (lane == 0) {
        uint32_t slot = atomicAdd(a.overflow_count, 1u);
        a.overflow_list[slot] = qi;
      }
      continue;  // the second pass (exact walk, large table) will produce this query's results
    }

    // ---- result: the first min(k, |W|) entries, already in (distance, admission) order --------
    const uint32_t n_out = min(a.k, cnt);
    __syncwarp();
    for (uint32_t i = lane; i < a.k; i += 32) {
      uint64_t id = 0;
      float d = INF;
      if (i < n_out) {
        id = a.ids[wrow[i]];
        d = ord_f32((uint32_t)(wkey[i] >> 32));
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

template <int METRIC, int GM>
static int32_t launch_hnsw(scn_store* s_, HnswArgs& a, int sms, cudaStream_t stream, Scratch& scratch, Profiler* prof) {
  // pass 1: shared-memory visited table (or, global_first, a table of the same size in global memory)
  const int warps = GM ? 1 : HNSW_MAX_WARPS;
  const uint32_t stages = gm_bytes(GM);
  const uint32_t blocks_needed = (a.nq + warps - 1) / warps;
  SCN_TRY(scratch.alloc(&a.overflow_list, (size_t)a.nq));
  SCN_TRY(scratch.alloc(&a.overflow_count, 1));
  SCN_CUDA(cudaMemsetAsync(a.overflow_count, 0, sizeof(uint32_t), stream));
  if (!a.global_first) {
    const size_t smem = hnsw_warp_bytes(a.pitch, a.ef_pad, a.hash_size, false, stages) * warps;
    SCN_ALLOW_SMEM((hnsw_search_kernel<METRIC, false, GM>), smem);
    int per_sm = 0;
    SCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hnsw_search_kernel<METRIC, false, GM>, warps * 32, smem));
    if (per_sm < 1) return fail(SCN_ERR_INVALID_PARAMETERS, "ef=%u / dim=%u need more shared memory than one SM has", a.ef, a.dim);
    const int grid = (int)std::min<uint32_t>(blocks_needed, (uint32_t)(sms * per_sm));
    if (prof) prof->begin("hnsw_search");
    hnsw_search_kernel<METRIC, false, GM><<<grid, warps * 32, smem, stream>>>(a);
    SCN_LAUNCHED();
    if (prof) prof->end();
  } else {
    const size_t smem = hnsw_warp_bytes(a.pitch, a.ef_pad, 0, true, stages) * warps;
    SCN_ALLOW_SMEM((hnsw_search_kernel<METRIC, true, GM>), smem);
    int per_sm = 0;
    SCN_CUDA(cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, hnsw_search_kernel<METRIC, true, GM>, warps * 32, smem));
    if (per_sm < 1) return fail(SCN_ERR_INVALID_PARAMETERS, "ef=%u / dim=%u need more shared memory than one SM has", a.ef, a.dim);
    if (a.max_per_sm > 0) per_sm = std::min<int>(per_sm, (int)a.max_per_sm);
    const int grid = (int)std::min<uint32_t>(blocks_needed, (uint32_t)(sms * per_sm));
    SCN_TRY(scratch.alloc(&a.ghash, (size_t)grid * warps * a.hash_size));
    if (prof) prof->begin("hnsw_search");
    hnsw_search_kernel<METRIC, true, GM><<<grid, warps * 32, smem, stream>>>(a);
    SCN_LAUNCHED();
    if (prof) prof->end();
  }
  // pass 2 (always enqueued, exits at once when its list is empty): the exact walk kernel of hnsw_build.cu for
  // the queries whose visited table overflowed or whose walk met a distance tie at the edge of W. Table: 2x
  // the row count — one that cannot fill up — capped at 4 M entries (16 MB per resident query; a walk that
  // visits more than 3.6 M rows is reported as failed, scn_search_hnsw returns 5001).
  uint32_t hash2 = (uint32_t)std::min<uint64_t>((uint64_t)1 << 22, next_pow2(a.n_rows) * 2ull);
  hash2 = std::max<uint32_t>(hash2, 1024u);
  if (prof) prof->begin("hnsw_search_exact");
  SCN_TRY(hnsw_search_exact(s_, a.q, a.overflow_list, a.overflow_count, a.nq, a.k, a.ef, hash2, a.out_ids, a.out_dist, a.out_counts, a.failed,
                            stream, scratch));
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
  a.has_deleted = (s->live != s->rows) ? 1u : 0u;
  // gather mode (see gm_ch). Auto: the smallest stage layout (most resident queries) for rows of up
  // to 512 bytes; longer rows: see opt_hnsw_gather_long.
  int gm = (int)s->opt_hnsw_gather;
  if (gm < 0 || gm > 3) gm = (a.pitch * 4 > 512) ? (int)s->opt_hnsw_gather_long : 3;
  if (gm < 0 || gm > 3) gm = 2;
  if (gm == 2 && a.pitch * 4 <= 512) gm = 1;  // a single stage needs no second buffer
  a.global_first = s->opt_hnsw_global ? 1u : 0u;
  a.early_issue = s->opt_hnsw_early ? 1u : 0u;
  a.exact_ties = s->opt_hnsw_exact_ties ? 1u : 0u;
  a.max_per_sm = (uint32_t)std::max<int64_t>(0, s->opt_hnsw_per_sm);
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
  // global tables cost no shared memory: 4x the entries keep the groups sparse (fewer second probe rounds)
  if (a.global_first) a.hash_size = (uint32_t)std::min<uint64_t>((uint64_t)next_pow2(a.hash_size) * 2, 1u << 16);
  if (s->opt_hnsw_hash > 0) a.hash_size = round_up((uint32_t)std::min<int64_t>(s->opt_hnsw_hash, 1 << 16), 512);
  // visited entries: row + 1 in the low row_bits bits, the query tag above
  if (a.n_rows >= (1u << 31)) return fail(SCN_ERR_INVALID_PARAMETERS, "HNSW search supports at most 2^31 - 1 rows per device");
  a.row_bits = 1;
  while ((1ull << a.row_bits) <= (uint64_t)a.n_rows) ++a.row_bits;   // row + 1 <= n_rows < 2^row_bits
  a.tag_max = (uint32_t)((1ull << (32 - a.row_bits)) - 1);           // >= 1
  a.out_ids = d_out_ids;
  a.out_dist = d_out_dist;
  a.out_counts = d_out_counts;
  SCN_CUDA(cudaMemsetAsync(s->d_counters, 0, 4 * sizeof(unsigned long long), stream));
  a.stats = s->opt_profile ? s->d_counters : nullptr;
  a.failed = s->d_counters + 3;
  // shrink the table until at least one block fits
  while (!a.global_first && hnsw_warp_bytes(a.pitch, a.ef_pad, a.hash_size, false, gm_bytes(gm)) * (gm ? 1 : HNSW_MAX_WARPS) > 200 * 1024 &&
         a.hash_size > 1024)
    a.hash_size = round_up(a.hash_size / 2, 512);
  int32_t rc;
#define HN(MT)                                                                  \
  switch (gm) {                                                                 \
    case 0: rc = launch_hnsw<MT, 0>(s, a, sms, stream, scratch, prof); break;      \
    case 1: rc = launch_hnsw<MT, 1>(s, a, sms, stream, scratch, prof); break;      \
    case 2: rc = launch_hnsw<MT, 2>(s, a, sms, stream, scratch, prof); break;      \
    default: rc = launch_hnsw<MT, 3>(s, a, sms, stream, scratch, prof); break;     \
  }
  switch (s->metric) {
    case M_L2: HN(M_L2); break;
    case M_COS: HN(M_COS); break;
    default: HN(M_IP); break;
  }
#undef HN
  SCN_TRY(rc);
  return SCN_OK;
}

}  // namespace scn
