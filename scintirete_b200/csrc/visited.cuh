// visited.cuh — the per-warp visited set of the graph walks (hnsw_search.cu, hnsw_build.cu).
#pragma once

#include "common.cuh"

namespace scn {

constexpr uint32_t HASH_EMPTY = 0u;

__device__ __forceinline__ uint32_t home_group(uint32_t row, uint32_t n_groups) { return __umulhi(row * 2654435761u, n_groups); }

// visited set: open addressing over groups of four 32-bit slots (one 128-bit load per probe). Slots
// of a group fill in order and are never emptied within a query (entries of earlier queries carry
// another tag and count as empty: the entries of the current query form a prefix of the group), so a
// group that still has a free slot and does not hold the key proves the key absent. Returns true
// if `row` was inserted (first visit), false if it was already there.
template <bool GLOBAL>
__device__ __forceinline__ uint4 ld_group(const uint32_t* p) {
  uint4 v;
  if (GLOBAL) {
    // (the table is private to one warp, written with L2 atomics: only L1 must be bypassed)
    asm volatile("ld.relaxed.gpu.global.v4.u32 {%0,%1,%2,%3}, [%4];" : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w) : "l"(p) : "memory");
  } else {
    asm volatile("ld.volatile.shared.v4.u32 {%0,%1,%2,%3}, [%4];"
                 : "=r"(v.x), "=r"(v.y), "=r"(v.z), "=r"(v.w)
                 : "r"((uint32_t)__cvta_generic_to_shared(p)));
  }
  return v;
}

// `key` = (tag << row_bits) | (row + 1); a slot is in use iff it carries the current tag.
template <bool GLOBAL>
__device__ __forceinline__ bool visited_insert(uint32_t* tab, uint32_t n_groups, uint32_t row, uint32_t key, uint32_t tag,
                                               uint32_t row_bits) {
  uint32_t g = home_group(row, n_groups);
  for (uint32_t probes = 0; probes <= n_groups;) {
    uint32_t* grp = tab + g * 4;
    const uint4 v = ld_group<GLOBAL>(grp);
    if (v.x == key || v.y == key || v.z == key || v.w == key) return false;
    const int e = ((v.x >> row_bits) != tag) ? 0 : ((v.y >> row_bits) != tag) ? 1 : ((v.z >> row_bits) != tag) ? 2 : ((v.w >> row_bits) != tag) ? 3 : -1;
    if (e >= 0) {
      const uint32_t seen = (e == 0) ? v.x : (e == 1) ? v.y : (e == 2) ? v.z : v.w;  // a stale entry of an earlier query (or 0)
      const uint32_t old = atomicCAS(grp + e, seen, key);
      if (old == seen) return true;
      if (old == key) return false;
      continue;  // another lane took the slot: look at the same group again
    }
    ++probes;
    if (++g == n_groups) g = 0;
  }
  return false;  // table full (guarded against by the overflow check)
}

// Warp-collective insert for a table in GLOBAL memory, where every dependent access is an L2 round
// trip: one group load per probe round and NO atomic. The table is private to the warp, so the
// lanes settle among themselves who takes which slot (match.any on the group index: the lanes
// that want a slot of the same group take consecutive ones, those that do not fit move on to the
// next group in the next round) and write with plain stores, which nobody waits for. `pre` may hold
// the home group fetched ahead of time (valid only if nothing was inserted since). Lanes with
// want == false only take part in the collectives. Returns true on a first visit.
__device__ __forceinline__ bool visited_insert_warp(uint32_t* tab, uint32_t n_groups, uint32_t row, bool want, bool have_pre,
                                                    uint4 pre, uint32_t lane, uint32_t tag, uint32_t row_bits) {
  // (An adjacency list never names a row twice: scn_graph_upload drops repeats, which the
  // reference would skip as visited anyway. So the lanes of a batch hold distinct rows.)
  const uint32_t key = (tag << row_bits) | (row + 1);
  bool pending = want;
  bool fresh = false;
  uint32_t g = home_group(row, n_groups);
  for (uint32_t round = 0; round <= n_groups; ++round) {
    if (!__any_sync(0xffffffffu, pending)) break;
    int e = 4;
    if (pending) {
      const uint4 v = (have_pre && round == 0) ? pre : ld_group<true>(tab + g * 4);
      if (v.x == key || v.y == key || v.z == key || v.w == key) pending = false;
      else e = ((v.x >> row_bits) != tag) ? 0 : ((v.y >> row_bits) != tag) ? 1 : ((v.z >> row_bits) != tag) ? 2 : ((v.w >> row_bits) != tag) ? 3 : 4;
    }
    // rank among the lower lanes that want a slot of the same group (31 shuffles: a third of the
    // latency of MATCH.ANY on ~26 distinct values)
    const uint32_t gi = (pending && e < 4) ? g : 0xFFFFFFFFu;
    uint32_t rank = 0;
#pragma unroll
    for (uint32_t j = 0; j < 31; ++j) {
      const uint32_t gj = __shfl_sync(0xffffffffu, gi, j);
      rank += (j < lane && gj == g) ? 1u : 0u;
    }
    if (pending) {
      const uint32_t slot = (uint32_t)e + rank;
      if (slot < 4) {
        asm volatile("st.relaxed.gpu.global.u32 [%0], %1;" ::"l"(tab + g * 4 + slot), "r"(key) : "memory");
        pending = false;
        fresh = true;
      } else if (++g == n_groups) {
        g = 0;
      }
    }
    __syncwarp();  // this round's stores are ordered before the next round's (and the next batch's) loads
  }
  return fresh;
}

}  // namespace scn
