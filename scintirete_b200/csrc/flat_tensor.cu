// flat_tensor.cu — K2: tensor-core candidate filter for large query batches, plus the certified
// exact rerank that makes its results identical to the exact scan.
//
// Pipeline (all asynchronous on one stream):
//   prep_queries      q -> bf16 q~ (tensor operand), ||q||, ||q~||, ||q - q~||
//   tensor_filter     tcgen05 GEMM  S = aux[x] + c * (q~ . x~)   (c = -2 for L2, -1 for IP/cosine),
//                     128 queries stationary in TMEM (A operand), bf16 mirror rows streamed by TMA
//                     (B operand, 128B-swizzled K-major tiles), fp32 accumulators in TMEM (double
//                     buffered); the epilogue warps read the accumulators with tcgen05.ld and keep,
//                     per query, the k' best (score,row) of their column chunk — the score matrix
//                     is never materialised.
//   merge_candidates  per query: union of the chunk lists -> best k'' rows + tau (a lower bound on
//                     the filter score of every row NOT passed on)
//   rerank_rows       exact reference arithmetic on the k'' rows -> exact top-k keys
//   certify           proves from tau and rigorous rounding bounds that no other row can enter the
//                     top-k (ties included); queries that cannot be certified are appended to a list
//   flat_search_exact exact scan for the listed queries (normally none)
//
// Roofline of tensor_filter: tensor pipe, 2*nq*N*D flop; HBM side reads the bf16 mirror once per
// wave of query blocks (L2 keeps the tile window shared by the CTAs that walk the same chunk).
#include <cuda.h>

#include <cstdio>
#include <cstdlib>

#include "store.h"

namespace scn {

constexpr int TF_BM = 128;        // queries per CTA (UMMA M)
constexpr int TF_BK = 64;         // bf16 per smem K block = one 128-byte swizzle row
constexpr int TF_THREADS = 192;   // warp 0: TMA, warp 1: MMA, warps 2..: epilogue, 4*EW warps (TMEM lane quarters 2,3,0,1,...)
constexpr int TF_MAX_KPAD = 768;  // widest A operand that fits TMEM next to an accumulator (TMEM-stationary queries)
constexpr int TF_MAX_KPAD_STREAM = 8192;  // beyond 768 the query block is streamed through shared memory with the rows

// ---- PTX helpers --------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count));
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
__device__ __forceinline__ void mbar_wait(uint64_t* bar, uint32_t parity) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "WAIT_%=:\n\t"
      "mbarrier.try_wait.parity.shared::cta.b64 p, [%0], %1;\n\t"
      "@p bra DONE_%=;\n\t"
      "bra WAIT_%=;\n\t"
      "DONE_%=:\n\t}" ::"r"(smem_u32(bar)),
      "r"(parity)
      : "memory");
}
__device__ __forceinline__ void tma_load_2d(void* smem_dst, const CUtensorMap* tmap, uint64_t* bar, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T, kind::f16 (bf16 in, fp32 accumulate), M=128
__device__ __forceinline__ void tc_mma_ts(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// D[tmem] (+)= A[smem] * B[smem]^T (both operands K-major 128B-swizzled tiles)
__device__ __forceinline__ void tc_mma_ss(uint32_t d_tmem, uint64_t a_desc, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::f16 [%0], %1, %2, %3, p;\n\t}" ::"r"(d_tmem),
      "l"(a_desc), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}
__device__ __forceinline__ void tc_st4(uint32_t taddr, uint32_t a, uint32_t b, uint32_t c, uint32_t d) {
  asm volatile("tcgen05.st.sync.aligned.32x32b.x4.b32 [%0], {%1, %2, %3, %4};" ::"r"(taddr), "r"(a), "r"(b), "r"(c), "r"(d)
               : "memory");
}
__device__ __forceinline__ void tc_wait_st() { asm volatile("tcgen05.wait::st.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_wait_ld() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }
__device__ __forceinline__ void tc_ld32(uint32_t taddr, float v[32]) {
  uint32_t* u = reinterpret_cast<uint32_t*>(v);
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(u[0]), "=r"(u[1]), "=r"(u[2]), "=r"(u[3]), "=r"(u[4]), "=r"(u[5]), "=r"(u[6]), "=r"(u[7]), "=r"(u[8]),
        "=r"(u[9]), "=r"(u[10]), "=r"(u[11]), "=r"(u[12]), "=r"(u[13]), "=r"(u[14]), "=r"(u[15]), "=r"(u[16]),
        "=r"(u[17]), "=r"(u[18]), "=r"(u[19]), "=r"(u[20]), "=r"(u[21]), "=r"(u[22]), "=r"(u[23]), "=r"(u[24]),
        "=r"(u[25]), "=r"(u[26]), "=r"(u[27]), "=r"(u[28]), "=r"(u[29]), "=r"(u[30]), "=r"(u[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void epi_bar_sync() { asm volatile("bar.sync 1, 128;" ::: "memory"); }

// 128B-swizzled K-major smem operand descriptor (rows of 128 B, 8-row atoms of 1024 B)
__device__ __forceinline__ uint64_t make_b_desc(uint32_t smem_addr) {
  uint64_t d = 0;
  d |= (uint64_t)((smem_addr & 0x3FFFFu) >> 4);  // start address
  d |= (uint64_t)1 << 16;                          // leading byte offset (unused for swizzled K-major)
  d |= (uint64_t)(1024 >> 4) << 32;                // stride byte offset: next 8-row atom
  d |= (uint64_t)1 << 46;                          // descriptor version (Blackwell)
  d |= (uint64_t)2 << 61;                          // SWIZZLE_128B
  return d;
}

__device__ __forceinline__ bool elect_one() {
  uint32_t pred;
  asm volatile("{\n\t.reg .pred p;\n\telect.sync _|p, 0xffffffff;\n\tselp.u32 %0, 1, 0, p;\n\t}" : "=r"(pred));
  return pred != 0;
}

struct FilterArgs {
  const __nv_bfloat16* qb;  // [n_qblocks*128][kpad]
  const float* aux;         // [n_rows]
  uint32_t nq, n_rows, kpad;
  uint32_t n_qblocks, n_chunks, tiles_per_chunk, n_tiles;
  float coef;               // -2 (L2) or -1 (IP, cosine)
  float* cand_score;        // [nq][n_chunks][KP]
  uint32_t* cand_row;       // [nq][n_chunks][KP]
  float* chunk_tau;         // [nq][n_chunks]
  float* dbg_scores;        // optional [nq][n_rows]
  uint32_t* hint;           // [nq] ord(lowest threshold published by a finished list of the query), 0xFFFFFFFF = none
  uint32_t hint_target;     // rows of the whole shard that should beat a published threshold (see publish_value)
  float* pub;               // optional float2 [lists_pad][nq_stride]: the two best scores every list of a query has found so far (see publish_two_best)
  uint32_t lists_pad;       // lists per query in `pub`, rounded up to 16 (never-written entries stay NaN)
  uint32_t nq_stride;       // queries per list row of `pub` (whole query blocks)
};


// KP   : candidates kept per (query, chunk, column slice); scores live in registers, rows in smem
// NBUF : accumulator buffers in TMEM (2 when the A operand leaves room, else 1)
// EW   : epilogue warps per TMEM lane quarter; each owns a slice of BN/EW columns of every tile.
//        Short K (small dim) makes the MMA of a tile cheaper than its gate, so more warps gate.
// ASM  : kpad > 768: a 128-query block no longer fits TMEM next to an accumulator, so the A operand
//        is streamed by TMA with the rows (one 16 KB A block + one 16 KB B block per stage, both
//        128B-swizzled K-major) and the MMAs take both operands from shared memory (bound by the
//        L2 -> SM fill rate at about 0.7 of the tensor roofline).
// BN   : database rows per tile (UMMA N). 128, or 64 for kpad in {640, 768}: with the 384-column A
//        operand only 128 TMEM columns are left, and ONE 128-column accumulator makes the MMA wait
//        while the epilogue pulls every finished tile out of TMEM (~12 % of the tile time). Two
//        64-column accumulators restore the overlap; a stage then holds 64 rows x 128 K (two TMA
//        boxes), so the MMA warp still issues 8 instructions per barrier round trip.
template <int KP, int NBUF, int EW, bool DBG, bool ASM, int BN>
struct FilterCfg {
  static constexpr int KSTEP = (BN == 64) ? 128 : 64;              // K elements per stage
  static constexpr int B_BYTES = BN * KSTEP * 2;                    // 16 KB either way
  static constexpr int STAGE_BYTES = (ASM ? TF_BM * TF_BK * 2 : 0) + B_BYTES;
  static constexpr int STAGES = ASM ? (KP > 16 ? 5 : 6) : ((KP > 16 && EW == 2) ? 9 : 10);  // what fits 227 KB next to the lists
  static_assert(!ASM || BN == 128, "the streamed-A variant uses 128-row tiles");
};

// packed fp32x2 FMA: {d0, d1} = c * {a0, a1} + {b0, b1} (one FFMA2; each half an IEEE fma)
__device__ __forceinline__ void fma2(unsigned long long c2, float a0, float a1, float b0, float b1, float& d0, float& d1) {
  unsigned long long r;
  asm("fma.rn.f32x2 %0, %1, %2, %3;" : "=l"(r) : "l"(c2), "l"(pack_f32x2(a0, a1)), "l"(pack_f32x2(b0, b1)));
  asm("mov.b64 {%0, %1}, %2;" : "=f"(d0), "=f"(d1) : "l"(r));
}
__device__ __forceinline__ float min3(float a, float b, float c) {
  float r;
  asm("min.f32 %0, %1, %2, %3;" : "=f"(r) : "f"(a), "f"(b), "f"(c));
  return r;
}

// What a finished list publishes as the starting threshold of the query's other lists. ANY value is
// admissible — a list's own final threshold stays a lower bound on every row it turned away, and the
// certificate decides from those bounds whether the reranked rows are provably the exact top-k. The
// default is the list's own k'-th best (it never costs a candidate). Option "tensor_hint_target" = T
// publishes the j-th best instead, j = ceil(T * share of the shard's rows the list saw): an estimate
// of the score that only T rows of the whole shard beat. Measured (C2, T = 96): the filter gets 8 %
// faster, but the minimum over two dozen such estimates is biased low, 131 of 10 000 queries lose
// their certificate and fall back to the exact scan (+18 ms) — hence off by default.
template <int KP>
__device__ __forceinline__ float publish_value(const float (&sc)[KP], uint32_t rows_list, uint32_t n_rows, uint32_t target) {
  const uint32_t j = min((uint32_t)KP, max(1u, (uint32_t)(((uint64_t)target * rows_list + n_rows - 1) / n_rows)));
  float pub = __int_as_float(0x7f800000);
#pragma unroll
  for (int t = 0; t < KP; ++t) {
    uint32_t rank = 0;
#pragma unroll
    for (int u = 0; u < KP; ++u) rank += (sc[u] < sc[t] || (sc[u] == sc[t] && u < t)) ? 1u : 0u;
    if (rank == j - 1) pub = sc[t];
  }
  return pub;   // +Inf when the list holds fewer than j rows
}

// Thresholds shared between the lists of a query WHILE they are being built. With few queries a query block is
// spread over dozens of row chunks that all start at the same time: every list converges on its own from +Inf
// (k' ln(rows / k') insertions), and because a warp gates 32 queries at once, almost every 16-column segment of
// such a list takes the insertion path — the filter of 256..1024 queries ran at half the rate its MMAs allow.
// Each list therefore publishes its two best scores whenever they may have changed (`pub`: one float2 per list
// and query, [lists_pad][nq_stride], so that the 32 queries of a warp store / load one 256-byte segment), and
// ONE extra warp per CTA — the bound helper, which has nothing else to do — keeps turning the table into a
// bound per query of the block the CTA is working on: list l feeds groups (l mod 16, rank) — 32 groups, each
// holding the score of a row of its own (a row belongs to one list, and the two best of a list are two rows).
// B = the maximum over the groups of the group minimum: 32 distinct rows score <= B, so the query's 32nd best
// score over the whole shard is <= B, and rows that score >= B cannot be among the k'' = 32 rows the rerank gets:
// B is an admissible threshold for every list of the query, and no theta ever drops below the final 32nd best
// score, which is (within one rank) the tau the certificate works with anyway. The helper leaves (item, B) pairs
// in shared memory (one 8-byte store each); an epilogue lane takes a bound only if it carries the item the lane
// is working on. Everything is racy on purpose: a value that was ever published stays the score of an existing
// row. Fewer than 16 lists: some group stays empty, B = +Inf.
template <int KP>
__device__ __forceinline__ void publish_two_best(const float (&sc)[KP], float* dst) {
  const float INF = __int_as_float(0x7f800000);
  float m1 = sc[0];
#pragma unroll
  for (int j = 1; j < KP; ++j) m1 = fminf(m1, sc[j]);
  float m2 = INF;   // (equal scores of two rows count twice)
  bool skipped = false;
#pragma unroll
  for (int j = 0; j < KP; ++j) {
    const bool skip = !skipped && sc[j] == m1;
    m2 = skip ? m2 : fminf(m2, sc[j]);
    skipped = skipped || skip;
  }
  asm volatile("st.global.cg.v2.f32 [%0], {%1, %2};" ::"l"(dst), "f"(m1), "f"(m2) : "memory");
}
constexpr uint32_t ITEM_NONE = 0xFFFFFFFEu, ITEM_EXIT = 0xFFFFFFFFu;
__device__ __forceinline__ uint32_t ld_shared_volatile_u32(const uint32_t* p) {
  uint32_t v;
  asm volatile("ld.volatile.shared.u32 %0, [%1];" : "=r"(v) : "r"(smem_u32(p)) : "memory");
  return v;
}
__device__ __forceinline__ void st_shared_volatile_u32(uint32_t* p, uint32_t v) {
  asm volatile("st.volatile.shared.u32 [%0], %1;" ::"r"(smem_u32(p)), "r"(v) : "memory");
}
// the (item, bound) pair of a query row of the block; +Inf unless it belongs to `item`
__device__ __forceinline__ float helper_bound_for(const uint2* s_hb, uint32_t qrow, uint32_t item) {
  uint32_t tag, bits;
  asm volatile("ld.volatile.shared.v2.u32 {%0, %1}, [%2];" : "=r"(tag), "=r"(bits) : "r"(smem_u32(s_hb + qrow)) : "memory");
  return tag == item ? __uint_as_float(bits) : __int_as_float(0x7f800000);
}
// The helper warp: until the CTA has finished its last item, bounds for the 128 queries starting at q0(item).
template <class ItemToQuery>
__device__ __forceinline__ void bound_helper_loop(const FilterArgs& a, const uint32_t* s_cur_item, uint2* s_hb, uint32_t lane, ItemToQuery q0_of) {
  const float INF = __int_as_float(0x7f800000);
  for (;;) {
    const uint32_t item = ld_shared_volatile_u32(s_cur_item);
    if (item == ITEM_EXIT) break;
    if (item == ITEM_NONE) {
      __nanosleep(500);
      continue;
    }
    const uint32_t q0 = q0_of(item);
    for (uint32_t r = 0; r < 4; ++r) {
      const float2* src = reinterpret_cast<const float2*>(a.pub) + (q0 + r * 32 + lane);   // (< nq_stride: the table is padded to whole blocks)
      float g0[16], g1[16];
#pragma unroll
      for (int u = 0; u < 16; ++u) g0[u] = g1[u] = INF;
      for (uint32_t l0 = 0; l0 < a.lists_pad; l0 += 16) {
        float2 v[16];
#pragma unroll
        for (int u = 0; u < 16; ++u)
          asm volatile("ld.global.cg.v2.f32 {%0, %1}, [%2];" : "=f"(v[u].x), "=f"(v[u].y) : "l"(src + (size_t)(l0 + u) * a.nq_stride));
#pragma unroll
        for (int u = 0; u < 16; ++u) {   // fminf drops the NaN of entries nobody has written yet
          g0[u] = fminf(g0[u], v[u].x);
          g1[u] = fminf(g1[u], v[u].y);
        }
      }
      float bound = fmaxf(g0[0], g1[0]);
#pragma unroll
      for (int u = 1; u < 16; ++u) bound = fmaxf(bound, fmaxf(g0[u], g1[u]));
      asm volatile("st.volatile.shared.v2.u32 [%0], {%1, %2};" ::"r"(smem_u32(s_hb + r * 32 + lane)), "r"(item), "r"(__float_as_uint(bound)) : "memory");
    }
    __nanosleep(3000);
  }
}

// SHARE: the instantiation with the bound helper warp and the publication code (see publish_two_best); without it the
//        kernel is exactly the one of the first half of round 2 (the extra warp and code cost the gate-bound short rows
//        about 15 % at big batches, where sharing is off anyway)
template <int KP, int NBUF, int EW, bool DBG, bool ASM, int BN, bool SHARE>
__global__ void __launch_bounds__(64 + 128 * EW + (SHARE ? 32 : 0), 1)
    tensor_filter_kernel(const __grid_constant__ CUtensorMap tmap_b, const __grid_constant__ CUtensorMap tmap_a, FilterArgs a) {
  using Cfg = FilterCfg<KP, NBUF, EW, DBG, ASM, BN>;
  constexpr int TF_BN = BN;
  constexpr int TF_STAGES = Cfg::STAGES;
  constexpr int TF_STAGE_BYTES = Cfg::STAGE_BYTES;
  constexpr int KSTEP = Cfg::KSTEP;
  constexpr int TF_B_OFF = ASM ? (TF_BM * TF_BK * 2) : 0;  // B block within a stage (A block first when streamed)
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  // carve-up: [B stages][cand_row 128*EW*KP][queue 2*16*128*EW][aux 8*BN][barriers][tmem ptr]
  // 1024-byte alignment for the 128B-swizzled tiles; plain pointer arithmetic on the __shared__
  // array keeps the shared address space visible to the compiler (LDS/STS, not generic LD/ST)
  unsigned char* sb = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  uint32_t* s_crow = reinterpret_cast<uint32_t*>(sb + TF_STAGES * TF_STAGE_BYTES);
  constexpr int ET = 128 * EW;                                     // epilogue threads
  float* s_qs = reinterpret_cast<float*>(s_crow + ET * KP);        // [16][ET] queued scores
  uint32_t* s_qc = reinterpret_cast<uint32_t*>(s_qs + 16 * ET);    // [16][ET] queued columns
  float* s_aux = reinterpret_cast<float*>(s_qc + 16 * ET);
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_aux + 8 * TF_BN);  // 4*EW warps x 2 buffers x BN/EW columns
  uint64_t* full = bars;                     // [STAGES]  TMA -> MMA
  uint64_t* empty = full + TF_STAGES;        // [STAGES]  MMA -> TMA
  uint64_t* acc_full = empty + TF_STAGES;    // [2]       MMA -> epilogue
  uint64_t* acc_empty = acc_full + 2;        // [2]       epilogue -> MMA
  uint64_t* a_ready = acc_empty + 2;         // [1]       epilogue (A stored) -> MMA
  uint2* s_hb = reinterpret_cast<uint2*>(a_ready + 1);            // [128] (item, bound) per query row, written by the bound helper
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_hb + 128);
  uint32_t* s_cur_item = s_tmem + 1;                              // the item the epilogue is working on (for the bound helper)

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t KB = a.kpad / KSTEP;   // stages per tile

  if (threadIdx.x < 128) s_hb[threadIdx.x] = make_uint2(ITEM_NONE, 0x7f800000u);
  if (threadIdx.x == 0) {
    *s_cur_item = ITEM_NONE;
    for (int i = 0; i < TF_STAGES; ++i) {
      mbar_init(full + i, 1);
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, 4 * EW);  // one arrive per epilogue warp
    }
    mbar_init(a_ready, 4 * EW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // TMEM: all 512 columns (one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_acc = tmem_base;                    // columns [0, NBUF*BN)
  const uint32_t tmem_a = tmem_base + NBUF * TF_BN;       // columns [NBUF*BN, NBUF*BN + kpad/2)

  const uint32_t n_items = a.n_qblocks * a.n_chunks;
  pdl_trigger();   // the kernels behind the filter may become resident as SMs free up (they wait for this grid to complete)

  if (warp == 0) {
    // ===== TMA producer (warp-uniform control flow; one elected lane issues) =====
    // (the mirror is not written by the kernel before this one: the row tiles start to arrive at once)
    if (ASM) pdl_wait();   // streamed queries: the A tiles are the bf16 rows prep_queries writes
    uint32_t stage = 0, phase = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint32_t chunk = item / a.n_qblocks;
      const uint32_t qblk = item - chunk * a.n_qblocks;
      const uint32_t t0 = chunk * a.tiles_per_chunk;
      const uint32_t t1 = min(a.n_tiles, t0 + a.tiles_per_chunk);
      for (uint32_t t = t0; t < t1; ++t) {
        for (uint32_t kb = 0; kb < KB; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          if (elect_one()) {
            mbar_expect_tx(full + stage, TF_STAGE_BYTES);
            if (ASM) tma_load_2d(sb + stage * TF_STAGE_BYTES, &tmap_a, full + stage, (int)(kb * TF_BK), (int)(qblk * TF_BM));
#pragma unroll
            for (int h = 0; h < KSTEP / TF_BK; ++h)  // one 128-byte-wide box per 64 K elements
              tma_load_2d(sb + stage * TF_STAGE_BYTES + TF_B_OFF + h * (TF_BN * TF_BK * 2), &tmap_b, full + stage,
                          (int)(kb * KSTEP + h * TF_BK), (int)(t * TF_BN));
          }
          __syncwarp();
          if (++stage == TF_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (warp-uniform control flow; one elected lane issues) =====
    // instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=128
    const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TF_BN >> 3) << 17) | ((uint32_t)(TF_BM >> 4) << 24);
    const uint64_t bdesc0 = make_b_desc(smem_u32(sb));
    uint32_t stage = 0, phase = 0, acc_it = 0, a_phase = 0;
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint32_t chunk = item / a.n_qblocks;
      const uint32_t t0 = chunk * a.tiles_per_chunk;
      const uint32_t t1 = min(a.n_tiles, t0 + a.tiles_per_chunk);
      if (!ASM) {
        mbar_wait(a_ready, a_phase);
        a_phase ^= 1;
        tc_fence_after();
      }
      for (uint32_t t = t0; t < t1; ++t, ++acc_it) {
        const uint32_t buf = (NBUF == 2) ? (acc_it & 1) : 0;
        const uint32_t use = (NBUF == 2) ? (acc_it >> 1) : acc_it;
        mbar_wait(acc_empty + buf, (use & 1) ^ 1);
        tc_fence_after();
        const uint32_t d_tmem = tmem_acc + buf * TF_BN;
        for (uint32_t kb = 0; kb < KB; ++kb) {
          mbar_wait(full + stage, phase);
          tc_fence_after();
          if (elect_one()) {
            const uint64_t adesc = bdesc0 + (uint64_t)(stage * (TF_STAGE_BYTES >> 4));
            const uint64_t bdesc = adesc + (uint64_t)(TF_B_OFF >> 4);
            const uint32_t a_col = tmem_a + kb * (KSTEP / 2);
#pragma unroll
            for (uint32_t k = 0; k < KSTEP / 16; ++k) {
              // A: 16 bf16 of K = 8 TMEM columns (or 32 bytes along the swizzled smem row);
              // B: 32 bytes further along the swizzled row, next 64-K box after four steps
              const uint64_t boff = (uint64_t)((k >> 2) * ((TF_BN * TF_BK * 2) >> 4) + (k & 3) * 2);
              if (ASM) tc_mma_ss(d_tmem, adesc + (uint64_t)(k * 2), bdesc + boff, idesc, (kb | k) != 0);
              else tc_mma_ts(d_tmem, a_col + k * 8, bdesc + boff, idesc, (kb | k) != 0);
            }
            tc_commit(empty + stage);                     // smem slot free once these MMAs retire
            if (kb == KB - 1) tc_commit(acc_full + buf);  // accumulator ready for the epilogue
          }
          __syncwarp();
          if (++stage == TF_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (SHARE && warp == 2 + 4 * EW) {
    // ===== bound helper (see publish_two_best) =====
    if (a.pub) {
      pdl_wait();   // the table is initialised by prep_queries
      bound_helper_loop(a, s_cur_item, s_hb, (uint32_t)lane, [&](uint32_t item) { return (item % a.n_qblocks) * (uint32_t)TF_BM; });
    }
  } else {
    // ===== epilogue warps: thread <-> query (TMEM lane) =====
    const uint32_t quarter = warp & 3;                 // TMEM lane quarter this warp may access
    const uint32_t qrow = quarter * 32 + lane;         // query row within the block == TMEM lane
    const uint32_t et = (warp - 2) * 32 + lane;        // 0..ET-1 among epilogue threads
    const uint32_t slice = (warp - 2) >> 2;            // which 128/EW columns of a tile this warp gates
    constexpr int CW = TF_BN / EW;                     // columns per warp
    const uint32_t lane_addr = (quarter * 32) << 16;
    const float INF = __int_as_float(0x7f800000);
    pdl_wait();   // bf16 queries and hints come from prep_queries (chained launch, common.cuh)
    uint32_t* my_row = s_crow + slice * 128 + qrow;    // slot j at [j*ET]
    float* my_qs = s_qs + slice * 128 + qrow;
    uint32_t* my_qc = s_qc + slice * 128 + qrow;
    uint32_t acc_it = 0;
    const unsigned long long coef2 = pack_f32x2(a.coef, a.coef);
    for (uint32_t item = blockIdx.x; item < n_items; item += gridDim.x) {
      const uint32_t chunk = item / a.n_qblocks;
      const uint32_t qblk = item - chunk * a.n_qblocks;
      const uint32_t t0 = chunk * a.tiles_per_chunk;
      const uint32_t t1 = min(a.n_tiles, t0 + a.tiles_per_chunk);
      const uint32_t q_global = qblk * TF_BM + qrow;
      // ---- A operand: this thread's query row -> its TMEM lane, packed bf16 pairs ----
      if (!ASM) {
        const uint4* src = reinterpret_cast<const uint4*>(a.qb + (size_t)q_global * a.kpad);
        const uint32_t n16 = a.kpad / 8;  // 16-byte groups = 4 TMEM columns each
        for (uint32_t i0 = slice; i0 < n16; i0 += 8 * EW) {  // the EW warps of a quarter share the copy; 8 loads in flight per lane
          uint4 v[8];
#pragma unroll
          for (int u = 0; u < 8; ++u) v[u] = (i0 + u * EW < n16) ? __ldg(src + i0 + u * EW) : make_uint4(0, 0, 0, 0);
#pragma unroll
          for (int u = 0; u < 8; ++u)
            if (i0 + u * EW < n16) tc_st4(tmem_a + lane_addr + (i0 + u * EW) * 4, v[u].x, v[u].y, v[u].z, v[u].w);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(a_ready);
      }
      // ---- per-query candidate set: KP scores in registers (unsorted), rows in smem ----
      float sc[KP];
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        sc[j] = INF;
        my_row[j * ET] = ROW_NONE;
      }
      // Admission threshold. A list that has finished publishes its final threshold (its KP-th best
      // score): the query's overall KP-th best can only be lower, so that value is a valid starting
      // threshold for every other list of the same query — rows it rejects could never have made
      // the final cut, and `theta` stays a lower bound on the score of every rejected row, which is
      // all the certificate needs. Lists of later waves therefore admit only a handful of rows
      // instead of re-converging from +Inf (KP*ln(rows/KP) insertions each).
      float hint = INF;
      if (a.hint && q_global < a.nq) {
        uint32_t h;
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(h) : "l"(a.hint + q_global));
        if (h != 0xFFFFFFFFu) hint = ord_f32(h);
      }
      float theta = hint;  // min(hint, max of sc[]): the threshold every admitted row must beat
      int imax = 0;        // a slot holding theta
      const bool sharing = SHARE && a.pub != nullptr;   // (see publish_two_best)
      float* my_pub = a.pub + ((size_t)(chunk * EW + slice) * a.nq_stride + q_global) * 2;   // (q_global < nq_stride)
      bool dirty = false;                      // the list changed since it was last published
      if (sharing && warp == 2 && lane == 0) st_shared_volatile_u32(s_cur_item, item);
      // additive per-column term for the first tile, fetched one tile ahead from here on
      // Each warp stages the term of its own CW columns (lanes 0..CW/4-1, one float4 each) in a
      // private double buffer: no block barrier per tile, so the epilogue warps are coupled only
      // through the accumulator barriers and one warp's insertions no longer stall the other seven.
      float* my_aux = s_aux + (warp - 2) * 2 * CW;
      auto load_aux = [&](uint32_t t) -> float4 {
        float4 r = make_float4(INF, INF, INF, INF);
        if (lane < CW / 4 && t < t1) {
          const uint32_t c = t * TF_BN + slice * CW + lane * 4;
          if (c + 3 < a.n_rows) {
            r = __ldg(reinterpret_cast<const float4*>(a.aux + c));
          } else {
            if (c + 0 < a.n_rows) r.x = __ldg(a.aux + c + 0);
            if (c + 1 < a.n_rows) r.y = __ldg(a.aux + c + 1);
            if (c + 2 < a.n_rows) r.z = __ldg(a.aux + c + 2);
          }
        }
        return r;
      };
      float4 aux_next = load_aux(t0);
      for (uint32_t t = t0; t < t1; ++t, ++acc_it) {
        const uint32_t buf = (NBUF == 2) ? (acc_it & 1) : 0;
        const uint32_t use = (NBUF == 2) ? (acc_it >> 1) : acc_it;
        const uint32_t col0 = t * TF_BN;
        float* aux_t = my_aux + (acc_it & 1) * CW;
        // per-column additive term (||x~||^2, 0, or +Inf for deleted / out-of-range rows)
        if (lane < CW / 4) reinterpret_cast<float4*>(aux_t)[lane] = aux_next;
        aux_next = load_aux(t + 1);
        __syncwarp();
        mbar_wait(acc_full + buf, use & 1);
        tc_fence_after();
        float v[CW];
#pragma unroll
        for (int g = 0; g < CW / 32; ++g) tc_ld32(tmem_acc + lane_addr + buf * TF_BN + slice * CW + g * 32, v + g * 32);
        tc_wait_ld();
        // the accumulator now lives in registers: hand the TMEM buffer back before processing
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive(acc_empty + buf);
        // Gate four columns at a time (min of 4 against theta); survivors are queued in shared
        // memory and inserted by one loop per 16-column segment, so the unrolled code stays small.
        const float4* aux4 = reinterpret_cast<const float4*>(aux_t);
        const uint32_t colw = col0 + slice * CW;  // first column of this warp's slice
        float4 ax[4];
#pragma unroll
        for (int i = 0; i < 4; ++i) ax[i] = aux4[i];
#pragma unroll
        for (int seg = 0; seg < CW / 16; ++seg) {
          float4 axn[4];
          if (seg + 1 < CW / 16) {
#pragma unroll
            for (int i = 0; i < 4; ++i) axn[i] = aux4[(seg + 1) * 4 + i];
          }
          // All 16 scores of the segment first (8 FFMA2), then the minimum of each group of four
          // and of the segment (FMNMX3): straight-line code with ONE branch per segment, taken only
          // when some column beats the threshold (per warp about once per tile in steady state).
          float sg[16], mg[4];
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const int j = seg * 16 + g4 * 4;
            fma2(coef2, v[j + 0], v[j + 1], ax[g4].x, ax[g4].y, sg[g4 * 4 + 0], sg[g4 * 4 + 1]);
            fma2(coef2, v[j + 2], v[j + 3], ax[g4].z, ax[g4].w, sg[g4 * 4 + 2], sg[g4 * 4 + 3]);
            if (DBG) {
              v[j + 0] = sg[g4 * 4 + 0];
              v[j + 1] = sg[g4 * 4 + 1];
              v[j + 2] = sg[g4 * 4 + 2];
              v[j + 3] = sg[g4 * 4 + 3];
            }
            mg[g4] = fminf(min3(sg[g4 * 4 + 0], sg[g4 * 4 + 1], sg[g4 * 4 + 2]), sg[g4 * 4 + 3]);
          }
          if (fminf(min3(mg[0], mg[1], mg[2]), mg[3]) < theta) {
            uint32_t cnt = 0;
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const int j = seg * 16 + g4 * 4;
              if (mg[g4] < theta) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (sg[g4 * 4 + e] < theta) {
                    my_qs[cnt * ET] = sg[g4 * 4 + e];
                    my_qc[cnt * ET] = colw + j + e;
                    ++cnt;
                  }
                }
              }
            }
#pragma unroll 1
            for (uint32_t i = 0; i < cnt; ++i) {
              const float s = my_qs[i * ET];
              if (s < theta) {
                // replace the current worst, then find the new worst by a tree arg-max (no ordering
                // is needed here: the merge kernel sorts; rows with s >= theta are exactly the rejected ones)
                my_row[imax * ET] = my_qc[i * ET];
                float mv[KP];
                int mi[KP];
#pragma unroll
                for (int t2 = 0; t2 < KP; ++t2) {
                  sc[t2] = (t2 == imax) ? s : sc[t2];
                  mv[t2] = sc[t2];
                  mi[t2] = t2;
                }
#pragma unroll
                for (int w = KP / 2; w >= 1; w >>= 1) {
#pragma unroll
                  for (int t2 = 0; t2 < w; ++t2) {
                    const bool gt = mv[t2 + w] > mv[t2];
                    mv[t2] = gt ? mv[t2 + w] : mv[t2];
                    mi[t2] = gt ? mi[t2 + w] : mi[t2];
                  }
                }
                theta = fminf(mv[0], hint);
                imax = mi[0];
                if (SHARE) dirty = true;
              }
            }
          }
          if (seg + 1 < CW / 16) {
#pragma unroll
            for (int i = 0; i < 4; ++i) ax[i] = axn[i];
          }
        }
        if (DBG && a.dbg_scores && q_global < a.nq) {
#pragma unroll
          for (int j = 0; j < CW; ++j) {
            uint32_t c = colw + j;
            if (c < a.n_rows) a.dbg_scores[(size_t)q_global * a.n_rows + c] = v[j];
          }
        }
        if (sharing) {   // what the list found goes out, what the helper made of everybody's lists comes in
          if (dirty && ((t - t0) & 1u)) {
            publish_two_best<KP>(sc, my_pub);
            dirty = false;
          }
          hint = fminf(hint, helper_bound_for(s_hb, qrow, item));
          theta = fminf(theta, hint);
        }
      }
      // ---- chunk result ----
      if (q_global < a.nq) {
        // every (chunk, column slice) pair is its own candidate list: n_chunks*EW lists per query
        const size_t vchunk = (size_t)chunk * EW + slice;
        size_t base = ((size_t)q_global * a.n_chunks * EW + vchunk) * KP;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          a.cand_score[base + j] = sc[j];
          a.cand_row[base + j] = my_row[j * ET];
        }
        a.chunk_tau[(size_t)q_global * a.n_chunks * EW + vchunk] = theta;
        if (sharing) publish_two_best<KP>(sc, my_pub);   // the list's final two best, for the lists of later waves
        if (a.hint) {
          const float pub = a.hint_target == 0 ? theta
                                               : fminf(theta, publish_value<KP>(sc, (min(a.n_rows, t1 * (uint32_t)TF_BN) - t0 * (uint32_t)TF_BN) / EW,
                                                                                a.n_rows, a.hint_target));
          if (pub < INF) atomicMin(a.hint + q_global, f32_ord(pub));
        }
      }
      // all MMAs of this item have retired (the last acc_full was waited on), so A may be rewritten
    }
    if (SHARE && warp == 2 && lane == 0) st_shared_volatile_u32(s_cur_item, ITEM_EXIT);   // releases the bound helper
  }

  tc_fence_before();
  __syncthreads();
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

// ================================================================================================
// tensor_filter2_kernel — the same filter on CTA PAIRS (tcgen05 cta_group::2), for kpad in {512, 640, 768}.
// Measured with tools/ubench/mma_rate.cu: an M=128 N=64 K=16 instruction takes 44.6 cycles (72 % of the
// tensor rate), N=128 takes exactly 64 (100 %). Next to the 384-column A operand a single CTA has room
// for two 64-column accumulators or one of 128 columns, and with N = 128 it would have to pull 64 B/clk
// of rows out of L2, more than one SM gets. A pair of CTAs (one cluster, one TPC) removes both limits:
// each CTA keeps ITS 128 queries stationary in its own TMEM and loads only HALF of every 128-row tile
// (64 rows per stage, 32 B/clk at full rate); one lane of the leader CTA issues
// tcgen05.mma.cta_group::2 M=256 N=128 K=16 — 64 cycles on both SMs, each reading the other's rows
// inside the TPC — and each CTA finds the scores of its 128 queries against all 128 rows in its own
// 128-column accumulator.
//   barriers   full[s]      leader CTA only; both CTAs' TMA transactions and one arrive each land here
//              empty[s]     one per CTA; the MMA commit is multicast to both
//              acc_full     one per CTA; multicast commit after the last K stage of a tile
//              acc_empty    leader CTA only; the 4 epilogue warps of BOTH CTAs arrive there once the
//                           accumulator is in their registers (one accumulator: the next tile's MMAs
//                           wait for this, ~15 % of a tile; the gate itself overlaps the next tile)
//              a_ready      leader CTA only; both CTAs' epilogue warps arrive after storing A
// The epilogue is the one above (threshold gate, k' best per query and chunk).
// stages of 16 KB each (64 rows x 128 K bf16, two 64-K boxes): what fits 227 KB next to the lists and queues
__host__ __device__ constexpr int filter2_stages(int kp, int ew) { return (ew == 2 && kp > 16) ? 9 : 10; }
constexpr int TF2_STAGE_BYTES = 64 * 128 * 2;
constexpr int TF2_BN = 128;                                // rows per pair tile (UMMA N); 64 loaded per CTA
constexpr int TF2_KSTEP = 128;

__device__ __forceinline__ uint32_t cluster_ctarank() {
  uint32_t r;
  asm volatile("mov.u32 %0, %%cluster_ctarank;" : "=r"(r));
  return r;
}
__device__ __forceinline__ void cluster_sync_all() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n\tbarrier.cluster.wait.acquire.aligned;" ::: "memory");
}
// shared::cluster address of `p` (a shared::cta address of this CTA) in CTA `rank` of the cluster
__device__ __forceinline__ uint32_t mapa_u32(uint32_t addr, uint32_t rank) {
  uint32_t r;
  asm volatile("mapa.shared::cluster.u32 %0, %1, %2;" : "=r"(r) : "r"(addr), "r"(rank));
  return r;
}
// (default semantics on purpose: `.release.cluster` compiles to MEMBAR.ALL.GPU and `.acquire.cluster` on the
// waiting side to CCTL.IVALL — an L1 flush per barrier round trip, measured at 40 % of the kernel. What the
// barriers order here is TMEM / async-proxy traffic, which tcgen05.fence and tcgen05.commit cover.)
__device__ __forceinline__ void mbar_arrive_cluster(uint32_t cluster_addr) {
  asm volatile("mbarrier.arrive.shared::cluster.b64 _, [%0];" ::"r"(cluster_addr) : "memory");
}
// TMA tile load whose completion is signalled on a barrier that may live in the peer CTA of the pair
__device__ __forceinline__ void tma_load_2d_pair(void* smem_dst, const CUtensorMap* tmap, uint32_t bar_cluster_addr, int c0, int c1) {
  asm volatile(
      "cp.async.bulk.tensor.2d.cta_group::2.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(smem_dst)),
      "l"(tmap), "r"(bar_cluster_addr), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_commit_pair(uint64_t* bar) {   // arrives on `bar` in BOTH CTAs of the pair
  asm volatile("tcgen05.commit.cta_group::2.mbarrier::arrive::one.shared::cluster.multicast::cluster.b64 [%0], %1;" ::"r"(smem_u32(bar)),
               "h"((uint16_t)3)
               : "memory");
}
// D[tmem] (+)= A[tmem] * B[smem]^T on both SMs of the pair: M = 256 (128 TMEM lanes per CTA), B split by rows
__device__ __forceinline__ void tc_mma_ts_pair(uint32_t d_tmem, uint32_t a_tmem, uint64_t b_desc, uint32_t idesc, uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::2.kind::f16 [%0], [%1], %2, %3, p;\n\t}" ::"r"(d_tmem),
      "r"(a_tmem), "l"(b_desc), "r"(idesc), "r"(accumulate)
      : "memory");
}

// NBUF: accumulators per CTA — 2 when the A operand leaves room (kpad <= 512), else 1
// EW  : epilogue warps per TMEM lane quarter, each gating 128 / EW columns of every tile into a list of its own. An
//       epilogue warp is alone on its scheduler and runs one dependent chain (ncu: one instruction per six cycles), so
//       a tile costs it 2200 cycles when hardly any row beats the thresholds and two to three times that while the
//       lists are cold — more than the 3072 cycles of the tile's MMAs. Two warps per quarter halve that: for work items
//       that are short (small shards) or few (a few hundred queries), where cold lists are the rule.
template <int KP, int NBUF, int EW>
__global__ void __cluster_dims__(2, 1, 1) __launch_bounds__(96 + 128 * EW, 1)
    tensor_filter2_kernel(const __grid_constant__ CUtensorMap tmap_b, FilterArgs a) {
  extern __shared__ __align__(1024) unsigned char smem_raw[];
  unsigned char* sb = smem_raw + ((1024u - (smem_u32(smem_raw) & 1023u)) & 1023u);
  constexpr int ET = 128 * EW;                                             // epilogue threads per CTA
  constexpr int TF2_STAGES = filter2_stages(KP, EW);
  constexpr int CW = TF2_BN / EW;                                          // columns of a tile per epilogue warp
  uint32_t* s_crow = reinterpret_cast<uint32_t*>(sb + TF2_STAGES * TF2_STAGE_BYTES);
  float* s_qs = reinterpret_cast<float*>(s_crow + ET * KP);                // [16][ET] queued scores
  uint32_t* s_qc = reinterpret_cast<uint32_t*>(s_qs + 16 * ET);            // [16][ET] queued columns
  float* s_aux = reinterpret_cast<float*>(s_qc + 16 * ET);                 // [4 EW warps][2 buffers][CW]
  uint64_t* bars = reinterpret_cast<uint64_t*>(s_aux + 4 * 2 * TF2_BN);
  uint64_t* full = bars;                       // [STAGES]
  uint64_t* empty = full + TF2_STAGES;         // [STAGES]
  uint64_t* acc_full = empty + TF2_STAGES;     // [2]
  uint64_t* acc_empty = acc_full + 2;          // [2]
  uint64_t* a_ready = acc_empty + 2;           // [1]
  uint2* s_hb = reinterpret_cast<uint2*>(a_ready + 1);            // [128] (item, bound) per query row, written by the bound helper (the last warp)
  uint32_t* s_tmem = reinterpret_cast<uint32_t*>(s_hb + 128);
  uint32_t* s_cur_item = s_tmem + 1;

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const uint32_t rank = cluster_ctarank();     // 0 = leader (issues the MMAs)
  const uint32_t KB = a.kpad / TF2_KSTEP;      // stages per tile
  const uint32_t n_pairs = (a.n_qblocks + 1) / 2;
  const uint32_t n_items = n_pairs * a.n_chunks;
  const uint32_t cluster_id = blockIdx.x >> 1, n_clusters = gridDim.x >> 1;

  if (threadIdx.x < 128) s_hb[threadIdx.x] = make_uint2(ITEM_NONE, 0x7f800000u);
  if (threadIdx.x == 0) {
    *s_cur_item = ITEM_NONE;
    for (int i = 0; i < TF2_STAGES; ++i) {
      mbar_init(full + i, 2);                  // one arrive per CTA of the pair (+ the transaction bytes of both)
      mbar_init(empty + i, 1);
    }
    for (int i = 0; i < 2; ++i) {
      mbar_init(acc_full + i, 1);
      mbar_init(acc_empty + i, 8 * EW);        // the epilogue warps of both CTAs
    }
    mbar_init(a_ready, 8 * EW);
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 2) {  // TMEM: all 512 columns of both SMs of the pair
    asm volatile("tcgen05.alloc.cta_group::2.sync.aligned.shared::cta.b32 [%0], %1;" ::"r"(smem_u32(s_tmem)), "r"(512u) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::2.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  cluster_sync_all();                          // the peer's barriers are initialised before anything is signalled there
  tc_fence_after();
  const uint32_t tmem_base = *s_tmem;
  const uint32_t tmem_acc = tmem_base;                  // columns [0, NBUF*128)
  const uint32_t tmem_a = tmem_base + NBUF * TF2_BN;    // columns [NBUF*128, NBUF*128 + kpad/2)
  pdl_trigger();

  if (warp == 0) {
    // ===== TMA producer: this CTA's 64 rows of every 128-row tile =====
    uint32_t stage = 0, phase = 0;
    for (uint32_t item = cluster_id; item < n_items; item += n_clusters) {
      const uint32_t chunk = item / n_pairs;
      const uint32_t t0 = chunk * a.tiles_per_chunk;
      const uint32_t t1 = min(a.n_tiles, t0 + a.tiles_per_chunk);
      for (uint32_t t = t0; t < t1; ++t) {
        const int x_row0 = (int)(t * TF2_BN + rank * 64);
        for (uint32_t kb = 0; kb < KB; ++kb) {
          mbar_wait(empty + stage, phase ^ 1);
          if (elect_one()) {
            const uint32_t full_leader = mapa_u32(smem_u32(full + stage), 0);
            if (rank == 0) mbar_expect_tx(full + stage, 2 * TF2_STAGE_BYTES);   // both CTAs' bytes
            else mbar_arrive_cluster(full_leader);
            unsigned char* st = sb + stage * TF2_STAGE_BYTES;
#pragma unroll
            for (int h = 0; h < TF2_KSTEP / TF_BK; ++h)   // one 128-byte-wide box per 64 K elements
              tma_load_2d_pair(st + h * (64 * TF_BK * 2), &tmap_b, full_leader, (int)(kb * TF2_KSTEP + h * TF_BK), x_row0);
          }
          __syncwarp();
          if (++stage == TF2_STAGES) {
            stage = 0;
            phase ^= 1;
          }
        }
      }
    }
  } else if (warp == 1) {
    if (rank == 0) {
      // ===== MMA issuer: one lane of the leader CTA drives the tensor cores of both SMs =====
      // instruction descriptor: D=f32, A=B=bf16, both K-major, N=128, M=256
      const uint32_t idesc = (1u << 4) | (1u << 7) | (1u << 10) | ((uint32_t)(TF2_BN >> 3) << 17) | ((uint32_t)(256 >> 4) << 24);
      const uint64_t desc0 = make_b_desc(smem_u32(sb));
      uint32_t stage = 0, phase = 0, acc_it = 0, a_phase = 0;
      for (uint32_t item = cluster_id; item < n_items; item += n_clusters) {
        const uint32_t chunk = item / n_pairs;
        const uint32_t t0 = chunk * a.tiles_per_chunk;
        const uint32_t t1 = min(a.n_tiles, t0 + a.tiles_per_chunk);
        mbar_wait(a_ready, a_phase);   // both CTAs have their query blocks in TMEM
        a_phase ^= 1;
        tc_fence_after();
        for (uint32_t t = t0; t < t1; ++t, ++acc_it) {
          const uint32_t buf = (NBUF == 2) ? (acc_it & 1) : 0;
          const uint32_t use = (NBUF == 2) ? (acc_it >> 1) : acc_it;
          mbar_wait(acc_empty + buf, (use & 1) ^ 1);
          tc_fence_after();
          const uint32_t d_tmem = tmem_acc + buf * TF2_BN;
          for (uint32_t kb = 0; kb < KB; ++kb) {
            mbar_wait(full + stage, phase);
            tc_fence_after();
            if (elect_one()) {
              const uint64_t bdesc = desc0 + (uint64_t)(stage * (TF2_STAGE_BYTES >> 4));
              const uint32_t a_col = tmem_a + kb * (TF2_KSTEP / 2);
#pragma unroll
              for (uint32_t k = 0; k < TF2_KSTEP / 16; ++k) {
                // A: 16 bf16 of K = 8 TMEM columns; B: 32 bytes further along the swizzled row, next 64-K box after four steps
                const uint64_t boff = (uint64_t)((k >> 2) * ((64 * TF_BK * 2) >> 4) + (k & 3) * 2);
                tc_mma_ts_pair(d_tmem, a_col + k * 8, bdesc + boff, idesc, (kb | k) != 0);
              }
              tc_commit_pair(empty + stage);                 // the slot is free in both CTAs once these MMAs retire
              if (kb == KB - 1) tc_commit_pair(acc_full + buf);   // both epilogues may read their accumulators
            }
            __syncwarp();
            if (++stage == TF2_STAGES) {
              stage = 0;
              phase ^= 1;
            }
          }
        }
      }
    }
  } else if (warp == 2 + 4 * EW) {
    // ===== bound helper (see publish_two_best): the 128 queries of THIS CTA =====
    if (a.pub) {
      pdl_wait();   // the table is initialised by prep_queries
      bound_helper_loop(a, s_cur_item, s_hb, (uint32_t)lane, [&](uint32_t item) { return ((item % n_pairs) * 2 + rank) * (uint32_t)TF_BM; });
    }
  } else {
    // ===== epilogue warps: thread <-> query (TMEM lane of this CTA) =====
    const uint32_t quarter = warp & 3;
    const uint32_t qrow = quarter * 32 + lane;
    const uint32_t slice = (warp - 2) >> 2;    // which CW columns of a tile this warp gates
    const uint32_t lane_addr = (quarter * 32) << 16;
    const float INF = __int_as_float(0x7f800000);
    pdl_wait();   // bf16 queries and hints come from prep_queries (chained launch, common.cuh)
    uint32_t* my_row = s_crow + slice * 128 + qrow;   // slot j at [j*ET]
    float* my_qs = s_qs + slice * 128 + qrow;
    uint32_t* my_qc = s_qc + slice * 128 + qrow;
    float* my_aux = s_aux + (warp - 2) * 2 * CW;
    const uint32_t acc_empty_leader = mapa_u32(smem_u32(acc_empty), 0);
    const uint32_t a_ready_leader = mapa_u32(smem_u32(a_ready), 0);
    uint32_t acc_it = 0;
    const unsigned long long coef2 = pack_f32x2(a.coef, a.coef);
    for (uint32_t item = cluster_id; item < n_items; item += n_clusters) {
      const uint32_t chunk = item / n_pairs;
      const uint32_t pair = item - chunk * n_pairs;
      const uint32_t t0 = chunk * a.tiles_per_chunk;
      const uint32_t t1 = min(a.n_tiles, t0 + a.tiles_per_chunk);
      const uint32_t q_global = (pair * 2 + rank) * TF_BM + qrow;   // (< n_qblocks_pad * 128: the bf16 query array is padded to whole pairs)
      // ---- A operand: this thread's query row -> its TMEM lane, packed bf16 pairs. All MMAs of the previous
      // item have retired (its last acc_full was waited on), so A may be rewritten.
      {
        const uint4* src = reinterpret_cast<const uint4*>(a.qb + (size_t)q_global * a.kpad);
        const uint32_t n16 = a.kpad / 8;   // 16-byte groups = 4 TMEM columns each; kpad is a multiple of 128: n16 of 16
        for (uint32_t i0 = slice * 16; i0 < n16; i0 += 16 * EW) {   // 16 independent loads in flight per lane before the first store; the warps of a quarter share the copy
          uint4 v[16];
#pragma unroll
          for (int u = 0; u < 16; ++u) v[u] = __ldg(src + i0 + u);
#pragma unroll
          for (int u = 0; u < 16; ++u) tc_st4(tmem_a + lane_addr + (i0 + u) * 4, v[u].x, v[u].y, v[u].z, v[u].w);
        }
        tc_wait_st();
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(a_ready_leader);
      }
      float sc[KP];
#pragma unroll
      for (int j = 0; j < KP; ++j) {
        sc[j] = INF;
        my_row[j * ET] = ROW_NONE;
      }
      float hint = INF;
      if (a.hint && q_global < a.nq) {
        uint32_t h;
        asm volatile("ld.volatile.global.u32 %0, [%1];" : "=r"(h) : "l"(a.hint + q_global));
        if (h != 0xFFFFFFFFu) hint = ord_f32(h);
      }
      float theta = hint;
      int imax = 0;
      const bool sharing = a.pub != nullptr;   // (see publish_two_best)
      float* my_pub = a.pub + ((size_t)(chunk * EW + slice) * a.nq_stride + q_global) * 2;   // (q_global < nq_stride)
      bool dirty = false;                      // the list changed since it was last published
      if (sharing && warp == 2 && lane == 0) st_shared_volatile_u32(s_cur_item, item);
      // additive term of this warp's CW columns of a tile: one float4 per lane, fetched one tile ahead
      auto load_aux = [&](uint32_t t) -> float4 {
        float4 r = make_float4(INF, INF, INF, INF);
        if (lane < CW / 4 && t < t1) {
          const uint32_t c = t * TF2_BN + slice * CW + lane * 4;
          if (c + 3 < a.n_rows) {
            r = __ldg(reinterpret_cast<const float4*>(a.aux + c));
          } else {
            if (c + 0 < a.n_rows) r.x = __ldg(a.aux + c + 0);
            if (c + 1 < a.n_rows) r.y = __ldg(a.aux + c + 1);
            if (c + 2 < a.n_rows) r.z = __ldg(a.aux + c + 2);
          }
        }
        return r;
      };
      float4 aux_next = load_aux(t0);
      for (uint32_t t = t0; t < t1; ++t, ++acc_it) {
        float* aux_t = my_aux + (acc_it & 1) * CW;
        if (lane < CW / 4) reinterpret_cast<float4*>(aux_t)[lane] = aux_next;
        aux_next = load_aux(t + 1);
        __syncwarp();
        const uint32_t buf = (NBUF == 2) ? (acc_it & 1) : 0;
        const uint32_t use = (NBUF == 2) ? (acc_it >> 1) : acc_it;
        mbar_wait(acc_full + buf, use & 1);
        tc_fence_after();
        float v[CW];
#pragma unroll
        for (int g = 0; g < CW / 32; ++g) tc_ld32(tmem_acc + lane_addr + buf * TF2_BN + slice * CW + g * 32, v + g * 32);
        tc_wait_ld();
        // the accumulator now lives in registers: hand it back to the MMA warp of the leader before gating
        tc_fence_before();
        __syncwarp();
        if (lane == 0) mbar_arrive_cluster(acc_empty_leader + buf * 8);
        const float4* aux4 = reinterpret_cast<const float4*>(aux_t);
        const uint32_t colw = t * TF2_BN + slice * CW;
#pragma unroll
        for (int seg = 0; seg < CW / 16; ++seg) {
          float4 ax[4];
#pragma unroll
          for (int i = 0; i < 4; ++i) ax[i] = aux4[seg * 4 + i];
          float sg[16], mg[4];
#pragma unroll
          for (int g4 = 0; g4 < 4; ++g4) {
            const int j = seg * 16 + g4 * 4;
            fma2(coef2, v[j + 0], v[j + 1], ax[g4].x, ax[g4].y, sg[g4 * 4 + 0], sg[g4 * 4 + 1]);
            fma2(coef2, v[j + 2], v[j + 3], ax[g4].z, ax[g4].w, sg[g4 * 4 + 2], sg[g4 * 4 + 3]);
            mg[g4] = fminf(min3(sg[g4 * 4 + 0], sg[g4 * 4 + 1], sg[g4 * 4 + 2]), sg[g4 * 4 + 3]);
          }
          if (fminf(min3(mg[0], mg[1], mg[2]), mg[3]) < theta) {
            uint32_t cnt = 0;
#pragma unroll
            for (int g4 = 0; g4 < 4; ++g4) {
              const int j = seg * 16 + g4 * 4;
              if (mg[g4] < theta) {
#pragma unroll
                for (int e = 0; e < 4; ++e) {
                  if (sg[g4 * 4 + e] < theta) {
                    my_qs[cnt * ET] = sg[g4 * 4 + e];
                    my_qc[cnt * ET] = colw + j + e;
                    ++cnt;
                  }
                }
              }
            }
#pragma unroll 1
            for (uint32_t i = 0; i < cnt; ++i) {
              const float s = my_qs[i * ET];
              if (s < theta) {
                my_row[imax * ET] = my_qc[i * ET];
                float mv[KP];
                int mi[KP];
#pragma unroll
                for (int t2 = 0; t2 < KP; ++t2) {
                  sc[t2] = (t2 == imax) ? s : sc[t2];
                  mv[t2] = sc[t2];
                  mi[t2] = t2;
                }
#pragma unroll
                for (int w = KP / 2; w >= 1; w >>= 1) {
#pragma unroll
                  for (int t2 = 0; t2 < w; ++t2) {
                    const bool gt = mv[t2 + w] > mv[t2];
                    mv[t2] = gt ? mv[t2 + w] : mv[t2];
                    mi[t2] = gt ? mi[t2 + w] : mi[t2];
                  }
                }
                theta = fminf(mv[0], hint);
                imax = mi[0];
                dirty = true;
              }
            }
          }
        }
        if (sharing) {   // what the list found goes out, what the helper made of everybody's lists comes in
          if (dirty && ((t - t0) & 1u)) {
            publish_two_best<KP>(sc, my_pub);
            dirty = false;
          }
          hint = fminf(hint, helper_bound_for(s_hb, qrow, item));
          theta = fminf(theta, hint);
        }
      }
      // ---- chunk result: one candidate list per (query, chunk, column slice) ----
      if (q_global < a.nq) {
        const size_t vchunk = (size_t)chunk * EW + slice;
        const size_t base = ((size_t)q_global * a.n_chunks * EW + vchunk) * KP;
#pragma unroll
        for (int j = 0; j < KP; ++j) {
          a.cand_score[base + j] = sc[j];
          a.cand_row[base + j] = my_row[j * ET];
        }
        a.chunk_tau[(size_t)q_global * a.n_chunks * EW + vchunk] = theta;
        if (sharing) publish_two_best<KP>(sc, my_pub);   // the list's final two best, for the lists of later waves
        if (a.hint) {
          const float pub = a.hint_target == 0 ? theta
                                               : fminf(theta, publish_value<KP>(sc, (min(a.n_rows, t1 * (uint32_t)TF2_BN) - t0 * (uint32_t)TF2_BN) / EW,
                                                                                a.n_rows, a.hint_target));
          if (pub < INF) atomicMin(a.hint + q_global, f32_ord(pub));
        }
      }
    }
    if (warp == 2 && lane == 0) st_shared_volatile_u32(s_cur_item, ITEM_EXIT);   // releases the bound helper
  }

  tc_fence_before();
  __syncthreads();
  cluster_sync_all();   // nobody of the pair still reads TMEM or signals a barrier of the peer
  if (warp == 2) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::2.sync.aligned.b32 %0, %1;" ::"r"(tmem_base), "r"(512u) : "memory");
  }
}

template <int KP, int EW>
static size_t filter2_smem() {
  return 1024 + (size_t)filter2_stages(KP, EW) * TF2_STAGE_BYTES + (size_t)128 * EW * KP * 4 + (size_t)2 * 16 * 128 * EW * 4 +
         (size_t)4 * 2 * TF2_BN * 4 + (size_t)(2 * filter2_stages(KP, EW) + 5) * 8 + 128 * 8 + 16;
}

// CTA pairs that are resident at the same time (not every SM of the part has a free partner in its
// TPC): the persistent grid is sized to it, so that no pair waits for a second wave.
static int32_t filter2_resident_pairs(int* out) {
  static int cached = 0;   // per process: one part per box (the variants all take a whole SM per CTA)
  if (cached == 0) {
    SCN_ALLOW_SMEM((tensor_filter2_kernel<16, 1, 1>), (filter2_smem<16, 1>()));
    cudaLaunchConfig_t cfg{};
    cfg.gridDim = dim3(2 * 74);
    cfg.blockDim = dim3(224);
    cfg.dynamicSmemBytes = filter2_smem<16, 1>();
    int n = 0;
    if (cudaOccupancyMaxActiveClusters(&n, tensor_filter2_kernel<16, 1, 1>, &cfg) != cudaSuccess || n < 1) {
      cudaGetLastError();
      n = 64;
    }
    cached = n;
    if (getenv("SCN_DEBUG")) fprintf(stderr, "[scn] tensor_filter2: %d CTA pairs resident\n", n);
  }
  *out = cached;
  return SCN_OK;
}

template <int KP, int NBUF, int EW>
static int32_t launch_filter2(const CUtensorMap& tmap_b, const FilterArgs& fa, int pairs, cudaStream_t stream, bool pdl) {
  SCN_ALLOW_SMEM((tensor_filter2_kernel<KP, NBUF, EW>), (filter2_smem<KP, EW>()));
  SCN_CUDA(launch_chained(tensor_filter2_kernel<KP, NBUF, EW>, dim3(2 * pairs), dim3(96 + 128 * EW), filter2_smem<KP, EW>(), stream, pdl, tmap_b,
                          fa));   // (cluster dims are compiled in)
  SCN_LAUNCHED();
  return SCN_OK;
}

// ---- query preparation ---------------------------------------------------------------------------
// One warp per query: bf16 row (zero padded to kpad), and the norms the certificate needs.
// qstat[q] = {||q||, ||q~||, ||q - q~||, ||q~||^2}
// The per-call initialisations ride along (they were three memsets): the published thresholds `hint`, the
// two failure counts, and — first batch of a call — the device counters.
__global__ void __launch_bounds__(128) prep_queries_kernel(const float* __restrict__ q, uint32_t nq, uint32_t nq_pad, uint32_t dim,
                                                           uint32_t kpad, __nv_bfloat16* __restrict__ qb,
                                                           float4* __restrict__ qstat, uint32_t* __restrict__ hint,
                                                           uint32_t* __restrict__ nfail, unsigned long long* __restrict__ counters,
                                                           float* __restrict__ pub, uint32_t lists_pad) {
  uint32_t qi = blockIdx.x * 4 + (threadIdx.x >> 5);
  const int lane = threadIdx.x & 31;
  pdl_trigger();   // the filter may set itself up (barriers, TMEM, first row tiles) while the queries are converted
  if (blockIdx.x == 0 && threadIdx.x < 4) {
    if (threadIdx.x < 2) nfail[threadIdx.x] = 0;
    if (counters) counters[threadIdx.x] = 0;
  }
  if (qi >= nq_pad) return;
  float nn = 0.f, mm = 0.f, ee = 0.f;
  for (uint32_t i = lane; i < kpad; i += 32) {
    float v = (qi < nq && i < dim) ? q[(size_t)qi * dim + i] : 0.0f;
    __nv_bfloat16 b = __float2bfloat16_rn(v);
    float bv = __bfloat162float(b);
    qb[(size_t)qi * kpad + i] = b;
    nn = fmaf(v, v, nn);
    mm = fmaf(bv, bv, mm);
    ee = fmaf(v - bv, v - bv, ee);
  }
  for (int o = 16; o > 0; o >>= 1) {
    nn += __shfl_xor_sync(0xffffffffu, nn, o);
    mm += __shfl_xor_sync(0xffffffffu, mm, o);
    ee += __shfl_xor_sync(0xffffffffu, ee, o);
  }
  if (lane == 0 && qi < nq) {
    qstat[qi] = make_float4(sqrtf(nn), sqrtf(mm), sqrtf(ee), mm);
    hint[qi] = 0xFFFFFFFFu;
  }
  if (pub)   // the table of published bests, float2 [lists_pad][nq_pad]: NaN = nothing published yet
    for (uint32_t l = lane; l < lists_pad; l += 32)
      reinterpret_cast<float2*>(pub)[(size_t)l * nq_pad + qi] = make_float2(__int_as_float(0x7fc00000), __int_as_float(0x7fc00000));
}

// ---- candidate merge: chunk lists -> k'' rows + tau -------------------------------------------------
// Bitonic sort of 64 keys by one warp: two keys per lane (elements lane and lane + 32), exchanges through
// shuffles — no block barrier per step.
__device__ __forceinline__ void warp_sort64(uint64_t* a, uint32_t lane) {
  uint64_t x0 = a[lane], x1 = a[lane + 32];
#pragma unroll
  for (uint32_t size = 2; size <= 64; size <<= 1) {
#pragma unroll
    for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
      if (stride == 32) {  // the partner is this lane's other element; size == 64: ascending
        const uint64_t lo = x0 < x1 ? x0 : x1, hi = x0 < x1 ? x1 : x0;
        x0 = lo;
        x1 = hi;
      } else {
        const uint64_t y0 = __shfl_xor_sync(0xffffffffu, x0, stride), y1 = __shfl_xor_sync(0xffffffffu, x1, stride);
        const bool lower = (lane & stride) == 0;                 // this element is the lower index of its pair
        const bool up0 = (lane & size) == 0, up1 = ((lane + 32) & size) == 0;
        const uint64_t mn0 = x0 < y0 ? x0 : y0, mx0 = x0 < y0 ? y0 : x0;
        const uint64_t mn1 = x1 < y1 ? x1 : y1, mx1 = x1 < y1 ? y1 : x1;
        x0 = (lower == up0) ? mn0 : mx0;
        x1 = (lower == up1) ? mn1 : mx1;
      }
    }
  }
  a[lane] = x0;
  a[lane + 32] = x1;
}

// keys[0..max(n_pad, 64)) hold the n candidate keys of a query, KEY_NONE padded (the array has room for at least
// 256 keys). On return keys[0..n_sort) are in ascending order and contain the kpp + 1 smallest keys (or all of
// them); returns n_sort. Block-wide: every thread of the block calls it.
// Long lists (small batches spread a query block over ~148 row chunks: thousands of candidates, of which
// k'' + 1 matter): radix-select the score of rank k'' (4 passes over the 32 score bits), move the keys up to
// that score into a 256-entry buffer and sort only those — by one warp when at most 64 are left, which is the
// normal case (k'' + 1 = 33 unless scores tie). Many keys tied at the boundary (more than the buffer holds)
// fall through to the full sort.
__device__ __forceinline__ uint32_t select_and_sort(uint64_t* keys, uint32_t n, uint32_t n_pad, uint32_t kpp) {
  uint32_t n_sort = n_pad < 64 ? 64u : n_pad;
  if (n_pad > 128) {
    __shared__ uint32_t hist[256];
    __shared__ uint32_t s_prefix, s_rank, s_cnt, s_total;
    __shared__ uint64_t buf[256];
    if (threadIdx.x == 0) {
      s_prefix = 0;
      s_rank = kpp;
      s_cnt = 0;
    }
    uint32_t mask = 0;
    bool all = false;  // fewer than k'' + 1 valid keys: take them all
    for (int shift = 24; shift >= 0 && !all; shift -= 8) {
      for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) hist[i] = 0;
      __syncthreads();
      const uint32_t prefix = s_prefix;
      for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
        const uint64_t key = keys[i];
        const uint32_t o = (uint32_t)(key >> 32);
        if (key != KEY_NONE && (o & mask) == prefix) atomicAdd(&hist[(o >> shift) & 255u], 1u);
      }
      __syncthreads();
      if (threadIdx.x < 32) {
        const uint32_t lane = threadIdx.x;
        uint32_t part = 0;
#pragma unroll
        for (int b = 0; b < 8; ++b) part += hist[lane * 8 + b];
        uint32_t incl = part;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const uint32_t up = __shfl_up_sync(0xffffffffu, incl, o);
          if (lane >= (uint32_t)o) incl += up;
        }
        const uint32_t total = __shfl_sync(0xffffffffu, incl, 31);
        const uint32_t rank = s_rank;
        const uint32_t excl = incl - part;
        if (lane == 0) s_total = total;
        if (rank < total && rank >= excl && rank < incl) {  // the bin of rank `rank` is one of this lane's eight
          uint32_t cum = excl;
          for (int b = 0; b < 8; ++b) {
            const uint32_t h = hist[lane * 8 + b];
            if (rank < cum + h) {
              s_prefix = prefix | ((uint32_t)(lane * 8 + b) << shift);
              s_rank = rank - cum;
              break;
            }
            cum += h;
          }
        }
      }
      __syncthreads();
      if (s_rank >= s_total && shift == 24) all = true;  // (only the first pass sees every valid key)
      mask |= 255u << shift;
    }
    const uint32_t vmax = all ? 0xFFFFFFFFu : s_prefix;  // score (ord) of the key of rank k''
    for (uint32_t i = threadIdx.x; i < 256; i += blockDim.x) buf[i] = KEY_NONE;
    __syncthreads();
    for (uint32_t i = threadIdx.x; i < n; i += blockDim.x) {
      const uint64_t key = keys[i];
      if (key != KEY_NONE && (uint32_t)(key >> 32) <= vmax) {
        const uint32_t pos = atomicAdd(&s_cnt, 1u);
        if (pos < 256) buf[pos] = key;
      }
    }
    __syncthreads();
    const uint32_t cnt = s_cnt;
    if (cnt <= 256) {
      n_sort = cnt <= 64 ? 64u : 256u;
      for (uint32_t i = threadIdx.x; i < n_sort; i += blockDim.x) keys[i] = buf[i];
    }
  }
  __syncthreads();
  if (n_sort == 64) {
    if (threadIdx.x < 32) warp_sort64(keys, threadIdx.x);
  } else {
    // block bitonic sort (ascending)
    for (uint32_t size = 2; size <= n_sort; size <<= 1) {
      for (uint32_t stride = size >> 1; stride > 0; stride >>= 1) {
        for (uint32_t t = threadIdx.x; t < n_sort / 2; t += blockDim.x) {
          uint32_t lo = 2 * t - (t & (stride - 1)), hi = lo + stride;
          bool up = ((lo & size) == 0);
          uint64_t x = keys[lo], y = keys[hi];
          if ((x > y) == up) {
            keys[lo] = y;
            keys[hi] = x;
          }
        }
        __syncthreads();
      }
    }
  }
  __syncthreads();
  return n_sort;
}

// (key of a candidate, KEY_NONE for an empty slot)
__device__ __forceinline__ uint64_t candidate_key(const float* __restrict__ cand_score, const uint32_t* __restrict__ cand_row, size_t i) {
  const uint32_t row = cand_row[i];
  return row != ROW_NONE ? make_key(cand_score[i], row) : KEY_NONE;
}

__global__ void __launch_bounds__(512) merge_candidates_kernel(const float* __restrict__ cand_score,
                                                               const uint32_t* __restrict__ cand_row,
                                                               const float* __restrict__ chunk_tau, uint32_t n_chunks,
                                                               uint32_t kprime, uint32_t n_pad, uint32_t kpp,
                                                               uint32_t* __restrict__ out_rows, float* __restrict__ out_tau,
                                                               float* __restrict__ out_tau_chunks) {
  extern __shared__ __align__(16) unsigned char smem_merge[];
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_merge);  // [max(n_pad, 256)]
  const uint32_t q = blockIdx.x;
  const uint32_t n = n_chunks * kprime;
  pdl_trigger();
  pdl_wait();
  for (uint32_t i = threadIdx.x; i < max(n_pad, 64u); i += blockDim.x) keys[i] = (i < n) ? candidate_key(cand_score, cand_row, (size_t)q * n + i) : KEY_NONE;
  __syncthreads();
  const uint32_t n_sort = select_and_sort(keys, n, n_pad, kpp);
  for (uint32_t i = threadIdx.x; i < kpp; i += blockDim.x) {
    uint64_t key = (i < n_sort) ? keys[i] : KEY_NONE;
    out_rows[(size_t)q * kpp + i] = (key == KEY_NONE) ? ROW_NONE : (uint32_t)key;
  }
  if (threadIdx.x == 0) {
    // tau = lower bound on the filter score of every row that is not handed to the rerank:
    // rows never kept by a chunk (>= that chunk's tau) and rows dropped here (>= keys[kpp])
    float tau = __int_as_float(0x7f800000);
    for (uint32_t c = 0; c < n_chunks; ++c) tau = fminf(tau, chunk_tau[(size_t)q * n_chunks + c]);
    out_tau_chunks[q] = tau;  // bound on rows no chunk kept: valid when ALL kept candidates are reranked
    if (kpp < n_sort && keys[kpp] != KEY_NONE) tau = fminf(tau, ord_f32((uint32_t)(keys[kpp] >> 32)));
    out_tau[q] = tau;
  }
}

// ---- certificate -------------------------------------------------------------------------------------
// For query q let d_k be the k-th exact distance found by the rerank. Every row x that was not
// reranked has filter score s~(x) >= tau. With
//   E  = ||q - q~||*Mmax + ||q||*Emax + e_tc     (bf16 rounding of both operands; Mmax = max ||mirror row||,
//                                                 Emax = max ||row - mirror row||; e_tc = fp32 accumulation
//                                                 error of the tensor pipe, bounded by 8*K*2^-24*||q~||*Mmax)
//   g  = 1.01*(D+2)*2^-24                        (the reference's own sequential fp32 summation error)
// the reference distance of such a row is bounded below by LB (per metric, see below). The query is
// certified iff LB > d_k strictly, so neither a closer row nor a tie with a lower row index can have
// been missed. Anything else goes to the exact scan.
__device__ __forceinline__ bool certificate_holds(uint64_t kth_key, float t, const float4 st, const float* __restrict__ bounds,
                                                  uint32_t dim, uint32_t kpad, int metric) {
  const float INF = __int_as_float(0x7f800000);
  const float dk = (kth_key == KEY_NONE) ? INF : ord_f32((uint32_t)(kth_key >> 32));
  if (t == INF) return true;      // nothing was left out: the rerank saw every live row
  if (!(dk < INF)) return false;  // fewer than k finite results although rows were left out
  const float qn = st.x, qtn = st.y, eq = st.z, qtn2 = st.w;
  const float Mmax = bounds[0], Emax = bounds[1], Xmax = bounds[2];
  const float u = 5.9604645e-8f;  // 2^-24
  const float etc = 8.0f * (float)kpad * u * qtn * Mmax;
  const float E = (eq * Mmax + qn * Emax + etc) * 1.0001f;
  const float g = 1.01f * (float)(dim + 2) * u;
  float lb;
  if (metric == M_IP) {
    // d_ref(x) = -fl(q.x) >= -(q.x) - g*||q||*||x|| ;  -(q.x) >= s~ - E >= tau - E
    lb = t - E - g * qn * Xmax;
    lb -= fabsf(lb) * 4.0f * u;
  } else if (metric == M_COS) {
    // mirror rows are x/||x|| (fp32 divide, rel. error <= 2^-23 per element): -(q.x^) >= tau - E - ||q||*2^-22
    // cos(x) = (q.x^)/||q|| <= (E' - tau)/||q||;  d_ref >= 1 - cos - (2g + 2^-21)
    float Ep = E + qn * 4.0f * u;
    float c = (Ep - t) / qn;
    lb = 1.0f - c - (2.0f * g + 8.0f * u);
    lb -= fabsf(lb) * 4.0f * u + 4.0f * u;
    if (!(qn > 0.0f)) lb = -INF;  // zero query: every distance is exactly 1 -> ties everywhere
  } else {
    // ||q - x|| >= ||q~ - x~|| - ||q - q~|| - ||x - x~||, and ||q~ - x~||^2 = ||q~||^2 + s~ (s~ = ||x~||^2 - 2 q~.x~)
    // fp32 error of s~: 2*e_tc plus the rounding of aux and of ||q~||^2 (<= K*2^-23 relative each)
    float es = 2.0f * etc + (float)kpad * 2.0f * u * (Mmax * Mmax + qtn2);
    float r2 = qtn2 + t - es;
    float r = r2 > 0.0f ? sqrtf(r2) * (1.0f - 2.0f * u) : 0.0f;
    lb = r - eq * 1.0001f - Emax;
    lb = lb * (1.0f - g) * (1.0f - 4.0f * u);
  }
  return lb > dk;
}

__global__ void certify_kernel(const uint64_t* __restrict__ keys, const float* __restrict__ tau, const float4* __restrict__ qstat,
                               const float* __restrict__ bounds, uint32_t nq, uint32_t k, uint32_t dim, uint32_t kpad,
                               int metric, const uint32_t* __restrict__ qlist_in, const uint32_t* __restrict__ nq_in_dev,
                               uint32_t* __restrict__ fail_list, uint32_t* __restrict__ fail_count,
                               unsigned long long* __restrict__ counters, int counter_slot) {
  pdl_trigger();
  pdl_wait();
  const uint32_t n_slots = nq_in_dev ? min(*nq_in_dev, nq) : nq;
  uint32_t slot = blockIdx.x * blockDim.x + threadIdx.x;
  if (slot >= n_slots) return;
  const uint32_t q = qlist_in ? qlist_in[slot] : slot;
  const bool ok = certificate_holds(keys[(size_t)q * k + (k - 1)], tau[q], qstat[q], bounds, dim, kpad, metric);
  if (!ok) {
    uint32_t slot = atomicAdd(fail_count, 1u);
    fail_list[slot] = q;
  }
  if (counters) {
    if (counter_slot == 1) atomicAdd(counters + 0, 1ull);
    if (!ok) atomicAdd(counters + counter_slot, 1ull);
  }
}

// ---- fused tail: merge -> exact rerank -> certificate, one block per query ----------------------------
// What merge_candidates, rerank_kernel and certify_kernel do in three launches (each a grid of small blocks
// plus a launch gap), for batches where those gaps are what the caller waits for: behind a 0.25 ms pass over
// the mirror the three cost 0.04 ms, a sixth of the call. One block of 256 threads per query:
//   1. the query's candidate keys -> shared memory; select_and_sort -> the k'' best rows and tau;
//   2. the k'' rows are copied into shared memory by the whole block (cp.async, 16 bytes per thread and
//      request, chunks of FIN_CHB bytes per row, two chunks in flight); thread j walks row j in the
//      reference's sequential fp32 order (acc_step4), thread k'' the query's own norm (cosine);
//   3. one warp sorts the k'' exact keys, thread 0 emits the first k distinct ones and checks the certificate.
// Results are those of the three kernels bit for bit (same keys, same arithmetic, same order).
constexpr uint32_t FIN_THREADS = 256;
__host__ __device__ inline uint32_t fin_chunk_bytes(uint32_t kpp) { return kpp <= 32 ? 1024u : 512u; }
__host__ __device__ inline size_t fin_smem_bytes(uint32_t n_pad, uint32_t pitch, uint32_t kpp) {
  return (size_t)(n_pad < 256 ? 256 : n_pad) * 8 + (size_t)pitch * 4 + (size_t)2 * kpp * (fin_chunk_bytes(kpp) + 16) + 64 * 8 + (size_t)kpp * 4;
}

struct FinishArgs {
  const float* cand_score;
  const uint32_t* cand_row;
  const float* chunk_tau;
  uint32_t n_lists, kprime, n_pad, kpp;
  const float* vec;
  const float* norm;
  const uint32_t* deleted;
  uint32_t pitch, dim, n_rows, kpad;
  const float* q;
  const float4* qstat;
  const float* bounds;
  uint32_t k, row_base;
  uint64_t* out_keys;
  float* out_tau_chunks;
  uint32_t* fail_list;
  uint32_t* fail_count;
  unsigned long long* counters;
};

template <int METRIC>
__global__ void __launch_bounds__(FIN_THREADS) finish_queries_kernel(FinishArgs a) {
  extern __shared__ __align__(16) unsigned char smem_fin[];
  const uint32_t kpp = a.kpp, CHB = fin_chunk_bytes(kpp), ROWB = CHB + 16;   // stage rows CHB + 16 bytes apart: conflict-free LDS.128
  uint64_t* keys = reinterpret_cast<uint64_t*>(smem_fin);                      // [max(n_pad, 256)]
  float* s_q = reinterpret_cast<float*>(keys + (a.n_pad < 256 ? 256 : a.n_pad));   // [pitch]
  unsigned char* s_stage = reinterpret_cast<unsigned char*>(s_q + a.pitch);    // [2][kpp][ROWB]
  uint64_t* xkeys = reinterpret_cast<uint64_t*>(s_stage + (size_t)2 * kpp * ROWB);  // [64]
  uint32_t* s_rows = reinterpret_cast<uint32_t*>(xkeys + 64);                  // [kpp]
  __shared__ float s_red[FIN_THREADS / 32];
  __shared__ float s_tau[2];
  __shared__ float s_qnorm;
  const uint32_t q = blockIdx.x, tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const uint32_t n = a.n_lists * a.kprime;
  const float INF = __int_as_float(0x7f800000);
  pdl_trigger();
  pdl_wait();

  // 1. candidates, query, chunk thresholds
  for (uint32_t i = tid; i < max(a.n_pad, 64u); i += FIN_THREADS) keys[i] = (i < n) ? candidate_key(a.cand_score, a.cand_row, (size_t)q * n + i) : KEY_NONE;
  stage_query(s_q, a.q + (size_t)q * a.dim, a.dim, a.pitch, tid, FIN_THREADS);
  float tmin = INF;
  for (uint32_t c = tid; c < a.n_lists; c += FIN_THREADS) tmin = fminf(tmin, a.chunk_tau[(size_t)q * a.n_lists + c]);
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) tmin = fminf(tmin, __shfl_xor_sync(0xffffffffu, tmin, o));
  if (lane == 0) s_red[warp] = tmin;
  if (tid < 64) xkeys[tid] = KEY_NONE;
  __syncthreads();
  const uint32_t n_sort = select_and_sort(keys, n, a.n_pad, kpp);
  if (tid < kpp) {
    const uint64_t key = (tid < n_sort) ? keys[tid] : KEY_NONE;
    uint32_t row = (key == KEY_NONE) ? ROW_NONE : (uint32_t)key;
    if (row != ROW_NONE && (row >= a.n_rows || bit_test(a.deleted, row))) row = ROW_NONE;
    s_rows[tid] = row;
  }
  if (tid == 0) {
    float tau = INF;
#pragma unroll
    for (int w = 0; w < (int)(FIN_THREADS / 32); ++w) tau = fminf(tau, s_red[w]);
    s_tau[1] = tau;  // bound on rows no chunk kept
    if (kpp < n_sort && keys[kpp] != KEY_NONE) tau = fminf(tau, ord_f32((uint32_t)(keys[kpp] >> 32)));
    s_tau[0] = tau;  // bound on every row that is not reranked below
  }
  __syncthreads();

  // 2. exact distances of the k'' rows
  const uint32_t row_bytes = a.pitch * 4, n_ch = (row_bytes + CHB - 1) / CHB, pieces = CHB / 16;
  auto issue = [&](uint32_t ch) {
    if (ch < n_ch) {
      unsigned char* dst = s_stage + (size_t)(ch & 1) * kpp * ROWB;
      for (uint32_t p = tid; p < kpp * pieces; p += FIN_THREADS) {
        const uint32_t r = p / pieces, piece = p - r * pieces;
        const uint32_t off = ch * CHB + piece * 16;
        const uint32_t row = s_rows[r];
        const bool valid = row != ROW_NONE && off < row_bytes;
        const unsigned char* src = reinterpret_cast<const unsigned char*>(a.vec) + (valid ? (size_t)row * row_bytes + off : 0);
        cp_async16(dst + (size_t)r * ROWB + piece * 16, src, valid);
      }
    }
    cp_async_commit();
  };
  issue(0);
  issue(1);
  float acc = 0.0f;
  const uint32_t my_row = tid < kpp ? s_rows[tid] : ROW_NONE;
  const float xn = (METRIC == M_COS && my_row != ROW_NONE) ? __ldg(a.norm + my_row) : 0.0f;
  for (uint32_t ch = 0; ch < n_ch; ++ch) {
    cp_async_wait<1>();   // chunk ch has landed (chunk ch + 1 may still be in flight)
    __syncthreads();
    const uint32_t n4 = min(CHB, row_bytes - ch * CHB) / 16;
    const float4* q4 = reinterpret_cast<const float4*>(s_q) + ch * pieces;
    if (tid < kpp) {
      const float4* x4 = reinterpret_cast<const float4*>(s_stage + ((size_t)(ch & 1) * kpp + tid) * ROWB);
#pragma unroll 4
      for (uint32_t i = 0; i < n4; ++i) acc = acc_step4<METRIC>(acc, q4[i], x4[i]);
    } else if (METRIC == M_COS && tid == kpp) {   // the query's norm, the reference's sequential sum (distance.go:58-66)
      for (uint32_t i = 0; i < n4; ++i) {
        const float4 v = q4[i];
        acc = __fadd_rn(acc, __fmul_rn(v.x, v.x));
        acc = __fadd_rn(acc, __fmul_rn(v.y, v.y));
        acc = __fadd_rn(acc, __fmul_rn(v.z, v.z));
        acc = __fadd_rn(acc, __fmul_rn(v.w, v.w));
      }
    }
    __syncthreads();      // buffer ch & 1 is free again
    issue(ch + 2);
  }
  if (METRIC == M_COS && tid == kpp) s_qnorm = __fsqrt_rn(acc);
  __syncthreads();
  if (tid < kpp && my_row != ROW_NONE) xkeys[tid] = make_key(finish_distance<METRIC>(acc, METRIC == M_COS ? s_qnorm : 0.0f, xn), my_row + a.row_base);
  __syncthreads();

  // 3. order, emit, certify
  if (warp == 0) {
    warp_sort64(xkeys, lane);
    __syncwarp();
    for (uint32_t i = lane; i < a.k; i += 32) a.out_keys[(size_t)q * a.k + i] = KEY_NONE;
    __syncwarp();
    if (lane == 0) {
      uint32_t w = 0;
      uint64_t prev = KEY_NONE, kth = KEY_NONE;
      for (uint32_t i = 0; i < 64 && w < a.k; ++i) {   // (identical keys would be adjacent; merged candidates are distinct rows)
        const uint64_t v = xkeys[i];
        if (v == KEY_NONE) break;
        if (v != prev) {
          a.out_keys[(size_t)q * a.k + w++] = v;
          if (w == a.k) kth = v;
        }
        prev = v;
      }
      a.out_tau_chunks[q] = s_tau[1];
      const bool ok = certificate_holds(kth, s_tau[0], a.qstat[q], a.bounds, a.dim, a.kpad, METRIC);
      if (!ok) a.fail_list[atomicAdd(a.fail_count, 1u)] = q;
      if (a.counters) {
        atomicAdd(a.counters + 0, 1ull);
        if (!ok) atomicAdd(a.counters + 1, 1ull);
      }
    }
  }
}

__global__ void set_aux_kernel(float* aux, const uint32_t* rows, uint32_t n, float value) {
  uint32_t i = blockIdx.x * blockDim.x + threadIdx.x;
  if (i < n) aux[rows[i]] = value;
}

// ---- host side -----------------------------------------------------------------------------------------
typedef CUresult (*EncodeTiledFn)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*, const cuuint64_t*,
                                  const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave, CUtensorMapSwizzle,
                                  CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

static EncodeTiledFn encode_tiled_fn() {
  static EncodeTiledFn fn = nullptr;
  if (!fn) {
    void* p = nullptr;
    cudaDriverEntryPointQueryResult qres;
    if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &qres) == cudaSuccess &&
        qres == cudaDriverEntryPointSuccess)
      fn = reinterpret_cast<EncodeTiledFn>(p);
  }
  return fn;
}

bool tensor_path_supported(const scn_store* s, uint32_t k) { return s->kpad <= TF_MAX_KPAD_STREAM && k <= 24 && s->rows >= 1; }

// Row chunks per query block: work items = n_qb * chunks are dealt round robin to `sms` persistent
// workers, so the kernel lasts waves * (tiles of a chunk + what an item costs besides its tiles: the
// query block goes into TMEM, the lists are written out). The cheapest chunk count under that model.
static uint32_t pick_chunks(uint32_t n_qb, uint32_t n_tiles, uint32_t sms, uint32_t item_overhead_tiles = 3) {
  const uint32_t cmax = std::max(1u, std::min(256u, n_tiles / 8));
  uint32_t best_c = 1;
  uint64_t best_cost = ~0ull;
  for (uint32_t c = 1; c <= cmax; ++c) {
    const uint64_t items = (uint64_t)n_qb * c;
    const uint64_t waves = (items + sms - 1) / sms;
    const uint64_t cost = waves * ((n_tiles + c - 1) / c + item_overhead_tiles);
    if (cost < best_cost) {
      best_cost = cost;
      best_c = c;
    }
  }
  return best_c;
}

template <int KP, int NBUF, int EW, bool DBG, bool ASM, int BN, bool SHARE = false>
static int32_t launch_filter(const CUtensorMap& tmap_b, const CUtensorMap& tmap_a, const FilterArgs& fa, int grid, cudaStream_t stream, bool pdl) {
  using Cfg = FilterCfg<KP, NBUF, EW, DBG, ASM, BN>;
  size_t smem = 1024 + (size_t)Cfg::STAGES * Cfg::STAGE_BYTES + (size_t)128 * EW * KP * 4 + (size_t)2 * 16 * 128 * EW * 4 +
                (size_t)8 * BN * 4 + (size_t)(2 * Cfg::STAGES + 5) * 8 + 128 * 8 + 16;
  SCN_ALLOW_SMEM((tensor_filter_kernel<KP, NBUF, EW, DBG, ASM, BN, SHARE>), smem);
  SCN_CUDA(launch_chained(tensor_filter_kernel<KP, NBUF, EW, DBG, ASM, BN, SHARE>, dim3(grid), dim3(64 + 128 * EW + (SHARE ? 32 : 0)), smem, stream, pdl, tmap_b,
                          tmap_a, fa));
  SCN_LAUNCHED();
  return SCN_OK;
}

static int32_t flat_search_tensor_batch(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint64_t row_base,
                                        uint64_t* d_out_keys, cudaStream_t stream, Profiler* prof, float* dbg_scores, bool first_batch) {
  int sms = 0;
  SCN_CUDA(cudaDeviceGetAttribute(&sms, cudaDevAttrMultiProcessorCount, s->device));
  EncodeTiledFn enc = encode_tiled_fn();
  if (!enc) return fail(SCN_ERR_INTERNAL, "cuTensorMapEncodeTiled is not available from this driver");
  // CTA pairs (tensor_filter2_kernel) for long rows and batches of at least one full pair of query blocks
  const bool pair_kernel = s->opt_tensor_pair != 0 && (s->kpad == 512 || s->kpad == 640 || s->kpad == 768) && nq >= 256 && !dbg_scores &&
                           (sms % 2) == 0;
  // 64-row tiles with two accumulators where one 128-column accumulator is all that fits next to A
  const bool half_tiles = !pair_kernel && (s->kpad == 640 || s->kpad == 768) && !dbg_scores && s->opt_tensor_bn != 128;
  const uint32_t BN = pair_kernel ? (uint32_t)TF2_BN : half_tiles ? 64u : 128u;
  const uint32_t n_rows = (uint32_t)s->rows;
  const uint32_t n_tiles = (n_rows + BN - 1) / BN;
  const uint32_t n_qb = (uint32_t)((nq + TF_BM - 1) / TF_BM);
  const uint32_t nq_pad = (pair_kernel ? (n_qb + 1) / 2 * 2 : n_qb) * TF_BM;   // whole pairs of query blocks
  // candidates kept per (query, chunk): 16 covers k <= 10 with a 60 % margin, 32 covers k <= 24
  uint32_t kprime = (k <= 10 && s->opt_overfetch <= 16) ? 16u : 32u;
  int resident_pairs = 0;
  if (pair_kernel) SCN_TRY(filter2_resident_pairs(&resident_pairs));
  // (an item of the pair kernel costs about 24 tiles besides its own: measured at C2, 11 chunks 10.9 ms, 24: 11.5, 48: 12.7, 96: 14.0 —
  // every item stages a query block and starts its candidate lists over)
  uint32_t n_chunks = pair_kernel ? pick_chunks((n_qb + 1) / 2, n_tiles, (uint32_t)resident_pairs, 24) : pick_chunks(n_qb, n_tiles, (uint32_t)sms);
  if (s->opt_tensor_chunks > 0) n_chunks = (uint32_t)std::min<int64_t>(s->opt_tensor_chunks, n_tiles);
  uint32_t tiles_per_chunk = (n_tiles + n_chunks - 1) / n_chunks;
  n_chunks = (n_tiles + tiles_per_chunk - 1) / tiles_per_chunk;
  // epilogue warps per TMEM lane quarter: the gate of a 128x128 tile costs ~1600 issue cycles with
  // one warp per quarter; the MMA of the tile takes 4*kpad cycles
  const bool stream_a = s->kpad > TF_MAX_KPAD;  // query block too wide for TMEM: stream it with the rows
  // Pair kernel: two warps per quarter. Measured (1 M x 768, cosine; one warp -> two): 256 queries 0.465 -> 0.379 ms, 1024: 1.73 -> 1.22,
  // 4096: 5.50 -> 4.88, C2 (10 000): filter 11.71 -> 10.99 ms, a 125 k-row shard (8 GPUs): 2.00 -> 1.38 ms.
  // (Single CTAs on short rows with FOUR warps per quarter were measured too — 1 M x 128, 16 384 queries: 7.78 -> 7.57 ms, small
  // batches slower, twice the candidates to merge — and dropped: short rows are not bound by the latency of the gate's chain.)
  const uint32_t pair_ew = s->opt_tensor_pair_ew > 0 ? (uint32_t)std::min<int64_t>(s->opt_tensor_pair_ew, 2) : 2u;
  const uint32_t ew = pair_kernel ? pair_ew : (s->kpad >= 640 || dbg_scores) ? 1u : 2u;
  const uint32_t n_lists = n_chunks * ew;  // candidate lists per query
  // few lists per query (huge batches) concentrate the global top-k in one list: at wide rows, where
  // the certificate needs a bigger margin, keep 32 per list so that its threshold stays far below
  if (stream_a && n_lists < 8 && !dbg_scores) kprime = 32u;
  const uint32_t n_cand = n_lists * kprime;
  const uint32_t n_pad = std::max(32u, next_pow2(n_cand));
  // rows handed to the exact rerank. The certificate's worst-case rounding bound grows like D while
  // the gaps between order statistics grow like sqrt(D): wide rows get a deeper cut.
  const uint32_t kpp = std::min(n_pad, s->kpad > TF_MAX_KPAD ? std::max(64u, next_pow2(4 * k)) : std::max(32u, next_pow2(2 * k)));

  // B operand tensor map: mirror [rows][kpad] bf16, box {64, BN}, 128B swizzle, OOB rows read as zero
  CUtensorMap tmap;
  cuuint64_t gdim[2] = {s->kpad, n_rows};
  cuuint64_t gstr[1] = {(cuuint64_t)s->kpad * 2};
  cuuint32_t box[2] = {TF_BK, pair_kernel ? 64u : BN};   // (a CTA of a pair loads 64 of the tile's 128 rows)
  cuuint32_t estr[2] = {1, 1};
  CUresult cr = enc(&tmap, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, s->d_mirror, gdim, gstr, box, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
                    CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (cr != CUDA_SUCCESS) return fail(SCN_ERR_INTERNAL, "cuTensorMapEncodeTiled failed (%d)", (int)cr);

  // merge -> rerank -> certificate in one launch (finish_queries_kernel) where the launch gaps behind the filter are
  // what the caller waits for; big batches keep the three grids, which use the machine better (measured: even at 1024
  // queries, 0.085 ms either way; 2048: 0.16 ms fused, about 0.10 in three grids). Auto: up to 512 queries.
  const size_t fin_smem = fin_smem_bytes(n_pad, s->pitch, kpp);
  const bool fused_tail = !dbg_scores && fin_smem <= 160 * 1024 && kpp <= 64 &&
                          (s->opt_tensor_fused > 0 || (s->opt_tensor_fused < 0 && nq <= 512));

  // lists of a query exchange bounds while they are built (publish_two_best) when there are at least 16 of them — few
  // queries spread over many row chunks; the deep cut of wide rows (k'' = 64) keeps its own, looser thresholds
  // Where it pays (measured, 1 M rows): where the gate is what bounds the kernel — short rows (128 elements: 16 queries 0.220 ->
  // 0.171 ms, 1024 queries 0.99 -> 0.64 ms) — while the lists are cold for most of their life (at most two waves of work items:
  // 16 384 queries 7.1 -> 7.8 ms) and a warp gates more than a few real queries (one query: 0.125 -> 0.150 ms). Long rows do not
  // need it: small batches stream the mirror at HBM rate (768 elements, <= 128 queries: +4 %), and the pair kernel with two warps per
  // quarter keeps up with its MMAs anyway (256 .. 4096 queries: +-1 %).
  const uint32_t lists_pad = (n_lists + 15) / 16 * 16;
  const uint32_t items = (pair_kernel ? (n_qb + 1) / 2 : n_qb) * n_chunks;
  const bool share = !stream_a && !dbg_scores && n_lists >= 16 && (pair_kernel || ew == 2) &&
                     (s->opt_tensor_share > 1 || (s->opt_tensor_share == 1 && !pair_kernel && s->kpad <= 256 && items <= 2 * (uint32_t)sms && nq >= 8));

  Scratch scratch(stream);
  __nv_bfloat16* d_qb = nullptr;
  float4* d_qstat = nullptr;
  float *d_cscore = nullptr, *d_ctau = nullptr, *d_tau = nullptr, *d_pub = nullptr;
  uint32_t *d_crow = nullptr, *d_rows = nullptr, *d_fail = nullptr, *d_nfail = nullptr;
  {  // one allocation for the whole call
    const size_t sizes[] = {(size_t)nq_pad * s->kpad * 2, nq * sizeof(float4), (size_t)nq * n_cand * 4, (size_t)nq * n_cand * 4,
                            (size_t)nq * n_lists * 4, nq * 4, (size_t)nq * kpp * 4, nq * 4, nq * 4, nq * 4, 2 * 4, nq * 4,
                            share ? (size_t)nq_pad * lists_pad * 8 : 0};
    size_t total = 0;
    for (size_t b : sizes) total += Scratch::padded(b);
    SCN_TRY(scratch.reserve(total));
  }
  SCN_TRY(scratch.alloc(&d_qb, (size_t)nq_pad * s->kpad));
  SCN_TRY(scratch.alloc(&d_qstat, nq));
  SCN_TRY(scratch.alloc(&d_cscore, (size_t)nq * n_cand));
  SCN_TRY(scratch.alloc(&d_crow, (size_t)nq * n_cand));
  SCN_TRY(scratch.alloc(&d_ctau, (size_t)nq * n_lists));
  SCN_TRY(scratch.alloc(&d_tau, nq));
  SCN_TRY(scratch.alloc(&d_rows, (size_t)nq * kpp));
  float* d_tau_chunks = nullptr;
  uint32_t *d_fail2 = nullptr;
  SCN_TRY(scratch.alloc(&d_tau_chunks, nq));
  SCN_TRY(scratch.alloc(&d_fail, nq));
  SCN_TRY(scratch.alloc(&d_fail2, nq));
  SCN_TRY(scratch.alloc(&d_nfail, 2));
  uint32_t* d_hint = nullptr;
  SCN_TRY(scratch.alloc(&d_hint, nq));
  if (share) SCN_TRY(scratch.alloc(&d_pub, (size_t)nq_pad * lists_pad * 2));

  if (prof) prof->begin("prep_queries");
  prep_queries_kernel<<<(nq_pad + 3) / 4, 128, 0, stream>>>(d_q, (uint32_t)nq, nq_pad, s->dim, s->kpad, d_qb, d_qstat, d_hint, d_nfail,
                                                            first_batch ? s->d_counters : nullptr, d_pub, lists_pad);
  SCN_LAUNCHED();
  if (prof) prof->end();

  // A operand tensor map (streamed variant only): bf16 queries [nq_pad][kpad], same tiling as the rows
  CUtensorMap tmap_a = tmap;
  if (stream_a) {
    cuuint64_t adim[2] = {s->kpad, nq_pad};
    cuuint32_t abox[2] = {TF_BK, TF_BM};
    cr = enc(&tmap_a, CU_TENSOR_MAP_DATA_TYPE_BFLOAT16, 2, d_qb, adim, gstr, abox, estr, CU_TENSOR_MAP_INTERLEAVE_NONE,
             CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B, CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
    if (cr != CUDA_SUCCESS) return fail(SCN_ERR_INTERNAL, "cuTensorMapEncodeTiled (queries) failed (%d)", (int)cr);
  }

  FilterArgs fa;
  fa.qb = d_qb;
  fa.aux = s->d_aux;
  fa.nq = (uint32_t)nq;
  fa.n_rows = n_rows;
  fa.kpad = s->kpad;
  fa.n_qblocks = n_qb;
  fa.n_chunks = n_chunks;
  fa.tiles_per_chunk = tiles_per_chunk;
  fa.n_tiles = n_tiles;
  fa.coef = (s->metric == M_L2) ? -2.0f : -1.0f;
  fa.cand_score = d_cscore;
  fa.cand_row = d_crow;
  fa.chunk_tau = d_ctau;
  fa.dbg_scores = dbg_scores;
  fa.hint = s->opt_tensor_hint ? d_hint : nullptr;
  fa.hint_target = (uint32_t)std::max<int64_t>(s->opt_tensor_hint_target, 0);   // 0: publish the list's own k'-th best (never costs a candidate)
  fa.pub = d_pub;
  fa.lists_pad = lists_pad;
  fa.nq_stride = nq_pad;
  int grid = (int)std::min<uint32_t>((uint32_t)sms, n_qb * n_chunks);
  const bool pdl = s->opt_pdl != 0 && !(prof && prof->on) && !dbg_scores;   // chained launches (common.cuh): prep -> filter -> tail
  if (prof) prof->begin("tensor_filter");
  const bool two_buf = s->kpad <= 512;
  int32_t rc;
  if (pair_kernel) {
    const int pairs = (int)std::min<uint32_t>((uint32_t)resident_pairs, ((n_qb + 1) / 2) * n_chunks);
    if (ew == 2) {
      if (s->kpad <= 512) rc = (kprime == 16) ? launch_filter2<16, 2, 2>(tmap, fa, pairs, stream, pdl) : launch_filter2<32, 2, 2>(tmap, fa, pairs, stream, pdl);
      else rc = (kprime == 16) ? launch_filter2<16, 1, 2>(tmap, fa, pairs, stream, pdl) : launch_filter2<32, 1, 2>(tmap, fa, pairs, stream, pdl);
    } else {
      if (s->kpad <= 512) rc = (kprime == 16) ? launch_filter2<16, 2, 1>(tmap, fa, pairs, stream, pdl) : launch_filter2<32, 2, 1>(tmap, fa, pairs, stream, pdl);
      else rc = (kprime == 16) ? launch_filter2<16, 1, 1>(tmap, fa, pairs, stream, pdl) : launch_filter2<32, 1, 1>(tmap, fa, pairs, stream, pdl);
    }
  } else if (stream_a) {
    if (dbg_scores) rc = launch_filter<16, 2, 1, true, true, 128>(tmap, tmap_a, fa, grid, stream, pdl);
    else rc = (kprime == 16) ? launch_filter<16, 2, 1, false, true, 128>(tmap, tmap_a, fa, grid, stream, pdl)
                             : launch_filter<32, 2, 1, false, true, 128>(tmap, tmap_a, fa, grid, stream, pdl);
  } else if (dbg_scores) {
    rc = two_buf ? launch_filter<16, 2, 1, true, false, 128>(tmap, tmap_a, fa, grid, stream, pdl)
                 : launch_filter<16, 1, 1, true, false, 128>(tmap, tmap_a, fa, grid, stream, pdl);
  } else if (half_tiles) {
    rc = (kprime == 16) ? launch_filter<16, 2, 1, false, false, 64>(tmap, tmap_a, fa, grid, stream, pdl)
                        : launch_filter<32, 2, 1, false, false, 64>(tmap, tmap_a, fa, grid, stream, pdl);
  } else if (ew == 1) {
    rc = (kprime == 16) ? launch_filter<16, 1, 1, false, false, 128>(tmap, tmap_a, fa, grid, stream, pdl)
                        : launch_filter<32, 1, 1, false, false, 128>(tmap, tmap_a, fa, grid, stream, pdl);
  } else if (share) {
    rc = (kprime == 16) ? launch_filter<16, 2, 2, false, false, 128, true>(tmap, tmap_a, fa, grid, stream, pdl)
                        : launch_filter<32, 2, 2, false, false, 128, true>(tmap, tmap_a, fa, grid, stream, pdl);
  } else {
    rc = (kprime == 16) ? launch_filter<16, 2, 2, false, false, 128>(tmap, tmap_a, fa, grid, stream, pdl)
                        : launch_filter<32, 2, 2, false, false, 128>(tmap, tmap_a, fa, grid, stream, pdl);
  }
  if (prof) prof->end();
  SCN_TRY(rc);

  if (fused_tail) {
    FinishArgs fin;
    fin.cand_score = d_cscore;
    fin.cand_row = d_crow;
    fin.chunk_tau = d_ctau;
    fin.n_lists = n_lists;
    fin.kprime = kprime;
    fin.n_pad = n_pad;
    fin.kpp = kpp;
    fin.vec = s->d_vec;
    fin.norm = s->d_norm;
    fin.deleted = s->d_deleted;
    fin.pitch = s->pitch;
    fin.dim = s->dim;
    fin.n_rows = n_rows;
    fin.kpad = s->kpad;
    fin.q = d_q;
    fin.qstat = d_qstat;
    fin.bounds = s->d_bounds;
    fin.k = k;
    fin.row_base = (uint32_t)row_base;
    fin.out_keys = d_out_keys;
    fin.out_tau_chunks = d_tau_chunks;
    fin.fail_list = d_fail;
    fin.fail_count = d_nfail;
    fin.counters = s->d_counters;
    if (prof) prof->begin("finish_queries");
#define FIN(MT)                                                                              \
  do {                                                                                       \
    SCN_ALLOW_SMEM((finish_queries_kernel<MT>), fin_smem);                                   \
    SCN_CUDA(launch_chained(finish_queries_kernel<MT>, dim3((unsigned)nq), dim3(FIN_THREADS), fin_smem, stream, pdl, fin)); \
  } while (0)
    switch (s->metric) {
      case M_L2: FIN(M_L2); break;
      case M_COS: FIN(M_COS); break;
      default: FIN(M_IP); break;
    }
#undef FIN
    SCN_LAUNCHED();
    if (prof) prof->end();
  } else {
    const size_t merge_smem = (size_t)std::max(n_pad, 256u) * 8;
    if (merge_smem > 200 * 1024) return fail(SCN_ERR_INVALID_PARAMETERS, "too many candidate lists (%u) for one merge block", n_lists);
    SCN_ALLOW_SMEM((merge_candidates_kernel), merge_smem);
    if (prof) prof->begin("merge_candidates");
    // one block per query; long candidate lists (small batches: many chunks) get more threads per sort
    merge_candidates_kernel<<<(unsigned)nq, std::min(512u, std::max(128u, n_pad / 4)), merge_smem, stream>>>(d_cscore, d_crow, d_ctau, n_lists, kprime,
                                                                                                              n_pad, kpp, d_rows, d_tau, d_tau_chunks);
    SCN_LAUNCHED();
    if (prof) prof->end();

    if (prof) prof->begin("rerank_exact");
    SCN_TRY(rerank_rows(s, d_q, nq, d_rows, kpp, k, row_base, d_out_keys, stream));
    if (prof) prof->end();

    if (prof) prof->begin("certify");
    certify_kernel<<<(unsigned)((nq + 127) / 128), 128, 0, stream>>>(d_out_keys, d_tau, d_qstat, s->d_bounds, (uint32_t)nq, k, s->dim,
                                                                     s->kpad, s->metric, nullptr, nullptr, d_fail, d_nfail,
                                                                     s->d_counters, 1);
    SCN_LAUNCHED();
    if (prof) prof->end();
  }

  // second chance for the few queries whose k'' rows were not enough: rerank every candidate the
  // chunks kept (n_chunks * k') and certify against the chunk thresholds alone
  const uint32_t* d_nfail_final = d_nfail + 1;
  if (n_cand > kpp) {
    if (prof) prof->begin("rerank_exact_wide");
    SCN_TRY(rerank_rows(s, d_q, nq, d_crow, n_cand, k, row_base, d_out_keys, stream, d_fail, d_nfail));
    if (prof) prof->end();
    if (prof) prof->begin("certify_wide");
    SCN_CUDA(launch_chained(certify_kernel, dim3((unsigned)((nq + 127) / 128)), dim3(128), 0, stream, pdl, (const uint64_t*)d_out_keys,
                            (const float*)d_tau_chunks, (const float4*)d_qstat, (const float*)s->d_bounds, (uint32_t)nq, k, s->dim, s->kpad,
                            (int)s->metric, (const uint32_t*)d_fail, (const uint32_t*)d_nfail, d_fail2, d_nfail + 1, s->d_counters, 2));
    SCN_LAUNCHED();
    if (prof) prof->end();
  } else {
    d_fail2 = d_fail;  // nothing wider to look at: the first list of failures is the final one
    d_nfail_final = d_nfail;
  }

  // exact scan for whatever could still not be certified (normally nothing: the kernel exits at once)
  return flat_search_exact(s, d_q, d_fail2, d_nfail_final, nq, k, row_base, d_out_keys, stream, prof);
}

int32_t flat_search_tensor(scn_store* s, const float* d_q, uint64_t nq, uint32_t k, uint64_t row_base, uint64_t* d_out_keys,
                           cudaStream_t stream, Profiler* prof) {
  // (the device counters are cleared by the first batch's prep_queries)
  // bound the scratch of the (normally idle) exact fallback: sms * nq * k * 8 bytes of partial lists
  const uint64_t max_batch = std::max<uint64_t>(1024, (512ull << 20) / ((uint64_t)160 * k * 8));
  for (uint64_t q0 = 0; q0 < nq; q0 += max_batch) {
    uint64_t n = std::min(max_batch, nq - q0);
    SCN_TRY(flat_search_tensor_batch(s, d_q + q0 * s->dim, n, k, row_base, d_out_keys + q0 * k, stream, prof, nullptr, q0 == 0));
  }
  return SCN_OK;
}

int32_t tensor_debug_scores(scn_store* s, const float* d_q, uint64_t nq, float* d_scores, cudaStream_t stream) {
  Scratch scratch(stream);
  uint64_t* d_keys = nullptr;
  SCN_TRY(scratch.alloc(&d_keys, nq * 10));
  return flat_search_tensor_batch(s, d_q, nq, 10, 0, d_keys, stream, nullptr, d_scores, true);
}

int32_t mark_aux_deleted(scn_store* s, const uint32_t* h_rows, uint32_t n, cudaStream_t stream) {
  if (!n) return SCN_OK;
  Scratch scratch(stream);
  uint32_t* d_rows = nullptr;
  SCN_TRY(scratch.alloc(&d_rows, n));
  SCN_CUDA(cudaMemcpyAsync(d_rows, h_rows, n * sizeof(uint32_t), cudaMemcpyHostToDevice, stream));
  set_aux_kernel<<<(n + 127) / 128, 128, 0, stream>>>(s->d_aux, d_rows, n, INFINITY);
  SCN_LAUNCHED();
  SCN_CUDA(cudaStreamSynchronize(stream));
  return SCN_OK;
}

}  // namespace scn
