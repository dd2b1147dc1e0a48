// Full-catalog scoring without the (users x items) score matrix: tensor cores FILTER, fp32 decides.
//
// The reference scores every item for every eval user, masks the seen ones and sorts the row once per
// metric object (revisit_bpr/models/bpr/model.py:43-47,131-145, experiments/bpr/exp.py:369-374,
// revisit_bpr/metrics/metric.py:110-113).  Only the top-k of a row matters, and ranks must be the
// fp32 ranks.  So:
//   pack      item rows (and the bias as one extra K column) / gathered user rows into K-padded
//             fp32 panels, with their norms;
//   pass A    S~ = U V^T on tcgen05 (kind::tf32, operands by TMA with 128-byte swizzle, fp32
//             accumulators in TMEM, 128x128 tiles, double-buffered accumulators so the epilogue of a
//             tile overlaps the MMAs of the next).  The epilogue — 8 warps, one thread per (user row,
//             half of the tile's columns), straight from TMEM — keeps only the MAX of every 16
//             consecutive items.  Item 0 and the rows padding the catalogue to a multiple of 128 carry
//             -1e30 in an extra K column, so they never are a maximum; seen items are NOT masked here
//             (a per-element select would triple the epilogue), except for the few users with more
//             than 64 seen items;
//   select    tau~ = the (k + n_seen)-th largest group maximum of the user: the group maxima above
//             it are k + n_seen distinct items, at most n_seen of them seen, so at least k unseen items
//             score tau~ or more — a lower bound of the k-th largest unseen score (users with masked
//             maxima: simply the k-th largest);
//   pass B    the same contraction again (cheaper than storing 200 M scores); the epilogue emits the
//             items with S~ >= tau~ - 2 eps.  The panels hold operands ROUNDED to TF32 (cvt.rna, so the
//             tensor core's own fp32->tf32 handling is exact): every product is off by at most
//             (2 * 2^-11 + 2^-22) |a||b|, hence |S~ - S| <= eps = 1.01 * 2^-10 |u| max|v| by
//             Cauchy-Schwarz, and the emitted set is a superset of the exact top-k (ties included),
//             ~1.2 k items per user;
//   rescore   exact fp32 FMA scores of the candidates (same arithmetic as score_gemm), sort by (score
//             desc, item asc), hits / NDCG / Recall / Precision / MAP (score_common.cuh).
// DRAM traffic: the panels, a bitmask of the seen items, 1/16 of the score matrix as group maxima,
// the candidate lists — against 2 x 800 MB written and ~6 passes re-read by the dense path at 10 k
// users of the ML-20M shape.  Users whose candidate list overflows (mass ties, e.g. an all-zero
// user row) and shapes outside the tensor path (tiny catalogues, D + bias > 288) take the dense
// fp32 path of score.cu; `RBPR_NO_TC_SCORE=1` forces it everywhere.
#include <cuda.h>

#include "score_common.cuh"

namespace {

constexpr int TM = 128;            // users per tile (UMMA M)
constexpr int TN = 128;            // items per tile (UMMA N)
constexpr int KBLK = 32;           // fp32 per 128-byte swizzle row = one K block
constexpr int KB_MAX = 9;          // K blocks of the resident user panel (D + bias <= 288)
constexpr int GROUP = 16;          // items per group maximum
constexpr int NGT = TN / GROUP;    // group maxima per tile and user
constexpr int CAND_CAP = 1024;     // candidates per user the ranking kernel takes
constexpr int LIST_CAP = 256;      // slots of one private candidate list (user, CTA segment, column half)
constexpr int kEpiWarps = 8;       // two warps per TMEM lane quarter, 64 of the tile's 128 columns each
constexpr int kTcThreads = 64 + 32 * kEpiWarps;  // warp 0: TMA producer, warp 1: MMA issuer, warps 2..9: epilogue
constexpr float kNever = -1e30f;   // score offset of item 0 / padding rows (exact in TF32)
constexpr uint32_t kStageBytes = TN * KBLK * 4;  // 16 KiB: one K block of a 128-row panel
constexpr int kMinItems = 8192;    // below this the dense path is used

struct TcParams {
  int n_users;       // rows of this block of users
  int n_utiles, n_itiles, KB, stages;
  int pass;          // 0: group maxima, 1: candidates
  int segs;          // CTA segments a user tile can be split into (private candidate lists)
  int pair;          // 1: a unit is TWO user tiles x one item tile (every item K-block feeds two accumulators:
                     //    half the L2 traffic of the item panel, which is what bounds the passes)
  const uint2* mask;       // (n_users, n_itiles, 2) 64 bits per (user, tile, column half): 1 = masked
  float* gmax;             // (n_users, n_itiles * NGT)
  const float* thr;        // (n_users)
  const uint8_t* heavy;    // (n_users) 1: many seen items -> this user's group maxima are taken over UNSEEN items only
  int32_t* cnt;            // (n_users, segs, 2) candidates in each private list (> LIST_CAP: overflow)
  int32_t* cand;           // (n_users, segs, 2, LIST_CAP)
  int32_t* err;
};

// ---- PTX wrappers ---------------------------------------------------------------------------------
__device__ __forceinline__ uint32_t smem_u32(const void* p) { return (uint32_t)__cvta_generic_to_shared(p); }

__device__ __forceinline__ void mbar_init(uint64_t* bar, uint32_t count) {
  asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(smem_u32(bar)), "r"(count) : "memory");
}
__device__ __forceinline__ void mbar_expect_tx(uint64_t* bar, uint32_t bytes) {
  asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(smem_u32(bar)), "r"(bytes) : "memory");
}
__device__ __forceinline__ void mbar_arrive(uint64_t* bar) {
  asm volatile("mbarrier.arrive.shared::cta.b64 _, [%0];" ::"r"(smem_u32(bar)) : "memory");
}
// Bounded wait: a pipeline bug must end in an error flag, never in a hung GPU.
__device__ __forceinline__ bool mbar_wait(uint64_t* bar, uint32_t parity, int32_t* err) {
  const uint32_t addr = smem_u32(bar);
  for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
    uint32_t ok;
    asm volatile(
        "{\n\t.reg .pred p;\n\t"
        "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n\t"
        "selp.u32 %0, 1, 0, p;\n\t}"
        : "=r"(ok)
        : "r"(addr), "r"(parity)
        : "memory");
    if (ok) return true;
    if (spin > 1024) __nanosleep(64);
  }
  atomicExch(err, 11);
  return false;
}
__device__ __forceinline__ void tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1, uint64_t* bar) {
  asm volatile(
      "cp.async.bulk.tensor.2d.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1, {%3, %4}], [%2];" ::"r"(
          smem_u32(dst)),
      "l"(reinterpret_cast<uint64_t>(map)), "r"(smem_u32(bar)), "r"(c0), "r"(c1)
      : "memory");
}
__device__ __forceinline__ void tc_fence_before() { asm volatile("tcgen05.fence::before_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_fence_after() { asm volatile("tcgen05.fence::after_thread_sync;" ::: "memory"); }
__device__ __forceinline__ void tc_commit(uint64_t* bar) {
  asm volatile("tcgen05.commit.cta_group::1.mbarrier::arrive::one.shared::cluster.b64 [%0];" ::"r"(smem_u32(bar))
               : "memory");
}
__device__ __forceinline__ void tc_mma_tf32(uint32_t tmem_d, uint64_t adesc, uint64_t bdesc, uint32_t idesc,
                                            uint32_t accumulate) {
  asm volatile(
      "{\n\t.reg .pred p;\n\t"
      "setp.ne.b32 p, %4, 0;\n\t"
      "tcgen05.mma.cta_group::1.kind::tf32 [%0], %1, %2, %3, p;\n\t}" ::"r"(tmem_d),
      "l"(adesc), "l"(bdesc), "r"(idesc), "r"(accumulate)
      : "memory");
}
// 32 lanes x 32 consecutive fp32 columns -> 32 registers per thread (thread i <- TMEM lane base + i);
// issue only: tc_ld_wait() before the values are used
__device__ __forceinline__ void tc_ld32(uint32_t taddr, uint32_t* r) {
  asm volatile(
      "tcgen05.ld.sync.aligned.32x32b.x32.b32 "
      "{%0, %1, %2, %3, %4, %5, %6, %7, %8, %9, %10, %11, %12, %13, %14, %15, "
      "%16, %17, %18, %19, %20, %21, %22, %23, %24, %25, %26, %27, %28, %29, %30, %31}, [%32];"
      : "=r"(r[0]), "=r"(r[1]), "=r"(r[2]), "=r"(r[3]), "=r"(r[4]), "=r"(r[5]), "=r"(r[6]), "=r"(r[7]), "=r"(r[8]),
        "=r"(r[9]), "=r"(r[10]), "=r"(r[11]), "=r"(r[12]), "=r"(r[13]), "=r"(r[14]), "=r"(r[15]), "=r"(r[16]),
        "=r"(r[17]), "=r"(r[18]), "=r"(r[19]), "=r"(r[20]), "=r"(r[21]), "=r"(r[22]), "=r"(r[23]), "=r"(r[24]),
        "=r"(r[25]), "=r"(r[26]), "=r"(r[27]), "=r"(r[28]), "=r"(r[29]), "=r"(r[30]), "=r"(r[31])
      : "r"(taddr)
      : "memory");
}
__device__ __forceinline__ void tc_ld_wait() { asm volatile("tcgen05.wait::ld.sync.aligned;" ::: "memory"); }

// Shared-memory matrix descriptor of a K-major panel written by TMA with 128-byte swizzle: rows of
// 128 B, 8-row groups 1024 B apart (SBO), descriptor version 1 (sm_100), layout SWIZZLE_128B.
__device__ __forceinline__ uint64_t umma_desc(uint32_t saddr) {
  const uint32_t lo = ((saddr & 0x3FFFFu) >> 4) | (1u << 16);
  const uint32_t hi = (1024u >> 4) | (1u << 14) | (2u << 29);
  return (uint64_t)lo | ((uint64_t)hi << 32);
}
// kind::tf32, fp32 accumulate, A and B K-major, M = 128, N = 128
constexpr uint32_t kIdesc = (1u << 4) | (2u << 7) | (2u << 10) | ((uint32_t)(TN >> 3) << 17) | ((uint32_t)(TM >> 4) << 24);

__device__ __forceinline__ float round_tf32(float x) {
  uint32_t r;
  asm("cvt.rna.tf32.f32 %0, %1;" : "=r"(r) : "f"(x));
  return __uint_as_float(r);
}

struct Work {  // the CTA's contiguous range of (user tile, item tile) units, user-tile major
  long long begin, end;
};
// first CTA whose range contains unit x, for ranges [total*b/G, total*(b+1)/G)
__host__ __device__ inline long long first_cta_of(long long x, long long total, long long G) {
  return ((x + 1) * G - 1) / total;
}

__device__ __forceinline__ float max16(const uint32_t* r) {  // tree: independent FMNMX chains
  float a = fmaxf(__uint_as_float(r[0]), __uint_as_float(r[1])), b = fmaxf(__uint_as_float(r[2]), __uint_as_float(r[3]));
  float c = fmaxf(__uint_as_float(r[4]), __uint_as_float(r[5])), d = fmaxf(__uint_as_float(r[6]), __uint_as_float(r[7]));
  float e = fmaxf(__uint_as_float(r[8]), __uint_as_float(r[9])), f = fmaxf(__uint_as_float(r[10]), __uint_as_float(r[11]));
  float g = fmaxf(__uint_as_float(r[12]), __uint_as_float(r[13])), h = fmaxf(__uint_as_float(r[14]), __uint_as_float(r[15]));
  return fmaxf(fmaxf(fmaxf(a, b), fmaxf(c, d)), fmaxf(fmaxf(e, f), fmaxf(g, h)));
}
__device__ __forceinline__ float max16_masked(const uint32_t* r, uint32_t bits) {  // bit j set: column j is masked
  float m = -INFINITY;
#pragma unroll
  for (int j = 0; j < 16; ++j) m = fmaxf(m, ((bits >> j) & 1u) ? -INFINITY : __uint_as_float(r[j]));
  return m;
}
__device__ __forceinline__ uint32_t ge_mask32(const uint32_t* r, float thr) {  // bit j = r[j] >= thr
  uint32_t m0 = 0u, m1 = 0u, m2 = 0u, m3 = 0u;  // four independent chains
#pragma unroll
  for (int j = 0; j < 8; ++j) {
    m0 |= (__uint_as_float(r[j]) >= thr ? 1u : 0u) << j;
    m1 |= (__uint_as_float(r[8 + j]) >= thr ? 1u : 0u) << (8 + j);
    m2 |= (__uint_as_float(r[16 + j]) >= thr ? 1u : 0u) << (16 + j);
    m3 |= (__uint_as_float(r[24 + j]) >= thr ? 1u : 0u) << (24 + j);
  }
  return (m0 | m1) | (m2 | m3);
}

// ---- the contraction + filter kernel --------------------------------------------------------------
__global__ void __launch_bounds__(kTcThreads, 1)
score_tc(const __grid_constant__ CUtensorMap tmA, const __grid_constant__ CUtensorMap tmB, const TcParams p) {
  extern __shared__ __align__(1024) uint8_t smem_raw[];
  uint8_t* smem = reinterpret_cast<uint8_t*>((reinterpret_cast<uintptr_t>(smem_raw) + 1023) & ~(uintptr_t)1023);
  const int UT = 1 + p.pair;                            // user tiles per unit
  uint8_t* sA = smem;                                   // UT x KB x 16 KiB: the user panel(s) of the current unit
  uint8_t* sB = smem + (size_t)UT * p.KB * kStageBytes; // stages x 16 KiB: K blocks of item tiles
  uint64_t* bars = reinterpret_cast<uint64_t*>(sB + (size_t)p.stages * kStageBytes);
  uint64_t* full = bars;                 // [stages]  TMA -> MMA
  uint64_t* empty = bars + p.stages;     // [stages]  MMA -> TMA
  uint64_t* a_full = empty + p.stages;   // user panel landed
  uint64_t* a_free = a_full + 1;         // MMAs reading the user panel retired
  uint64_t* t_full = a_free + 1;         // [2] accumulator ready (MMA -> epilogue)
  uint64_t* t_empty = t_full + 2;        // [2] accumulator drained (epilogue -> MMA)
  uint32_t* tmem_slot = reinterpret_cast<uint32_t*>(t_empty + 2);

  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  if (threadIdx.x == 0) {
    for (int s = 0; s < p.stages; ++s) {
      mbar_init(full + s, 1);
      mbar_init(empty + s, 1);
    }
    mbar_init(a_full, 1);
    mbar_init(a_free, 1);
    for (int s = 0; s < 2; ++s) {
      mbar_init(t_full + s, 1);
      mbar_init(t_empty + s, kEpiWarps);
    }
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
  }
  if (warp == 1) {  // 2 accumulator stages x UT user tiles x 128 fp32 columns (all 512 allocated: one CTA per SM)
    asm volatile("tcgen05.alloc.cta_group::1.sync.aligned.shared::cta.b32 [%0], 512;" ::"r"(smem_u32(tmem_slot)) : "memory");
    asm volatile("tcgen05.relinquish_alloc_permit.cta_group::1.sync.aligned;" ::: "memory");
  }
  tc_fence_before();
  __syncthreads();
  tc_fence_after();
  const uint32_t tmem_base = *tmem_slot;

  const int n_uunits = (p.n_utiles + UT - 1) / UT;       // user tiles (or pairs of them)
  const long long total = (long long)n_uunits * p.n_itiles;
  Work w;
  w.begin = total * blockIdx.x / gridDim.x;
  w.end = total * (blockIdx.x + 1) / gridDim.x;

  if (warp == 0) {
    // ===== TMA producer =====
    if (lane == 0) {
      uint32_t q = 0, seg = 0;
      bool ok = true;
      for (long long unit = w.begin; unit < w.end && ok;) {
        const int ut = (int)(unit / p.n_itiles), it0 = (int)(unit % p.n_itiles);
        const long long left = w.end - unit;
        const int it1 = (int)((long long)(p.n_itiles - it0) < left ? p.n_itiles : it0 + left);
        ok = mbar_wait(a_free, (seg & 1u) ^ 1u, p.err);
        if (!ok) break;
        mbar_expect_tx(a_full, (uint32_t)(UT * p.KB) * kStageBytes);
        for (int h = 0; h < UT; ++h)
          for (int kb = 0; kb < p.KB; ++kb)
            tma_load_2d(sA + (size_t)(h * p.KB + kb) * kStageBytes, &tmA, kb * KBLK, (ut * UT + h) * TM, a_full);
        for (int it = it0; it < it1 && ok; ++it)
          for (int kb = 0; kb < p.KB; ++kb, ++q) {
            const uint32_t st = q % (uint32_t)p.stages, ph = (q / (uint32_t)p.stages) & 1u;
            ok = mbar_wait(empty + st, ph ^ 1u, p.err);
            if (!ok) break;
            mbar_expect_tx(full + st, kStageBytes);
            tma_load_2d(sB + (size_t)st * kStageBytes, &tmB, kb * KBLK, it * TN, full + st);
          }
        unit += it1 - it0;
        ++seg;
      }
    }
  } else if (warp == 1) {
    // ===== MMA issuer (one thread) =====
    if (lane == 0) {
      uint32_t q = 0, seg = 0, t = 0;
      bool ok = true;
      for (long long unit = w.begin; unit < w.end && ok;) {
        const int it0 = (int)(unit % p.n_itiles);
        const long long left = w.end - unit;
        const int it1 = (int)((long long)(p.n_itiles - it0) < left ? p.n_itiles : it0 + left);
        ok = mbar_wait(a_full, seg & 1u, p.err);
        if (!ok) break;
        tc_fence_after();
        for (int it = it0; it < it1 && ok; ++it, ++t) {
          const uint32_t acc = t & 1u, aph = (t >> 1) & 1u;
          ok = mbar_wait(t_empty + acc, aph ^ 1u, p.err);
          if (!ok) break;
          tc_fence_after();
          for (int kb = 0; kb < p.KB; ++kb, ++q) {
            const uint32_t st = q % (uint32_t)p.stages, ph = (q / (uint32_t)p.stages) & 1u;
            ok = mbar_wait(full + st, ph, p.err);
            if (!ok) break;
            tc_fence_after();
            const uint64_t bd = umma_desc(smem_u32(sB + (size_t)st * kStageBytes));
            for (int h = 0; h < UT; ++h) {  // the same item K-block against every resident user tile
              const uint32_t d = tmem_base + (acc * (uint32_t)UT + (uint32_t)h) * (uint32_t)TN;
              const uint64_t ad = umma_desc(smem_u32(sA + (size_t)(h * p.KB + kb) * kStageBytes));
#pragma unroll
              for (int k4 = 0; k4 < KBLK / 8; ++k4)  // UMMA K = 8 fp32 = 32 bytes inside the swizzled row
                tc_mma_tf32(d, ad + (uint64_t)(k4 * 2), bd + (uint64_t)(k4 * 2), kIdesc, (kb | k4) != 0 ? 1u : 0u);
            }
            tc_commit(empty + st);  // frees the stage when these MMAs have read it
          }
          if (ok) tc_commit(t_full + acc);
        }
        if (ok) tc_commit(a_free);
        unit += it1 - it0;
        ++seg;
      }
    }
  } else {
    // ===== epilogue: 8 warps; thread <-> (TMEM lane = user row of the tile, half of the columns) =====
    const int quarter = warp & 3;          // the TMEM lane quarter this warp may read
    const int half = (warp - 2) >> 2;      // columns [64*half, 64*half + 64) of the tile
    const int row = quarter * 32 + lane;
    uint32_t t = 0;
    bool ok = true;
    for (long long unit = w.begin; unit < w.end && ok;) {
      const int ut = (int)(unit / p.n_itiles), it0 = (int)(unit % p.n_itiles);
      const long long left = w.end - unit;
      const int it1 = (int)((long long)(p.n_itiles - it0) < left ? p.n_itiles : it0 + left);
      const int sg = (int)((long long)blockIdx.x - first_cta_of((long long)ut * p.n_itiles, total, gridDim.x));
      // per resident user tile: the row this thread owns, its threshold, its private candidate list
      int u[2];
      bool valid[2], need_mask[2];
      float thr[2] = {0.f, 0.f};
      int n_cand[2] = {0, 0};
      size_t list[2] = {0, 0};
      uint2 mw_next[2];
#pragma unroll
      for (int h = 0; h < 2; ++h) {
        u[h] = (ut * UT + h) * TM + row;
        valid[h] = h < UT && u[h] < p.n_users;
        mw_next[h] = make_uint2(~0u, ~0u);
        need_mask[h] = false;
        if (valid[h]) {
          list[h] = ((size_t)u[h] * p.segs + (size_t)sg) * 2 + (size_t)half;
          if (p.pass == 1) thr[h] = __ldg(p.thr + u[h]);
          need_mask[h] = p.pass == 1 || __ldg(p.heavy + u[h]) != 0;
          if (need_mask[h]) mw_next[h] = __ldg(p.mask + ((size_t)u[h] * p.n_itiles + it0) * 2 + half);
        }
      }
      for (int it = it0; it < it1 && ok; ++it, ++t) {
        const uint32_t acc = t & 1u, aph = (t >> 1) & 1u;
        uint2 mw[2];
#pragma unroll
        for (int h = 0; h < 2; ++h) {
          mw[h] = mw_next[h];
          if (need_mask[h] && it + 1 < it1)  // one tile ahead
            mw_next[h] = __ldg(p.mask + ((size_t)u[h] * p.n_itiles + it + 1) * 2 + half);
        }
        ok = mbar_wait(t_full + acc, aph, p.err);
        if (!ok) break;
        tc_fence_after();
        for (int h = 0; h < UT; ++h) {
          const uint32_t taddr = tmem_base + ((uint32_t)(quarter * 32) << 16) + (acc * (uint32_t)UT + (uint32_t)h) * (uint32_t)TN +
                                 (uint32_t)(half * 64);
          uint32_t r0[32], r1[32];
          tc_ld32(taddr, r0);
          tc_ld32(taddr + 32u, r1);
          tc_ld_wait();
          if (h == UT - 1) {  // both accumulators of the stage are in registers: hand it back to the MMA warp
            tc_fence_before();
            __syncwarp();
            if (lane == 0) mbar_arrive(t_empty + acc);
          }
          if (p.pass == 0) {
            if (valid[h]) {
              float4* dst = reinterpret_cast<float4*>(p.gmax + (size_t)u[h] * p.n_itiles * NGT + (size_t)it * NGT + half * 4);
              if (need_mask[h])  // heavy user: maxima over unseen items only (the selection then asks for rank k)
                *dst = make_float4(max16_masked(r0, mw[h].x), max16_masked(r0 + 16, mw[h].x >> 16),
                                   max16_masked(r1, mw[h].y), max16_masked(r1 + 16, mw[h].y >> 16));
              else
                *dst = make_float4(max16(r0), max16(r0 + 16), max16(r1), max16(r1 + 16));
            }
          } else {
            uint32_t pm0 = ge_mask32(r0, thr[h]) & ~mw[h].x, pm1 = ge_mask32(r1, thr[h]) & ~mw[h].y;
            const int base = it * TN + half * 64;
            int32_t* my_cand = p.cand + list[h] * LIST_CAP;
            int n = n_cand[h];
            while (pm0 != 0u) {  // rare: ~1 % of the scores
              const int j = __ffs(pm0) - 1;
              pm0 &= pm0 - 1u;
              if (n < LIST_CAP) my_cand[n] = base + j;
              ++n;
            }
            while (pm1 != 0u) {
              const int j = __ffs(pm1) - 1;
              pm1 &= pm1 - 1u;
              if (n < LIST_CAP) my_cand[n] = base + 32 + j;
              ++n;
            }
            n_cand[h] = n;
          }
        }
      }
      if (p.pass == 1 && ok) {
#pragma unroll
        for (int h = 0; h < 2; ++h)
          if (valid[h]) p.cnt[list[h]] = n_cand[h];
      }
      unit += it1 - it0;
    }
  }
  tc_fence_before();
  __syncthreads();
  if (warp == 1) {
    tc_fence_after();
    asm volatile("tcgen05.dealloc.cta_group::1.sync.aligned.b32 %0, 512;" ::"r"(tmem_base) : "memory");
  }
}

// ---- small kernels around it ----------------------------------------------------------------------
// K-padded TF32 panel of the item table: [v_i, bias_i, never_i, 0 ...] with never_i = -1e30 for item 0
// and for the rows padding the catalogue to a multiple of 128 (the users' panel holds 1 in that
// column), 0 otherwise.  One warp per row; also max |[v_i, bias_i]| over the real rows.
__global__ void pack_items(const float* __restrict__ item_emb, const float* __restrict__ item_bias, int I, int D,
                           int with_bias, int Kp, int rows, float* __restrict__ out, uint32_t* __restrict__ vmax_bits) {
  const int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  float sq = 0.f;
  for (int c = lane; c < Kp; c += 32) {
    float v = 0.f;
    if (r < I) {
      if (c < D) v = item_emb[(int64_t)r * D + c];
      else if (c == D && with_bias) v = item_bias[r];
    }
    sq += v * v;
    if (c == D + with_bias && (r == 0 || r >= I)) v = kNever;
    out[(int64_t)r * Kp + c] = round_tf32(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0 && r < I) atomicMax(vmax_bits, __float_as_uint(sqrtf(sq)));  // non-negative floats order as uints
}

// Gathered users' panel [u, 1 (bias column), 1 (never column), 0 ...]; |[u, 1]| for the error bound.
__global__ void pack_users(const float* __restrict__ user_emb, const int64_t* __restrict__ users, int n_users, int D,
                           int Kp, int rows, int with_bias, int64_t U, float* __restrict__ out,
                           float* __restrict__ unorm, int32_t* __restrict__ err) {
  const int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= rows) return;
  int64_t u = 0;
  if (r < n_users) {
    u = users[r];
    if (u < 0 || u >= U) {
      if (lane == 0) atomicExch(err, 7);
      u = 0;
    }
  }
  float sq = 0.f;
  for (int c = lane; c < Kp; c += 32) {
    float v = 0.f;
    if (r < n_users) {
      if (c < D) v = user_emb[u * D + c];
      else if (c == D && with_bias) v = 1.0f;
    }
    sq += v * v;
    if (r < n_users && c == D + with_bias) v = 1.0f;
    out[(int64_t)r * Kp + c] = round_tf32(v);
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) sq += __shfl_xor_sync(0xffffffffu, sq, o);
  if (lane == 0 && r < n_users) unorm[r] = sqrtf(sq);
}

// 128 bits per (user, item tile): item 0, the user's seen items and columns >= I are masked (pass B).
constexpr int kHeavySeen = 64;  // users with more seen items take masked group maxima (pass A)
__global__ void build_mask(const int64_t* __restrict__ seen_indptr, const int32_t* __restrict__ seen_indices,
                           int64_t row0, int n_users, int I, int n_itiles, uint32_t* __restrict__ mask,
                           uint8_t* __restrict__ heavy) {
  const int r = (int)(((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  const int lane = threadIdx.x & 31;
  if (r >= n_users) return;
  if (lane == 0)
    heavy[r] = (seen_indptr != nullptr && seen_indptr[row0 + r + 1] - seen_indptr[row0 + r] > kHeavySeen) ? 1 : 0;
  uint32_t* m = mask + (size_t)r * n_itiles * 4;
  if (lane == 0) atomicOr(m, 1u);
  for (int c = I + lane; c < n_itiles * TN; c += 32) atomicOr(m + (c >> 5), 1u << (c & 31));
  if (seen_indptr != nullptr) {
    const int64_t lo = seen_indptr[row0 + r], hi = seen_indptr[row0 + r + 1];
    for (int64_t q = lo + lane; q < hi; q += 32) {
      const int32_t it = seen_indices[q];
      if (it >= 0 && it < I) atomicOr(m + (it >> 5), 1u << (it & 31));
    }
  }
}

// tau~ = the (k + n_seen)-th largest group maximum of a user, by bisection over the monotone uint32
// float keys held in shared memory (32 rounds, each a strided count + one block reduction), then the
// pass-B threshold tau~ - 2 eps.  One CTA of 128 threads per user.
constexpr int kSelThreads = 128;
__global__ void __launch_bounds__(kSelThreads)
select_threshold(const float* __restrict__ gmax, int G, int k, const int64_t* __restrict__ seen_indptr, int64_t row0,
                 const uint8_t* __restrict__ heavy, const float* __restrict__ unorm, const uint32_t* __restrict__ vmax_bits,
                 float* __restrict__ thr) {
  extern __shared__ uint32_t keys[];  // G
  __shared__ int warp_cnt[kSelThreads / 32];
  const int tid = threadIdx.x;
  const float* row = gmax + (size_t)blockIdx.x * G;
  for (int i = tid; i < G; i += kSelThreads) keys[i] = fkey(row[i]);
  int64_t n_seen = 0;
  if (seen_indptr != nullptr && heavy[blockIdx.x] == 0)  // heavy users' maxima already exclude their seen items
    n_seen = seen_indptr[row0 + blockIdx.x + 1] - seen_indptr[row0 + blockIdx.x];
  const int64_t want64 = (int64_t)k + n_seen;
  __syncthreads();
  if (want64 > (int64_t)G) {  // not enough groups to bound the k-th unseen score: everything is a candidate
    if (tid == 0) thr[blockIdx.x] = -INFINITY;
    return;
  }
  const int want = (int)want64;
  // largest key T with #{keys >= T} >= want  ==  the want-th largest key
  uint32_t T = 0u;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
    for (int i = tid; i < G; i += kSelThreads) c += (keys[i] >= cand);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if ((tid & 31) == 0) warp_cnt[tid >> 5] = c;
    __syncthreads();
    int tot = 0;
#pragma unroll
    for (int q = 0; q < kSelThreads / 32; ++q) tot += warp_cnt[q];
    if (tot >= want) T = cand;
    __syncthreads();
  }
  if (tid == 0) {
    const float tau = ikey(T);
    const float eps2 = 0.002f * unorm[blockIdx.x] * __uint_as_float(*vmax_bits);  // >= 2 * 1.01 * 2^-10 |u'| max|v'|
    thr[blockIdx.x] = tau - eps2;
  }
}

// The same selection with one WARP per user and the keys in registers (G <= 32 * KPL): no block
// barriers, 32 bisection rounds of KPL compares and one warp reduction.
template <int KPL>
__global__ void __launch_bounds__(128)
select_threshold_warp(const float* __restrict__ gmax, int G, int k, int n_users, const int64_t* __restrict__ seen_indptr,
                      int64_t row0, const uint8_t* __restrict__ heavy, const float* __restrict__ unorm,
                      const uint32_t* __restrict__ vmax_bits, float* __restrict__ thr) {
  const int lane = threadIdx.x & 31;
  const int u = (int)((blockIdx.x * blockDim.x + threadIdx.x) >> 5);
  if (u >= n_users) return;
  const float* row = gmax + (size_t)u * G;
  uint32_t key[KPL];
#pragma unroll
  for (int e = 0; e < KPL; ++e) {
    const int i = lane + 32 * e;
    key[e] = i < G ? fkey(__ldg(row + i)) : 0u;  // 0 is below every real key
  }
  int64_t n_seen = 0;
  if (seen_indptr != nullptr && heavy[u] == 0) n_seen = seen_indptr[row0 + u + 1] - seen_indptr[row0 + u];
  const int64_t want64 = (int64_t)k + n_seen;
  if (want64 > (int64_t)G) {
    if (lane == 0) thr[u] = -INFINITY;
    return;
  }
  const int want = (int)want64;
  uint32_t T = 0u;
  for (int bit = 31; bit >= 0; --bit) {
    const uint32_t cand = T | (1u << bit);
    int c = 0;
#pragma unroll
    for (int e = 0; e < KPL; ++e) c += (key[e] >= cand);
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) c += __shfl_xor_sync(0xffffffffu, c, o);
    if (c >= want) T = cand;
  }
  if (lane == 0) thr[u] = ikey(T) - 0.002f * unorm[u] * __uint_as_float(*vmax_bits);
}

// Exact fp32 scores of the candidates, ranking, outputs.  One CTA of 256 threads per user.
constexpr int kRankThreads = 256;
__global__ void __launch_bounds__(kRankThreads)
rescore_rank(const float* __restrict__ user_emb, const float* __restrict__ item_emb, const float* __restrict__ item_bias,
             const int64_t* __restrict__ users, int D, int segs, const int32_t* __restrict__ cnt,
             const int32_t* __restrict__ cand, int32_t* __restrict__ overflow_rows, int32_t* __restrict__ overflow_count,
             TopkParams p) {
  __shared__ unsigned long long key_in[CAND_CAP];
  __shared__ unsigned long long sel[KCAP];
  __shared__ int32_t items[CAND_CAP];
  __shared__ __align__(16) float s_u[1024];
  __shared__ int s_n, s_bad;
  const int tid = threadIdx.x;
  const int64_t urow = blockIdx.x;
  const int n_lists = segs * 2;
  const int32_t* c_row = cnt + (size_t)urow * n_lists;
  // gather the private lists (thread 0 walks the counts: a handful of lists, 2*segs)
  if (tid == 0) {
    int n = 0, bad = 0;
    for (int l = 0; l < n_lists; ++l) {
      const int c = c_row[l];
      bad |= (c > LIST_CAP);
      n += c;
    }
    s_n = n;
    s_bad = bad | (n > CAND_CAP);
  }
  __syncthreads();
  if (s_bad) {  // mass ties: this user goes through the dense path afterwards
    if (tid == 0) overflow_rows[atomicAdd(overflow_count, 1)] = (int32_t)urow;
    return;
  }
  const int n = s_n;
  {
    int off = 0;
    for (int l = 0; l < n_lists; ++l) {
      const int c = c_row[l];
      const int32_t* src = cand + ((size_t)urow * n_lists + l) * LIST_CAP;
      for (int i = tid; i < c; i += kRankThreads) items[off + i] = src[i];
      off += c;
    }
  }
  const int64_t u = users[urow];
  for (int c = tid; c < D; c += kRankThreads) s_u[c] = user_emb[u * D + c];
  for (int i = tid; i < KCAP; i += kRankThreads) sel[i] = 0ull;
  __syncthreads();
  for (int q = tid; q < n; q += kRankThreads) {
    const int32_t it = items[q];
    const float* v = item_emb + (int64_t)it * D;
    float acc = 0.f;  // k ascending fmaf chain: the arithmetic of score_gemm
    int c = 0;
    for (; c + 32 <= D; c += 32) {  // 8 independent 128-bit loads in flight, then the sequential chain
      float4 x[8];
#pragma unroll
      for (int e = 0; e < 8; ++e) x[e] = __ldg(reinterpret_cast<const float4*>(v + c) + e);
#pragma unroll
      for (int e = 0; e < 8; ++e) {
        acc = fmaf(s_u[c + 4 * e], x[e].x, acc);
        acc = fmaf(s_u[c + 4 * e + 1], x[e].y, acc);
        acc = fmaf(s_u[c + 4 * e + 2], x[e].z, acc);
        acc = fmaf(s_u[c + 4 * e + 3], x[e].w, acc);
      }
    }
    for (; c < D; c += 4) {
      const float4 x = *reinterpret_cast<const float4*>(v + c);
      acc = fmaf(s_u[c], x.x, acc);
      acc = fmaf(s_u[c + 1], x.y, acc);
      acc = fmaf(s_u[c + 2], x.z, acc);
      acc = fmaf(s_u[c + 3], x.w, acc);
    }
    if (item_bias != nullptr) acc += __ldg(item_bias + it);
    key_in[q] = ((unsigned long long)fkey(acc) << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)it);
  }
  __syncthreads();
  // Ranks (keys are distinct: the item id is part of the key).  Up to 256 candidates (the usual case):
  // every warp sorts its 32 keys in registers (bitonic network over shuffles), publishes them, and a
  // key's rank is its place in its own warp plus, for every other warp, the number of larger keys
  // there (5-step binary search in that warp's sorted list): ~250 instructions per thread instead of
  // an n-step counting loop.  Above 256 candidates: counting.
  if (n <= kRankThreads) {
    __shared__ unsigned long long sorted[kRankThreads];
    const int lane = tid & 31, wrp = tid >> 5;
    unsigned long long mine = tid < n ? key_in[tid] : 0ull;  // 0 sorts last
#pragma unroll
    for (int size = 2; size <= 32; size <<= 1) {
#pragma unroll
      for (int stride = size >> 1; stride > 0; stride >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, mine, stride);
        const bool upper = (lane & stride) != 0;          // holds the second element of the pair
        const bool desc = (lane & size) == 0;             // this sub-sequence sorts descending
        const bool take_max = (upper != desc);            // first element of a descending pair keeps the larger key
        const unsigned long long hi = mine > other ? mine : other, lo = mine > other ? other : mine;
        mine = take_max ? hi : lo;
      }
    }
    // (lane & 32) == 0 for every lane, so the last merge sorted all 32 keys descending
    sorted[tid] = mine;
    __syncthreads();
    int r = lane;
    for (int w = 0; w < kRankThreads / 32; ++w) {
      if (w == wrp) continue;
      const unsigned long long* lst = sorted + w * 32;
      int lo = 0, hi = 32;  // first position whose key is <= mine == number of keys > mine
      while (lo < hi) {
        const int mid = (lo + hi) >> 1;
        if (lst[mid] > mine) lo = mid + 1; else hi = mid;
      }
      r += lo;
    }
    if (mine != 0ull && r < KCAP) sel[r] = mine;
  } else {
    for (int q = tid; q < n; q += kRankThreads) {
      const unsigned long long mine = key_in[q];
      int r = 0;
      for (int j = 0; j < n; ++j) r += (key_in[j] > mine);  // same address across the warp: broadcast
      if (r < KCAP) sel[r] = mine;
    }
  }
  __syncthreads();
  topk_emit_outputs(p, urow, sel, min(min(p.k_max, p.I), n));
}

typedef CUresult (*fn_encode_tiled)(CUtensorMap*, CUtensorMapDataType, cuuint32_t, void*, const cuuint64_t*,
                                    const cuuint64_t*, const cuuint32_t*, const cuuint32_t*, CUtensorMapInterleave,
                                    CUtensorMapSwizzle, CUtensorMapL2promotion, CUtensorMapFloatOOBfill);

int make_panel_map(rbpr_ctx* ctx, CUtensorMap* map, float* base, int64_t rows, int Kp) {
  static fn_encode_tiled fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    RBPR_CUDA(ctx, cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &sym, cudaEnableDefault, &q));
    if (!sym) RBPR_FAIL(ctx, RBPR_ERR_CUDA, "cuTensorMapEncodeTiled not available");
    fn = (fn_encode_tiled)sym;
  }
  const cuuint64_t gdim[2] = {(cuuint64_t)Kp, (cuuint64_t)rows};
  const cuuint64_t gstride[1] = {(cuuint64_t)Kp * sizeof(float)};
  const cuuint32_t box[2] = {(cuuint32_t)KBLK, (cuuint32_t)TN};
  const cuuint32_t estr[2] = {1, 1};
  const CUresult r = fn(map, CU_TENSOR_MAP_DATA_TYPE_FLOAT32, 2, base, gdim, gstride, box, estr,
                        CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_128B, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
                        CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE);
  if (r != CUDA_SUCCESS) RBPR_FAIL(ctx, RBPR_ERR_CUDA, "cuTensorMapEncodeTiled failed (%d)", (int)r);
  return 0;
}

int grow(rbpr_ctx* ctx, void** ptr, size_t* have, size_t need) {
  if (need <= *have) return 0;
  cudaFree(*ptr);
  *ptr = nullptr;
  *have = 0;
  RBPR_CUDA(ctx, cudaMalloc(ptr, need));
  *have = need;
  return 0;
}

}  // namespace

// Is the tensor path applicable to this context's tables and this request?
bool rbpr_score_tc_eligible(const rbpr_ctx* ctx, int k_max) {
  if (getenv("RBPR_NO_TC_SCORE") != nullptr) return false;
  const int kp = ((ctx->D + (ctx->item_bias ? 1 : 0) + 1 + KBLK - 1) / KBLK) * KBLK;
  return ctx->I >= kMinItems && kp / KBLK <= KB_MAX && k_max <= KCAP && ctx->D % 4 == 0 && ctx->D <= 1024;
}

// One block of users (n_users <= 16384) through the tensor path.  Outputs as TopkParams describes;
// `overflow_host` receives the number of users left to the dense path (their local rows are in
// ctx->tc_overflow_rows).  Synchronises the stream once (to read that count).
int rbpr_score_tc_block(rbpr_ctx* ctx, const int64_t* users, int n_users, const int64_t* seen_indptr,
                        const int32_t* seen_indices, int64_t row0, const TopkParams& tp_in, int* overflow_host,
                        cudaStream_t st) {
  NvtxRange nvtx("rbpr.score_tc_block (pack, pass A, select, pass B, rescore)");
  const int D = ctx->D, I = (int)ctx->I;
  const int with_bias = ctx->item_bias ? 1 : 0;
  const int Kp = ((D + with_bias + 1 + KBLK - 1) / KBLK) * KBLK, KB = Kp / KBLK;  // + the "never" column
  const int n_utiles = (n_users + TM - 1) / TM, n_itiles = (I + TN - 1) / TN;
  // pair mode: two user tiles resident per CTA (2 x KB x 16 KiB) + >= 4 item stages must fit 227 KiB
  const int pair = (n_utiles >= 2 && (size_t)(2 * KB + 4) * kStageBytes + 2048 <= 227 * 1024 && getenv("RBPR_TC_NO_PAIR") == nullptr) ? 1 : 0;
  const int UT = 1 + pair;
  const int n_uunits = (n_utiles + UT - 1) / UT;
  const int urows = n_uunits * UT * TM, irows = n_itiles * TN;
  const int G = n_itiles * NGT;
  // persistent grid: one CTA per SM, contiguous ranges of (user tile [pair], item tile) units; a user
  // tile split over several CTAs gets one private candidate list per CTA segment
  const long long units = (long long)n_uunits * n_itiles;
  const int grid = (int)(units < ctx->sm_count ? units : ctx->sm_count);
  int segs = 1;
  for (int ut = 0; ut < n_uunits; ++ut) {
    const long long a = first_cta_of((long long)ut * n_itiles, units, grid);
    const long long b = first_cta_of((long long)(ut + 1) * n_itiles - 1, units, grid);
    if ((int)(b - a + 1) > segs) segs = (int)(b - a + 1);
  }
  const size_t n_lists = (size_t)n_users * segs * 2;
  // scratch
  int rc = grow(ctx, (void**)&ctx->tc_items, &ctx->tc_items_bytes, (size_t)irows * Kp * sizeof(float));
  if (rc) return rc;
  rc = grow(ctx, (void**)&ctx->tc_users, &ctx->tc_users_bytes, (size_t)urows * Kp * sizeof(float));
  if (rc) return rc;
  rc = grow(ctx, (void**)&ctx->tc_mask, &ctx->tc_mask_bytes, (size_t)n_users * n_itiles * 16);
  if (rc) return rc;
  rc = grow(ctx, (void**)&ctx->tc_gmax, &ctx->tc_gmax_bytes, (size_t)n_users * G * sizeof(float));
  if (rc) return rc;
  rc = grow(ctx, (void**)&ctx->tc_cand, &ctx->tc_cand_bytes, n_lists * LIST_CAP * sizeof(int32_t));
  if (rc) return rc;
  // small arrays in one allocation: thr | unorm | overflow rows | {overflow count, vmax} | list counts
  const size_t small = (size_t)n_users * 3 * sizeof(float) + 256 + n_lists * sizeof(int32_t) + (size_t)n_users + 16;
  rc = grow(ctx, (void**)&ctx->tc_small, &ctx->tc_small_bytes, small);
  if (rc) return rc;
  float* thr = (float*)ctx->tc_small;
  float* unorm = thr + n_users;
  int32_t* ovf_rows = (int32_t*)(unorm + n_users);
  int32_t* ovf_count = ovf_rows + n_users;
  uint32_t* vmax_bits = (uint32_t*)(ovf_count + 1);
  int32_t* cnt = ovf_count + 64;
  uint8_t* heavy = (uint8_t*)(cnt + n_lists);
  ctx->tc_overflow_rows = ovf_rows;

  RBPR_CUDA(ctx, cudaMemsetAsync(ctx->tc_mask, 0, (size_t)n_users * n_itiles * 16, st));
  RBPR_CUDA(ctx, cudaMemsetAsync(ovf_count, 0, 256 + n_lists * sizeof(int32_t), st));  // count, vmax, list counts
  pack_items<<<(unsigned)(((int64_t)irows * 32 + 255) / 256), 256, 0, st>>>(ctx->item_emb, ctx->item_bias, I, D, with_bias,
                                                                            Kp, irows, ctx->tc_items, vmax_bits);
  pack_users<<<(unsigned)(((int64_t)urows * 32 + 255) / 256), 256, 0, st>>>(ctx->user_emb, users, n_users, D, Kp, urows,
                                                                            with_bias, ctx->U, ctx->tc_users, unorm, ctx->flag);
  build_mask<<<(unsigned)(((int64_t)n_users * 32 + 255) / 256), 256, 0, st>>>(seen_indptr, seen_indices, row0, n_users, I,
                                                                              n_itiles, (uint32_t*)ctx->tc_mask, heavy);
  ctx->launches += 3;
  CUtensorMap tmA, tmB;
  rc = make_panel_map(ctx, &tmA, ctx->tc_users, urows, Kp);
  if (rc) return rc;
  rc = make_panel_map(ctx, &tmB, ctx->tc_items, irows, Kp);
  if (rc) return rc;
  int stages = (int)((225 * 1024 - (size_t)UT * KB * kStageBytes) / kStageBytes);
  if (stages > 8) stages = 8;
  if (stages < 2) RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_tc: dim too large for the tensor path");
  const size_t smem = (size_t)(UT * KB + stages) * kStageBytes + (2 * stages + 6) * sizeof(uint64_t) + 16 + 1024;
  static bool attr_set = false;
  if (!attr_set) {
    RBPR_CUDA(ctx, cudaFuncSetAttribute(score_tc, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024));
    RBPR_CUDA(ctx, cudaFuncSetAttribute(select_threshold, cudaFuncAttributeMaxDynamicSharedMemorySize, 96 * 1024));
    attr_set = true;
  }
  if ((size_t)G * sizeof(uint32_t) > 96 * 1024) RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_tc: catalogue too large for the tensor path");
  TcParams p;
  memset(&p, 0, sizeof(p));
  p.n_users = n_users;
  p.n_utiles = n_utiles;
  p.n_itiles = n_itiles;
  p.KB = KB;
  p.stages = stages;
  p.segs = segs;
  p.pair = pair;
  p.mask = (const uint2*)ctx->tc_mask;
  p.gmax = ctx->tc_gmax;
  p.thr = thr;
  p.heavy = heavy;
  p.cnt = cnt;
  p.cand = ctx->tc_cand;
  p.err = ctx->flag;
  p.pass = 0;
  score_tc<<<grid, kTcThreads, smem, st>>>(tmA, tmB, p);
  const int k = tp_in.k_max < I ? tp_in.k_max : I;
  if (G <= 32 * 48)
    select_threshold_warp<48><<<(n_users * 32 + 127) / 128, 128, 0, st>>>(ctx->tc_gmax, G, k, n_users, seen_indptr, row0, heavy,
                                                                          unorm, vmax_bits, thr);
  else
    select_threshold<<<n_users, kSelThreads, (size_t)G * sizeof(uint32_t), st>>>(ctx->tc_gmax, G, k, seen_indptr, row0, heavy,
                                                                                  unorm, vmax_bits, thr);
  p.pass = 1;
  score_tc<<<grid, kTcThreads, smem, st>>>(tmA, tmB, p);
  TopkParams tp = tp_in;
  tp.row0 = row0;
  rescore_rank<<<n_users, kRankThreads, 0, st>>>(ctx->user_emb, ctx->item_emb, ctx->item_bias, users, D, segs, cnt,
                                                 ctx->tc_cand, ovf_rows, ovf_count, tp);
  ctx->launches += 4;
  ctx->topk_launches++;
  ctx->tc_passes++;
  RBPR_CUDA(ctx, cudaGetLastError());
  int32_t h = 0;
  RBPR_CUDA(ctx, cudaMemcpyAsync(&h, ovf_count, sizeof(h), cudaMemcpyDeviceToHost, st));
  RBPR_CUDA(ctx, cudaStreamSynchronize(st));
  *overflow_host = (int)h;
  return 0;
}
