// Device code of the fused BPR training step (sm_100a).  See train.cu for the step structure.
//
// Row geometry: a row of D floats is held by a GROUP of LANES lanes (power of two, D/16 when D is
// a multiple of 16), each lane keeping NV (<=4, or up to 8 for D>512) float4 vectors; column of
// vector v on group lane gl is 4*(gl + LANES*v).  A warp therefore works on 32/LANES triples at
// once (4 for D=128), which amortises the lane-uniform work (Philox, CSR probe, index math) and
// keeps 32/LANES * 2 * NV 128-bit row loads in flight per warp.
#pragma once
#include "common.cuh"
#include "philox.cuh"

namespace rbpr_dev {

struct TrainParams {
  float* __restrict__ user_emb;
  const float* __restrict__ item_emb;
  const float* __restrict__ item_bias;
  float* __restrict__ user_m;
  float* __restrict__ user_v;
  int32_t* __restrict__ user_last;
  float* __restrict__ item_grad;  // (I,D) dense accumulator
  float* __restrict__ user_grad;  // (U,D) dense accumulator, used only by users that occur more
                                  // than once in a step
  float* __restrict__ bias_grad;  // (I) or null
  uint32_t* __restrict__ touched;
  const int64_t* __restrict__ indptr;
  const int32_t* __restrict__ indices;
  const int32_t* __restrict__ coo_user;
  const uint32_t* __restrict__ bloom;  // (U, 8) 256-bit membership filter of every CSR row, or null
  const int64_t* __restrict__ triple_idx;  // the wave's triple ids, input order
  const uint32_t* __restrict__ cnt;        // (steps in wave, U) occurrences of each user per step
  uint32_t* __restrict__ icnt;             // (steps in wave, I) item occurrences per step, or null (small-batch path only)
  int32_t* __restrict__ mh_list;           // (wave triples) per step: users flagged kRecMultiHead, compacted; or null
  uint32_t* __restrict__ mh_count;         // (steps in wave) entries of each step's list
  const uint32_t* __restrict__ ord;        // arrival rank of each slot among its user's slots
  int64_t batch;                           // triples per step
  int64_t U;
  int64_t nnz;
  const int64_t* __restrict__ neg_in;
  int64_t* __restrict__ neg_out;
  const float* __restrict__ alias_prob;
  const int32_t* __restrict__ alias_idx;
  float4* __restrict__ partials;  // one float4 of step statistics per warp
  float2* __restrict__ logit_out;     // (logits_pos, logits_neg) per original position, or null
  const int32_t* __restrict__ step_pos;  // original position of each sorted slot of this step
  int32_t* __restrict__ flag;
  int n;      // triples in this step
  int D;
  uint32_t I;
  uint32_t draw_n;       // I-1 (uniform) or I (alias)
  uint32_t draw_thresh;  // 2^32 mod draw_n (Lemire rejection threshold)
  uint32_t seed_lo, seed_hi;
  uint64_t step;  // global 0-based step -> sampler counter; Adam applies step+1
  int sampler;
  float lr, beta1, beta2, eps;
  float reg_user, reg_item, reg_neg;
  const float2* __restrict__ adam_tab;  // per-step Adam scalars (see adam_catchup4), or null for SGD
  // fused peer-memory exchange (exchange.cu): before touching item rows, wait until every rank has
  // published the rows of the previous step (flag words in LOCAL memory, written by the peers)
  const uint32_t* __restrict__ xwait_flags;  // (world) or null
  int xwait_n;
  uint32_t xwait_epoch;
  int pdl;  // host side only: launch with programmatic stream serialization (see launch_phase_a)
};

// Every CTA waits (thread 0 polls, bounded) until all `n` flag words have reached `epoch`.
__device__ __forceinline__ void wait_peer_flags(const uint32_t* flags, int n, uint32_t epoch, int32_t* err) {
  if (flags == nullptr) return;
  if (threadIdx.x == 0) {
    for (int q = 0; q < n; ++q) {
      bool ok = false;
      for (uint32_t spin = 0; spin < (1u << 22); ++spin) {
        uint32_t seen;
        asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(flags + q) : "memory");
        if ((int32_t)(seen - epoch) >= 0) {
          ok = true;
          break;
        }
        __nanosleep(32);
      }
      if (!ok) atomicExch(err, 10);
    }
  }
  __syncthreads();
}

struct ApplyParams {
  float* __restrict__ item_emb;
  float* __restrict__ item_bias;
  float* __restrict__ item_m;
  float* __restrict__ item_v;
  float* __restrict__ bias_m;
  float* __restrict__ bias_v;
  float* __restrict__ item_grad;
  float* __restrict__ bias_grad;
  uint32_t* __restrict__ touched;
  int64_t I;
  int D;
  int dense;  // 1: ignore touched flags (multi-GPU / Adam)
  uint64_t step;
  float lr, beta1, beta2, eps;
  const float2* __restrict__ adam_tab;
  // user part: users occurring more than once in the step (records flagged kRecMultiHead)
  int do_items, do_users;
  const int4* __restrict__ records;
  int n;
  const int32_t* __restrict__ mh_list;   // the step's multi-occurrence users (from the sampler), or null: scan records
  const uint32_t* __restrict__ mh_count;
  float* __restrict__ user_emb;
  float* __restrict__ user_grad;
  float* __restrict__ user_m;
  float* __restrict__ user_v;
  int32_t* __restrict__ user_last;
  // step statistics: block 0 sums phase A's per-warp partials in double (no extra launch)
  const float4* __restrict__ partials;
  int n_partials;
  double* __restrict__ stats_out;  // (RBPR_STATS_PER_STEP) or null
};

// record.w flags
constexpr int kRecHead = 1;       // first triple of a user run inside its step
constexpr int kRecSingle = 2;     // the user occurs exactly once in the step
constexpr int kRecMultiHead = 4;  // head of a run of length >= 2
constexpr int kRecApplyPos = 8;   // small-batch path: this slot applies the item row of its positive this step
constexpr int kRecApplyNeg = 16;  // ... of its negative (exactly one slot per touched item and step)

template <int LANES>
struct Group {
  int gl;         // lane within group
  unsigned mask;  // lanes of this group
  int shift;      // first lane of the group within the warp
  __device__ __forceinline__ Group() {
    const int lane = threadIdx.x & 31;
    gl = lane & (LANES - 1);
    shift = lane - gl;
    mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << shift);
  }
  __device__ __forceinline__ float sum(float v) const {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
  }
};

// Is item j present in the ascending row idx[lo,hi)?  (LANES+1)-ary cooperative search;
// 32-bit offsets (nnz < 2^32 is enforced at bind time).
template <int LANES>
__device__ __forceinline__ bool row_contains(const int32_t* __restrict__ idx, uint32_t lo,
                                             uint32_t hi, int32_t j, const Group<LANES>& g) {
  while (true) {
    if (hi <= lo) return false;
    const uint32_t n = hi - lo;
    if (n <= (uint32_t)LANES) {
      const int32_t v = ((uint32_t)g.gl < n) ? __ldg(idx + lo + g.gl) : -1;
      return __ballot_sync(g.mask, v == j) != 0u;
    }
    // pivots p_l = lo + (l+1)*n/(LANES+1); n < 2^32/(LANES+1) is guaranteed for rows (< num_items)
    const uint32_t p = lo + ((uint32_t)(g.gl + 1) * n) / (uint32_t)(LANES + 1);
    const int32_t v = __ldg(idx + p);
    if (__ballot_sync(g.mask, v == j) != 0u) return true;
    const unsigned lt = __ballot_sync(g.mask, v < j) >> g.shift;
    const uint32_t k = (uint32_t)__popc(lt);
    const uint32_t plo = lo + (k * n) / (uint32_t)(LANES + 1);
    const uint32_t phi = lo + ((k + 1u) * n) / (uint32_t)(LANES + 1);
    const uint32_t nlo = (k == 0u) ? lo : plo + 1u;
    const uint32_t nhi = (k == (uint32_t)LANES) ? hi : phi;
    lo = nlo;
    hi = nhi;
  }
}

// 256-bit Bloom filter per user (two hash bits per seen item, built at bind time, L2-resident:
// 32 B per user).  No false negatives, so accepting a candidate whose bits are not both set is
// exact; only "maybe seen" candidates pay for the probe of the CSR row in DRAM.
__device__ __forceinline__ uint32_t bloom_h1(int32_t j) { return ((uint32_t)j * 0x9E3779B1u) >> 24; }
__device__ __forceinline__ uint32_t bloom_h2(int32_t j) { return ((uint32_t)j * 0x85EBCA77u) >> 24; }

// Counter-based negative draw.  All lanes of the group evaluate the same Philox
// block, so control flow is group-uniform.  Returns -1 after 256 blocks of failed attempts.
template <int LANES>
__device__ __forceinline__ int32_t draw_negative(const TrainParams& p, uint32_t t, int32_t uu,
                                                 uint32_t lo, uint32_t hi, const Group<LANES>& g) {
  const uint32_t n = p.draw_n, thresh = p.draw_thresh;
  // every lane of the group reads the two filter words of a candidate itself (same address:
  // one broadcast transaction, L2-resident table)
  const uint32_t* bl = (p.bloom != nullptr) ? p.bloom + (size_t)uu * 8u : nullptr;
  auto maybe_seen = [&](int32_t j) -> bool {
    if (bl == nullptr) return true;
    const uint32_t h1 = bloom_h1(j), h2 = bloom_h2(j);
    return ((__ldg(bl + (h1 >> 5)) >> (h1 & 31u)) & (__ldg(bl + (h2 >> 5)) >> (h2 & 31u)) & 1u) != 0u;
  };
  const uint32_t step_lo = (uint32_t)(p.step << 8), step_hi = (uint32_t)(p.step >> 24);
  for (uint32_t blk = 0; blk < 256u; ++blk) {
    const philox4 r = philox4x32_10(step_lo | blk, step_hi, t, 0u, p.seed_lo, p.seed_hi);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    if (p.sampler == RBPR_SAMPLER_UNIFORM) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        const uint32_t mlo = w[a] * n, mhi = __umulhi(w[a], n);
        if (mlo < thresh) continue;  // Lemire rejection: exactly uniform
        const int32_t j = 1 + (int32_t)mhi;
        if (!maybe_seen(j) || !row_contains<LANES>(p.indices, lo, hi, j, g)) return j;
      }
    } else {  // Walker alias over [0,I), two words per attempt
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        const uint32_t mlo = w[2 * a] * n, mhi = __umulhi(w[2 * a], n);
        if (mlo < thresh) continue;
        const int32_t col = (int32_t)mhi;
        const float uf = (float)(w[2 * a + 1] >> 8) * (1.0f / 16777216.0f);
        const int32_t j = (uf < __ldg(p.alias_prob + col)) ? col : __ldg(p.alias_idx + col);
        if (j == 0) continue;
        if (!maybe_seen(j) || !row_contains<LANES>(p.indices, lo, hi, j, g)) return j;
      }
    }
  }
  return -1;
}

__device__ __forceinline__ float4 ld4(const float* p) { return *reinterpret_cast<const float4*>(p); }
__device__ __forceinline__ void st4(float* p, float4 v) { *reinterpret_cast<float4*>(p) = v; }
__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y),
               "f"(v.z), "f"(v.w)
               : "memory");
}
__device__ __forceinline__ float dot4(float4 a, float4 b) {
  return a.x * b.x + a.y * b.y + a.z * b.z + a.w * b.w;
}
__device__ __forceinline__ float4 f4zero() { return make_float4(0.f, 0.f, 0.f, 0.f); }

// torch.optim.Adam (no amsgrad, no weight decay) arithmetic for one element:
// m.lerp_(g, 1-b1); v = b2 v + (1-b2) g²; p -= (lr/bc1) * m / (sqrt(v)/sqrt(bc2) + eps)
__device__ __forceinline__ void adam1(float& p, float& m, float& v, float g, float b1, float b2,
                                      float eps, float step_size, float bc2_sqrt) {
  m = m + (g - m) * (1.0f - b1);
  v = v * b2 + (1.0f - b2) * g * g;
  const float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p - step_size * (m / denom);
}
// ---- optimizer arithmetic, one element ------------------------------------------------------------
// OPT selects torch.optim.{SGD, Adam, SGD(momentum[, nesterov]), RMSprop(momentum=0)}.  s1 / s2 are the
// optimizer state of the element (Adam: exp_avg / exp_avg_sq; SGDM: momentum_buffer; RMSprop:
// square_avg).  Hyper-parameter slots (rbpr_hparams): SGDM beta1 = momentum, beta2 != 0 = nesterov;
// RMSprop beta2 = alpha.  A zero-gradient step is the same function with g = 0 (dense semantics).
struct OptScalars {
  float lr, b1, b2, eps, step_size, bc2_sqrt;
};
template <int OPT>
__device__ __forceinline__ void opt1(float& p, float& s1, float& s2, float g, const OptScalars& h) {
  if (OPT == RBPR_OPT_SGD) {
    p -= h.lr * g;
  } else if (OPT == RBPR_OPT_ADAM) {
    adam1(p, s1, s2, g, h.b1, h.b2, h.eps, h.step_size, h.bc2_sqrt);
  } else if (OPT == RBPR_OPT_SGDM) {  // buf = mu*buf + g ; d = nesterov ? g + mu*buf : buf ; p -= lr*d
    s1 = s1 * h.b1 + g;
    p -= h.lr * ((h.b2 != 0.f) ? g + h.b1 * s1 : s1);
  } else {  // RMSprop: sq = alpha*sq + (1-alpha) g^2 ; p -= lr * g / (sqrt(sq) + eps)
    s1 = s1 * h.b2 + (1.0f - h.b2) * g * g;
    p -= h.lr * (g / (sqrtf(s1) + h.eps));
  }
}
template <int OPT>
__device__ __forceinline__ void opt4(float4& p, float4& s1, float4& s2, float4 g, const OptScalars& h) {
  opt1<OPT>(p.x, s1.x, s2.x, g.x, h);
  opt1<OPT>(p.y, s1.y, s2.y, g.y, h);
  opt1<OPT>(p.z, s1.z, s2.z, g.z, h);
  opt1<OPT>(p.w, s1.w, s2.w, g.w, h);
}
// Replay optimizer steps from+1..to with zero gradient (a row that received no gradient in those
// steps still moves under the dense torch optimizers).
//
// SGD-momentum / RMSprop: the steps are replayed one by one (a multiply-add per element and step).
//
// Adam: j zero-gradient steps from (m, v) give m_j = b1^j m, v_j = b2^j v and
//   dp = sum_j ss_t b1^j m / (b2^(j/2) sqrt(v) / bc2_t + eps),  t = from + j
//      = m * sum_j c_j / (sqrt(v) + e_j),   c_j = ss_t bc2_t (b1/sqrt(b2))^j,  e_j = eps bc2_t b2^(-j/2).
// Replaying that per element costs a sqrt and a division per element AND step: at 65 536 triples per
// step on the MSD shape a touched user row is ~10 steps behind and the replay made phase A
// compute-bound (0.24 ms per step).  Instead the steps are grouped into SEGMENTS over which e_j grows
// by at most 5 % (one segment in steady state; a few while the bias correction is still moving,
// t < ~100), and a segment is applied in closed form with row-level scalars that cost three
// multiply-adds per step: R = sum c_j, e~ = sum c_j e_j / R (the c-weighted mean of e_j, which
// cancels the first-order error), then per element ONE sqrt and ONE division:
//   p -= m R / (sqrt(v) + e~);  m *= b1^n;  v *= b2^n.
// Deviation from the step-by-step replay: second order in the 5 % spread and only where
// sqrt(v) ~ eps = 1e-8 (checked on 2e4 random states: < 3e-6 absolute at lr 1e-3; rounding elsewhere).
template <int OPT>
struct Catchup {
  int64_t from, to;
  const float2* tab;
  OptScalars h;
  __device__ __forceinline__ Catchup(int64_t from_, int64_t to_, const float2* __restrict__ tab_, const OptScalars& h_)
      : from(from_), to(to_), tab(tab_), h(h_) {}
  static __device__ __forceinline__ void seg1(float& p, float& s1, float& s2, float R, float ebar, float pb1, float pb2) {
    p -= R * (s1 / (sqrtf(s2) + ebar));
    s1 *= pb1;
    s2 *= pb2;
  }
  __device__ __forceinline__ void apply(float4& p, float4& s1, float4& s2) const {
    if (OPT == RBPR_OPT_ADAM) {
      const float ihb = rsqrtf(h.b2), rho = h.b1 * ihb;
      float pw = 1.f, hb = 1.f, R = 0.f, E = 0.f, pb1 = 1.f, pb2 = 1.f, e0 = 0.f;
      for (int64_t s = from + 1; s <= to; ++s) {
        const float2 t = __ldg(tab + s);  // {lr_s / (1 - b1^s), sqrt(1 - b2^s)}
        float e = h.eps * t.y * (hb * ihb);
        if (R > 0.f && e > 1.05f * e0) {  // close the segment before this step
          const float ebar = E / R;
          seg1(p.x, s1.x, s2.x, R, ebar, pb1, pb2);
          seg1(p.y, s1.y, s2.y, R, ebar, pb1, pb2);
          seg1(p.z, s1.z, s2.z, R, ebar, pb1, pb2);
          seg1(p.w, s1.w, s2.w, R, ebar, pb1, pb2);
          pw = hb = pb1 = pb2 = 1.f;
          R = E = 0.f;
          e = h.eps * t.y * ihb;
        }
        if (R == 0.f) e0 = e;
        pw *= rho;
        hb *= ihb;
        pb1 *= h.b1;
        pb2 *= h.b2;
        const float c = t.x * t.y * pw;
        R += c;
        E += c * e;
      }
      if (R > 0.f) {
        const float ebar = E / R;
        seg1(p.x, s1.x, s2.x, R, ebar, pb1, pb2);
        seg1(p.y, s1.y, s2.y, R, ebar, pb1, pb2);
        seg1(p.z, s1.z, s2.z, R, ebar, pb1, pb2);
        seg1(p.w, s1.w, s2.w, R, ebar, pb1, pb2);
      }
    } else {
      for (int64_t s = from + 1; s <= to; ++s) opt4<OPT>(p, s1, s2, f4zero(), h);
    }
  }
};
__host__ __device__ constexpr bool opt_has_s2(int opt) { return opt == RBPR_OPT_ADAM; }

constexpr int kPhaseAThreads = 128;

// P1 — negative sampling for every step of a WAVE in one launch.  A group of kSampleLanes lanes per slot
// resolves (user, item), draws the negative with a cooperative (lanes+1)-ary CSR probe, and emits a
// 16-byte record {u, i+, i-, flags} in input order.  The run flags come from the per-step user
// occurrence counts built by count_users (train.cu): kRecSingle if the user occurs once in the
// step, kRecMultiHead for ONE designated slot (arrival rank 0) of a user that occurs several times.
// The static samplers depend only on (seed, step, triple, CSR), never on the model, so a whole
// wave's dependent-load chains run here, one wave ahead of the row-gather kernel.
#ifndef RBPR_SAMPLE_LANES
#define RBPR_SAMPLE_LANES 2  // measured 1/2/4/8 lanes: 152/147/152/165 us per 262144-triple step (profiles/r01v)
#endif
constexpr int kSampleLanes = RBPR_SAMPLE_LANES;  // lanes per slot (instruction-bound kernel: fewer lanes = more slots per warp)

// (sl = step of the wave the slot belongs to; kernels take it from blockIdx.y — a 64-bit division
// per slot was a third of the sampler's instructions)
__device__ __forceinline__ int32_t slot_flags(const TrainParams& p, uint64_t k, uint64_t sl, int32_t uu) {
  const uint32_t total = __ldg(p.cnt + sl * (uint64_t)p.U + (uint64_t)uu);
  if (total <= 1u) return kRecHead | kRecSingle;
  return (__ldg(p.ord + k) == 0u) ? (kRecHead | kRecMultiHead) : 0;
}

static __global__ void __launch_bounds__(256) bpr_sample(TrainParams p, int4* __restrict__ records,
                                                         uint64_t n_slots, uint64_t step0) {
  // grid: x over the slots of one step, y = step of the wave
  const Group<kSampleLanes> g;
  const uint32_t local = (blockIdx.x * blockDim.x + threadIdx.x) / kSampleLanes;
  const uint64_t sl = blockIdx.y;
  if (local >= (uint32_t)p.batch) return;  // whole group leaves together
  const uint64_t k = sl * (uint64_t)p.batch + local;
  if (k >= n_slots) return;
  int64_t t64 = __ldg(p.triple_idx + k);
  if (t64 < 0 || t64 >= p.nnz) {
    if (g.gl == 0) atomicExch(p.flag, 2);
    t64 = 0;
  }
  const uint32_t t = (uint32_t)t64;
  const int32_t uu = __ldg(p.coo_user + t);
  const int32_t i = __ldg(p.indices + t);
  const int32_t flags = (records != nullptr) ? slot_flags(p, k, sl, uu) : 0;  // sampler-only calls keep no counts
  int32_t j;
  if (p.sampler == RBPR_SAMPLER_INJECTED) {
    j = (int32_t)__ldg(p.neg_in + k);
  } else {
    p.step = step0 + sl;
    const uint32_t lo = (uint32_t)__ldg(p.indptr + uu), hi = (uint32_t)__ldg(p.indptr + uu + 1);
    j = draw_negative<kSampleLanes>(p, t, uu, lo, hi, g);
    if (j < 0) {
      if (g.gl == 0) atomicExch(p.flag, 1);
      j = 1;
    }
  }
  if (g.gl == 0) {
    int32_t fl = flags;
    if (p.icnt != nullptr) {  // designate ONE slot per touched item and step (first to arrive here)
      uint32_t* row = p.icnt + sl * (uint64_t)p.I;
      if (i > 0 && (uint32_t)i < p.I && atomicAdd(row + i, 1u) == 0u) fl |= kRecApplyPos;
      if (j > 0 && (uint32_t)j < p.I && atomicAdd(row + j, 1u) == 0u) fl |= kRecApplyNeg;
    }
    if (records != nullptr) records[k] = make_int4(uu, i, j, fl);
    if (p.mh_list != nullptr && (fl & kRecMultiHead) != 0 && uu != 0)  // compact list for bpr_apply's user half
      p.mh_list[sl * (uint64_t)p.batch + atomicAdd(p.mh_count + sl, 1u)] = uu;
    if (p.neg_out != nullptr) p.neg_out[k] = (int64_t)j;
  }
}

// P2 — row gather / loss / gradients, one lane group per triple, every triple independent.
// Measured on B200 (profiles/r01i_gather_floor.txt): for 512-byte rows the
// memory system serves this pattern fastest from plain 128-bit row loads at high occupancy
// (one LDG.128 per lane fetches a whole D=128 row per warp instruction); the TMA/mbarrier-staged
// variants of this kernel (git history, profiles/r01[cfgh]_*) were 2x slower because of their
// per-triple bookkeeping.  So the kernel is deliberately minimal:
//   * the three rows are read with 128-bit loads, the dot is reduced with shuffles;
//   * the two item-row gradients go to the dense accumulator with red.global.add.v4.f32
//     (item rows are only READ here, so every triple sees pre-step values: exact minibatch);
//   * a user that occurs once in the step (kRecSingle, the common case) is updated in place
//     by its only triple; a user with several triples accumulates into the dense user-gradient
//     buffer and is updated by bpr_apply (its row is not written here, so its other triples
//     still read the pre-step value).
template <int LANES, int NV, int OPT>
__global__ void __launch_bounds__(kPhaseAThreads) bpr_phase_a(const TrainParams p,
                                                              const int4* __restrict__ records) {
  const Group<LANES> g;
  const int D = p.D;
  const uint32_t n = (uint32_t)p.n;
  const uint32_t ngroups = (gridDim.x * kPhaseAThreads) / LANES;
  float loss_acc = 0.f, absx_acc = 0.f, l2_acc = 0.f, cnt_acc = 0.f;
  wait_peer_flags(p.xwait_flags, p.xwait_n, p.xwait_epoch, p.flag);  // data parallel: last step's rows have landed

  bool colok[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) colok[v] = 4 * (g.gl + LANES * v) < D;

  OptScalars h = {p.lr, p.beta1, p.beta2, p.eps, 0.f, 1.f};
  if (OPT == RBPR_OPT_ADAM) {
    const float2 t = __ldg(p.adam_tab + (p.step + 1));  // 1-based optimizer step being applied
    h.step_size = t.x;
    h.bc2_sqrt = t.y;
  }

  // the next record of the group is fetched one iteration ahead: its latency hides behind this
  // triple's row loads instead of heading the next iteration's dependent chain
  uint32_t k = (blockIdx.x * kPhaseAThreads + threadIdx.x) / LANES;
  int4 rec_next = (k < n) ? __ldg(records + k) : make_int4(0, 0, 0, 0);
  for (; k < n; k += ngroups) {
    const int4 rec = rec_next;
    if (k + ngroups < n) rec_next = __ldg(records + k + ngroups);
    const int32_t uu = rec.x, i = rec.y, j = rec.z;
    const bool single = (rec.w & kRecSingle) != 0;
    const float* urow = p.user_emb + (size_t)uu * D;
    const float* irow = p.item_emb + (size_t)i * D;
    const float* jrow = p.item_emb + (size_t)j * D;
    float4 u[NV], vi[NV], vj[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (g.gl + LANES * v);
      u[v] = colok[v] ? ld4(urow + c) : f4zero();
      vi[v] = colok[v] ? ld4(irow + c) : f4zero();
      vj[v] = colok[v] ? ld4(jrow + c) : f4zero();
    }
    float4 m[NV], vv[NV];  // optimizer state of the user row (stateful optimizers only)
    if (OPT != RBPR_OPT_SGD) {
      // Lazy dense-optimizer catch-up of the user row to the optimizer steps already taken globally:
      // the caught-up row is what every triple of the user must see.  (bpr_apply redoes it for
      // users with several triples, whose row and moments are not written here.)
      const int64_t last = p.user_last[uu];
      const int64_t upto = (int64_t)p.step;
      const bool behind = last > 0 && last < upto;
      if (single || behind) {
        const float* mrow = p.user_m + (size_t)uu * D;
        const float* vrow = p.user_v + (size_t)uu * D;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = 4 * (g.gl + LANES * v);
          m[v] = colok[v] ? ld4(mrow + c) : f4zero();
          vv[v] = (opt_has_s2(OPT) && colok[v]) ? ld4(vrow + c) : f4zero();
        }
        if (behind) {
          const Catchup<OPT> cu(last, upto, p.adam_tab, h);
#pragma unroll
          for (int v = 0; v < NV; ++v) cu.apply(u[v], m[v], vv[v]);
        }
      }
    }
    float pp = 0.f, pn = 0.f, sq = 0.f, usq = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      pp += dot4(u[v], vi[v]);
      pn += dot4(u[v], vj[v]);
      sq += p.reg_item * dot4(vi[v], vi[v]) + p.reg_neg * dot4(vj[v], vj[v]);
      usq += dot4(u[v], u[v]);
    }
    float x;
    if (p.logit_out != nullptr) {  // drop-in Model.forward output: logits_pos / logits_neg
      float xp = g.sum(pp), xn = g.sum(pn);
      if (p.item_bias != nullptr) {
        xp += __ldg(p.item_bias + i);
        xn += __ldg(p.item_bias + j);
      }
      x = xp - xn;
      if (g.gl == 0)
        p.logit_out[p.step_pos != nullptr ? __ldg(p.step_pos + k) : (int32_t)k] = make_float2(xp, xn);
    } else {
      x = g.sum(pp - pn);
      if (p.item_bias != nullptr) x += __ldg(p.item_bias + i) - __ldg(p.item_bias + j);
    }
    // softplus(-x) and c = sigmoid(-x), overflow-safe; fast intrinsics: abs err ~1e-7 on a
    // per-triple loss of O(1), inside the 1e-4 parity budget
    const float e = __expf(-fabsf(x));
    const float sp = fmaxf(-x, 0.f) + __logf(1.0f + e);
    const float inv = __fdividef(1.0f, 1.0f + e);
    const float c = (x >= 0.f) ? e * inv : inv;
    l2_acc += 0.5f * (sq + p.reg_user * usq);
    if (g.gl == 0) {
      loss_acc += sp;
      absx_acc += fabsf(x);
      cnt_acc += 1.f;
      p.touched[i] = 1u;
      p.touched[j] = 1u;
      if (p.bias_grad != nullptr) {
        atomicAdd(p.bias_grad + i, -c);
        atomicAdd(p.bias_grad + j, c);
      }
    }
    // row 0 of either table is the padding row: nn.Embedding(padding_idx=0) blocks its gradient
    float* gi = (i != 0) ? p.item_grad + (size_t)i * D : nullptr;
    float* gj = (j != 0) ? p.item_grad + (size_t)j * D : nullptr;
    float* gurow = p.user_grad + (size_t)uu * D;
    float* uout = p.user_emb + (size_t)uu * D;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!colok[v]) continue;
      const int cidx = 4 * (g.gl + LANES * v);
      const float4 cu = make_float4(c * u[v].x, c * u[v].y, c * u[v].z, c * u[v].w);
      float4 a, b, gu;
      a.x = p.reg_item * vi[v].x - cu.x;
      a.y = p.reg_item * vi[v].y - cu.y;
      a.z = p.reg_item * vi[v].z - cu.z;
      a.w = p.reg_item * vi[v].w - cu.w;
      b.x = p.reg_neg * vj[v].x + cu.x;
      b.y = p.reg_neg * vj[v].y + cu.y;
      b.z = p.reg_neg * vj[v].z + cu.z;
      b.w = p.reg_neg * vj[v].w + cu.w;
      if (gi != nullptr) red4(gi + cidx, a);
      if (gj != nullptr) red4(gj + cidx, b);
      gu.x = p.reg_user * u[v].x - c * (vi[v].x - vj[v].x);
      gu.y = p.reg_user * u[v].y - c * (vi[v].y - vj[v].y);
      gu.z = p.reg_user * u[v].z - c * (vi[v].z - vj[v].z);
      gu.w = p.reg_user * u[v].w - c * (vi[v].w - vj[v].w);
      if (uu == 0) continue;
      if (!single) {
        red4(gurow + cidx, gu);
      } else if (OPT == RBPR_OPT_SGD) {
        float4 o;
        o.x = u[v].x - p.lr * gu.x;
        o.y = u[v].y - p.lr * gu.y;
        o.z = u[v].z - p.lr * gu.z;
        o.w = u[v].w - p.lr * gu.w;
        st4(uout + cidx, o);
      } else {
        float4 pp4 = u[v];
        opt4<OPT>(pp4, m[v], vv[v], gu, h);
        st4(uout + cidx, pp4);
        st4(p.user_m + (size_t)uu * D + cidx, m[v]);
        if (opt_has_s2(OPT)) st4(p.user_v + (size_t)uu * D + cidx, vv[v]);
      }
    }
    if (OPT != RBPR_OPT_SGD && single && uu != 0) {
      __syncwarp(g.mask);  // every lane has read user_last before it moves
      if (g.gl == 0) p.user_last[uu] = (int32_t)(p.step + 1);
    }
  }

  // per-warp statistics partial (no block barrier: warps retire independently)
  float a = loss_acc, b = l2_acc, cabs = absx_acc, d = cnt_acc;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    cabs += __shfl_xor_sync(0xffffffffu, cabs, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
  if ((threadIdx.x & 31) == 0)
    p.partials[blockIdx.x * (kPhaseAThreads / 32) + (threadIdx.x >> 5)] = make_float4(a, b, cabs, d);
}

// ---- pieces of phase B, shared by bpr_apply and the fused exchange kernel (exchange.cu) ------------

// users: rows of users with several triples in the step (records flagged kRecMultiHead): their summed
// gradient sits in user_grad; single-occurrence users were finished by phase A.  One lane group
// finishes one user.
template <int LANES, int NV, int OPT>
__device__ __forceinline__ void apply_users(const ApplyParams& p, const OptScalars& h, int64_t gid, int64_t groups) {
  const Group<LANES> g;
  const int D = p.D;
  auto finish_user = [&](int64_t r) {
    float* grow = p.user_grad + r * D;
    float* prow = p.user_emb + r * D;
    int64_t last = 0;
    if (OPT != RBPR_OPT_SGD) last = p.user_last[r];
    const bool behind = OPT != RBPR_OPT_SGD && last > 0 && last < (int64_t)p.step;
    const Catchup<OPT> cu(last, behind ? (int64_t)p.step : last, p.adam_tab, h);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (g.gl + LANES * v);
      if (c >= D) continue;
      const float4 gr = ld4(grow + c);
      float4 pp = ld4(prow + c);
      if (OPT == RBPR_OPT_SGD) {
        pp.x -= p.lr * gr.x;
        pp.y -= p.lr * gr.y;
        pp.z -= p.lr * gr.z;
        pp.w -= p.lr * gr.w;
      } else {
        float4 m = ld4(p.user_m + r * D + c);
        float4 vv = opt_has_s2(OPT) ? ld4(p.user_v + r * D + c) : f4zero();
        if (behind) cu.apply(pp, m, vv);
        opt4<OPT>(pp, m, vv, gr, h);
        st4(p.user_m + r * D + c, m);
        if (opt_has_s2(OPT)) st4(p.user_v + r * D + c, vv);
      }
      st4(prow + c, pp);
      st4(grow + c, f4zero());
    }
    if (OPT != RBPR_OPT_SGD) {
      __syncwarp(g.mask);
      if (g.gl == 0) p.user_last[r] = (int32_t)(p.step + 1);
    }
  };
  if (p.mh_list != nullptr) {
    // the sampler compacted the step's multi-occurrence users: no scan, one round trip per user
    const int64_t n_mh = (int64_t)__ldg(p.mh_count);
    for (int64_t e = gid; e < n_mh; e += groups) finish_user((int64_t)__ldg(p.mh_list + e));
  } else {
    for (int64_t k = gid; k < p.n; k += groups) {
      const int4 rec = __ldg(p.records + k);
      if ((rec.w & kRecMultiHead) == 0 || rec.x == 0) continue;
      finish_user((int64_t)rec.x);
    }
  }
}

// step statistics: one block sums phase A's per-warp partials in double (no extra launch)
__device__ __forceinline__ void apply_stats(const ApplyParams& p) {
  __shared__ double red[4][8];
  double a = 0.0, b = 0.0, c = 0.0, d = 0.0;
  for (int i = threadIdx.x; i < p.n_partials; i += blockDim.x) {
    const float4 v = p.partials[i];
    a += (double)v.x;
    b += (double)v.y;
    c += (double)v.z;
    d += (double)v.w;
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(0xffffffffu, a, o);
    b += __shfl_xor_sync(0xffffffffu, b, o);
    c += __shfl_xor_sync(0xffffffffu, c, o);
    d += __shfl_xor_sync(0xffffffffu, d, o);
  }
  const int warp = threadIdx.x >> 5;
  if ((threadIdx.x & 31) == 0) {
    red[0][warp] = a;
    red[1][warp] = b;
    red[2][warp] = c;
    red[3][warp] = d;
  }
  __syncthreads();
  if (threadIdx.x < 4) {
    double t = 0.0;
    for (int w = 0; w < (int)(blockDim.x >> 5); ++w) t += red[threadIdx.x][w];
    p.stats_out[threadIdx.x] = t;
  }
}

// Phase B — apply the accumulated gradients and clear the accumulators.
//   items: SGD touches only the rows flagged by phase A; Adam (and the multi-GPU path, where the
//          accumulator holds the all-reduced gradient) sweeps every row, which is torch.optim's
//          dense semantics for the replicated item table.
//   users: apply_users above.
template <int LANES, int NV, int OPT>
__global__ void __launch_bounds__(256) bpr_apply(const ApplyParams p) {
  const Group<LANES> g;
  const int D = p.D;
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / LANES;
  const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  OptScalars h = {p.lr, p.beta1, p.beta2, p.eps, 0.f, 1.f};
  if (OPT == RBPR_OPT_ADAM) {
    const float2 t = __ldg(p.adam_tab + (p.step + 1));
    h.step_size = t.x;
    h.bc2_sqrt = t.y;
  }
  if (p.do_users) apply_users<LANES, NV, OPT>(p, h, gid, groups);
  for (int64_t r = gid; p.do_items && r < p.I; r += groups) {
    if (!p.dense) {
      if (p.touched[r] == 0u) continue;
    }
    float* grow = p.item_grad + r * D;
    float* prow = p.item_emb + r * D;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (g.gl + LANES * v);
      if (c >= D) continue;
      const float4 gr = ld4(grow + c);
      float4 pp = ld4(prow + c);
      if (OPT == RBPR_OPT_SGD) {
        pp.x -= p.lr * gr.x;
        pp.y -= p.lr * gr.y;
        pp.z -= p.lr * gr.z;
        pp.w -= p.lr * gr.w;
      } else {
        float4 m = ld4(p.item_m + r * D + c);
        float4 vv = opt_has_s2(OPT) ? ld4(p.item_v + r * D + c) : f4zero();
        opt4<OPT>(pp, m, vv, gr, h);
        st4(p.item_m + r * D + c, m);
        if (opt_has_s2(OPT)) st4(p.item_v + r * D + c, vv);
      }
      st4(prow + c, pp);
      st4(grow + c, f4zero());
    }
    if (g.gl == 0) {
      p.touched[r] = 0u;
      if (p.bias_grad != nullptr) {
        const float gb = p.bias_grad[r];
        float b = p.item_bias[r];
        if (OPT == RBPR_OPT_SGD) {
          b -= p.lr * gb;
        } else {
          float m = p.bias_m[r], vv = opt_has_s2(OPT) ? p.bias_v[r] : 0.f;
          opt1<OPT>(b, m, vv, gb, h);
          p.bias_m[r] = m;
          if (opt_has_s2(OPT)) p.bias_v[r] = vv;
        }
        p.item_bias[r] = b;
        p.bias_grad[r] = 0.f;
      }
    }
  }
  if (blockIdx.x == 0 && p.stats_out != nullptr) apply_stats(p);
}

// Bring all user rows to `step` applied optimizer steps (dense semantics), grid-stride over rows.
template <int LANES, int NV, int OPT>
__global__ void __launch_bounds__(256) bpr_flush_users(float* __restrict__ user_emb,
                                                       float* __restrict__ user_m,
                                                       float* __restrict__ user_v,
                                                       int32_t* __restrict__ user_last, int64_t U,
                                                       int D, int64_t step,
                                                       const float2* __restrict__ tab, OptScalars h) {
  const Group<LANES> g;
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / LANES;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; r < U; r += groups) {
    const int64_t last = user_last[r];
    if (last <= 0 || last >= step) continue;
    const Catchup<OPT> cu(last, step, tab, h);
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (g.gl + LANES * v);
      if (c >= D) continue;
      float4 pp = ld4(user_emb + r * D + c), m = ld4(user_m + r * D + c);
      float4 vv = opt_has_s2(OPT) ? ld4(user_v + r * D + c) : f4zero();
      cu.apply(pp, m, vv);
      st4(user_emb + r * D + c, pp);
      st4(user_m + r * D + c, m);
      if (opt_has_s2(OPT)) st4(user_v + r * D + c, vv);
    }
    __syncwarp(g.mask);
    if (g.gl == 0) user_last[r] = (int32_t)step;
  }
}

}  // namespace rbpr_dev

namespace rbpr_dev {
// Launch of phase A.  p.pdl (data parallel, fused exchange): the previous kernel of the stream is the
// exchange kernel, whose CTAs all run `griddepcontrol.launch_dependents` first thing, so this
// grid's CTAs are placed as soon as SM resources free up instead of after the exchange grid has
// drained and a launch latency has passed; nothing of the previous step is read before
// wait_peer_flags has seen every rank's B2 flag, this rank's own included (its last CTA sets it after
// the local work, the table slice and the accumulator clear are all fenced).
template <typename Kernel>
inline void launch_phase_a(Kernel kern, int blocks, cudaStream_t st, const TrainParams& p, const int4* records) {
  if (p.pdl) {
    cudaLaunchConfig_t cfg = {};
    cfg.gridDim = dim3((unsigned)blocks);
    cfg.blockDim = dim3(kPhaseAThreads);
    cfg.dynamicSmemBytes = 0;
    cfg.stream = st;
    cudaLaunchAttribute at[1];
    at[0].id = cudaLaunchAttributeProgrammaticStreamSerialization;
    at[0].val.programmaticStreamSerializationAllowed = 1;
    cfg.attrs = at;
    cfg.numAttrs = 1;
    cudaLaunchKernelEx(&cfg, kern, p, records);
  } else {
    kern<<<blocks, kPhaseAThreads, 0, st>>>(p, records);
  }
}
}  // namespace rbpr_dev

// Per-optimizer launchers (train_sgd.cu / train_adam.cu), one translation unit each so the
// template instantiations compile in parallel.
int rbpr_launch_phase_a_sgd(rbpr_ctx* ctx, const rbpr_dev::TrainParams& p, int lanes, int nv,
                            const int4* records, int blocks, cudaStream_t st);
int rbpr_launch_phase_a_adam(rbpr_ctx* ctx, const rbpr_dev::TrainParams& p, int lanes, int nv,
                             const int4* records, int blocks, cudaStream_t st);
int rbpr_phase_a_prepare_sgd(rbpr_ctx* ctx, int dim, int lanes, int nv, int* blocks_per_sm);
int rbpr_phase_a_prepare_adam(rbpr_ctx* ctx, int dim, int lanes, int nv, int* blocks_per_sm);
int rbpr_launch_apply_sgd(rbpr_ctx* ctx, const rbpr_dev::ApplyParams& p, int lanes, int nv,
                          cudaStream_t st);
int rbpr_launch_apply_adam(rbpr_ctx* ctx, const rbpr_dev::ApplyParams& p, int lanes, int nv,
                           cudaStream_t st);
int rbpr_launch_flush_users_adam(rbpr_ctx* ctx, int64_t step, const rbpr_hparams* hp, int lanes, int nv,
                                 cudaStream_t st);
#define RBPR_DECLARE_OPT_LAUNCHERS(sfx)                                                                     \
  int rbpr_launch_phase_a_##sfx(rbpr_ctx* ctx, const rbpr_dev::TrainParams& p, int lanes, int nv,           \
                                const int4* records, int blocks, cudaStream_t st);                          \
  int rbpr_phase_a_prepare_##sfx(rbpr_ctx* ctx, int dim, int lanes, int nv, int* blocks_per_sm);            \
  int rbpr_launch_apply_##sfx(rbpr_ctx* ctx, const rbpr_dev::ApplyParams& p, int lanes, int nv,             \
                              cudaStream_t st);                                                             \
  int rbpr_launch_flush_users_##sfx(rbpr_ctx* ctx, int64_t step, const rbpr_hparams* hp, int lanes, int nv, \
                                    cudaStream_t st);
RBPR_DECLARE_OPT_LAUNCHERS(sgdm)
RBPR_DECLARE_OPT_LAUNCHERS(rms)

// Instantiate every (LANES, NV) pair rbpr_geometry can return for D in [4,1024], D % 4 == 0.
#define RBPR_FOR_EACH_GEOMETRY(X) \
  X(1, 1) X(1, 2) X(2, 2) X(4, 2) X(8, 2) X(16, 2) X(32, 2) X(32, 3) X(32, 4) X(32, 5) X(32, 6) X(32, 7) X(32, 8)
