// Fused BPR training step for sm_100a: on-device negative sampling, (u, i+, i-) row gather,
// dot, log-sigmoid loss + L2, exact synchronous-minibatch gradients, in-place update.
//
// One step = two phases on one stream:
//   phase A  bpr_phase_a   — the batch is sorted by triple id (== grouped by user, because
//            triples are the COO flattening of the CSR).  A lane group owns every user run
//            that STARTS inside its chunk: the user row is staged once in registers, each
//            triple of the run samples its negative, gathers the two item rows with 128-bit
//            loads, reduces the dot with shuffles, and pushes the two item-row gradients
//            into the dense item-gradient accumulator with vector red.global.add.v4.f32.
//            The user row is updated in place when the run ends (no atomics: one owner).
//            Item rows are only READ in this phase, so every triple sees pre-step values.
//   phase B  bpr_apply_items — applies the accumulated item gradient (SGD: touched rows
//            only; Adam: every row, which is torch.optim.Adam's dense semantics) and
//            clears the accumulator.
// Reference call sites replaced: see include/rbpr.h (rbpr_train_steps).
#include <cub/device/device_radix_sort.cuh>

#include "common.cuh"
#include "philox.cuh"

namespace {

struct TrainParams {
  // tables
  float* __restrict__ user_emb;
  const float* __restrict__ item_emb;
  const float* __restrict__ item_bias;
  float* __restrict__ user_m;
  float* __restrict__ user_v;
  int32_t* __restrict__ user_last;
  // accumulators
  float* __restrict__ item_grad;  // (I,D)
  float* __restrict__ bias_grad;  // (I) or null
  uint32_t* __restrict__ touched;
  // data
  const int64_t* __restrict__ indptr;
  const int32_t* __restrict__ indices;
  const int32_t* __restrict__ coo_user;
  const uint64_t* __restrict__ keys;  // sorted (step<<32|t) for this step
  const int32_t* __restrict__ pos;    // original positions (or null)
  const int64_t* __restrict__ neg_in;
  int64_t* __restrict__ neg_out;
  const float* __restrict__ alias_prob;
  const int32_t* __restrict__ alias_idx;
  double* __restrict__ stats;  // 4 doubles of this step
  int32_t* __restrict__ flag;
  int n;      // triples in this step
  int chunk;  // triples per group
  int D;
  uint32_t I;
  uint32_t seed_lo, seed_hi;
  uint64_t step;  // global step (0-based) -> sampler; Adam uses step+1
  int sampler;
  float lr, beta1, beta2, eps;
  float reg_user, reg_item, reg_neg;
};

template <int LANES>
struct Group {
  int gl;          // lane within group
  unsigned mask;   // participating lanes
  int shift;       // first lane of the group within the warp
  __device__ Group() {
    int lane = threadIdx.x & 31;
    gl = lane % LANES;
    shift = lane - gl;
    mask = (LANES == 32) ? 0xffffffffu : (((1u << LANES) - 1u) << shift);
  }
  __device__ float sum(float v) const {
#pragma unroll
    for (int o = LANES / 2; o > 0; o >>= 1) v += __shfl_xor_sync(mask, v, o);
    return v;
  }
};

// Is item j present in the ascending row idx[lo,hi)?  (LANES+1)-ary cooperative search.
template <int LANES>
__device__ __forceinline__ bool row_contains(const int32_t* __restrict__ idx, int64_t lo,
                                             int64_t hi, int32_t j, const Group<LANES>& g) {
  while (true) {
    int64_t n = hi - lo;
    if (n <= 0) return false;
    if (n <= LANES) {
      int32_t v = (g.gl < n) ? __ldg(idx + lo + g.gl) : -1;
      return __ballot_sync(g.mask, v == j) != 0u;
    }
    int64_t p = lo + ((int64_t)(g.gl + 1) * n) / (LANES + 1);
    int32_t v = __ldg(idx + p);
    if (__ballot_sync(g.mask, v == j) != 0u) return true;
    unsigned lt = __ballot_sync(g.mask, v < j) >> g.shift;
    int k = __popc(lt);
    int64_t plo = lo + ((int64_t)k * n) / (LANES + 1);
    int64_t phi = lo + ((int64_t)(k + 1) * n) / (LANES + 1);
    int64_t nlo = (k == 0) ? lo : plo + 1;
    int64_t nhi = (k == LANES) ? hi : phi;
    lo = nlo;
    hi = nhi;
  }
}

// Counter-based negative draw.  All lanes of the group evaluate the same
// Philox block, so control flow is group-uniform.  Returns -1 after 1024 failed attempts.
template <int LANES>
__device__ __forceinline__ int32_t draw_negative(const TrainParams& p, uint64_t step, uint64_t t,
                                                 int64_t lo, int64_t hi, const Group<LANES>& g) {
  const uint32_t t_lo = (uint32_t)t, t_hi = (uint32_t)(t >> 32);
  if (p.sampler == RBPR_SAMPLER_UNIFORM) {
    const uint32_t n = p.I - 1u;
    const uint32_t thresh = (0u - n) % n;
    for (uint32_t blk = 0; blk < 256u; ++blk) {
      uint64_t off = (step << 8) | blk;
      philox4 r = philox4x32_10((uint32_t)off, (uint32_t)(off >> 32), t_lo, t_hi, p.seed_lo,
                                p.seed_hi);
      uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        uint64_t m = (uint64_t)w[a] * (uint64_t)n;
        if ((uint32_t)m < thresh) continue;  // Lemire rejection: exactly uniform
        int32_t j = 1 + (int32_t)(m >> 32);
        if (!row_contains<LANES>(p.indices, lo, hi, j, g)) return j;
      }
    }
    return -1;
  } else {  // RBPR_SAMPLER_WEIGHTED: Walker alias over [0,I), two words per attempt
    const uint32_t n = p.I;
    const uint32_t thresh = (0u - n) % n;
    for (uint32_t blk = 0; blk < 256u; ++blk) {
      uint64_t off = (step << 8) | blk;
      philox4 r = philox4x32_10((uint32_t)off, (uint32_t)(off >> 32), t_lo, t_hi, p.seed_lo,
                                p.seed_hi);
      uint32_t w[4] = {r.x, r.y, r.z, r.w};
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        uint64_t m = (uint64_t)w[2 * a] * (uint64_t)n;
        if ((uint32_t)m < thresh) continue;
        int32_t col = (int32_t)(m >> 32);
        float uf = (float)(w[2 * a + 1] >> 8) * (1.0f / 16777216.0f);
        int32_t j = (uf < __ldg(p.alias_prob + col)) ? col : __ldg(p.alias_idx + col);
        if (j == 0) continue;
        if (!row_contains<LANES>(p.indices, lo, hi, j, g)) return j;
      }
    }
    return -1;
  }
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

// One Adam update of 4 lanes worth of parameters, torch.optim.Adam (non-amsgrad, no weight
// decay) arithmetic: m.lerp_(g, 1-b1); v = b2 v + (1-b2) g²; p -= (lr/bc1) m / (sqrt(v)/sqrt(bc2)+eps)
__device__ __forceinline__ void adam1(float& p, float& m, float& v, float g, float b1, float b2,
                                      float eps, float step_size, float bc2_sqrt) {
  m = m + (g - m) * (1.0f - b1);
  v = v * b2 + (1.0f - b2) * g * g;
  float denom = sqrtf(v) / bc2_sqrt + eps;
  p = p - step_size * (m / denom);
}
__device__ __forceinline__ void adam4(float4& p, float4& m, float4& v, float4 g, float b1,
                                      float b2, float eps, float step_size, float bc2_sqrt) {
  adam1(p.x, m.x, v.x, g.x, b1, b2, eps, step_size, bc2_sqrt);
  adam1(p.y, m.y, v.y, g.y, b1, b2, eps, step_size, bc2_sqrt);
  adam1(p.z, m.z, v.z, g.z, b1, b2, eps, step_size, bc2_sqrt);
  adam1(p.w, m.w, v.w, g.w, b1, b2, eps, step_size, bc2_sqrt);
}

// Replay Adam steps (from+1 .. to) with zero gradient on registers (dense-Adam semantics for a
// row that received no gradient in those steps).
__device__ __forceinline__ void adam_catchup4(float4& p, float4& m, float4& v, int64_t from,
                                              int64_t to, float lr, float b1, float b2,
                                              float eps) {
  if (to <= from) return;
  double b1p = pow((double)b1, (double)from), b2p = pow((double)b2, (double)from);
  for (int64_t s = from + 1; s <= to; ++s) {
    b1p *= (double)b1;
    b2p *= (double)b2;
    float step_size = (float)((double)lr / (1.0 - b1p));
    float bc2_sqrt = (float)sqrt(1.0 - b2p);
    adam4(p, m, v, make_float4(0.f, 0.f, 0.f, 0.f), b1, b2, eps, step_size, bc2_sqrt);
  }
}

template <int LANES, int NV, int OPT>
__global__ void __launch_bounds__(256) bpr_phase_a(const TrainParams p) {
  const Group<LANES> g;
  const int D = p.D;
  const int groups_per_block = blockDim.x / LANES;
  const int64_t gid = (int64_t)blockIdx.x * groups_per_block + threadIdx.x / LANES;
  const int64_t start = gid * p.chunk;
  const int64_t end = min(start + (int64_t)p.chunk, (int64_t)p.n);

  float loss_acc = 0.f, absx_acc = 0.f, l2_acc = 0.f;  // loss/absx: lane gl==0; l2: per lane
  int cnt = 0;

  bool colok[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) colok[v] = 4 * (g.gl + LANES * v) < D;

  int64_t k = start;
  if (start < p.n && start > 0) {
    int32_t uprev = __ldg(p.coo_user + (uint32_t)__ldg(p.keys + start - 1));
    while (k < end && __ldg(p.coo_user + (uint32_t)__ldg(p.keys + k)) == uprev) ++k;
  }

  int32_t cur_u = -1;
  float4 u[NV], gu[NV];
  int nocc = 0;
  int64_t row_lo = 0, row_hi = 0;

  auto flush_user = [&]() {
    if (cur_u < 0) return;
    float* urow = p.user_emb + (int64_t)cur_u * D;
    const float rn = p.reg_user * (float)nocc;
    if (OPT == RBPR_OPT_SGD) {
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!colok[v]) continue;
        float4 o;
        o.x = u[v].x - p.lr * (gu[v].x + rn * u[v].x);
        o.y = u[v].y - p.lr * (gu[v].y + rn * u[v].y);
        o.z = u[v].z - p.lr * (gu[v].z + rn * u[v].z);
        o.w = u[v].w - p.lr * (gu[v].w + rn * u[v].w);
        st4(urow + 4 * (g.gl + LANES * v), o);
      }
    } else {
      const int64_t s = (int64_t)p.step + 1;  // 1-based optimizer step being applied
      const double b1p = pow((double)p.beta1, (double)s), b2p = pow((double)p.beta2, (double)s);
      const float step_size = (float)((double)p.lr / (1.0 - b1p));
      const float bc2_sqrt = (float)sqrt(1.0 - b2p);
      float* mrow = p.user_m + (int64_t)cur_u * D;
      float* vrow = p.user_v + (int64_t)cur_u * D;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!colok[v]) continue;
        const int c = 4 * (g.gl + LANES * v);
        float4 m = ld4(mrow + c), vv = ld4(vrow + c);
        float4 grad;
        grad.x = gu[v].x + rn * u[v].x;
        grad.y = gu[v].y + rn * u[v].y;
        grad.z = gu[v].z + rn * u[v].z;
        grad.w = gu[v].w + rn * u[v].w;
        float4 pp = u[v];  // already caught up to step s-1 at load time
        adam4(pp, m, vv, grad, p.beta1, p.beta2, p.eps, step_size, bc2_sqrt);
        st4(urow + c, pp);
        st4(mrow + c, m);
        st4(vrow + c, vv);
      }
      if (g.gl == 0) p.user_last[cur_u] = (int32_t)s;
    }
  };

  while (k < p.n) {
    const uint64_t key = __ldg(p.keys + k);
    const uint32_t t = (uint32_t)key;
    const int32_t uu = __ldg(p.coo_user + t);
    if (uu != cur_u) {
      if (k >= end) break;  // that run belongs to the next group
      flush_user();
      cur_u = uu;
      nocc = 0;
      row_lo = __ldg(p.indptr + uu);
      row_hi = __ldg(p.indptr + uu + 1);
      const float* urow = p.user_emb + (int64_t)uu * D;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        u[v] = colok[v] ? ld4(urow + 4 * (g.gl + LANES * v)) : make_float4(0.f, 0.f, 0.f, 0.f);
        gu[v] = make_float4(0.f, 0.f, 0.f, 0.f);
      }
      if (OPT == RBPR_OPT_ADAM) {
        const int64_t last = p.user_last[uu];
        const int64_t upto = (int64_t)p.step;  // steps 1..step already taken globally
        if (last > 0 && last < upto) {
          float* mrow = p.user_m + (int64_t)uu * D;
          float* vrow = p.user_v + (int64_t)uu * D;
#pragma unroll
          for (int v = 0; v < NV; ++v) {
            if (!colok[v]) continue;
            const int c = 4 * (g.gl + LANES * v);
            float4 m = ld4(mrow + c), vv = ld4(vrow + c);
            adam_catchup4(u[v], m, vv, last, upto, p.lr, p.beta1, p.beta2, p.eps);
            st4(mrow + c, m);
            st4(vrow + c, vv);
          }
        }
      }
    }
    const int32_t i = __ldg(p.indices + t);
    int32_t j;
    if (p.sampler == RBPR_SAMPLER_INJECTED) {
      j = (int32_t)__ldg(p.neg_in + __ldg(p.pos + k));
    } else {
      j = draw_negative<LANES>(p, p.step, (uint64_t)t, row_lo, row_hi, g);
      if (j < 0) {
        if (g.gl == 0) atomicExch(p.flag, 1);
        j = 1;
      }
    }
    if (p.neg_out != nullptr && g.gl == 0) p.neg_out[__ldg(p.pos + k)] = (int64_t)j;

    const float* irow = p.item_emb + (int64_t)i * D;
    const float* jrow = p.item_emb + (int64_t)j * D;
    float4 vi[NV], vj[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (g.gl + LANES * v);
      vi[v] = colok[v] ? ld4(irow + c) : make_float4(0.f, 0.f, 0.f, 0.f);
      vj[v] = colok[v] ? ld4(jrow + c) : make_float4(0.f, 0.f, 0.f, 0.f);
    }
    float part = 0.f, sq = 0.f;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      part += dot4(u[v], vi[v]) - dot4(u[v], vj[v]);
      sq += p.reg_item * dot4(vi[v], vi[v]) + p.reg_neg * dot4(vj[v], vj[v]) +
            p.reg_user * dot4(u[v], u[v]);
    }
    float x = g.sum(part);
    if (p.item_bias != nullptr) x += __ldg(p.item_bias + i) - __ldg(p.item_bias + j);
    // softplus(-x) and c = sigmoid(-x), overflow-safe
    const float e = expf(-fabsf(x));
    const float sp = fmaxf(-x, 0.f) + log1pf(e);
    const float c = (x >= 0.f) ? e / (1.f + e) : 1.f / (1.f + e);
    l2_acc += 0.5f * sq;
    if (g.gl == 0) {
      loss_acc += sp;
      absx_acc += fabsf(x);
      ++cnt;
      p.touched[i] = 1u;
      p.touched[j] = 1u;
      if (p.bias_grad != nullptr) {
        atomicAdd(p.bias_grad + i, -c);
        atomicAdd(p.bias_grad + j, c);
      }
    }
    float* gi = p.item_grad + (int64_t)i * D;
    float* gj = p.item_grad + (int64_t)j * D;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!colok[v]) continue;
      const int cidx = 4 * (g.gl + LANES * v);
      float4 a, b;
      a.x = -c * u[v].x + p.reg_item * vi[v].x;
      a.y = -c * u[v].y + p.reg_item * vi[v].y;
      a.z = -c * u[v].z + p.reg_item * vi[v].z;
      a.w = -c * u[v].w + p.reg_item * vi[v].w;
      b.x = c * u[v].x + p.reg_neg * vj[v].x;
      b.y = c * u[v].y + p.reg_neg * vj[v].y;
      b.z = c * u[v].z + p.reg_neg * vj[v].z;
      b.w = c * u[v].w + p.reg_neg * vj[v].w;
      red4(gi + cidx, a);
      red4(gj + cidx, b);
      gu[v].x -= c * (vi[v].x - vj[v].x);
      gu[v].y -= c * (vi[v].y - vj[v].y);
      gu[v].z -= c * (vi[v].z - vj[v].z);
      gu[v].w -= c * (vi[v].w - vj[v].w);
    }
    ++nocc;
    ++k;
  }
  flush_user();

  // block reduction of the step statistics -> one atomic set per CTA
  if (p.stats != nullptr) {
    float a = loss_acc, b = l2_acc, cabs = absx_acc, d = (float)cnt;
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
      a += __shfl_xor_sync(0xffffffffu, a, o);
      b += __shfl_xor_sync(0xffffffffu, b, o);
      cabs += __shfl_xor_sync(0xffffffffu, cabs, o);
      d += __shfl_xor_sync(0xffffffffu, d, o);
    }
    __shared__ float red[4][8];
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    if (lane == 0) {
      red[0][warp] = a;
      red[1][warp] = b;
      red[2][warp] = cabs;
      red[3][warp] = d;
    }
    __syncthreads();
    if (threadIdx.x < 4) {
      double s = 0.0;
      const int nw = blockDim.x >> 5;
      for (int w = 0; w < nw; ++w) s += (double)red[threadIdx.x][w];
      if (s != 0.0) atomicAdd(p.stats + threadIdx.x, s);
    }
  }
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
};

template <int LANES, int NV, int OPT>
__global__ void __launch_bounds__(256) bpr_apply_items(const ApplyParams p) {
  const Group<LANES> g;
  const int D = p.D;
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / LANES;
  int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  float step_size = 0.f, bc2_sqrt = 1.f;
  if (OPT == RBPR_OPT_ADAM) {
    const double s = (double)(p.step + 1);
    step_size = (float)((double)p.lr / (1.0 - pow((double)p.beta1, s)));
    bc2_sqrt = (float)sqrt(1.0 - pow((double)p.beta2, s));
  }
  for (; r < p.I; r += groups) {
    if (!p.dense) {
      if (p.touched[r] == 0u) continue;
    }
    float* grow = p.item_grad + r * D;
    float* prow = p.item_emb + r * D;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (g.gl + LANES * v);
      if (c >= D) continue;
      float4 gr = ld4(grow + c);
      float4 pp = ld4(prow + c);
      if (OPT == RBPR_OPT_SGD) {
        pp.x -= p.lr * gr.x;
        pp.y -= p.lr * gr.y;
        pp.z -= p.lr * gr.z;
        pp.w -= p.lr * gr.w;
      } else {
        float4 m = ld4(p.item_m + r * D + c), vv = ld4(p.item_v + r * D + c);
        adam4(pp, m, vv, gr, p.beta1, p.beta2, p.eps, step_size, bc2_sqrt);
        st4(p.item_m + r * D + c, m);
        st4(p.item_v + r * D + c, vv);
      }
      st4(prow + c, pp);
      st4(grow + c, make_float4(0.f, 0.f, 0.f, 0.f));
    }
    if (g.gl == 0) {
      p.touched[r] = 0u;
      if (p.bias_grad != nullptr) {
        float gb = p.bias_grad[r];
        float b = p.item_bias[r];
        if (OPT == RBPR_OPT_SGD) {
          b -= p.lr * gb;
        } else {
          float m = p.bias_m[r], vv = p.bias_v[r];
          adam1(b, m, vv, gb, p.beta1, p.beta2, p.eps, step_size, bc2_sqrt);
          p.bias_m[r] = m;
          p.bias_v[r] = vv;
        }
        p.item_bias[r] = b;
        p.bias_grad[r] = 0.f;
      }
    }
  }
}

// Bring all user rows to `step` applied Adam steps (dense semantics), grid-stride over rows.
template <int LANES, int NV>
__global__ void __launch_bounds__(256) bpr_flush_users(float* __restrict__ user_emb,
                                                       float* __restrict__ user_m,
                                                       float* __restrict__ user_v,
                                                       int32_t* __restrict__ user_last, int64_t U,
                                                       int D, int64_t step, float lr, float b1,
                                                       float b2, float eps) {
  const Group<LANES> g;
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / LANES;
  for (int64_t r = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES; r < U; r += groups) {
    const int64_t last = user_last[r];
    if (last <= 0 || last >= step) continue;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (g.gl + LANES * v);
      if (c >= D) continue;
      float4 pp = ld4(user_emb + r * D + c), m = ld4(user_m + r * D + c),
             vv = ld4(user_v + r * D + c);
      adam_catchup4(pp, m, vv, last, step, lr, b1, b2, eps);
      st4(user_emb + r * D + c, pp);
      st4(user_m + r * D + c, m);
      st4(user_v + r * D + c, vv);
    }
    __syncwarp(g.mask);
    if (g.gl == 0) user_last[r] = (int32_t)step;
  }
}

__global__ void make_keys(const int64_t* __restrict__ triple_idx, int64_t n, int64_t batch,
                          int64_t nnz, uint64_t* __restrict__ keys, int32_t* __restrict__ pos,
                          int32_t* __restrict__ flag) {
  int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t t = triple_idx[k];
  if (t < 0 || t >= nnz) {
    atomicExch(flag, 2);
    t = 0;
  }
  keys[k] = ((uint64_t)(k / batch) << 32) | (uint64_t)(uint32_t)t;
  if (pos != nullptr) pos[k] = (int32_t)k;
}

__global__ void expand_rows(const int64_t* __restrict__ indptr, int64_t U, int64_t nnz,
                            uint32_t I, int32_t* __restrict__ coo_user,
                            const int32_t* __restrict__ indices, int32_t* __restrict__ flag) {
  // one warp per row
  int64_t row = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  int lane = threadIdx.x & 31;
  if (row >= U) return;
  int64_t lo = indptr[row], hi = indptr[row + 1];
  if (lo > hi || lo < 0 || hi > nnz) {
    if (lane == 0) atomicExch(flag, 3);
    return;
  }
  if (lane == 0 && hi - lo >= (int64_t)I - 1) atomicExch(flag, 4);
  for (int64_t q = lo + lane; q < hi; q += 32) {
    coo_user[q] = (int32_t)row;
    int32_t it = indices[q];
    if (it <= 0 || (uint32_t)it >= I) atomicExch(flag, 5);
    if (q > lo && indices[q - 1] >= it) atomicExch(flag, 6);
  }
}

// Standalone sampler: one group of 8 lanes per triple.
__global__ void sample_only(const TrainParams p, const int64_t* __restrict__ triple_idx,
                            int64_t n64, int64_t* __restrict__ out) {
  const Group<8> g;
  int64_t k = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8;
  if (k >= n64) return;  // whole group exits together
  int64_t t = triple_idx[k];
  int32_t uu = p.coo_user[t];
  int64_t lo = p.indptr[uu], hi = p.indptr[uu + 1];
  int32_t j = draw_negative<8>(p, p.step, (uint64_t)t, lo, hi, g);
  if (j < 0) {
    if (g.gl == 0) atomicExch(p.flag, 1);
    j = 1;
  }
  if (g.gl == 0) out[k] = (int64_t)j;
}

template <int OPT>
int launch_phase_a_opt(rbpr_ctx* ctx, const TrainParams& p, int lanes, int nv, cudaStream_t st) {
  const int groups_per_block = 256 / lanes;
  const int64_t groups = ((int64_t)p.n + p.chunk - 1) / p.chunk;
  const int blocks = (int)((groups + groups_per_block - 1) / groups_per_block);
#define RBPR_CASE(L, V)                                      \
  if (lanes == L && nv == V) {                               \
    bpr_phase_a<L, V, OPT><<<blocks, 256, 0, st>>>(p);       \
    return 0;                                                \
  }
  RBPR_CASE(1, 1) RBPR_CASE(2, 1) RBPR_CASE(4, 1) RBPR_CASE(8, 1) RBPR_CASE(16, 1)
  RBPR_CASE(32, 1) RBPR_CASE(32, 2) RBPR_CASE(32, 3) RBPR_CASE(32, 4) RBPR_CASE(32, 5)
  RBPR_CASE(32, 6) RBPR_CASE(32, 7) RBPR_CASE(32, 8)
#undef RBPR_CASE
  RBPR_FAIL(ctx, RBPR_ERR_ARG, "unsupported dim geometry lanes=%d nv=%d", lanes, nv);
}

template <int OPT>
int launch_apply_opt(rbpr_ctx* ctx, const ApplyParams& p, int lanes, int nv, cudaStream_t st) {
  const int groups_per_block = 256 / lanes;
  int64_t blocks64 = (p.I + groups_per_block - 1) / groups_per_block;
  const int64_t maxb = (int64_t)ctx->sm_count * 8;
  const int blocks = (int)(blocks64 < maxb ? blocks64 : maxb);
#define RBPR_CASE(L, V)                                      \
  if (lanes == L && nv == V) {                               \
    bpr_apply_items<L, V, OPT><<<blocks, 256, 0, st>>>(p);   \
    return 0;                                                \
  }
  RBPR_CASE(1, 1) RBPR_CASE(2, 1) RBPR_CASE(4, 1) RBPR_CASE(8, 1) RBPR_CASE(16, 1)
  RBPR_CASE(32, 1) RBPR_CASE(32, 2) RBPR_CASE(32, 3) RBPR_CASE(32, 4) RBPR_CASE(32, 5)
  RBPR_CASE(32, 6) RBPR_CASE(32, 7) RBPR_CASE(32, 8)
#undef RBPR_CASE
  RBPR_FAIL(ctx, RBPR_ERR_ARG, "unsupported dim geometry lanes=%d nv=%d", lanes, nv);
}

int flush_users_launch(rbpr_ctx* ctx, int64_t step, const rbpr_hparams* hp, int lanes, int nv,
                       cudaStream_t st) {
  const int blocks = ctx->sm_count * 8;
#define RBPR_CASE(L, V)                                                                      \
  if (lanes == L && nv == V) {                                                               \
    bpr_flush_users<L, V><<<blocks, 256, 0, st>>>(ctx->user_emb, ctx->user_m, ctx->user_v,   \
                                                  ctx->user_last, ctx->U, ctx->D, step,      \
                                                  hp->lr, hp->beta1, hp->beta2, hp->eps);    \
    return 0;                                                                                \
  }
  RBPR_CASE(1, 1) RBPR_CASE(2, 1) RBPR_CASE(4, 1) RBPR_CASE(8, 1) RBPR_CASE(16, 1)
  RBPR_CASE(32, 1) RBPR_CASE(32, 2) RBPR_CASE(32, 3) RBPR_CASE(32, 4) RBPR_CASE(32, 5)
  RBPR_CASE(32, 6) RBPR_CASE(32, 7) RBPR_CASE(32, 8)
#undef RBPR_CASE
  RBPR_FAIL(ctx, RBPR_ERR_ARG, "unsupported dim geometry lanes=%d nv=%d", lanes, nv);
}

int check_ready(rbpr_ctx* ctx, const rbpr_hparams* hp) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!hp) RBPR_FAIL(ctx, RBPR_ERR_ARG, "hparams is NULL");
  if (!ctx->user_emb || !ctx->item_emb) RBPR_FAIL(ctx, RBPR_ERR_STATE, "tables not bound");
  if (!ctx->indptr || !ctx->indices) RBPR_FAIL(ctx, RBPR_ERR_STATE, "CSR not bound");
  if (ctx->csr_users != ctx->U)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "CSR has %lld rows but the user table has %lld",
              (long long)ctx->csr_users, (long long)ctx->U);
  if (hp->optimizer != RBPR_OPT_SGD && hp->optimizer != RBPR_OPT_ADAM)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "unknown optimizer %d", hp->optimizer);
  if (hp->optimizer == RBPR_OPT_ADAM) {
    if (!ctx->user_m || !ctx->user_v || !ctx->item_m || !ctx->item_v || !ctx->user_last)
      RBPR_FAIL(ctx, RBPR_ERR_STATE, "Adam requested but Adam state not bound");
    if (ctx->item_bias && (!ctx->bias_m || !ctx->bias_v))
      RBPR_FAIL(ctx, RBPR_ERR_STATE, "Adam requested with item bias but bias state not bound");
  }
  if (hp->sampler < RBPR_SAMPLER_UNIFORM || hp->sampler > RBPR_SAMPLER_INJECTED)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "unknown sampler %d", hp->sampler);
  if (hp->sampler == RBPR_SAMPLER_WEIGHTED && (!ctx->alias_prob || !ctx->alias_idx))
    RBPR_FAIL(ctx, RBPR_ERR_STATE, "weighted sampler requested but alias table not bound");
  return 0;
}

int ensure_capacity(rbpr_ctx* ctx, int64_t n, int64_t steps) {
  if (n > ctx->cap) {
    cudaFree(ctx->keys_in);
    cudaFree(ctx->keys_out);
    cudaFree(ctx->pos_in);
    cudaFree(ctx->pos_out);
    ctx->keys_in = ctx->keys_out = nullptr;
    ctx->pos_in = ctx->pos_out = nullptr;
    ctx->cap = 0;
    RBPR_CUDA(ctx, cudaMalloc(&ctx->keys_in, n * sizeof(uint64_t)));
    RBPR_CUDA(ctx, cudaMalloc(&ctx->keys_out, n * sizeof(uint64_t)));
    RBPR_CUDA(ctx, cudaMalloc(&ctx->pos_in, n * sizeof(int32_t)));
    RBPR_CUDA(ctx, cudaMalloc(&ctx->pos_out, n * sizeof(int32_t)));
    ctx->cap = n;
  }
  if (steps > ctx->stats_cap) {
    cudaFree(ctx->stats);
    ctx->stats = nullptr;
    ctx->stats_cap = 0;
    RBPR_CUDA(ctx, cudaMalloc(&ctx->stats, steps * RBPR_STATS_PER_STEP * sizeof(double)));
    ctx->stats_cap = steps;
  }
  return 0;
}

void fill_train_params(rbpr_ctx* ctx, TrainParams& p, uint64_t seed, const rbpr_hparams* hp) {
  memset(&p, 0, sizeof(p));
  p.user_emb = ctx->user_emb;
  p.item_emb = ctx->item_emb;
  p.item_bias = ctx->item_bias;
  p.user_m = ctx->user_m;
  p.user_v = ctx->user_v;
  p.user_last = ctx->user_last;
  p.item_grad = ctx->item_grad;
  p.bias_grad = ctx->item_bias ? ctx->item_grad + ctx->I * ctx->D : nullptr;
  p.touched = ctx->touched;
  p.indptr = ctx->indptr;
  p.indices = ctx->indices;
  p.coo_user = ctx->coo_user;
  p.alias_prob = ctx->alias_prob;
  p.alias_idx = ctx->alias_idx;
  p.flag = ctx->flag;
  p.D = ctx->D;
  p.I = (uint32_t)ctx->I;
  p.seed_lo = (uint32_t)seed;
  p.seed_hi = (uint32_t)(seed >> 32);
  if (hp) {
    p.sampler = hp->sampler;
    p.lr = hp->lr;
    p.beta1 = hp->beta1;
    p.beta2 = hp->beta2;
    p.eps = hp->eps;
    p.reg_user = hp->reg_user;
    p.reg_item = hp->reg_item;
    p.reg_neg = hp->reg_neg;
  }
}

int pick_chunk(rbpr_ctx* ctx, int64_t n, int lanes) {
  static int env_chunk = -1;
  if (env_chunk < 0) {
    const char* e = getenv("RBPR_CHUNK");
    env_chunk = e ? atoi(e) : 0;
  }
  if (env_chunk > 0) return env_chunk;
  // aim for >= 2 waves of resident groups (48 warps/SM), chunk in [1, 16]
  const int64_t resident_groups = (int64_t)ctx->sm_count * 48 * (32 / lanes);
  int64_t c = n / (2 * resident_groups);
  if (c < 1) c = 1;
  if (c > 16) c = 16;
  return (int)c;
}

cudaEvent_t next_event(rbpr_ctx* ctx) {
  if (ctx->ev_used == ctx->ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->ev.push_back(e);
  }
  return ctx->ev[ctx->ev_used++];
}

// Sort (step,t) keys for n triples split into batches; fills keys_out (+pos_out).
int sort_batches(rbpr_ctx* ctx, const int64_t* triple_idx, int64_t n, int64_t batch,
                 bool need_pos, cudaStream_t st) {
  const int64_t steps = (n + batch - 1) / batch;
  int rc = ensure_capacity(ctx, n, steps);
  if (rc) return rc;
  const int threads = 256;
  const int blocks = (int)((n + threads - 1) / threads);
  make_keys<<<blocks, threads, 0, st>>>(triple_idx, n, batch, ctx->nnz, ctx->keys_in,
                                        need_pos ? ctx->pos_in : nullptr, ctx->flag);
  ctx->launches++;
  int step_bits = 0;
  while ((1ll << step_bits) < steps) ++step_bits;
  int nnz_bits = 1;
  while ((1ll << nnz_bits) < ctx->nnz) ++nnz_bits;
  // keys: low 32 bits = triple id (< nnz), high = step index. Sort bits [0,nnz_bits) and
  // [32, 32+step_bits): cub sorts a contiguous bit range, so sort [0, 32+step_bits).
  const int end_bit = (step_bits == 0) ? nnz_bits : 32 + step_bits;
  size_t need = 0;
  if (need_pos)
    cub::DeviceRadixSort::SortPairs(nullptr, need, ctx->keys_in, ctx->keys_out, ctx->pos_in,
                                    ctx->pos_out, n, 0, end_bit, st);
  else
    cub::DeviceRadixSort::SortKeys(nullptr, need, ctx->keys_in, ctx->keys_out, n, 0, end_bit, st);
  if (need > ctx->cub_tmp_bytes) {
    cudaFree(ctx->cub_tmp);
    ctx->cub_tmp = nullptr;
    ctx->cub_tmp_bytes = 0;
    RBPR_CUDA(ctx, cudaMalloc(&ctx->cub_tmp, need));
    ctx->cub_tmp_bytes = need;
  }
  size_t tb = ctx->cub_tmp_bytes;
  if (need_pos)
    RBPR_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->cub_tmp, tb, ctx->keys_in, ctx->keys_out,
                                                   ctx->pos_in, ctx->pos_out, n, 0, end_bit, st));
  else
    RBPR_CUDA(ctx, cub::DeviceRadixSort::SortKeys(ctx->cub_tmp, tb, ctx->keys_in, ctx->keys_out,
                                                  n, 0, end_bit, st));
  return 0;
}

int run_phase_a(rbpr_ctx* ctx, TrainParams& p, const rbpr_hparams* hp, cudaStream_t st) {
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  p.chunk = pick_chunk(ctx, p.n, lanes);
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  if (ctx->timing) {
    e0 = next_event(ctx);
    e1 = next_event(ctx);
    cudaEventRecord(e0, st);
  }
  int rc = (hp->optimizer == RBPR_OPT_SGD) ? launch_phase_a_opt<RBPR_OPT_SGD>(ctx, p, lanes, nv, st)
                                           : launch_phase_a_opt<RBPR_OPT_ADAM>(ctx, p, lanes, nv, st);
  if (rc) return rc;
  if (ctx->timing) cudaEventRecord(e1, st);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int run_apply(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, int dense, cudaStream_t st) {
  ApplyParams a;
  memset(&a, 0, sizeof(a));
  a.item_emb = ctx->item_emb;
  a.item_bias = ctx->item_bias;
  a.item_m = ctx->item_m;
  a.item_v = ctx->item_v;
  a.bias_m = ctx->bias_m;
  a.bias_v = ctx->bias_v;
  a.item_grad = ctx->item_grad;
  a.bias_grad = ctx->item_bias ? ctx->item_grad + ctx->I * ctx->D : nullptr;
  a.touched = ctx->touched;
  a.I = ctx->I;
  a.D = ctx->D;
  a.dense = dense || hp->optimizer == RBPR_OPT_ADAM;
  a.step = step;
  a.lr = hp->lr;
  a.beta1 = hp->beta1;
  a.beta2 = hp->beta2;
  a.eps = hp->eps;
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  int rc = (hp->optimizer == RBPR_OPT_SGD) ? launch_apply_opt<RBPR_OPT_SGD>(ctx, a, lanes, nv, st)
                                           : launch_apply_opt<RBPR_OPT_ADAM>(ctx, a, lanes, nv, st);
  if (rc) return rc;
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int check_flag(rbpr_ctx* ctx, cudaStream_t st) {
  int32_t f = 0;
  RBPR_CUDA(ctx, cudaMemcpyAsync(&f, ctx->flag, sizeof(f), cudaMemcpyDeviceToHost, st));
  RBPR_CUDA(ctx, cudaStreamSynchronize(st));
  if (f != 0) {
    cudaMemsetAsync(ctx->flag, 0, sizeof(int32_t), st);
    switch (f) {
      case 1: RBPR_FAIL(ctx, RBPR_ERR_DATA, "negative sampler exhausted 1024 attempts for a triple");
      case 2: RBPR_FAIL(ctx, RBPR_ERR_ARG, "triple index out of range [0,nnz)");
      case 3: RBPR_FAIL(ctx, RBPR_ERR_DATA, "CSR indptr is not monotone within [0,nnz]");
      case 4: RBPR_FAIL(ctx, RBPR_ERR_DATA, "a user has seen every non-padding item: no negative exists");
      case 5: RBPR_FAIL(ctx, RBPR_ERR_DATA, "CSR item id outside [1,num_items)");
      case 6: RBPR_FAIL(ctx, RBPR_ERR_DATA, "CSR row is not strictly ascending");
      default: RBPR_FAIL(ctx, RBPR_ERR_DATA, "device error flag %d", f);
    }
  }
  return 0;
}

}  // namespace

extern "C" {

int rbpr_bind_csr(rbpr_ctx* ctx, const int64_t* indptr, const int32_t* indices, int64_t num_users,
                  int64_t nnz, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!indptr || !indices || num_users <= 0 || nnz <= 0)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_csr: null pointer or empty matrix");
  if (nnz >= (1ll << 32)) RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_csr: nnz must be < 2^32");
  if (ctx->I <= 1) RBPR_FAIL(ctx, RBPR_ERR_STATE, "bind_csr: bind tables first");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaFree(ctx->coo_user);
  ctx->coo_user = nullptr;
  RBPR_CUDA(ctx, cudaMalloc(&ctx->coo_user, nnz * sizeof(int32_t)));
  const int64_t threads = num_users * 32;
  expand_rows<<<(int)((threads + 255) / 256), 256, 0, st>>>(indptr, num_users, nnz,
                                                            (uint32_t)ctx->I, ctx->coo_user,
                                                            indices, ctx->flag);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  int rc = check_flag(ctx, st);
  if (rc) return rc;
  ctx->indptr = indptr;
  ctx->indices = indices;
  ctx->csr_users = num_users;
  ctx->nnz = nnz;
  return 0;
}

int rbpr_bind_item_alias(rbpr_ctx* ctx, const float* prob, const int32_t* alias) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!prob || !alias) RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_item_alias: null pointer");
  ctx->alias_prob = prob;
  ctx->alias_idx = alias;
  return 0;
}

int rbpr_sample_negatives(rbpr_ctx* ctx, const int64_t* triple_idx, int64_t n, uint64_t seed,
                          uint64_t step, int32_t sampler, int64_t* neg_out, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  rbpr_hparams hp;
  memset(&hp, 0, sizeof(hp));
  hp.sampler = sampler;
  int rc = check_ready(ctx, &hp);
  if (rc) return rc;
  if (sampler == RBPR_SAMPLER_INJECTED) RBPR_FAIL(ctx, RBPR_ERR_ARG, "sampler must draw");
  if (n == 0) return 0;
  if (!triple_idx || !neg_out || n < 0) RBPR_FAIL(ctx, RBPR_ERR_ARG, "sample: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  TrainParams p;
  fill_train_params(ctx, p, seed, &hp);
  p.step = step;
  const int64_t threads = n * 8;
  sample_only<<<(int)((threads + 255) / 256), 256, 0, st>>>(p, triple_idx, n, neg_out);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return check_flag(ctx, st);
}

int rbpr_train_steps(rbpr_ctx* ctx, const int64_t* triple_idx, int64_t n, int64_t batch,
                     uint64_t seed, uint64_t step0, const rbpr_hparams* hp, const int64_t* neg_in,
                     int64_t* neg_out, double* stats_out, void* stream) {
  int rc = check_ready(ctx, hp);
  if (rc) return rc;
  if (n == 0) return 0;
  if (!triple_idx || n < 0 || batch <= 0) RBPR_FAIL(ctx, RBPR_ERR_ARG, "train: bad arguments");
  if (batch >= (1ll << 31) || n >= (1ll << 31))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "train: n and batch must be < 2^31 per call");
  if (hp->sampler == RBPR_SAMPLER_INJECTED && !neg_in)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "train: injected sampler needs neg_in");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  const bool need_pos = (neg_in != nullptr && hp->sampler == RBPR_SAMPLER_INJECTED) || neg_out;
  rc = sort_batches(ctx, triple_idx, n, batch, need_pos, st);
  if (rc) return rc;
  const int64_t steps = (n + batch - 1) / batch;
  RBPR_CUDA(ctx, cudaMemsetAsync(ctx->stats, 0, steps * RBPR_STATS_PER_STEP * sizeof(double), st));
  TrainParams p;
  fill_train_params(ctx, p, seed, hp);
  for (int64_t s = 0; s < steps; ++s) {
    const int64_t off = s * batch;
    p.n = (int)((n - off) < batch ? (n - off) : batch);
    p.keys = ctx->keys_out + off;
    p.pos = need_pos ? ctx->pos_out + off : nullptr;
    // positions stored are global (0..n): neg_in/neg_out are indexed globally
    p.neg_in = neg_in;
    p.neg_out = neg_out;
    p.stats = ctx->stats + s * RBPR_STATS_PER_STEP;
    p.step = step0 + (uint64_t)s;
    rc = run_phase_a(ctx, p, hp, st);
    if (rc) return rc;
    rc = run_apply(ctx, p.step, hp, 0, st);
    if (rc) return rc;
  }
  if (stats_out)
    RBPR_CUDA(ctx, cudaMemcpyAsync(stats_out, ctx->stats,
                                   steps * RBPR_STATS_PER_STEP * sizeof(double),
                                   cudaMemcpyDeviceToDevice, st));
  return 0;
}

int rbpr_train_steps_host(rbpr_ctx* ctx, const int64_t* triple_idx_host, int64_t n, int64_t batch,
                          uint64_t seed, uint64_t step0, const rbpr_hparams* hp,
                          const int64_t* neg_in_host, int64_t* neg_out_host,
                          double* stats_out_host, void* stream) {
  int rc = check_ready(ctx, hp);
  if (rc) return rc;
  if (n == 0) return 0;
  if (!triple_idx_host || n < 0 || batch <= 0)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "train_host: bad arguments");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n > ctx->stage_cap) {
    cudaFree(ctx->stage_idx);
    cudaFree(ctx->stage_neg);
    ctx->stage_idx = ctx->stage_neg = nullptr;
    ctx->stage_cap = 0;
    RBPR_CUDA(ctx, cudaMalloc(&ctx->stage_idx, n * sizeof(int64_t)));
    RBPR_CUDA(ctx, cudaMalloc(&ctx->stage_neg, n * sizeof(int64_t)));
    ctx->stage_cap = n;
  }
  RBPR_CUDA(ctx, cudaMemcpyAsync(ctx->stage_idx, triple_idx_host, n * sizeof(int64_t),
                                 cudaMemcpyHostToDevice, st));
  const bool inj = hp->sampler == RBPR_SAMPLER_INJECTED;
  if (inj) {
    if (!neg_in_host) RBPR_FAIL(ctx, RBPR_ERR_ARG, "train_host: injected sampler needs neg_in");
    RBPR_CUDA(ctx, cudaMemcpyAsync(ctx->stage_neg, neg_in_host, n * sizeof(int64_t),
                                   cudaMemcpyHostToDevice, st));
  }
  // neg_in and neg_out may alias the same staging buffer: each position is read before written
  rc = rbpr_train_steps(ctx, ctx->stage_idx, n, batch, seed, step0, hp,
                        inj ? ctx->stage_neg : nullptr, neg_out_host ? ctx->stage_neg : nullptr,
                        nullptr, st);
  if (rc) return rc;
  const int64_t steps = (n + batch - 1) / batch;
  if (stats_out_host)
    RBPR_CUDA(ctx, cudaMemcpyAsync(stats_out_host, ctx->stats,
                                   steps * RBPR_STATS_PER_STEP * sizeof(double),
                                   cudaMemcpyDeviceToHost, st));
  if (neg_out_host)
    RBPR_CUDA(ctx, cudaMemcpyAsync(neg_out_host, ctx->stage_neg, n * sizeof(int64_t),
                                   cudaMemcpyDeviceToHost, st));
  return check_flag(ctx, st);
}

int rbpr_grad_step(rbpr_ctx* ctx, const int64_t* triple_idx, int64_t n, uint64_t seed,
                   uint64_t step, const rbpr_hparams* hp, const int64_t* neg_in, int64_t* neg_out,
                   double* stats_out, void* stream) {
  int rc = check_ready(ctx, hp);
  if (rc) return rc;
  if (n < 0 || n >= (1ll << 31)) RBPR_FAIL(ctx, RBPR_ERR_ARG, "grad_step: bad n");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  rc = ensure_capacity(ctx, n > 0 ? n : 1, 1);
  if (rc) return rc;
  RBPR_CUDA(ctx, cudaMemsetAsync(ctx->stats, 0, RBPR_STATS_PER_STEP * sizeof(double), st));
  if (n > 0) {
    if (!triple_idx) RBPR_FAIL(ctx, RBPR_ERR_ARG, "grad_step: null triple_idx");
    if (hp->sampler == RBPR_SAMPLER_INJECTED && !neg_in)
      RBPR_FAIL(ctx, RBPR_ERR_ARG, "grad_step: injected sampler needs neg_in");
    const bool need_pos = (neg_in != nullptr && hp->sampler == RBPR_SAMPLER_INJECTED) || neg_out;
    rc = sort_batches(ctx, triple_idx, n, n, need_pos, st);
    if (rc) return rc;
    TrainParams p;
    fill_train_params(ctx, p, seed, hp);
    p.n = (int)n;
    p.keys = ctx->keys_out;
    p.pos = need_pos ? ctx->pos_out : nullptr;
    p.neg_in = neg_in;
    p.neg_out = neg_out;
    p.stats = ctx->stats;
    p.step = step;
    rc = run_phase_a(ctx, p, hp, st);
    if (rc) return rc;
  }
  if (stats_out)
    RBPR_CUDA(ctx, cudaMemcpyAsync(stats_out, ctx->stats, RBPR_STATS_PER_STEP * sizeof(double),
                                   cudaMemcpyDeviceToDevice, st));
  return 0;
}

int rbpr_item_grad_buffer(rbpr_ctx* ctx, float** ptr, int64_t* numel) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!ctx->item_grad) RBPR_FAIL(ctx, RBPR_ERR_STATE, "tables not bound");
  if (ptr) *ptr = ctx->item_grad;
  if (numel) *numel = ctx->I * ctx->D + (ctx->item_bias ? ctx->I : 0);
  return 0;
}

int rbpr_apply_item_grads(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, void* stream) {
  int rc = check_ready(ctx, hp);
  if (rc) return rc;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  return run_apply(ctx, step, hp, 1, (cudaStream_t)stream);
}

int rbpr_sync_check(rbpr_ctx* ctx, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  return check_flag(ctx, (cudaStream_t)stream);
}

int rbpr_flush_lazy(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!hp) RBPR_FAIL(ctx, RBPR_ERR_ARG, "hparams is NULL");
  if (hp->optimizer != RBPR_OPT_ADAM) return 0;
  if (!ctx->user_m || !ctx->user_v || !ctx->user_last)
    RBPR_FAIL(ctx, RBPR_ERR_STATE, "Adam state not bound");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  int rc = flush_users_launch(ctx, (int64_t)step, hp, lanes, nv, (cudaStream_t)stream);
  if (rc) return rc;
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // extern "C"
