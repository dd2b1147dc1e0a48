// Adaptive (factor / rank) negative sampling for sm_100a — DESIGN.md §3.3.
// Reference: AdaptiveSampler.sample / update_stats (revisit_bpr/modules/neg_samplers.py:74-132),
// BPRExperiment._adaptive_sampling / _update_adaptive_stats (experiments/bpr/exp.py:295-354).
//
//   adaptive_update_stats : snapshot the item table transposed (D,I), unbiased per-factor std over
//                           items 1..I-1, sort every factor column once (descending value, ties by
//                           ascending item id; one global radix sort on (factor, ~value) keys) -> order[f][q], and its inverse pos[f][item].
//   sample_adaptive       : per slot: factor ~ Categorical(|u_f| * std_f), rank ~ Geometric(p)
//                           clamped to the number of unseen items, item = rank-th unseen item in
//                           the factor's order (from the top if u_f > 0, else from the bottom).
//                           The rank-th UNSEEN position is the fixed point of
//                           q <- r + #{masked positions <= q} (masked = seen items and item 0).
#include <cub/device/device_radix_sort.cuh>

#include "train_kernels.cuh"

using namespace rbpr_dev;

namespace {

// snapshot[f][i] = item_emb[i][f]; ids[f][i] = i   (tile transpose through shared memory)
// key = (factor << 32) | ~monotone(value): ONE global ascending radix sort orders every factor
// column by descending value; radix sort is stable, so ties keep ascending item id.
__device__ __forceinline__ uint32_t desc_key(float x) {
  const uint32_t u = __float_as_uint(x == 0.f ? 0.f : x);  // -0.0 == +0.0
  const uint32_t asc = (u & 0x80000000u) ? ~u : (u | 0x80000000u);
  return ~asc;
}

__global__ void snapshot_transpose(const float* __restrict__ item_emb, int64_t I, int D,
                                   float* __restrict__ snap, int32_t* __restrict__ ids,
                                   uint64_t* __restrict__ keys) {
  __shared__ float tile[32][33];
  const int64_t i0 = (int64_t)blockIdx.x * 32;
  const int f0 = blockIdx.y * 32;
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int64_t i = i0 + r;
    const int f = f0 + threadIdx.x;
    tile[r][threadIdx.x] = (i < I && f < D) ? item_emb[i * D + f] : 0.f;
  }
  __syncthreads();
  for (int r = threadIdx.y; r < 32; r += blockDim.y) {
    const int f = f0 + r;
    const int64_t i = i0 + threadIdx.x;
    if (f < D && i < I) {
      const float v = tile[threadIdx.x][r];
      snap[(int64_t)f * I + i] = v;
      ids[(int64_t)f * I + i] = (int32_t)i;
      keys[(int64_t)f * I + i] = ((uint64_t)f << 32) | (uint64_t)desc_key(v);
    }
  }
}

// unbiased std of snapshot[f][1..I-1], one block per factor, two passes in double
__global__ void factor_std(const float* __restrict__ snap, int64_t I, float* __restrict__ out) {
  __shared__ double red[256];
  __shared__ double s_mean;
  const float* row = snap + (int64_t)blockIdx.x * I;
  const int64_t n = I - 1;
  double acc = 0.0;
  for (int64_t i = 1 + threadIdx.x; i < I; i += blockDim.x) acc += (double)row[i];
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) s_mean = red[0] / (double)n;
  __syncthreads();
  const double mean = s_mean;
  acc = 0.0;
  for (int64_t i = 1 + threadIdx.x; i < I; i += blockDim.x) {
    const double d = (double)row[i] - mean;
    acc += d * d;
  }
  __syncthreads();
  red[threadIdx.x] = acc;
  __syncthreads();
  for (int o = 128; o > 0; o >>= 1) {
    if ((int)threadIdx.x < o) red[threadIdx.x] += red[threadIdx.x + o];
    __syncthreads();
  }
  if (threadIdx.x == 0) out[blockIdx.x] = (float)sqrt(red[0] / (double)(n - 1));
}

__global__ void invert_order(const int32_t* __restrict__ order, int64_t I, int D,
                             int32_t* __restrict__ pos) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= I * D) return;
  const int64_t f = k / I, q = k % I;
  pos[f * I + order[k]] = (int32_t)q;
}

// 8 lanes per slot.  The factor draw is blocked over them: lane l owns the
// contiguous factors [l*blk, (l+1)*blk), blk = ceil(D/8) rounded up to a multiple of 4, so the
// inverse CDF costs 2*blk dependent adds per lane instead of 2*D (round 1: two sequential passes
// over D repeated by every lane: 72 us per 65 536 slots at D=64); the seen-row scans stride by 8.
constexpr int kAdaLanes = 8;

struct AdaptiveParams {
  const float* __restrict__ user_emb;
  const float* __restrict__ fstd;      // (D)
  const int32_t* __restrict__ order;   // (D,I)
  const int32_t* __restrict__ pos;     // (D,I)
  int64_t I;
  int D;
  double log1m_p;  // log1p(-p)
  uint32_t seed_lo, seed_hi;
  uint64_t step;
  int32_t* __restrict__ flag;
  // lazily updated user rows (stateful optimizers): the draw must see the row dense torch.optim would
  // hold after `opt_step` steps, so the missed zero-gradient steps are replayed in registers
  int opt;  // rbpr_optimizer; RBPR_OPT_SGD: rows are always current
  const float* __restrict__ user_m;
  const float* __restrict__ user_v;
  const int32_t* __restrict__ user_last;
  const float2* __restrict__ adam_tab;
  int64_t opt_step;
  float lr, beta1, beta2, eps;
};

// the current value of four consecutive factors of a user row
__device__ __forceinline__ float4 current_u4(const AdaptiveParams& p, int64_t user, int c, int64_t last) {
  float4 u = ld4(p.user_emb + user * p.D + c);
  if (p.opt != RBPR_OPT_SGD && last > 0 && last < p.opt_step) {
    float4 m = ld4(p.user_m + user * p.D + c);
    float4 v = (p.opt == RBPR_OPT_ADAM) ? ld4(p.user_v + user * p.D + c) : f4zero();
    const OptScalars h = {p.lr, p.beta1, p.beta2, p.eps, 0.f, 1.f};
    if (p.opt == RBPR_OPT_ADAM) Catchup<RBPR_OPT_ADAM>(last, p.opt_step, p.adam_tab, h).apply(u, m, v);
    else if (p.opt == RBPR_OPT_SGDM) Catchup<RBPR_OPT_SGDM>(last, p.opt_step, p.adam_tab, h).apply(u, m, v);
    else Catchup<RBPR_OPT_RMSPROP>(last, p.opt_step, p.adam_tab, h).apply(u, m, v);
  }
  return u;
}

// One group of kAdaLanes lanes per slot.  `masked(c)` enumerates the masked items of the slot's user:
// c in [0, n_mask) -> item id (0 = padding entries are ignored; item 0 itself is always masked).
template <typename RowFn>
__device__ __forceinline__ int32_t adaptive_draw(const AdaptiveParams& p, const Group<kAdaLanes>& g,
                                                 int64_t user, uint32_t sub_lo, uint32_t sub_hi,
                                                 int64_t n_entries, RowFn entry) {
  const int D = p.D;
  // number of unseen, non-padding items.  Rows of up to 32 entries (the common case: median degree
  // 5 on the Yelp shape) are read ONCE into registers, 4 per lane; longer rows are re-read.
  constexpr int kCache = 4;
  const bool cached = n_entries <= (int64_t)kAdaLanes * kCache;
  int32_t cent[kCache];
  int64_t n_seen = 0;
  if (cached) {
#pragma unroll
    for (int c = 0; c < kCache; ++c) {
      const int64_t e = g.gl + (int64_t)c * kAdaLanes;
      cent[c] = e < n_entries ? (int32_t)entry(e) : 0;
      n_seen += (cent[c] != 0);
    }
  } else {
    for (int64_t c = g.gl; c < n_entries; c += kAdaLanes) n_seen += (entry(c) != 0);
  }
#pragma unroll
  for (int o = kAdaLanes / 2; o > 0; o >>= 1) n_seen += __shfl_xor_sync(g.mask, n_seen, o);
  const int64_t n_unseen = (p.I - 1) - n_seen;
  if (n_unseen <= 0) {
    if (g.gl == 0) atomicExch(p.flag, 4);
    return 1;
  }
  const uint32_t step_lo = (uint32_t)(p.step << 8), step_hi = (uint32_t)(p.step >> 24);
  const philox4 r4 = philox4x32_10(step_lo, step_hi, sub_lo, sub_hi, p.seed_lo, p.seed_hi);
  // factor ~ Categorical(|u_f| * std_f), fp32 inverse CDF in a FIXED blocked order: lane l sums its
  // block sequentially (s_l), the block sums are accumulated in lane order (P_l), and the factor is the
  // first f, in natural order, with fl(... fl(fl(P_l + w_a) + w_b) ...) > target.
  const int blk = ((((D + kAdaLanes - 1) / kAdaLanes) + 3) / 4) * 4;
  const int f0 = g.gl * blk, f1 = min(D, f0 + blk);
  const int64_t last = (p.opt != RBPR_OPT_SGD) ? (int64_t)p.user_last[user] : 0;
  float s_l = 0.f;
  for (int c = f0; c < f1; c += 4) {
    const float4 u = current_u4(p, user, c, last);
    const float4 sd = ld4(p.fstd + c);
    s_l = __fadd_rn(s_l, __fmul_rn(fabsf(u.x), sd.x));
    s_l = __fadd_rn(s_l, __fmul_rn(fabsf(u.y), sd.y));
    s_l = __fadd_rn(s_l, __fmul_rn(fabsf(u.z), sd.z));
    s_l = __fadd_rn(s_l, __fmul_rn(fabsf(u.w), sd.w));
  }
  float before = 0.f, total = 0.f;  // P_l of this lane; P_8
#pragma unroll
  for (int l = 0; l < kAdaLanes; ++l) {
    const float sl = __shfl_sync(g.mask, s_l, g.shift + l);
    if (l == g.gl) before = total;
    total = __fadd_rn(total, sl);
  }
  if (!(total > 0.f)) {
    if (g.gl == 0) atomicExch(p.flag, 10);
    return 1;
  }
  const float target = __fmul_rn((float)(r4.x >> 8) * (1.0f / 16777216.0f), total);
  int hit = -1, lastpos = -1;
  float ufac = 0.f, ulast = 0.f;  // value of the user row at the chosen / last positive factor
  float cum = before;
  for (int c = f0; c < f1; c += 4) {
    const float4 u = current_u4(p, user, c, last);
    const float4 sd = ld4(p.fstd + c);
    const float uv[4] = {u.x, u.y, u.z, u.w};
    const float sv[4] = {sd.x, sd.y, sd.z, sd.w};
#pragma unroll
    for (int e = 0; e < 4; ++e) {
      const float w = __fmul_rn(fabsf(uv[e]), sv[e]);
      cum = __fadd_rn(cum, w);
      if (w > 0.f) {
        lastpos = c + e;
        ulast = uv[e];
      }
      if (hit < 0 && cum > target) {
        hit = c + e;
        ufac = uv[e];
      }
    }
  }
  // first lane (lowest block) that found a crossing; else the last factor with positive weight
  const unsigned found = __ballot_sync(g.mask, hit >= 0) >> g.shift;
  const unsigned pos_any = __ballot_sync(g.mask, lastpos >= 0) >> g.shift;
  const int src = found != 0u ? (__ffs(found) - 1) : (31 - __clz(pos_any));
  const int factor = __shfl_sync(g.mask, found != 0u ? hit : lastpos, g.shift + src);
  const float u_factor = __shfl_sync(g.mask, found != 0u ? ufac : ulast, g.shift + src);
  // rank ~ Geometric(p) on {1,2,...}, clamped to the number of unseen items
  const double u2 = ((double)(r4.y >> 8) + 1.0) * (1.0 / 16777216.0);
  double gq = ceil(log(u2) / p.log1m_p);
  if (!(gq >= 1.0)) gq = 1.0;
  const int64_t rank = gq > (double)n_unseen ? n_unseen : (int64_t)gq;
  const int64_t r = (u_factor > 0.f) ? rank - 1 : n_unseen - rank;
  // r-th unmasked position of the factor's order: fixed point of q <- r + #{masked pos <= q}
  const int32_t* prow = p.pos + (int64_t)factor * p.I;
  const int32_t pos0 = prow[0];
  int64_t q = r;
  if (cached) {  // positions of the masked items: one gather, then the fixed point runs on registers
    int32_t cpos[kCache];
#pragma unroll
    for (int c = 0; c < kCache; ++c) cpos[c] = cent[c] != 0 ? prow[cent[c]] : 0x7fffffff;
    while (true) {
      int64_t c = 0;
#pragma unroll
      for (int e = 0; e < kCache; ++e) c += ((int64_t)cpos[e] <= q);
#pragma unroll
      for (int o = kAdaLanes / 2; o > 0; o >>= 1) c += __shfl_xor_sync(g.mask, c, o);
      c += (pos0 <= q);
      const int64_t nq = r + c;
      if (nq == q) break;
      q = nq;
    }
  } else {
    while (true) {
      int64_t c = 0;
      for (int64_t e = g.gl; e < n_entries; e += kAdaLanes) {
        const int64_t it = entry(e);
        c += (it != 0 && (int64_t)prow[it] <= q);
      }
#pragma unroll
      for (int o = kAdaLanes / 2; o > 0; o >>= 1) c += __shfl_xor_sync(g.mask, c, o);
      c += (pos0 <= q);
      const int64_t nq = r + c;
      if (nq == q) break;
      q = nq;
    }
  }
  return p.order[(int64_t)factor * p.I + q];
}

__global__ void __launch_bounds__(256)
sample_adaptive_padded(const AdaptiveParams p, const int64_t* __restrict__ users,
                       const int64_t* __restrict__ seen, int64_t B, int64_t S, int64_t num,
                       int64_t U, int64_t* __restrict__ out) {
  const Group<kAdaLanes> g;
  const int64_t slot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / kAdaLanes;
  if (slot >= B * num) return;
  const int64_t row = slot / num;
  int64_t user = users[row];
  if (user < 0 || user >= U) {
    if (g.gl == 0) atomicExch(p.flag, 7);
    user = 0;
  }
  const int64_t* srow = seen + row * S;
  const int64_t I = p.I;
  auto entry = [&](int64_t c) -> int64_t {
    const int64_t v = srow[c];
    return (v > 0 && v < I) ? v : 0;
  };
  const int32_t j = adaptive_draw(p, g, user, (uint32_t)slot, (uint32_t)((uint64_t)slot >> 32), S, entry);
  if (g.gl == 0) out[slot] = (int64_t)j;
}

}  // namespace

// CSR variant used by bpr_sample-style preparation (train.cu): records for one step (tp.batch >= n).
__global__ void __launch_bounds__(256, 4)
rbpr_sample_adaptive_csr(const TrainParams tp, const float* fstd, const int32_t* order,
                         const int32_t* pos, double log1m_p, int4* __restrict__ records,
                         uint64_t n_slots, uint64_t step, int opt) {
  const Group<kAdaLanes> g;
  const uint64_t k = ((uint64_t)blockIdx.x * blockDim.x + threadIdx.x) / kAdaLanes;
  if (k >= n_slots) return;
  int64_t t64 = __ldg(tp.triple_idx + k);
  if (t64 < 0 || t64 >= tp.nnz) t64 = 0;  // flagged by count_users
  const uint32_t t = (uint32_t)t64;
  const int32_t uu = __ldg(tp.coo_user + t);
  const int32_t i = __ldg(tp.indices + t);
  const int32_t flags = slot_flags(tp, k, 0u, uu);  // one step per launch
  AdaptiveParams p;
  p.user_emb = tp.user_emb;
  p.fstd = fstd;
  p.order = order;
  p.pos = pos;
  p.I = tp.I;
  p.D = tp.D;
  p.log1m_p = log1m_p;
  p.seed_lo = tp.seed_lo;
  p.seed_hi = tp.seed_hi;
  p.step = step;
  p.flag = tp.flag;
  p.opt = opt;
  p.user_m = tp.user_m;
  p.user_v = tp.user_v;
  p.user_last = tp.user_last;
  p.adam_tab = tp.adam_tab;
  p.opt_step = (int64_t)step;  // optimizer steps applied before this step
  p.lr = tp.lr;
  p.beta1 = tp.beta1;
  p.beta2 = tp.beta2;
  p.eps = tp.eps;
  const int64_t lo = tp.indptr[uu], hi = tp.indptr[uu + 1];
  const int32_t* idx = tp.indices;
  auto entry = [&](int64_t c) -> int64_t { return (int64_t)__ldg(idx + lo + c); };
  const int32_t j = adaptive_draw(p, g, (int64_t)uu, t, 0u, hi - lo, entry);
  if (g.gl == 0) {
    records[k] = make_int4(uu, i, j, flags);
    if (tp.neg_out != nullptr) tp.neg_out[k] = (int64_t)j;
  }
}

int rbpr_internal_adaptive_ready(rbpr_ctx* ctx);
int rbpr_internal_ensure_adam_table(rbpr_ctx* ctx, const rbpr_hparams* hp, int64_t first, int64_t last, cudaStream_t st);

// Records of ONE step from sorted keys, negatives drawn adaptively from the current user rows.
int rbpr_internal_sample_adaptive_csr(rbpr_ctx* ctx, const TrainParams& tp, void* records, int64_t n,
                                      uint64_t step, double sampling_prob, int opt, cudaStream_t st) {
  int rc = rbpr_internal_adaptive_ready(ctx);
  if (rc) return rc;
  if (!(sampling_prob > 0.0 && sampling_prob < 1.0))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "adaptive sampler: adaptive_prob must be in (0,1)");
  const int64_t threads = n * kAdaLanes;
  rbpr_sample_adaptive_csr<<<(unsigned)((threads + 255) / 256), 256, 0, st>>>(
      tp, ctx->ad_std, ctx->ad_order, ctx->ad_pos, log1p(-sampling_prob),
      reinterpret_cast<int4*>(records), (uint64_t)n, step, opt);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int rbpr_internal_adaptive_ready(rbpr_ctx* ctx) {
  if (!ctx->ad_order || !ctx->ad_pos || !ctx->ad_std)
    RBPR_FAIL(ctx, RBPR_ERR_STATE, "adaptive sampler: call rbpr_adaptive_update_stats first");
  return 0;
}

extern "C" {

int rbpr_adaptive_update_stats(rbpr_ctx* ctx, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!ctx->item_emb) RBPR_FAIL(ctx, RBPR_ERR_STATE, "tables not bound");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t I = ctx->I;
  const int D = ctx->D;
  const size_t cells = (size_t)I * D;
  if (ctx->ad_cells != cells) {
    cudaFree(ctx->ad_snap); cudaFree(ctx->ad_keys); cudaFree(ctx->ad_keys_sorted); cudaFree(ctx->ad_ids);
    cudaFree(ctx->ad_order); cudaFree(ctx->ad_pos); cudaFree(ctx->ad_std);
    cudaFree(ctx->ad_tmp);
    ctx->ad_snap = ctx->ad_std = nullptr;
    ctx->ad_keys = ctx->ad_keys_sorted = nullptr;
    ctx->ad_ids = ctx->ad_order = ctx->ad_pos = nullptr;
    ctx->ad_tmp = nullptr;
    ctx->ad_tmp_bytes = 0;
    ctx->ad_cells = 0;
    RBPR_CUDA(ctx, cudaMalloc(&ctx->ad_snap, cells * 4));
    RBPR_CUDA(ctx, cudaMalloc(&ctx->ad_keys, cells * 8));
    RBPR_CUDA(ctx, cudaMalloc(&ctx->ad_keys_sorted, cells * 8));
    RBPR_CUDA(ctx, cudaMalloc(&ctx->ad_ids, cells * 4));
    RBPR_CUDA(ctx, cudaMalloc(&ctx->ad_order, cells * 4));
    RBPR_CUDA(ctx, cudaMalloc(&ctx->ad_pos, cells * 4));
    RBPR_CUDA(ctx, cudaMalloc(&ctx->ad_std, (size_t)D * 4));
    size_t need = 0;
    cub::DeviceRadixSort::SortPairs(nullptr, need, ctx->ad_keys, ctx->ad_keys_sorted, ctx->ad_ids,
                                    ctx->ad_order, (int64_t)cells, 0, 64, st);
    RBPR_CUDA(ctx, cudaMalloc(&ctx->ad_tmp, need > 0 ? need : 16));
    ctx->ad_tmp_bytes = need;
    ctx->ad_cells = cells;
  }
  dim3 grid((unsigned)((I + 31) / 32), (unsigned)((D + 31) / 32));
  snapshot_transpose<<<grid, dim3(32, 8), 0, st>>>(ctx->item_emb, I, D, ctx->ad_snap, ctx->ad_ids,
                                                   ctx->ad_keys);
  factor_std<<<D, 256, 0, st>>>(ctx->ad_snap, I, ctx->ad_std);
  size_t tb = ctx->ad_tmp_bytes;
  int fbits = 1;
  while ((1 << fbits) < D) ++fbits;
  RBPR_CUDA(ctx, cub::DeviceRadixSort::SortPairs(ctx->ad_tmp, tb, ctx->ad_keys, ctx->ad_keys_sorted,
                                                 ctx->ad_ids, ctx->ad_order, (int64_t)cells, 0,
                                                 32 + fbits, st));
  invert_order<<<(unsigned)((cells + 255) / 256), 256, 0, st>>>(ctx->ad_order, I, D, ctx->ad_pos);
  ctx->launches += 4;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int rbpr_adaptive_stats(rbpr_ctx* ctx, const float** factor_std_out, const int32_t** order_out,
                        const int32_t** pos_out) {
  if (!ctx) return RBPR_ERR_ARG;
  int rc = rbpr_internal_adaptive_ready(ctx);
  if (rc) return rc;
  if (factor_std_out) *factor_std_out = ctx->ad_std;
  if (order_out) *order_out = ctx->ad_order;
  if (pos_out) *pos_out = ctx->ad_pos;
  return 0;
}

int rbpr_sample_adaptive_padded(rbpr_ctx* ctx, const int64_t* users, const int64_t* seen,
                                int64_t batch, int64_t width, int64_t num, double sampling_prob,
                                uint64_t seed, uint64_t step, int64_t* neg_out, uint64_t opt_step,
                                const rbpr_hparams* hp, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!ctx->user_emb || !ctx->item_emb) RBPR_FAIL(ctx, RBPR_ERR_STATE, "tables not bound");
  int rc = rbpr_internal_adaptive_ready(ctx);
  if (rc) return rc;
  if (batch < 0 || width < 0 || num < 1) RBPR_FAIL(ctx, RBPR_ERR_ARG, "sample_adaptive: bad sizes");
  if (!(sampling_prob > 0.0 && sampling_prob < 1.0))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "sample_adaptive: sampling_prob must be in (0,1)");
  if (batch == 0) return 0;
  if (!users || (width > 0 && !seen) || !neg_out)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "sample_adaptive: null pointer");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  AdaptiveParams p;
  p.user_emb = ctx->user_emb;
  p.fstd = ctx->ad_std;
  p.order = ctx->ad_order;
  p.pos = ctx->ad_pos;
  p.I = ctx->I;
  p.D = ctx->D;
  p.log1m_p = log1p(-sampling_prob);
  p.seed_lo = (uint32_t)seed;
  p.seed_hi = (uint32_t)(seed >> 32);
  p.step = step;
  p.flag = ctx->flag;
  p.opt = RBPR_OPT_SGD;
  p.user_m = p.user_v = nullptr;
  p.user_last = nullptr;
  p.adam_tab = nullptr;
  p.opt_step = 0;
  p.lr = p.beta1 = p.beta2 = p.eps = 0.f;
  if (hp != nullptr && hp->optimizer != RBPR_OPT_SGD && ctx->user_last != nullptr && ctx->user_m != nullptr) {
    // lazily updated user rows: replay their missed zero-gradient steps in registers before drawing
    if (hp->optimizer == RBPR_OPT_ADAM) {
      rc = rbpr_internal_ensure_adam_table(ctx, hp, (int64_t)opt_step, (int64_t)opt_step + 1, (cudaStream_t)stream);
      if (rc) return rc;
    }
    p.opt = hp->optimizer;
    p.user_m = ctx->user_m;
    p.user_v = ctx->user_v;
    p.user_last = ctx->user_last;
    p.adam_tab = ctx->adam_tab;
    p.opt_step = (int64_t)opt_step;
    p.lr = hp->lr;
    p.beta1 = hp->beta1;
    p.beta2 = hp->beta2;
    p.eps = hp->eps;
  }
  const int64_t threads = batch * num * kAdaLanes;
  sample_adaptive_padded<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      p, users, seen, batch, width, num, ctx->U, neg_out);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // extern "C"
