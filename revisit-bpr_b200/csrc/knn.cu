// Item-neighbourhood logits models of the reference (revisit_bpr/models/bpr/model.py:156-251):
//   ItemKNN      logits[b,i] = w[item[b,i]] . SUM_{s kept} w[seen[b,s]]      (+ bias[item[b,i]])
//   FreeItemKNN  logits[b,i] = SUM_{s kept} W[item[b,i], seen[b,s]]          (+ bias[item[b,i]])
// where a seen entry is "kept" unless its id occurs among item[b,:] (model.py:184-190, 230-235).
// Forward and backward kernels; the backward accumulates into dense gradient buffers with
// atomics, which is what autograd's index_put_(accumulate=True) does for the reference.
// One block per batch row everywhere: rows are independent, and a train batch (256..65 536 rows)
// fills the 148 SMs many times over.
#include "common.cuh"

namespace {

constexpr int kKnnThreads = 256;
constexpr int kKnnScanLimit = 32;  // item lists up to this length are scanned, longer ones use a bitmap

__device__ __forceinline__ float warp_sum(float v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
  return v;
}

// keep[b,s] = 0 when seen[b,s] occurs in item[b,:], else 1.  Ids out of range raise flag 8 and
// are neutralised (keep = 0 for a seen entry; the consumers clamp item ids).
__global__ void __launch_bounds__(kKnnThreads) knn_keep(const int64_t* __restrict__ item, int64_t n_items,
                                                        const int64_t* __restrict__ seen, int64_t S,
                                                        int64_t I, int bitmap_words,
                                                        uint8_t* __restrict__ keep,
                                                        int32_t* __restrict__ flag) {
  extern __shared__ uint32_t bits[];
  const int64_t b = blockIdx.x;
  const int64_t* it = item + b * n_items;
  const int64_t* se = seen + b * S;
  uint8_t* kp = keep + b * S;
  const int tid = threadIdx.x;
  if (bitmap_words > 0) {
    for (int w = tid; w < bitmap_words; w += kKnnThreads) bits[w] = 0u;
    __syncthreads();
    for (int64_t i = tid; i < n_items; i += kKnnThreads) {
      const int64_t v = it[i];
      if (v < 0 || v >= I) {
        atomicExch(flag, 8);
        continue;
      }
      atomicOr(&bits[v >> 5], 1u << (v & 31));
    }
    __syncthreads();
    for (int64_t s = tid; s < S; s += kKnnThreads) {
      const int64_t v = se[s];
      if (v < 0 || v >= I) {
        atomicExch(flag, 8);
        kp[s] = 0;
        continue;
      }
      kp[s] = ((bits[v >> 5] >> (v & 31)) & 1u) ? 0 : 1;
    }
  } else {
    for (int64_t i = tid; i < n_items; i += kKnnThreads) {
      const int64_t v = it[i];
      if (v < 0 || v >= I) atomicExch(flag, 8);
    }
    for (int64_t s = tid; s < S; s += kKnnThreads) {
      const int64_t v = se[s];
      if (v < 0 || v >= I) {
        atomicExch(flag, 8);
        kp[s] = 0;
        continue;
      }
      bool hit = false;
      for (int64_t i = 0; i < n_items; ++i) hit |= (__ldg(it + i) == v);
      kp[s] = hit ? 0 : 1;
    }
  }
}

__device__ __forceinline__ int64_t clamp_id(int64_t v, int64_t I) { return (v < 0 || v >= I) ? 0 : v; }

// ItemKNN forward.  Shared: profile (H floats) + partial sums (kKnnThreads floats).
// The profile SUM_s w[seen_s] is built with thread = column (coalesced row reads); when H is below
// the block size the kept entries are split over blockDim/H thread groups and the group sums are
// combined in a fixed order, so the result does not depend on scheduling.
__global__ void __launch_bounds__(kKnnThreads) knn_forward(const float* __restrict__ w, int64_t I, int H,
                                                           const float* __restrict__ bias,
                                                           const int64_t* __restrict__ item, int64_t n_items,
                                                           const int64_t* __restrict__ seen, int64_t S,
                                                           const uint8_t* __restrict__ keep,
                                                           float* __restrict__ profile_out,
                                                           float* __restrict__ logits_out) {
  extern __shared__ float smem[];
  float* prof = smem;
  float* part = smem + H;
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t* se = seen + b * S;
  const uint8_t* kp = keep + b * S;
  const int Hc = H < kKnnThreads ? H : kKnnThreads;
  const int G = kKnnThreads / Hc;
  const int grp = tid / Hc, h0 = tid % Hc;
  for (int hb = 0; hb < H; hb += Hc) {
    const int h = hb + h0;
    float acc = 0.f;
    if (grp < G && h < H)
      for (int64_t s = grp; s < S; s += G)
        if (kp[s]) acc += __ldg(w + se[s] * H + h);
    if (grp < G) part[grp * Hc + h0] = acc;
    __syncthreads();
    if (grp == 0 && h < H) {
      float tot = 0.f;
      for (int g = 0; g < G; ++g) tot += part[g * Hc + h0];
      prof[h] = tot;
      profile_out[b * H + h] = tot;
    }
    __syncthreads();
  }
  const int warp = tid >> 5, lane = tid & 31;
  for (int64_t i = warp; i < n_items; i += kKnnThreads / 32) {
    const int64_t v = clamp_id(item[b * n_items + i], I);
    const float* row = w + v * H;
    float acc = 0.f;
    for (int h = lane; h < H; h += 32) acc += __ldg(row + h) * prof[h];
    acc = warp_sum(acc);
    if (lane == 0) logits_out[b * n_items + i] = acc + (bias != nullptr ? __ldg(bias + v) : 0.f);
  }
}

// ItemKNN backward: d w[item_i] += g_i * profile, d w[seen_s kept] += SUM_i g_i w[item_i],
// d bias[item_i] += g_i.
__global__ void __launch_bounds__(kKnnThreads) knn_backward(const float* __restrict__ w, int64_t I, int H,
                                                            const int64_t* __restrict__ item, int64_t n_items,
                                                            const int64_t* __restrict__ seen, int64_t S,
                                                            const uint8_t* __restrict__ keep,
                                                            const float* __restrict__ profile,
                                                            const float* __restrict__ grad_logits,
                                                            float* __restrict__ grad_w,
                                                            float* __restrict__ grad_bias) {
  extern __shared__ float smem[];
  float* dprof = smem;
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t* it = item + b * n_items;
  const int64_t* se = seen + b * S;
  const uint8_t* kp = keep + b * S;
  const float* g = grad_logits + b * n_items;
  for (int h = tid; h < H; h += kKnnThreads) {
    float acc = 0.f;
    for (int64_t i = 0; i < n_items; ++i) acc += g[i] * __ldg(w + clamp_id(it[i], I) * H + h);
    dprof[h] = acc;
  }
  __syncthreads();
  for (int64_t idx = tid; idx < n_items * H; idx += kKnnThreads) {
    const int64_t i = idx / H;
    const int h = (int)(idx - i * H);
    atomicAdd(grad_w + clamp_id(it[i], I) * H + h, g[i] * profile[b * H + h]);
  }
  for (int64_t idx = tid; idx < S * H; idx += kKnnThreads) {
    const int64_t s = idx / H;
    const int h = (int)(idx - s * H);
    if (kp[s]) atomicAdd(grad_w + se[s] * H + h, dprof[h]);
  }
  if (grad_bias != nullptr)
    for (int64_t i = tid; i < n_items; i += kKnnThreads) atomicAdd(grad_bias + clamp_id(it[i], I), g[i]);
}

// FreeItemKNN forward: one warp per (row, item), lanes over the seen entries.
__global__ void __launch_bounds__(kKnnThreads) freeknn_forward(const float* __restrict__ W, int64_t I,
                                                               const float* __restrict__ bias,
                                                               const int64_t* __restrict__ item, int64_t n_items,
                                                               const int64_t* __restrict__ seen, int64_t S,
                                                               const uint8_t* __restrict__ keep,
                                                               float* __restrict__ logits_out) {
  const int64_t b = blockIdx.x;
  const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
  const int64_t* se = seen + b * S;
  const uint8_t* kp = keep + b * S;
  for (int64_t i = warp; i < n_items; i += kKnnThreads / 32) {
    const int64_t v = clamp_id(item[b * n_items + i], I);
    const float* row = W + v * I;
    float acc = 0.f;
    for (int64_t s = lane; s < S; s += 32)
      if (kp[s]) acc += __ldg(row + se[s]);
    acc = warp_sum(acc);
    if (lane == 0) logits_out[b * n_items + i] = acc + (bias != nullptr ? __ldg(bias + v) : 0.f);
  }
}

__global__ void __launch_bounds__(kKnnThreads) freeknn_backward(int64_t I, const int64_t* __restrict__ item,
                                                                int64_t n_items,
                                                                const int64_t* __restrict__ seen, int64_t S,
                                                                const uint8_t* __restrict__ keep,
                                                                const float* __restrict__ grad_logits,
                                                                float* __restrict__ grad_W,
                                                                float* __restrict__ grad_bias) {
  const int64_t b = blockIdx.x;
  const int tid = threadIdx.x;
  const int64_t* it = item + b * n_items;
  const int64_t* se = seen + b * S;
  const uint8_t* kp = keep + b * S;
  const float* g = grad_logits + b * n_items;
  for (int64_t idx = tid; idx < n_items * S; idx += kKnnThreads) {
    const int64_t i = idx / S, s = idx - i * S;
    if (kp[s]) atomicAdd(grad_W + clamp_id(it[i], I) * I + se[s], g[i]);
  }
  if (grad_bias != nullptr)
    for (int64_t i = tid; i < n_items; i += kKnnThreads) atomicAdd(grad_bias + clamp_id(it[i], I), g[i]);
}

int launch_keep(rbpr_ctx* ctx, const int64_t* item, int64_t B, int64_t n_items, const int64_t* seen,
                int64_t S, int64_t I, uint8_t* keep, cudaStream_t st) {
  if (S == 0) return 0;
  int words = 0;
  if (n_items > kKnnScanLimit) {
    const int64_t need = (I + 31) / 32;
    if (need * 4 <= 200 * 1024) words = (int)need;  // else: the scan, slow but correct
  }
  if (words * 4 > 48 * 1024)
    RBPR_CUDA(ctx, cudaFuncSetAttribute(knn_keep, cudaFuncAttributeMaxDynamicSharedMemorySize, words * 4));
  knn_keep<<<(unsigned)B, kKnnThreads, (size_t)words * 4, st>>>(item, n_items, seen, S, I, words, keep,
                                                                ctx->flag);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int check_sizes(rbpr_ctx* ctx, const char* who, int64_t I, int64_t B, int64_t n_items, int64_t S) {
  if (I < 1 || B < 0 || n_items < 1 || S < 0) RBPR_FAIL(ctx, RBPR_ERR_ARG, "%s: bad sizes", who);
  if (B >= (1ll << 31)) RBPR_FAIL(ctx, RBPR_ERR_ARG, "%s: batch must be < 2^31", who);
  return 0;
}

}  // namespace

extern "C" {

int rbpr_knn_forward(rbpr_ctx* ctx, const float* weights, int64_t num_items, int32_t hidden,
                     const float* bias, const int64_t* item, int64_t batch, int64_t n_items,
                     const int64_t* seen, int64_t width, uint8_t* keep_out, float* profile_out,
                     float* logits_out, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (int rc = check_sizes(ctx, "knn_forward", num_items, batch, n_items, width)) return rc;
  if (hidden < 1 || hidden > 8192) RBPR_FAIL(ctx, RBPR_ERR_ARG, "knn_forward: hidden_dim must be in [1, 8192]");
  if (batch == 0) return 0;
  if (!weights || !item || (width > 0 && (!seen || !keep_out)) || !profile_out || !logits_out)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "knn_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = launch_keep(ctx, item, batch, n_items, seen, width, num_items, keep_out, st)) return rc;
  const size_t smem = ((size_t)hidden + kKnnThreads) * sizeof(float);
  knn_forward<<<(unsigned)batch, kKnnThreads, smem, st>>>(weights, num_items, hidden, bias, item, n_items,
                                                          seen, width, keep_out, profile_out, logits_out);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int rbpr_knn_backward(rbpr_ctx* ctx, const float* weights, int64_t num_items, int32_t hidden,
                      const int64_t* item, int64_t batch, int64_t n_items, const int64_t* seen,
                      int64_t width, const uint8_t* keep, const float* profile,
                      const float* grad_logits, float* grad_weights, float* grad_bias, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (int rc = check_sizes(ctx, "knn_backward", num_items, batch, n_items, width)) return rc;
  if (hidden < 1 || hidden > 8192) RBPR_FAIL(ctx, RBPR_ERR_ARG, "knn_backward: hidden_dim must be in [1, 8192]");
  if (batch == 0) return 0;
  if (!weights || !item || (width > 0 && (!seen || !keep)) || !profile || !grad_logits || !grad_weights)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "knn_backward: null pointer");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  knn_backward<<<(unsigned)batch, kKnnThreads, (size_t)hidden * sizeof(float), (cudaStream_t)stream>>>(
      weights, num_items, hidden, item, n_items, seen, width, keep, profile, grad_logits, grad_weights,
      grad_bias);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int rbpr_freeknn_forward(rbpr_ctx* ctx, const float* weights, int64_t num_items, const float* bias,
                         const int64_t* item, int64_t batch, int64_t n_items, const int64_t* seen,
                         int64_t width, uint8_t* keep_out, float* logits_out, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (int rc = check_sizes(ctx, "freeknn_forward", num_items, batch, n_items, width)) return rc;
  if (batch == 0) return 0;
  if (!weights || !item || (width > 0 && (!seen || !keep_out)) || !logits_out)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "freeknn_forward: null pointer");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  if (int rc = launch_keep(ctx, item, batch, n_items, seen, width, num_items, keep_out, st)) return rc;
  freeknn_forward<<<(unsigned)batch, kKnnThreads, 0, st>>>(weights, num_items, bias, item, n_items, seen,
                                                           width, keep_out, logits_out);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int rbpr_freeknn_backward(rbpr_ctx* ctx, int64_t num_items, const int64_t* item, int64_t batch,
                          int64_t n_items, const int64_t* seen, int64_t width, const uint8_t* keep,
                          const float* grad_logits, float* grad_weights, float* grad_bias,
                          void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (int rc = check_sizes(ctx, "freeknn_backward", num_items, batch, n_items, width)) return rc;
  if (batch == 0) return 0;
  if (!item || (width > 0 && (!seen || !keep)) || !grad_logits || !grad_weights)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "freeknn_backward: null pointer");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  freeknn_backward<<<(unsigned)batch, kKnnThreads, 0, (cudaStream_t)stream>>>(
      num_items, item, n_items, seen, width, keep, grad_logits, grad_weights, grad_bias);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // extern "C"
