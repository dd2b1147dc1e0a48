// Entry points shaped like the reference's Python call sites (batches of explicit ids handed over
// by a DataLoader) — the drop-in side of the C ABI.  The fast path (rbpr_train_steps) takes triple
// ids and samples on the device; these take what `Model.forward(batch)`, `Sampler.sample(batch)`
// and eval-mode `MF.forward(user, item)` receive.  See include/rbpr.h for the call sites replaced.
#include "train_kernels.cuh"

using namespace rbpr_dev;

// defined in train.cu
int rbpr_internal_phase_a_apply(rbpr_ctx* ctx, const rbpr_hparams* hp, const int4* records, int n,
                                uint64_t step, float2* logit_out, const int32_t* step_pos,
                                double* stats_out, cudaStream_t st);
int rbpr_internal_reserve_sort(rbpr_ctx* ctx, int64_t n);
int rbpr_internal_check_ready_tables(rbpr_ctx* ctx, const rbpr_hparams* hp);

namespace {

__global__ void explicit_count(const int64_t* __restrict__ users, int64_t n, int64_t U,
                               uint32_t* __restrict__ cnt, uint32_t* __restrict__ ord,
                               int32_t* __restrict__ flag) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t u = users[k];
  if (u < 0 || u >= U) {
    atomicExch(flag, 7);
    u = 0;
  }
  ord[k] = atomicAdd(cnt + u, 1u);
}

__global__ void explicit_records(const int64_t* __restrict__ users, const int64_t* __restrict__ items,
                                 const int64_t* __restrict__ negs, const uint32_t* __restrict__ cnt,
                                 const uint32_t* __restrict__ ord, int64_t n, int64_t U, int64_t I,
                                 int4* __restrict__ records, int32_t* __restrict__ flag) {
  const int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (k >= n) return;
  int64_t u = users[k], i = items[k], j = negs[k];
  if (u < 0 || u >= U) u = 0;  // flagged by explicit_count
  if (i < 0 || i >= I || j < 0 || j >= I) {
    atomicExch(flag, 8);
    i = j = 0;
  }
  const uint32_t total = cnt[u];
  const int32_t flags = total <= 1u ? (kRecHead | kRecSingle) : (ord[k] == 0u ? (kRecHead | kRecMultiHead) : 0);
  records[k] = make_int4((int32_t)u, (int32_t)i, (int32_t)j, flags);
}

// logits[b, k] = <U[users[b]], V[items[b, k]]> (+ item_bias[item]) (+ user_bias[user]); entries
// whose mask is 0 become -1e13 (Model.forward eval branch).  8 lanes per pair.
__global__ void __launch_bounds__(256)
pair_logits(const float* __restrict__ user_emb, const float* __restrict__ item_emb,
            const float* __restrict__ item_bias, const float* __restrict__ user_bias,
            const int64_t* __restrict__ users, const int64_t* __restrict__ items,
            const float* __restrict__ mask, int64_t n_users, int64_t per_user, int64_t U, int64_t I,
            int D, float* __restrict__ out, int32_t* __restrict__ flag) {
  const Group<8> g;
  const int64_t total = n_users * per_user;
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / 8;
  for (int64_t q = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8; q < total; q += groups) {
    int64_t u = users[q / per_user], it = items[q];
    if (u < 0 || u >= U || it < 0 || it >= I) {
      if (g.gl == 0) atomicExch(flag, 8);
      u = 0;
      it = 0;
    }
    const float* ur = user_emb + u * D;
    const float* ir = item_emb + it * D;
    float acc = 0.f;
    for (int c = 4 * g.gl; c < D; c += 32) acc += dot4(ld4(ur + c), ld4(ir + c));
    acc = g.sum(acc);
    if (g.gl == 0) {
      if (item_bias != nullptr) acc += item_bias[it];
      if (user_bias != nullptr) acc += user_bias[u];
      if (mask != nullptr && mask[q] == 0.f) acc = -1e13f;
      out[q] = acc;
    }
  }
}

// Negative sampling against the reference's padded seen matrix (B,S) int64, 0-padded, rows in any
// order.  Same counter-based draw as the CSR sampler with the batch row as the
// Philox subsequence; the membership test is a cooperative linear scan of the row.
__global__ void __launch_bounds__(256)
sample_padded(const int64_t* __restrict__ seen, int64_t B, int64_t S, uint32_t I,
              const float* __restrict__ alias_prob, const int32_t* __restrict__ alias_idx,
              int sampler, uint32_t seed_lo, uint32_t seed_hi, uint64_t step, int64_t num,
              int64_t* __restrict__ out, int32_t* __restrict__ flag) {
  const Group<8> g;
  const int64_t slot = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / 8;
  if (slot >= B * num) return;
  const int64_t row = slot / num;
  const int64_t* srow = seen + row * S;
  const uint32_t n = (sampler == RBPR_SAMPLER_WEIGHTED) ? I : I - 1u;
  const uint32_t thresh = (uint32_t)((1ull << 32) % (uint64_t)n);
  const uint32_t step_lo = (uint32_t)(step << 8), step_hi = (uint32_t)(step >> 24);
  auto contains = [&](int32_t j) -> bool {
    bool hit = false;
    for (int64_t c = g.gl; c < S; c += 8) hit |= (srow[c] == (int64_t)j);
    return __ballot_sync(g.mask, hit) != 0u;
  };
  int32_t res = -1;
  for (uint32_t blk = 0; blk < 256u && res < 0; ++blk) {
    const philox4 r = philox4x32_10(step_lo | blk, step_hi, (uint32_t)slot, (uint32_t)((uint64_t)slot >> 32),
                                    seed_lo, seed_hi);
    const uint32_t w[4] = {r.x, r.y, r.z, r.w};
    if (sampler == RBPR_SAMPLER_UNIFORM) {
#pragma unroll
      for (int a = 0; a < 4; ++a) {
        if (res >= 0) break;
        const uint32_t mlo = w[a] * n, mhi = __umulhi(w[a], n);
        if (mlo < thresh) continue;
        const int32_t j = 1 + (int32_t)mhi;
        if (!contains(j)) res = j;
      }
    } else {
#pragma unroll
      for (int a = 0; a < 2; ++a) {
        if (res >= 0) break;
        const uint32_t mlo = w[2 * a] * n, mhi = __umulhi(w[2 * a], n);
        if (mlo < thresh) continue;
        const int32_t col = (int32_t)mhi;
        const float uf = (float)(w[2 * a + 1] >> 8) * (1.0f / 16777216.0f);
        const int32_t j = (uf < __ldg(alias_prob + col)) ? col : __ldg(alias_idx + col);
        if (j == 0) continue;
        if (!contains(j)) res = j;
      }
    }
  }
  if (res < 0) {
    if (g.gl == 0) atomicExch(flag, 1);
    res = 1;
  }
  if (g.gl == 0) out[slot] = (int64_t)res;
}

// logits[b, seen[b, c]] = -1e13 for every entry of the padded seen matrix; logits[b, 0] = -1e13
// (BPRExperiment._remove_seen_items, experiments/bpr/exp.py:369-374).  Padding entries are 0.
__global__ void mask_seen_padded(float* __restrict__ logits, const int64_t* __restrict__ seen,
                                 int64_t B, int64_t S, int64_t I, int32_t* __restrict__ flag) {
  const int64_t q = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (q >= B * (S + 1)) return;
  const int64_t row = q / (S + 1), c = q % (S + 1);
  int64_t it = (c == S) ? 0 : seen[row * S + c];
  if (it < 0 || it >= I) {
    atomicExch(flag, 8);
    return;
  }
  logits[row * I + it] = -1e13f;
}

// RocAucManySlow (revisit_bpr/metrics/auc.py:149-166): per row, the fraction of (positive, negative)
// pairs with score[pos] > score[neg]; positives = target != 0, negatives = target == 0 and
// mask != 0.  One CTA per row: positives staged in shared memory in chunks, one pass over the row
// per chunk counting, for every negative, the positives that beat it.  0/0 -> NaN like the reference.
__global__ void __launch_bounds__(256)
auc_rows(const float* __restrict__ scores, const float* __restrict__ target,
         const float* __restrict__ mask, int64_t I, float* __restrict__ out) {
  constexpr int CH = 1024;
  __shared__ float pos_s[CH];
  __shared__ int s_npos_chunk;
  __shared__ unsigned long long s_wins, s_npos, s_nneg;
  const int64_t row = blockIdx.x;
  const float* sr = scores + row * I;
  const float* tr = target + row * I;
  const float* mr = mask ? mask + row * I : nullptr;
  if (threadIdx.x == 0) {
    s_wins = 0ull;
    s_npos = 0ull;
    s_nneg = 0ull;
  }
  __syncthreads();
  // count negatives once
  unsigned long long nneg = 0;
  for (int64_t i = threadIdx.x; i < I; i += blockDim.x)
    nneg += (tr[i] == 0.f && (mr == nullptr || mr[i] != 0.f));
  atomicAdd(&s_nneg, nneg);
  int64_t next = 0;  // first column not yet scanned for positives
  while (next < I) {
    if (threadIdx.x == 0) s_npos_chunk = 0;
    __syncthreads();
    // gather up to CH positives from columns [next, ...): sequential chunking by column blocks
    int64_t base = next;
    for (; base < I; base += blockDim.x) {
      const int64_t i = base + threadIdx.x;
      const bool is_pos = i < I && tr[i] != 0.f;
      if (is_pos) {
        const int slot = atomicAdd(&s_npos_chunk, 1);
        if (slot < CH) pos_s[slot] = sr[i];
      }
      __syncthreads();
      const int got = s_npos_chunk;
      __syncthreads();
      if (got > CH - (int)blockDim.x) {  // the next block of columns might overflow: stop here
        base += blockDim.x;
        break;
      }
    }
    next = base;
    const int np = min(s_npos_chunk, CH);
    unsigned long long wins = 0;
    for (int64_t i = threadIdx.x; i < I; i += blockDim.x) {
      if (!(tr[i] == 0.f && (mr == nullptr || mr[i] != 0.f))) continue;
      const float v = sr[i];
      int w = 0;
      for (int q = 0; q < np; ++q) w += (pos_s[q] > v);
      wins += (unsigned long long)w;
    }
    atomicAdd(&s_wins, wins);
    if (threadIdx.x == 0) s_npos += (unsigned long long)np;
    __syncthreads();
  }
  if (threadIdx.x == 0) out[row] = (float)((double)s_wins / ((double)s_npos * (double)s_nneg));
}

}  // namespace

extern "C" {

int rbpr_train_step_triples(rbpr_ctx* ctx, const int64_t* users, const int64_t* items,
                            const int64_t* negs, int64_t n, uint64_t step, const rbpr_hparams* hp,
                            float* logits_out, double* stats_out, void* stream) {
  int rc = rbpr_internal_check_ready_tables(ctx, hp);
  if (rc) return rc;
  if (n < 0 || n >= (1ll << 31)) RBPR_FAIL(ctx, RBPR_ERR_ARG, "train_step_triples: bad n");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  if (n > 0 && (!users || !items || !negs))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "train_step_triples: null id array");
  rc = rbpr_internal_reserve_sort(ctx, n > 0 ? n : 1);
  if (rc) return rc;
  if (n > 0) {
    const int threads = 256, blocks = (int)((n + threads - 1) / threads);
    RBPR_CUDA(ctx, cudaMemsetAsync(ctx->cnt, 0, (size_t)ctx->U * sizeof(uint32_t), st));
    explicit_count<<<blocks, threads, 0, st>>>(users, n, ctx->U, ctx->cnt, ctx->ord, ctx->flag);
    explicit_records<<<blocks, threads, 0, st>>>(users, items, negs, ctx->cnt, ctx->ord, n, ctx->U,
                                                 ctx->I, reinterpret_cast<int4*>(ctx->records[0]),
                                                 ctx->flag);
    ctx->launches += 2;
    RBPR_CUDA(ctx, cudaGetLastError());
  }
  return rbpr_internal_phase_a_apply(ctx, hp, reinterpret_cast<const int4*>(ctx->records[0]), (int)n,
                                     step, reinterpret_cast<float2*>(logits_out), nullptr,
                                     stats_out, st);
}

int rbpr_pair_logits(rbpr_ctx* ctx, const int64_t* users, const int64_t* items, const float* mask,
                     int64_t n_users, int64_t per_user, const float* user_bias, float* out,
                     void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!ctx->user_emb || !ctx->item_emb) RBPR_FAIL(ctx, RBPR_ERR_STATE, "tables not bound");
  if (n_users < 0 || per_user < 0) RBPR_FAIL(ctx, RBPR_ERR_ARG, "pair_logits: negative size");
  if (n_users == 0 || per_user == 0) return 0;
  if (!users || !items || !out) RBPR_FAIL(ctx, RBPR_ERR_ARG, "pair_logits: null pointer");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t total = n_users * per_user;
  int64_t blocks = (total * 8 + 255) / 256;
  const int64_t maxb = (int64_t)ctx->sm_count * 16;
  if (blocks > maxb) blocks = maxb;
  pair_logits<<<(int)blocks, 256, 0, (cudaStream_t)stream>>>(
      ctx->user_emb, ctx->item_emb, ctx->item_bias, user_bias, users, items, mask, n_users, per_user,
      ctx->U, ctx->I, ctx->D, out, ctx->flag);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int rbpr_sample_negatives_padded(rbpr_ctx* ctx, const int64_t* seen, int64_t batch, int64_t width,
                                 int64_t num_items, int64_t num, uint64_t seed, uint64_t step,
                                 int32_t sampler, int64_t* neg_out, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (batch < 0 || width < 0 || num < 1) RBPR_FAIL(ctx, RBPR_ERR_ARG, "sample_padded: bad sizes");
  if (num_items < 3 || num_items >= (1ll << 31))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "sample_padded: num_items must be in [3, 2^31)");
  if (sampler != RBPR_SAMPLER_UNIFORM && sampler != RBPR_SAMPLER_WEIGHTED)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "sample_padded: sampler must draw");
  if (sampler == RBPR_SAMPLER_WEIGHTED && (!ctx->alias_prob || !ctx->alias_idx))
    RBPR_FAIL(ctx, RBPR_ERR_STATE, "weighted sampler requested but alias table not bound");
  if (batch == 0) return 0;
  if ((width > 0 && !seen) || !neg_out) RBPR_FAIL(ctx, RBPR_ERR_ARG, "sample_padded: null pointer");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t threads = batch * num * 8;
  sample_padded<<<(int)((threads + 255) / 256), 256, 0, (cudaStream_t)stream>>>(
      seen, batch, width, (uint32_t)num_items, ctx->alias_prob, ctx->alias_idx, sampler,
      (uint32_t)seed, (uint32_t)(seed >> 32), step, num, neg_out, ctx->flag);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int rbpr_mask_seen_padded(rbpr_ctx* ctx, float* logits, const int64_t* seen, int64_t batch,
                          int64_t width, int64_t n_cols, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (batch < 0 || width < 0 || n_cols < 1) RBPR_FAIL(ctx, RBPR_ERR_ARG, "mask_seen_padded: bad sizes");
  if (batch == 0) return 0;
  if (!logits || (width > 0 && !seen)) RBPR_FAIL(ctx, RBPR_ERR_ARG, "mask_seen_padded: null pointer");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t total = batch * (width + 1);
  mask_seen_padded<<<(int)((total + 255) / 256), 256, 0, (cudaStream_t)stream>>>(logits, seen, batch, width,
                                                                             n_cols, ctx->flag);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

int rbpr_auc_dense(rbpr_ctx* ctx, const float* scores, const float* target, const float* mask,
                   int64_t n_rows, int64_t n_cols, float* auc_out, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (n_rows < 0 || n_cols < 1) RBPR_FAIL(ctx, RBPR_ERR_ARG, "auc_dense: bad sizes");
  if (n_rows == 0) return 0;
  if (!scores || !target || !auc_out) RBPR_FAIL(ctx, RBPR_ERR_ARG, "auc_dense: null pointer");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  auc_rows<<<(unsigned)n_rows, 256, 0, (cudaStream_t)stream>>>(scores, target, mask, n_cols, auc_out);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // extern "C"
