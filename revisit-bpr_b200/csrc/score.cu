// Full-catalog scoring + top-k + ranking metrics for sm_100a.
//
//   score_gemm   : S[u, i] = <U[users[u]], V[i]> (+ item_bias[i]) for a block of users, strict
//                  fp32 FMA (no TF32: ranks must match the reference's fp32 einsum), 128x128x8
//                  register-tiled CTA, column 0 (padding item) written as -1e13.
//   mask_seen    : S[u, seen(u)] = -1e13                     (exp.py:369-374)
//   topk_metrics : one CTA per user: 4-pass 8-bit radix select of the k-th largest score,
//                  gather of the top-k, bitonic sort by (score desc, item asc), hit flags against
//                  the user's held-out row, NDCG@k / Recall@k for every requested cut-off.
// Reference call sites replaced: see include/rbpr.h (rbpr_score_topk / rbpr_score_dense).
#include "score_common.cuh"

namespace {

constexpr int BM = 128, BN = 128, BK = 8;

__global__ void __launch_bounds__(256)
score_gemm(const float* __restrict__ user_emb, const float* __restrict__ item_emb,
           const float* __restrict__ item_bias, const int64_t* __restrict__ users, int n_users,
           int I, int D, float* __restrict__ S, int64_t ld) {
  __shared__ __align__(16) float As[2][BK][BM];
  __shared__ __align__(16) float Bs[2][BK][BN];
  const int tid = threadIdx.x;
  const int tx = tid & 15, ty = tid >> 4;
  const int m0 = blockIdx.y * BM, n0 = blockIdx.x * BN;

  // loader mapping: one float4 of A and one of B per thread per k-tile
  const int lrow = tid >> 1, lk = (tid & 1) * 4;
  const int arow = m0 + lrow, brow = n0 + lrow;
  const float* aptr = nullptr;
  const float* bptr = nullptr;
  if (arow < n_users) aptr = user_emb + users[arow] * (int64_t)D;
  if (brow < I) bptr = item_emb + (int64_t)brow * D;

  float acc[8][8];
#pragma unroll
  for (int a = 0; a < 8; ++a)
#pragma unroll
    for (int b = 0; b < 8; ++b) acc[a][b] = 0.f;

  auto gload = [&](int k0, float4& a, float4& b) {
    a = make_float4(0.f, 0.f, 0.f, 0.f);
    b = a;
    if (k0 + lk < D) {
      if (aptr) a = *reinterpret_cast<const float4*>(aptr + k0 + lk);
      if (bptr) b = *reinterpret_cast<const float4*>(bptr + k0 + lk);
    }
  };
  auto sstore = [&](int buf, const float4& a, const float4& b) {
    As[buf][lk + 0][lrow] = a.x;
    As[buf][lk + 1][lrow] = a.y;
    As[buf][lk + 2][lrow] = a.z;
    As[buf][lk + 3][lrow] = a.w;
    Bs[buf][lk + 0][lrow] = b.x;
    Bs[buf][lk + 1][lrow] = b.y;
    Bs[buf][lk + 2][lrow] = b.z;
    Bs[buf][lk + 3][lrow] = b.w;
  };

  float4 ra, rb;
  gload(0, ra, rb);
  sstore(0, ra, rb);
  __syncthreads();
  const int ktiles = (D + BK - 1) / BK;
  for (int kt = 0; kt < ktiles; ++kt) {
    const int buf = kt & 1;
    if (kt + 1 < ktiles) gload((kt + 1) * BK, ra, rb);
#pragma unroll
    for (int k = 0; k < BK; ++k) {
      float4 a0 = *reinterpret_cast<const float4*>(&As[buf][k][ty * 4]);
      float4 a1 = *reinterpret_cast<const float4*>(&As[buf][k][64 + ty * 4]);
      float4 b0 = *reinterpret_cast<const float4*>(&Bs[buf][k][tx * 4]);
      float4 b1 = *reinterpret_cast<const float4*>(&Bs[buf][k][64 + tx * 4]);
      const float av[8] = {a0.x, a0.y, a0.z, a0.w, a1.x, a1.y, a1.z, a1.w};
      const float bv[8] = {b0.x, b0.y, b0.z, b0.w, b1.x, b1.y, b1.z, b1.w};
#pragma unroll
      for (int a = 0; a < 8; ++a)
#pragma unroll
        for (int b = 0; b < 8; ++b) acc[a][b] = fmaf(av[a], bv[b], acc[a][b]);
    }
    if (kt + 1 < ktiles) {
      sstore(buf ^ 1, ra, rb);
      __syncthreads();
    }
  }

  const bool vec_ok = (ld % 4) == 0;
#pragma unroll
  for (int a = 0; a < 8; ++a) {
    const int row = m0 + ((a < 4) ? ty * 4 + a : 64 + ty * 4 + (a - 4));
    if (row >= n_users) continue;
#pragma unroll
    for (int h = 0; h < 2; ++h) {
      const int col = n0 + h * 64 + tx * 4;
      if (col >= I) continue;
      float o[4];
#pragma unroll
      for (int q = 0; q < 4; ++q) {
        const int c = col + q;
        float v = acc[a][h * 4 + q];
        if (item_bias != nullptr && c < I) v += __ldg(item_bias + c);
        if (c == 0) v = kMasked;
        o[q] = v;
      }
      float* dst = S + (int64_t)row * ld + col;
      if (vec_ok && col + 3 < I) {
        *reinterpret_cast<float4*>(dst) = make_float4(o[0], o[1], o[2], o[3]);
      } else {
#pragma unroll
        for (int q = 0; q < 4; ++q)
          if (col + q < I) dst[q] = o[q];
      }
    }
  }
}

__global__ void mask_seen(const int64_t* __restrict__ indptr, const int32_t* __restrict__ indices,
                          int64_t row0, const int32_t* __restrict__ row_map, int n_users, int I, float* __restrict__ S,
                          int64_t ld) {
  const int64_t w = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) >> 5;
  const int lane = threadIdx.x & 31;
  if (w >= n_users) return;
  const int64_t r = row0 + (row_map != nullptr ? (int64_t)row_map[w] : w);
  const int64_t lo = indptr[r], hi = indptr[r + 1];
  for (int64_t q = lo + lane; q < hi; q += 32) {
    const int32_t it = indices[q];
    if (it >= 0 && it < I) S[w * ld + it] = kMasked;
  }
}

__global__ void __launch_bounds__(256) topk_metrics(const TopkParams p) {
  __shared__ uint32_t hist[256];
  __shared__ uint32_t s_prefix, s_need, s_cnt, s_tie_cnt;
  __shared__ unsigned long long sel[KCAP];
  __shared__ uint32_t warp_tot[8];

  const int tid = threadIdx.x;
  const int64_t urow = blockIdx.x;
  const float* row = p.S + urow * p.ld;
  const int I = p.I;
  const int k = min(p.k_max, I);

  // ---- radix select: find key T of the k-th largest element ----
  uint32_t prefix = 0, prefix_mask = 0;
  uint32_t need = (uint32_t)k;  // rank (from the top) still to locate within the prefix class
  for (int pass = 0; pass < 4; ++pass) {
    const int shift = 24 - 8 * pass;
    hist[tid] = 0;
    __syncthreads();
    for (int i = tid; i < I; i += 256) {
      const uint32_t key = fkey(row[i]);
      if ((key & prefix_mask) == prefix) atomicAdd(&hist[(key >> shift) & 255u], 1u);
    }
    __syncthreads();
    if (tid == 0) {
      uint32_t acc = 0;
      int b = 255;
      for (; b > 0; --b) {
        if (acc + hist[b] >= need) break;
        acc += hist[b];
      }
      s_prefix = prefix | ((uint32_t)b << shift);
      s_need = need - acc;
    }
    __syncthreads();
    prefix = s_prefix;
    need = s_need;
    prefix_mask |= 255u << shift;
    __syncthreads();
  }
  const uint32_t T = prefix;       // key of the k-th largest
  const uint32_t need_ties = need; // how many elements == T belong to the top-k

  // ---- collect ----
  if (tid == 0) {
    s_cnt = 0;
    s_tie_cnt = 0;
  }
  for (int i = tid; i < KCAP; i += 256) sel[i] = 0ull;
  __syncthreads();
  for (int i = tid; i < I; i += 256) {
    const uint32_t key = fkey(row[i]);
    if (key > T) {
      const uint32_t slot = atomicAdd(&s_cnt, 1u);
      sel[slot] = ((unsigned long long)key << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
    }
  }
  __syncthreads();
  const uint32_t n_gt = s_cnt;  // == k - need_ties
  // ties in ascending item order (deterministic): ordered block scan over the row
  for (int base = 0; base < I && s_tie_cnt < need_ties; base += 256) {
    const int i = base + tid;
    const bool is_tie = (i < I) && (fkey(row[i]) == T);
    const unsigned bal = __ballot_sync(0xffffffffu, is_tie);
    const int warp = tid >> 5, lane = tid & 31;
    if (lane == 0) warp_tot[warp] = __popc(bal);
    __syncthreads();
    uint32_t before = s_tie_cnt;
    for (int w = 0; w < warp; ++w) before += warp_tot[w];
    before += __popc(bal & ((1u << lane) - 1u));
    if (is_tie && before < need_ties)
      sel[n_gt + before] =
          ((unsigned long long)T << 32) | (unsigned long long)(0xFFFFFFFFu - (uint32_t)i);
    __syncthreads();
    if (tid == 0) {
      uint32_t tot = 0;
      for (int w = 0; w < 8; ++w) tot += warp_tot[w];
      s_tie_cnt += tot;
    }
    __syncthreads();
  }

  // ---- bitonic sort of 128 composite keys, descending ----
  for (int size = 2; size <= KCAP; size <<= 1) {
    for (int stride = size >> 1; stride > 0; stride >>= 1) {
      if (tid < KCAP / 2) {
        const int lo = 2 * tid - (tid & (stride - 1));
        const int hi = lo + stride;
        const bool desc = ((lo & size) == 0);
        const unsigned long long a = sel[lo], b = sel[hi];
        if ((a < b) == desc) {
          sel[lo] = b;
          sel[hi] = a;
        }
      }
      __syncthreads();
    }
  }

  // ---- outputs ----
  __syncthreads();
  topk_emit_outputs(p, urow, sel, k);
}

int score_block(rbpr_ctx* ctx, const int64_t* users, int n_users, const int64_t* seen_indptr,
                const int32_t* seen_indices, int64_t row0, float* S, int64_t ld,
                cudaStream_t st, const int32_t* row_map = nullptr) {
  NvtxRange nvtx("rbpr.score_block (dense fp32)");
  dim3 grid((unsigned)((ctx->I + BN - 1) / BN), (unsigned)((n_users + BM - 1) / BM));
  score_gemm<<<grid, 256, 0, st>>>(ctx->user_emb, ctx->item_emb, ctx->item_bias, users, n_users,
                                   (int)ctx->I, ctx->D, S, ld);
  ctx->launches++;
  if (seen_indptr != nullptr) {
    const int64_t threads = (int64_t)n_users * 32;
    mask_seen<<<(int)((threads + 255) / 256), 256, 0, st>>>(seen_indptr, seen_indices, row0, row_map,
                                                            n_users, (int)ctx->I, S, ld);
    ctx->launches++;
  }
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // namespace

// defined in score_tc.cu
bool rbpr_score_tc_eligible(const rbpr_ctx* ctx, int k_max);
int rbpr_score_tc_block(rbpr_ctx* ctx, const int64_t* users, int n_users, const int64_t* seen_indptr,
                        const int32_t* seen_indices, int64_t row0, const TopkParams& tp_in, int* overflow_host,
                        cudaStream_t st);
constexpr int64_t kTcBlock = 16384;  // users per pass of the tensor path (bounds its scratch)

static __global__ void gather_users(const int64_t* __restrict__ users, const int32_t* __restrict__ rows, int n,
                                    int64_t* __restrict__ out) {
  const int k = blockIdx.x * blockDim.x + threadIdx.x;
  if (k < n) out[k] = users[rows[k]];
}

extern "C" {

int rbpr_score_dense(rbpr_ctx* ctx, const int64_t* users, int64_t n_users,
                     const int64_t* seen_indptr, const int32_t* seen_indices, float* out,
                     void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!ctx->user_emb || !ctx->item_emb) RBPR_FAIL(ctx, RBPR_ERR_STATE, "tables not bound");
  if (n_users == 0) return 0;
  if (!users || !out || n_users < 0 || n_users >= (1ll << 31))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_dense: bad arguments");
  if ((seen_indptr == nullptr) != (seen_indices == nullptr))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_dense: seen CSR must be both set or both NULL");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  return score_block(ctx, users, (int)n_users, seen_indptr, seen_indices, 0, out, ctx->I,
                     (cudaStream_t)stream);
}

// Shared body of rbpr_score_topk / rbpr_score_metrics.
static int score_topk_impl(rbpr_ctx* ctx, const int64_t* users, int64_t n_users,
                           const int64_t* seen_indptr, const int32_t* seen_indices,
                           const int64_t* held_indptr, const int32_t* held_indices, int32_t k_max,
                           const int32_t* ks, int32_t n_ks, int32_t* topk_items, float* topk_scores,
                           const rbpr_metric_outputs* mo, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!ctx->user_emb || !ctx->item_emb) RBPR_FAIL(ctx, RBPR_ERR_STATE, "tables not bound");
  if (n_users == 0) return 0;
  if (!users || n_users < 0 || n_users >= (1ll << 31))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_topk: bad users");
  if (k_max < 1 || k_max > RBPR_MAX_TOPK)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_topk: k_max=%d outside [1,%d]", k_max, RBPR_MAX_TOPK);
  if (n_ks < 0 || n_ks > 16 || (n_ks > 0 && !ks))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_topk: at most 16 cut-offs");
  for (int q = 0; q < n_ks; ++q)
    if (ks[q] < 1 || ks[q] > k_max)
      RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_topk: cut-off %d outside [1,k_max=%d]", ks[q], k_max);
  if ((seen_indptr == nullptr) != (seen_indices == nullptr))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_topk: seen CSR must be both set or both NULL");
  if ((held_indptr == nullptr) != (held_indices == nullptr))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_topk: held-out CSR must be both set or both NULL");
  const bool any_metric = mo && (mo->ndcg || mo->ndcg_linear || mo->recall || mo->precision || mo->map);
  if (any_metric && (!held_indptr || n_ks == 0))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_topk: metrics need the held-out CSR and cut-offs");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  const int64_t ld = (ctx->I + 3) & ~3ll;
  const bool tc = rbpr_score_tc_eligible(ctx, k_max);
  // user block sized so the score buffer stays <= 512 MB (tensor path: only overflow users need it)
  int64_t blk = ((tc ? 64ll : 512ll) << 20) / (ld * (int64_t)sizeof(float));
  blk = (blk / BM) * BM;
  if (blk < BM) blk = BM;
  if (blk > n_users) blk = n_users;
  const size_t need = (size_t)blk * ld * sizeof(float);
  if (need > ctx->score_buf_bytes) {
    cudaFree(ctx->score_buf);
    ctx->score_buf = nullptr;
    ctx->score_buf_bytes = 0;
    RBPR_CUDA(ctx, cudaMalloc(&ctx->score_buf, need));
    ctx->score_buf_bytes = need;
  }
  TopkParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.S = ctx->score_buf;
  tp.ld = ld;
  tp.I = (int)ctx->I;
  tp.k_max = k_max;
  tp.held_indptr = held_indptr;
  tp.held_indices = held_indices;
  tp.n_ks = n_ks;
  for (int q = 0; q < n_ks; ++q) tp.ks[q] = ks[q];
  tp.topk_items = topk_items;
  tp.topk_scores = topk_scores;
  if (mo) {
    tp.ndcg_out = mo->ndcg;
    tp.ndcg_linear_out = mo->ndcg_linear;
    tp.recall_out = mo->recall;
    tp.precision_out = mo->precision;
    tp.map_out = mo->map;
    tp.map_normalized = mo->map_normalized;
  }
  if (tc) {
    // tensor-core candidate filter + exact fp32 rescoring (score_tc.cu); users whose candidate list
    // overflowed (mass ties) are re-done by the dense path below
    for (int64_t r0 = 0; r0 < n_users; r0 += kTcBlock) {
      const int nb = (int)((n_users - r0) < kTcBlock ? (n_users - r0) : kTcBlock);
      int overflow = 0;
      int rc = rbpr_score_tc_block(ctx, users + r0, nb, seen_indptr, seen_indices, r0, tp, &overflow, st);
      if (rc) return rc;
      if (overflow > 0) {
        ctx->tc_overflow_users += overflow;
        if ((size_t)overflow * sizeof(int64_t) > ctx->tc_ovf_users_bytes) {
          cudaFree(ctx->tc_ovf_users);
          ctx->tc_ovf_users = nullptr;
          ctx->tc_ovf_users_bytes = 0;
          RBPR_CUDA(ctx, cudaMalloc(&ctx->tc_ovf_users, (size_t)nb * sizeof(int64_t)));
          ctx->tc_ovf_users_bytes = (size_t)nb * sizeof(int64_t);
        }
        gather_users<<<(overflow + 255) / 256, 256, 0, st>>>(users + r0, ctx->tc_overflow_rows, overflow, ctx->tc_ovf_users);
        ctx->launches++;
        for (int o0 = 0; o0 < overflow; o0 += (int)blk) {
          const int ob = (overflow - o0) < blk ? (overflow - o0) : (int)blk;
          rc = score_block(ctx, ctx->tc_ovf_users + o0, ob, seen_indptr, seen_indices, r0, ctx->score_buf, ld, st,
                           ctx->tc_overflow_rows + o0);
          if (rc) return rc;
          tp.row0 = r0;
          tp.row_map = ctx->tc_overflow_rows + o0;
          topk_metrics<<<ob, 256, 0, st>>>(tp);
          tp.row_map = nullptr;
          ctx->launches++;
          RBPR_CUDA(ctx, cudaGetLastError());
        }
      }
    }
    return 0;
  }
  for (int64_t r0 = 0; r0 < n_users; r0 += blk) {
    const int nb = (int)((n_users - r0) < blk ? (n_users - r0) : blk);
    int rc = score_block(ctx, users + r0, nb, seen_indptr, seen_indices, r0, ctx->score_buf, ld, st);
    if (rc) return rc;
    tp.row0 = r0;
    topk_metrics<<<nb, 256, 0, st>>>(tp);
    ctx->launches++;
    ctx->topk_launches++;
    RBPR_CUDA(ctx, cudaGetLastError());
  }
  return 0;
}

int rbpr_score_topk(rbpr_ctx* ctx, const int64_t* users, int64_t n_users,
                    const int64_t* seen_indptr, const int32_t* seen_indices,
                    const int64_t* held_indptr, const int32_t* held_indices, int32_t k_max,
                    const int32_t* ks, int32_t n_ks, int32_t* topk_items, float* topk_scores,
                    float* ndcg_out, float* recall_out, void* stream) {
  rbpr_metric_outputs mo;
  memset(&mo, 0, sizeof(mo));
  mo.ndcg = ndcg_out;
  mo.recall = recall_out;
  return score_topk_impl(ctx, users, n_users, seen_indptr, seen_indices, held_indptr, held_indices, k_max, ks,
                         n_ks, topk_items, topk_scores, &mo, stream);
}

int rbpr_score_metrics(rbpr_ctx* ctx, const int64_t* users, int64_t n_users,
                       const int64_t* seen_indptr, const int32_t* seen_indices,
                       const int64_t* held_indptr, const int32_t* held_indices, int32_t k_max,
                       const int32_t* ks, int32_t n_ks, const rbpr_metric_outputs* out, void* stream) {
  if (ctx && !out) RBPR_FAIL(ctx, RBPR_ERR_ARG, "score_metrics: null outputs");
  return score_topk_impl(ctx, users, n_users, seen_indptr, seen_indices, held_indptr, held_indices, k_max, ks,
                         n_ks, out ? out->topk_items : nullptr, nullptr, out, stream);
}

int64_t rbpr_topk_launch_count(const rbpr_ctx* ctx) { return ctx ? ctx->topk_launches : 0; }

int rbpr_score_path_counts(const rbpr_ctx* ctx, int64_t* tensor_passes, int64_t* overflow_users) {
  if (!ctx) return RBPR_ERR_ARG;
  if (tensor_passes) *tensor_passes = ctx->tc_passes;
  if (overflow_users) *overflow_users = ctx->tc_overflow_users;
  return 0;
}

int rbpr_topk_metrics_dense(rbpr_ctx* ctx, const float* scores, const float* target, int64_t n_rows,
                            int64_t n_cols, int32_t k_max, const int32_t* ks, int32_t n_ks,
                            int32_t linear_gain, float* ndcg_out, float* recall_out,
                            float* precision_out, float* map_out, int32_t map_normalized,
                            int32_t* topk_items, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (n_rows == 0) return 0;
  if (!scores || !target || n_rows < 0 || n_cols < 1 || n_cols >= (1ll << 31))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "topk_metrics_dense: bad arguments");
  if (k_max < 1 || k_max > RBPR_MAX_TOPK)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "topk_metrics_dense: k_max=%d outside [1,%d]", k_max, RBPR_MAX_TOPK);
  if (n_ks < 1 || n_ks > 16 || !ks) RBPR_FAIL(ctx, RBPR_ERR_ARG, "topk_metrics_dense: 1..16 cut-offs");
  for (int q = 0; q < n_ks; ++q)
    if (ks[q] < 1 || ks[q] > k_max)
      RBPR_FAIL(ctx, RBPR_ERR_ARG, "topk_metrics_dense: cut-off %d outside [1,k_max=%d]", ks[q], k_max);
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  TopkParams tp;
  memset(&tp, 0, sizeof(tp));
  tp.S = scores;
  tp.ld = n_cols;
  tp.I = (int)n_cols;
  tp.k_max = k_max;
  tp.n_ks = n_ks;
  for (int q = 0; q < n_ks; ++q) tp.ks[q] = ks[q];
  tp.topk_items = topk_items;
  tp.ndcg_out = ndcg_out;
  tp.recall_out = recall_out;
  tp.precision_out = precision_out;
  tp.map_out = map_out;
  tp.map_normalized = map_normalized;
  tp.target = target;
  tp.target_ld = n_cols;
  tp.linear_gain = linear_gain;
  tp.flag = ctx->flag;
  for (int64_t r0 = 0; r0 < n_rows; r0 += 32768) {  // grid.x limit is not the issue; keep launches modest
    const int nb = (int)((n_rows - r0) < 32768 ? (n_rows - r0) : 32768);
    tp.S = scores + r0 * n_cols;
    tp.row0 = r0;
    topk_metrics<<<nb, 256, 0, (cudaStream_t)stream>>>(tp);
    ctx->launches++;
    ctx->topk_launches++;
  }
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // extern "C"
