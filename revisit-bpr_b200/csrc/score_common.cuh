// Pieces shared by the two scoring paths (score.cu: fp32 SIMT GEMM + dense ranking; score_tc.cu:
// tensor-core candidate filter + exact fp32 rescoring): float ordering keys, the parameter block of
// the ranking kernels, and the metric arithmetic on a sorted top-k list.
#pragma once
#include "common.cuh"

constexpr float kMasked = -1e13f;

__device__ __forceinline__ uint32_t fkey(float x) {
  uint32_t u = __float_as_uint(x);
  return (u & 0x80000000u) ? ~u : (u | 0x80000000u);
}

constexpr int KCAP = RBPR_MAX_TOPK;  // 128

struct TopkParams {
  const float* __restrict__ S;
  int64_t ld;
  int I;
  int k_max;
  const int64_t* __restrict__ held_indptr;
  const int32_t* __restrict__ held_indices;
  int64_t row0;  // first local row of this block in the held CSR / outputs
  const int32_t* __restrict__ row_map;  // optional: block row -> local row (rows re-done by the dense path)
  int n_ks;
  int ks[16];
  int32_t* __restrict__ topk_items;
  float* __restrict__ topk_scores;
  float* __restrict__ ndcg_out;
  float* __restrict__ recall_out;
  float* __restrict__ precision_out;
  float* __restrict__ map_out;  // average precision @k (MAP.compute, revisit_bpr/metrics/map.py:45-64)
  int map_normalized;           // denominator min(n_pos, k) instead of hits@k
  // dense-target mode (revisit_bpr.metrics on (B,I) tensors): positives = target[row, item] > 0
  const float* __restrict__ target;
  int64_t target_ld;
  int linear_gain;  // NDCG gain_function="linear": discount 1/(rank+1) instead of 1/log2(rank+2)
  float* __restrict__ ndcg_linear_out;  // both gain functions from one pass (fused eval); needs !linear_gain
  int32_t* __restrict__ flag;
};


// inverse of fkey
__device__ __forceinline__ float ikey(uint32_t k) {
  return __uint_as_float((k & 0x80000000u) ? (k & 0x7FFFFFFFu) : ~k);
}

// Outputs of one user from its top-k list `sel` (shared memory, >= KCAP entries, sorted descending;
// entry = fkey(score) << 32 | (0xFFFFFFFF - item), only the first k are meaningful): top-k items /
// scores, hit flags against the held-out row (or the dense target), NDCG (both gains) / Recall /
// Precision / MAP at every cut-off.  Called by ALL threads of a block of >= 128 threads.
// Reference arithmetic: revisit_bpr/metrics/ndcg.py:8-24,69-78, recall.py:44-51, precision.py:44-51,
// map.py:45-64 (fp32 left-to-right sums in rank order).
__device__ __forceinline__ void topk_emit_outputs(const TopkParams& p, int64_t urow, const unsigned long long* sel,
                                                  int k) {
  __shared__ float disc_scan[KCAP], hit_scan[KCAP];
  __shared__ float disc_lin[KCAP], hit_lin[KCAP];
  __shared__ uint32_t hitbits[4];
  const int tid = threadIdx.x;
  const int64_t orow = p.row0 + (p.row_map != nullptr ? (int64_t)p.row_map[urow] : urow);
  int32_t item = -1;
  float hit = 0.f;
  int64_t hlo = 0, hhi = 0;
  if (p.held_indptr != nullptr) {
    hlo = p.held_indptr[orow];
    hhi = p.held_indptr[orow + 1];
  }
  int n_pos = (int)(hhi - hlo);
  if (p.target != nullptr) {  // count positives of the dense target row (and check it is binary)
    __shared__ int s_npos;
    if (tid == 0) s_npos = 0;
    __syncthreads();
    const float* trow = p.target + orow * p.target_ld;
    int local = 0;
    bool bad = false;
    for (int i = tid; i < p.I; i += (int)blockDim.x) {
      const float t = trow[i];
      local += (t == 1.0f);
      bad |= !(t == 0.0f || t == 1.0f);
    }
    if (bad) atomicExch(p.flag, 9);
    atomicAdd(&s_npos, local);
    __syncthreads();
    n_pos = s_npos;
  }
  // the user's held-out row in shared memory when it is short (the usual case): every ranked item then
  // checks its hit with a scan of shared memory instead of a chain of dependent global loads
  __shared__ int32_t s_held[KCAP];
  __shared__ float s_wtot[4][4];  // [array][warp] totals of the block scan below
  const bool held_small = p.target == nullptr && n_pos <= KCAP;
  if (held_small && tid < n_pos) s_held[tid] = p.held_indices[hlo + tid];
  __syncthreads();
  float v_disc = 0.f, v_hit = 0.f, v_dl = 0.f, v_hl = 0.f;
  if (tid < KCAP) {
    if (tid < k) {
      const unsigned long long c = sel[tid];
      item = (int32_t)(0xFFFFFFFFu - (uint32_t)(c & 0xFFFFFFFFull));
      if (p.topk_items) p.topk_items[orow * p.k_max + tid] = item;
      if (p.topk_scores) p.topk_scores[orow * p.k_max + tid] = ikey((uint32_t)(c >> 32));
      if (p.target != nullptr) {
        hit = (p.target[orow * p.target_ld + item] == 1.0f) ? 1.f : 0.f;
      } else if (held_small) {
        for (int q = 0; q < n_pos; ++q) hit = (s_held[q] == item) ? 1.f : hit;
      } else {
        int64_t lo = hlo, hi = hhi;
        while (lo < hi) {
          const int64_t mid = (lo + hi) >> 1;
          const int32_t v = p.held_indices[mid];
          if (v < item) lo = mid + 1; else hi = mid;
        }
        if (lo < hhi && p.held_indices[lo] == item) hit = 1.f;
      }
    } else if (tid < p.k_max) {
      if (p.topk_items) p.topk_items[orow * p.k_max + tid] = -1;
      if (p.topk_scores) p.topk_scores[orow * p.k_max + tid] = kMasked;
    }
    v_disc = p.linear_gain ? 1.0f / ((float)tid + 1.0f) : 1.0f / log2f((float)tid + 2.0f);
    v_hit = hit * v_disc;
    v_dl = 1.0f / ((float)tid + 1.0f);
    v_hl = hit * v_dl;
  }
  // inclusive prefix sums in rank order over the KCAP = 128 ranks: a shuffle scan inside each of the
  // four warps, then the totals of the warps before (fp32; both scoring paths share this code, so
  // they agree bit for bit; against the reference's torch.sum the order differs by ~1e-7 relative)
  {
    const int lane = tid & 31, wrp = tid >> 5;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const float a0 = __shfl_up_sync(0xffffffffu, v_disc, o), a1 = __shfl_up_sync(0xffffffffu, v_hit, o);
      const float a2 = __shfl_up_sync(0xffffffffu, v_dl, o), a3 = __shfl_up_sync(0xffffffffu, v_hl, o);
      if (lane >= o) {
        v_disc += a0;
        v_hit += a1;
        v_dl += a2;
        v_hl += a3;
      }
    }
    if (tid < KCAP && lane == 31) {
      s_wtot[0][wrp] = v_disc;
      s_wtot[1][wrp] = v_hit;
      s_wtot[2][wrp] = v_dl;
      s_wtot[3][wrp] = v_hl;
    }
    const unsigned bal = __ballot_sync(0xffffffffu, hit > 0.f);
    if (tid < KCAP && lane == 0) hitbits[wrp] = bal;
    __syncthreads();
    if (tid < KCAP) {
      for (int q = 0; q < wrp; ++q) {
        v_disc += s_wtot[0][q];
        v_hit += s_wtot[1][q];
        v_dl += s_wtot[2][q];
        v_hl += s_wtot[3][q];
      }
      disc_scan[tid] = v_disc;
      hit_scan[tid] = v_hit;
      disc_lin[tid] = v_dl;
      hit_lin[tid] = v_hl;
    }
  }
  __syncthreads();
  if (tid < p.n_ks) {
    const int kk = min(min(p.ks[tid], k), KCAP);
    float ndcg = 0.f, ndcg_lin = 0.f, recall = 0.f, precision = 0.f, ap = 0.f;
    if (kk > 0 && n_pos > 0) {
      const float dcg = hit_scan[kk - 1];
      const float idcg = disc_scan[min(kk, n_pos) - 1];
      ndcg = dcg / idcg;
      if (p.ndcg_linear_out != nullptr) ndcg_lin = hit_lin[kk - 1] / disc_lin[min(kk, n_pos) - 1];
      int hits = 0;
      for (int w = 0; w < 4; ++w) {
        const int lo = w * 32;
        if (kk <= lo) break;
        const int take = min(32, kk - lo);
        const uint32_t m = (take == 32) ? 0xffffffffu : ((1u << take) - 1u);
        hits += __popc(hitbits[w] & m);
      }
      recall = (float)hits / (float)n_pos;
      precision = (float)hits / (float)kk;
      if (p.map_out != nullptr) {
        float acc = 0.f;
        int cum = 0;
        for (int r = 0; r < kk; ++r) {
          if ((hitbits[r >> 5] >> (r & 31)) & 1u) {
            ++cum;
            acc += (float)cum / (float)(r + 1);
          }
        }
        const int denom = p.map_normalized ? min(n_pos, kk) : hits;
        ap = denom > 0 ? acc / (float)denom : 0.f;
      }
    }
    if (p.ndcg_out) p.ndcg_out[orow * p.n_ks + tid] = ndcg;
    if (p.ndcg_linear_out) p.ndcg_linear_out[orow * p.n_ks + tid] = ndcg_lin;
    if (p.recall_out) p.recall_out[orow * p.n_ks + tid] = recall;
    if (p.precision_out) p.precision_out[orow * p.n_ks + tid] = precision;
    if (p.map_out) p.map_out[orow * p.n_ks + tid] = ap;
  }
}

