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
  if (tid < KCAP) {
    if (tid < k) {
      const unsigned long long c = sel[tid];
      item = (int32_t)(0xFFFFFFFFu - (uint32_t)(c & 0xFFFFFFFFull));
      if (p.topk_items) p.topk_items[orow * p.k_max + tid] = item;
      if (p.topk_scores) p.topk_scores[orow * p.k_max + tid] = ikey((uint32_t)(c >> 32));
      if (p.target != nullptr) {
        hit = (p.target[orow * p.target_ld + item] == 1.0f) ? 1.f : 0.f;
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
    const float disc = p.linear_gain ? 1.0f / ((float)tid + 1.0f) : 1.0f / log2f((float)tid + 2.0f);
    disc_scan[tid] = disc;
    hit_scan[tid] = hit * disc;
    if (p.ndcg_linear_out != nullptr) {
      const float dl = 1.0f / ((float)tid + 1.0f);
      disc_lin[tid] = dl;
      hit_lin[tid] = hit * dl;
    }
  }
  __syncthreads();
  // sequential prefix sums in rank order (fp32, like a left-to-right sum) by two threads
  if (tid == 0) {
    float s = 0.f;
    for (int r = 0; r < KCAP; ++r) { s += disc_scan[r]; disc_scan[r] = s; }
  } else if (tid == 32) {
    float s = 0.f;
    for (int r = 0; r < KCAP; ++r) { s += hit_scan[r]; hit_scan[r] = s; }
  } else if (tid == 64 && p.ndcg_linear_out != nullptr) {
    float s = 0.f;
    for (int r = 0; r < KCAP; ++r) { s += disc_lin[r]; disc_lin[r] = s; }
  } else if (tid == 96 && p.ndcg_linear_out != nullptr) {
    float s = 0.f;
    for (int r = 0; r < KCAP; ++r) { s += hit_lin[r]; hit_lin[r] = s; }
  }
  // hit counts: reuse ballots
  
  {
    const unsigned bal = __ballot_sync(0xffffffffu, hit > 0.f);
    if (tid < KCAP && (tid & 31) == 0) hitbits[tid >> 5] = bal;
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

