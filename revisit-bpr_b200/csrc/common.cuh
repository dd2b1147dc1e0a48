// Shared context + helpers for librbpr.so (sm_100a only).
#pragma once
#include <cuda_runtime.h>
#include <stdint.h>
#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <string>
#include <vector>

#include <nvtx3/nvToolsExt.h>  // header-only: ranges cost a pointer check unless a profiler is attached

#include "../../include/rbpr.h"

struct rbpr_ctx {
  int device = 0;
  std::string err;
  // borrowed tables
  float* user_emb = nullptr;
  float* item_emb = nullptr;
  float* item_bias = nullptr;
  int64_t U = 0, I = 0;
  int D = 0;
  // borrowed Adam state
  float *user_m = nullptr, *user_v = nullptr, *item_m = nullptr, *item_v = nullptr;
  float *bias_m = nullptr, *bias_v = nullptr;
  int32_t* user_last = nullptr;
  // borrowed CSR
  const int64_t* indptr = nullptr;
  const int32_t* indices = nullptr;
  int64_t csr_users = 0, nnz = 0;
  // borrowed alias table
  const float* alias_prob = nullptr;
  const int32_t* alias_idx = nullptr;
  // owned scratch
  int32_t* coo_user = nullptr;  // (nnz) triple -> user
  uint32_t* bloom = nullptr;    // (U, 8) 256-bit membership filter per CSR row
  float* item_grad = nullptr;   // (I*D + I) dense item (+bias) gradient accumulator
  float* user_grad = nullptr;   // (U*D) dense user gradient accumulator (multi-occurrence users)
  uint32_t* touched = nullptr;  // (I) item touched in this step
  uint32_t* icnt = nullptr;     // (icnt_cap) small-batch path: per-step item occurrence counters of the wave being prepared
  int64_t icnt_cap = 0;
  int64_t small_launches = 0;   // launches of the persistent small-batch kernel
  uint32_t* ord = nullptr;  // (cap) arrival rank of each slot among its user's slots of the step
  int64_t cap = 0;
  uint32_t* cnt = nullptr;  // (cnt_cap) per-step user occurrence counters of the wave being prepared
  int64_t cnt_cap = 0;
  // per-wave scratch, double-buffered: wave w+1 is sorted/sampled on `aux` while wave w trains
  void* records[2] = {nullptr, nullptr};  // (records_cap) int4 {u, i+, i-, head}
  int64_t records_cap = 0;
  int32_t* mh_list[2] = {nullptr, nullptr};   // (records_cap) per step: multi-occurrence users, compacted by the sampler
  uint32_t* mh_count[2] = {nullptr, nullptr}; // (mh_steps_cap)
  int64_t mh_steps_cap = 0;
  float* partials[2] = {nullptr, nullptr};  // per-warp step statistics (float4 each)
  int64_t partials_cap = 0;
  cudaStream_t aux = nullptr;            // preparation stream (counting + negative sampling)
  cudaStream_t aux2 = nullptr;           // N>1: user half of bpr_apply, concurrent with the all-reduce
  cudaEvent_t ev_phase_a = nullptr, ev_users = nullptr;
  cudaEvent_t ev_inputs = nullptr;       // caller's stream -> aux: inputs of the call are ready
  cudaEvent_t ev_ready[2] = {nullptr, nullptr};  // aux -> main: records[b] are ready
  cudaEvent_t ev_free[2] = {nullptr, nullptr};   // main -> aux: records[b] may be overwritten
  double* stats = nullptr;  // (stats_cap steps, 4)
  int64_t stats_cap = 0;
  int32_t* flag = nullptr;  // device error flag
  int64_t* stage_idx = nullptr;  // staging for the *_host entry point
  int64_t* stage_neg = nullptr;
  int64_t stage_cap = 0;
  // adaptive sampler state (owned): transposed snapshot, per-factor std, sorted order + inverse
  float *ad_snap = nullptr, *ad_std = nullptr;
  uint64_t *ad_keys = nullptr, *ad_keys_sorted = nullptr;
  int32_t *ad_ids = nullptr, *ad_order = nullptr, *ad_pos = nullptr;
  void* ad_tmp = nullptr;
  size_t ad_tmp_bytes = 0, ad_cells = 0;
  // score scratch
  float* score_buf = nullptr;
  size_t score_buf_bytes = 0;
  // tensor-core scoring path (score_tc.cu): K-padded panels, seen bitmask, group maxima, candidates
  float *tc_items = nullptr, *tc_users = nullptr, *tc_gmax = nullptr;
  void *tc_mask = nullptr, *tc_small = nullptr;
  int32_t *tc_cand = nullptr, *tc_overflow_rows = nullptr;  // (tc_overflow_rows points into tc_small)
  size_t tc_items_bytes = 0, tc_users_bytes = 0, tc_gmax_bytes = 0, tc_mask_bytes = 0, tc_small_bytes = 0,
         tc_cand_bytes = 0;
  int64_t* tc_ovf_users = nullptr;  // (cap) users handed to the dense path
  size_t tc_ovf_users_bytes = 0;
  int64_t tc_passes = 0, tc_overflow_users = 0;
  // per-step Adam scalars {lr_s/(1-b1^s), sqrt(1-b2^s)}, index = 1-based optimizer step
  std::vector<float2> adam_host;
  float2* adam_tab = nullptr;
  int64_t adam_tab_cap = 0;
  // data-parallel communicator (NCCL, bound at run time; comm.cu)
  void* comm = nullptr;
  int world = 1, rank = 0;
  int64_t collectives = 0;
  // fused exchange over peer memory (exchange.cu): every rank's two gradient accumulators, item
  // table and bias mapped through cudaIpc, flag words for the cross-rank barrier
  bool fx_bound = false;
  void* fx_sym = nullptr;  // owned: buf[0] | buf[1] | flags
  float* fx_grad[RBPR_MAX_PEERS][2] = {};
  float* fx_item[RBPR_MAX_PEERS] = {};
  float* fx_bias[RBPR_MAX_PEERS] = {};
  uint32_t** fx_flags_dev = nullptr;  // device array of every rank's flag words
  std::vector<void*> fx_opened;       // cudaIpcOpenMemHandle mappings to close
  int fx_par = 0;                     // accumulator the current step uses
  uint32_t fx_epoch = 0;       // epoch of the exchanges' folded barriers (B1 / B2 flag words)
  uint32_t fx_bar_epoch = 0;   // epoch of the standalone barrier (bind time)
  uint32_t fx_wait_epoch = 0;  // B2 epoch the next reader of item rows must wait for (0: nothing pending)
  uint32_t* fx_flags_local = nullptr;  // this rank's flag words (inside fx_sym)
  uint32_t* fx_done = nullptr;         // CTA completion counter of the exchange kernel
  unsigned long long* fx_trace = nullptr;  // RBPR_FX_TRACE: globaltimer stamps of the exchange kernel (ring)
  int64_t fx_trace_n = 0;
  // symmetric-memory binding (rbpr_comm_symm_bind): host-owned buffer, optional NVSwitch multicast alias
  bool fx_symm_host = false;
  const float* fx_mc_grad[2] = {nullptr, nullptr};
  float* fx_mc_item = nullptr;
  float* fx_mc_bias = nullptr;
  float* fx_item_prev = nullptr;  // the caller's item table / bias before they moved into the symmetric buffer
  float* fx_bias_prev = nullptr;
  float* fx_item_grad_owned = nullptr;  // the library's own accumulator while the shared ones are in use
  int64_t fused_exchanges = 0;
  // instrumentation
  int64_t launches = 0;
  int64_t topk_launches = 0;  // launches of the ranking kernel (tests assert one per eval batch)
  bool timing = false;
  uint64_t timing_tick = 0;
  std::vector<cudaEvent_t> ev;  // pairs
  size_t ev_used = 0;
  double timed_ms = 0.0;
  int64_t timed_launches = 0;
  int sm_count = 148;
  int phase_a_blocks_per_sm[4] = {0, 0, 0, 0};  // per optimizer, for the bound dim (0 = not prepared)
};

// NVTX range over a host-side scope (what the kernels enqueued inside it belong to): preparation
// waves, phase A, the exchange, the apply, the scoring passes (tracing).
struct NvtxRange {
  explicit NvtxRange(const char* name) { nvtxRangePushA(name); }
  ~NvtxRange() { nvtxRangePop(); }
  NvtxRange(const NvtxRange&) = delete;
  NvtxRange& operator=(const NvtxRange&) = delete;
};

#define RBPR_FAIL(ctx, code, ...)                     \
  do {                                                \
    char _b[512];                                     \
    snprintf(_b, sizeof(_b), __VA_ARGS__);            \
    (ctx)->err = _b;                                  \
    return (code);                                    \
  } while (0)

#define RBPR_CUDA(ctx, expr)                                                            \
  do {                                                                                  \
    cudaError_t _e = (expr);                                                            \
    if (_e != cudaSuccess)                                                              \
      RBPR_FAIL(ctx, RBPR_ERR_CUDA, "%s failed: %s (%s:%d)", #expr, cudaGetErrorString(_e), \
                __FILE__, __LINE__);                                                    \
  } while (0)

// Group geometry for a row of D floats: LANES lanes (power of two, <=32) each holding NV float4
// vectors; column of vector v on lane gl is 4*(gl + LANES*v).  Two float4 per lane up to D=256
// (16 lanes at D=128: two triples per warp), more vectors per lane beyond.  Measured on B200 at
// D=128 (profiles/r01h_matrix.txt): 8x4, 16x2 and 32x1 are within 5% of each other, 16x2 best.
static inline void rbpr_geometry(int D, int* lanes, int* nv) {
  const int vecs = (D + 3) / 4;
  const int want = (vecs + 1) / 2;
  int l = 1;
  while (l < want && l < 32) l <<= 1;
  *lanes = l;
  *nv = (vecs + l - 1) / l;
}
