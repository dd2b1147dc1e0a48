// BPR training step for sm_100a: on-device negative sampling, (u, i+, i-) row gather, dot,
// log-sigmoid loss + L2, exact synchronous-minibatch gradients, in-place update.
//
// A call is cut into WAVES of whole steps.  Per wave, on the context's auxiliary stream:
//   P1  count_users + bpr_sample — per-step occurrence count of every user (one atomic per
//       triple into an L2-resident counter table; no sort), then a group of kSampleLanes (2) lanes
//       per slot draws the negative (Philox4x32-10, Bloom filter + CSR rejection) and emits a 16-byte record
//       {u, i+, i-, flags} where the flags say whether the user occurs once in its step.  The
//       static samplers depend on (seed, step, triple, CSR) only, so wave w+1 is prepared while
//       wave w trains.
// Per step, on the caller's stream (kernels in train_kernels.cuh):
//   P2  bpr_phase_a — one lane group per triple, 128-bit row loads, shuffle-reduced dot, the two
//       item-row gradients into the dense accumulator (vector red), single-occurrence users
//       updated in place, multi-occurrence users into the dense user-gradient accumulator.
//       Tables are only READ for rows that other triples of the step may read: exact minibatch.
//   P3  bpr_apply — accumulated item gradient (SGD: touched rows; Adam: every row, which is
//       torch.optim.Adam's dense semantics) and multi-occurrence users; clears the accumulators.
// Reference call sites replaced: see include/rbpr.h (rbpr_train_steps).
#include "train_kernels.cuh"

using namespace rbpr_dev;

namespace {

constexpr int64_t kWaveTriples = 1ll << 19;       // triples sorted + sampled per preparation wave
constexpr int64_t kFirstWaveTriples = 1ll << 17;  // ... of the first wave of a call
constexpr int64_t kCounterEntries = 1ll << 24;    // budget of the (steps in wave, U) counter table

// Per-step occurrence counts of users: slot k (step k / batch of the wave) takes a ticket from
// cnt[step][user]; its arrival rank is kept so that ONE slot of a repeated user can be designated.
__global__ void count_users(const int64_t* __restrict__ triple_idx, int64_t n, int64_t batch,
                            int64_t nnz, int64_t U, const int32_t* __restrict__ coo_user,
                            uint32_t* __restrict__ cnt, uint32_t* __restrict__ ord,
                            int32_t* __restrict__ flag) {
  // grid: x over the slots of one step, y = step of the wave
  const int64_t local = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
  if (local >= batch) return;
  const int64_t k = (int64_t)blockIdx.y * batch + local;
  if (k >= n) return;
  int64_t t = triple_idx[k];
  if (t < 0 || t >= nnz) {
    atomicExch(flag, 2);
    t = 0;
  }
  const int64_t u = coo_user[t];
  ord[k] = atomicAdd(cnt + (int64_t)blockIdx.y * U + u, 1u);
}

__global__ void expand_rows(const int64_t* __restrict__ indptr, int64_t U, int64_t nnz,
                            uint32_t I, int32_t* __restrict__ coo_user,
                            const int32_t* __restrict__ indices, uint32_t* __restrict__ bloom,
                            int32_t* __restrict__ flag) {
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
    const uint32_t h1 = bloom_h1(it), h2 = bloom_h2(it);
    atomicOr(bloom + row * 8 + (h1 >> 5), 1u << (h1 & 31u));
    atomicOr(bloom + row * 8 + (h2 >> 5), 1u << (h2 & 31u));
  }
}

int check_tables(rbpr_ctx* ctx, const rbpr_hparams* hp) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!hp) RBPR_FAIL(ctx, RBPR_ERR_ARG, "hparams is NULL");
  if (!ctx->user_emb || !ctx->item_emb) RBPR_FAIL(ctx, RBPR_ERR_STATE, "tables not bound");
  if (hp->optimizer < RBPR_OPT_SGD || hp->optimizer > RBPR_OPT_RMSPROP)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "unknown optimizer %d", hp->optimizer);
  if (hp->optimizer == RBPR_OPT_SGDM || hp->optimizer == RBPR_OPT_RMSPROP) {
    if (!ctx->user_m || !ctx->item_m || !ctx->user_last)
      RBPR_FAIL(ctx, RBPR_ERR_STATE, "momentum / RMSprop requested but optimizer state not bound");
    if (ctx->item_bias && !ctx->bias_m)
      RBPR_FAIL(ctx, RBPR_ERR_STATE, "optimizer state of the item bias not bound");
  }
  if (hp->optimizer == RBPR_OPT_ADAM) {
    if (!ctx->user_m || !ctx->user_v || !ctx->item_m || !ctx->item_v || !ctx->user_last)
      RBPR_FAIL(ctx, RBPR_ERR_STATE, "Adam requested but Adam state not bound");
    if (ctx->item_bias && (!ctx->bias_m || !ctx->bias_v))
      RBPR_FAIL(ctx, RBPR_ERR_STATE, "Adam requested with item bias but bias state not bound");
  }
  return 0;
}

int check_ready(rbpr_ctx* ctx, const rbpr_hparams* hp) {
  int rc = check_tables(ctx, hp);
  if (rc) return rc;
  if (!ctx->indptr || !ctx->indices) RBPR_FAIL(ctx, RBPR_ERR_STATE, "CSR not bound");
  if (ctx->csr_users != ctx->U)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "CSR has %lld rows but the user table has %lld",
              (long long)ctx->csr_users, (long long)ctx->U);
  if (hp->sampler < RBPR_SAMPLER_UNIFORM || hp->sampler > RBPR_SAMPLER_ADAPTIVE)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "unknown sampler %d", hp->sampler);
  if (hp->sampler == RBPR_SAMPLER_WEIGHTED && (!ctx->alias_prob || !ctx->alias_idx))
    RBPR_FAIL(ctx, RBPR_ERR_STATE, "weighted sampler requested but alias table not bound");
  return 0;
}

// Per-wave preparation scratch: arrival ranks (n slots) and the user counters (cnt_entries).
int ensure_capacity(rbpr_ctx* ctx, int64_t n, int64_t steps, int64_t cnt_entries) {
  if (n > ctx->cap) {
    cudaFree(ctx->ord);
    ctx->ord = nullptr;
    ctx->cap = 0;
    RBPR_CUDA(ctx, cudaMalloc(&ctx->ord, n * sizeof(uint32_t)));
    ctx->cap = n;
  }
  if (cnt_entries > ctx->cnt_cap) {
    cudaFree(ctx->cnt);
    ctx->cnt = nullptr;
    ctx->cnt_cap = 0;
    RBPR_CUDA(ctx, cudaMalloc(&ctx->cnt, cnt_entries * sizeof(uint32_t)));
    ctx->cnt_cap = cnt_entries;
  }
  if (steps > ctx->stats_cap) {
    cudaFree(ctx->stats);
    ctx->stats = nullptr;
    ctx->stats_cap = 0;
    const int64_t cap = steps > 4096 ? steps : 4096;
    RBPR_CUDA(ctx, cudaMalloc(&ctx->stats, cap * RBPR_STATS_PER_STEP * sizeof(double)));
    ctx->stats_cap = cap;
  }
  return 0;
}

// Make the device table of Adam scalars cover optimizer steps [1, last]; entries from `first` on are
// (re)computed with the CURRENT hyper-parameters (they describe steps that have not run yet),
// earlier ones keep the learning rate they ran with; holes (resume at a late step) take the current
// values.  Host double arithmetic == torch.optim.Adam's Python scalars.
int ensure_adam_table(rbpr_ctx* ctx, const rbpr_hparams* hp, int64_t first, int64_t last,
                      cudaStream_t st) {
  if (hp->optimizer != RBPR_OPT_ADAM) return 0;
  if (first < 1) first = 1;
  if (last < first) last = first;
  const int64_t old = (int64_t)ctx->adam_host.size();
  if (last + 1 > old) ctx->adam_host.resize((size_t)last + 1, make_float2(0.f, 1.f));
  const int64_t from = old < first ? (old < 1 ? 1 : old) : first;
  for (int64_t s = from; s <= last; ++s) {
    const double b1p = pow((double)hp->beta1, (double)s), b2p = pow((double)hp->beta2, (double)s);
    ctx->adam_host[(size_t)s] = make_float2((float)((double)hp->lr / (1.0 - b1p)), (float)sqrt(1.0 - b2p));
  }
  int64_t lo = from;
  if (last + 1 > ctx->adam_tab_cap) {
    int64_t cap = ctx->adam_tab_cap > 0 ? ctx->adam_tab_cap : 4096;
    while (cap < last + 1) cap *= 2;
    cudaFree(ctx->adam_tab);
    ctx->adam_tab = nullptr;
    ctx->adam_tab_cap = 0;
    RBPR_CUDA(ctx, cudaMalloc(&ctx->adam_tab, (size_t)cap * sizeof(float2)));
    ctx->adam_tab_cap = cap;
    lo = 0;  // re-upload everything
  }
  RBPR_CUDA(ctx, cudaMemcpyAsync(ctx->adam_tab + lo, ctx->adam_host.data() + lo,
                                 (size_t)(last + 1 - lo) * sizeof(float2), cudaMemcpyHostToDevice, st));
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
  p.user_grad = ctx->user_grad;
  p.adam_tab = ctx->adam_tab;
  p.bias_grad = ctx->item_bias ? ctx->item_grad + ctx->I * ctx->D : nullptr;
  p.touched = ctx->touched;
  p.indptr = ctx->indptr;
  p.indices = ctx->indices;
  p.coo_user = ctx->coo_user;
  p.bloom = ctx->bloom;
  p.cnt = ctx->cnt;
  p.ord = ctx->ord;
  p.U = ctx->U;
  p.nnz = ctx->nnz;
  p.alias_prob = ctx->alias_prob;
  p.alias_idx = ctx->alias_idx;
  p.flag = ctx->flag;
  if (ctx->fx_bound && ctx->fx_wait_epoch != 0) {  // item rows of the previous exchange must have landed
    p.xwait_flags = ctx->fx_flags_local + RBPR_MAX_PEERS;
    p.xwait_n = ctx->world;
    p.xwait_epoch = ctx->fx_wait_epoch;
  }
  p.D = ctx->D;
  p.I = (uint32_t)ctx->I;
  p.seed_lo = (uint32_t)seed;
  p.seed_hi = (uint32_t)(seed >> 32);
  if (hp) {
    p.draw_n = (hp->sampler == RBPR_SAMPLER_WEIGHTED) ? p.I : p.I - 1u;
    p.draw_thresh = (uint32_t)((1ull << 32) % (uint64_t)p.draw_n);
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

cudaEvent_t next_event(rbpr_ctx* ctx) {
  if (ctx->ev_used == ctx->ev.size()) {
    cudaEvent_t e;
    cudaEventCreate(&e);
    ctx->ev.push_back(e);
  }
  return ctx->ev[ctx->ev_used++];
}

// Occurrence counts for n triples split into steps of `batch` (capacity ensured by the caller).
int count_batches(rbpr_ctx* ctx, const int64_t* triple_idx, int64_t n, int64_t batch,
                  cudaStream_t st) {
  const int64_t steps = (n + batch - 1) / batch;
  RBPR_CUDA(ctx, cudaMemsetAsync(ctx->cnt, 0, (size_t)steps * ctx->U * sizeof(uint32_t), st));
  const int threads = 256;
  const int64_t per_step = batch < n ? batch : n;
  if (steps > 65535) RBPR_FAIL(ctx, RBPR_ERR_ARG, "more than 65535 steps in one preparation wave");
  const dim3 grid((unsigned)((per_step + threads - 1) / threads), (unsigned)steps);
  count_users<<<grid, threads, 0, st>>>(triple_idx, n, batch, ctx->nnz, ctx->U, ctx->coo_user, ctx->cnt,
                                        ctx->ord, ctx->flag);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

// CTAs of phase A: one lane group per triple, capped at ONE resident wave (the same number of
// CTAs on every SM; groups then stride over the step's records).
int pick_blocks(rbpr_ctx* ctx, const rbpr_hparams* hp, int64_t n, int lanes, int nv, int* blocks_out) {
  const int o = hp->optimizer;
  if (ctx->phase_a_blocks_per_sm[o] == 0) {
    int bps = 0;
    int rc = o == RBPR_OPT_ADAM   ? rbpr_phase_a_prepare_adam(ctx, ctx->D, lanes, nv, &bps)
             : o == RBPR_OPT_SGDM ? rbpr_phase_a_prepare_sgdm(ctx, ctx->D, lanes, nv, &bps)
             : o == RBPR_OPT_RMSPROP ? rbpr_phase_a_prepare_rms(ctx, ctx->D, lanes, nv, &bps)
                                     : rbpr_phase_a_prepare_sgd(ctx, ctx->D, lanes, nv, &bps);
    if (rc) return rc;
    if (bps < 1) RBPR_FAIL(ctx, RBPR_ERR_CUDA, "phase-A kernel does not fit on an SM (dim=%d)", ctx->D);
    const char* e = getenv("RBPR_BLOCKS_PER_SM");  // tuning override
    if (e && atoi(e) > 0 && atoi(e) < bps) bps = atoi(e);
    ctx->phase_a_blocks_per_sm[o] = bps;
  }
  const int gpb = kPhaseAThreads / lanes;
  const int64_t resident_blocks = (int64_t)ctx->sm_count * ctx->phase_a_blocks_per_sm[o];
  int64_t blocks = (n + gpb - 1) / gpb;
  if (blocks > resident_blocks) blocks = resident_blocks;
  if (blocks < 1) blocks = 1;
  *blocks_out = (int)blocks;
  return 0;
}

// Scratch for one wave (both buffers): 16-byte record per triple, `steps` x `stride` float4
// statistics partials.
int ensure_step_scratch(rbpr_ctx* ctx, int64_t n, int64_t steps, int stride) {
  if (n > ctx->records_cap) {
    for (int b = 0; b < 2; ++b) {
      cudaFree(ctx->records[b]);
      ctx->records[b] = nullptr;
    }
    ctx->records_cap = 0;
    for (int b = 0; b < 2; ++b) RBPR_CUDA(ctx, cudaMalloc(&ctx->records[b], (size_t)n * 16));
    for (int b = 0; b < 2; ++b) {
      cudaFree(ctx->mh_list[b]);
      ctx->mh_list[b] = nullptr;
      RBPR_CUDA(ctx, cudaMalloc(&ctx->mh_list[b], (size_t)n * sizeof(int32_t)));
    }
    ctx->records_cap = n;
  }
  if (steps > ctx->mh_steps_cap) {
    for (int b = 0; b < 2; ++b) {
      cudaFree(ctx->mh_count[b]);
      ctx->mh_count[b] = nullptr;
      RBPR_CUDA(ctx, cudaMalloc(&ctx->mh_count[b], (size_t)steps * sizeof(uint32_t)));
    }
    ctx->mh_steps_cap = steps;
  }
  const int64_t need = steps * stride;
  if (need > ctx->partials_cap) {
    for (int b = 0; b < 2; ++b) {
      cudaFree(ctx->partials[b]);
      ctx->partials[b] = nullptr;
    }
    ctx->partials_cap = 0;
    for (int b = 0; b < 2; ++b)
      RBPR_CUDA(ctx, cudaMalloc(&ctx->partials[b], need * 4 * sizeof(float)));
    ctx->partials_cap = need;
  }
  return 0;
}

// P1 over all sorted slots of a wave.
int run_sample(rbpr_ctx* ctx, const TrainParams& p, void* records, int64_t n, uint64_t step0,
               cudaStream_t st) {
  const int64_t per_step = p.batch < n ? p.batch : n;
  const int64_t steps = (n + p.batch - 1) / p.batch;
  const dim3 grid((unsigned)((per_step * kSampleLanes + 255) / 256), (unsigned)steps);
  bpr_sample<<<grid, 256, 0, st>>>(p, reinterpret_cast<int4*>(records), (uint64_t)n, step0);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

// P2 for one step: p.n / p.step / p.chunk set by the caller; records and partials of that step.
int run_phase_a(rbpr_ctx* ctx, TrainParams& p, const rbpr_hparams* hp, const int4* records,
                float4* partials, int blocks, cudaStream_t st, int* launched_blocks) {
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  p.partials = partials;
  cudaEvent_t e0 = nullptr, e1 = nullptr;
  // timing samples every 8th launch: event pairs around every launch perturb back-to-back steps
  const bool timed = ctx->timing && (ctx->timing_tick++ % 8 == 0);
  if (timed) {
    e0 = next_event(ctx);
    e1 = next_event(ctx);
    cudaEventRecord(e0, st);
    p.pdl = 0;  // a timed launch must not start (and spin) while the previous kernel still runs
  }
  // the last (short) step of a call may need fewer CTAs; never more than `blocks` (partials stride)
  const int64_t need = ((int64_t)p.n + (kPhaseAThreads / lanes) - 1) / (kPhaseAThreads / lanes);
  const int nb = (int)(need < blocks ? (need > 0 ? need : 1) : blocks);
  *launched_blocks = nb;
  int rc = hp->optimizer == RBPR_OPT_SGD    ? rbpr_launch_phase_a_sgd(ctx, p, lanes, nv, records, nb, st)
           : hp->optimizer == RBPR_OPT_ADAM ? rbpr_launch_phase_a_adam(ctx, p, lanes, nv, records, nb, st)
           : hp->optimizer == RBPR_OPT_SGDM ? rbpr_launch_phase_a_sgdm(ctx, p, lanes, nv, records, nb, st)
                                            : rbpr_launch_phase_a_rms(ctx, p, lanes, nv, records, nb, st);
  if (rc) return rc;
  if (timed) cudaEventRecord(e1, st);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

// do_items / do_users select the halves of bpr_apply; records/n are the step's records (users).
// partials/n_partials/stats_out: phase A's per-warp statistics of the step, summed by block 0.
ApplyParams make_apply_params(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, int dense, int do_items,
                              const int4* records, int n, const float4* partials, int n_partials,
                              double* stats_out, const int32_t* mh_list, const uint32_t* mh_count) {
  ApplyParams a;
  memset(&a, 0, sizeof(a));
  a.mh_list = mh_list;
  a.mh_count = mh_count;
  a.do_items = do_items;
  a.do_users = (records != nullptr && n > 0) ? 1 : 0;
  a.partials = partials;
  a.n_partials = n_partials;
  a.stats_out = stats_out;
  a.records = records;
  a.n = n;
  a.user_emb = ctx->user_emb;
  a.user_grad = ctx->user_grad;
  a.user_m = ctx->user_m;
  a.user_v = ctx->user_v;
  a.user_last = ctx->user_last;
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
  a.dense = dense || hp->optimizer != RBPR_OPT_SGD;  // stateful optimizers move every item row
  a.step = step;
  a.lr = hp->lr;
  a.beta1 = hp->beta1;
  a.beta2 = hp->beta2;
  a.eps = hp->eps;
  a.adam_tab = ctx->adam_tab;
  return a;
}

int run_apply(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, int dense, int do_items,
              const int4* records, int n, cudaStream_t st, const float4* partials = nullptr,
              int n_partials = 0, double* stats_out = nullptr, const int32_t* mh_list = nullptr,
              const uint32_t* mh_count = nullptr) {
  const ApplyParams a = make_apply_params(ctx, step, hp, dense, do_items, records, n, partials, n_partials,
                                          stats_out, mh_list, mh_count);
  if (!a.do_items && !a.do_users && !stats_out) return 0;
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  int rc = hp->optimizer == RBPR_OPT_SGD    ? rbpr_launch_apply_sgd(ctx, a, lanes, nv, st)
           : hp->optimizer == RBPR_OPT_ADAM ? rbpr_launch_apply_adam(ctx, a, lanes, nv, st)
           : hp->optimizer == RBPR_OPT_SGDM ? rbpr_launch_apply_sgdm(ctx, a, lanes, nv, st)
                                            : rbpr_launch_apply_rms(ctx, a, lanes, nv, st);
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
      case 7: RBPR_FAIL(ctx, RBPR_ERR_ARG, "user id outside [0,num_users)");
      case 8: RBPR_FAIL(ctx, RBPR_ERR_ARG, "item id outside [0,num_items)");
      case 9: RBPR_FAIL(ctx, RBPR_ERR_DATA, "metric target contains values outside of 0 and 1");
      case 11: RBPR_FAIL(ctx, RBPR_ERR_CUDA, "scoring pipeline stalled (mbarrier wait timed out): set RBPR_NO_TC_SCORE=1");
      case 10: RBPR_FAIL(ctx, RBPR_ERR_COMM, "cross-rank barrier timed out: the ranks are out of step");
      default: RBPR_FAIL(ctx, RBPR_ERR_DATA, "device error flag %d", f);
    }
  }
  return 0;
}

}  // namespace

// defined in adaptive.cu
int rbpr_internal_sample_adaptive_csr(rbpr_ctx* ctx, const TrainParams& tp, void* records, int64_t n,
                                      uint64_t step, double sampling_prob, int opt, cudaStream_t st);

// defined in comm.cu
int rbpr_internal_allreduce_item_grads(rbpr_ctx* ctx, cudaStream_t st);

// defined in train_small.cu
bool rbpr_small_batch_eligible(const rbpr_ctx* ctx, int64_t batch);
int rbpr_launch_small_steps(rbpr_ctx* ctx, const TrainParams& p, const int4* records, int64_t n, int n_steps,
                            double* stats, cudaStream_t st);

// defined in exchange.cu
int rbpr_internal_fused_exchange(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, cudaStream_t st,
                                 const ApplyParams* users);
int rbpr_internal_fx_wait(rbpr_ctx* ctx, cudaStream_t st);
int rbpr_internal_fx_trace_report(rbpr_ctx* ctx, cudaStream_t st);

// The step's one exchange on stream st: dense item gradient summed over ranks, then the (dense,
// identical on every rank) item update.
int rbpr_internal_exchange_apply(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, cudaStream_t st) {
  if (ctx->fx_bound) return rbpr_internal_fused_exchange(ctx, step, hp, st, nullptr);  // one kernel over peer memory
  int rc = rbpr_internal_allreduce_item_grads(ctx, st);
  if (rc) return rc;
  return run_apply(ctx, step, hp, 1, 1, nullptr, 0, st);
}

int rbpr_internal_ensure_adam_table(rbpr_ctx* ctx, const rbpr_hparams* hp, int64_t first, int64_t last,
                                    cudaStream_t st) {
  return ensure_adam_table(ctx, hp, first, last, st);
}

// ---- helpers shared with dropin.cu ---------------------------------------------------------------
int rbpr_internal_check_ready_tables(rbpr_ctx* ctx, const rbpr_hparams* hp) {
  return check_tables(ctx, hp);
}

int rbpr_internal_reserve_sort(rbpr_ctx* ctx, int64_t n) {
  int rc = ensure_capacity(ctx, n, 1, ctx->U);
  if (rc) return rc;
  return ensure_step_scratch(ctx, n, 1, 4 * ctx->sm_count * 16);
}

// One step over prepared records on the caller's stream: phase A, apply (items + users), stats.
int rbpr_internal_phase_a_apply(rbpr_ctx* ctx, const rbpr_hparams* hp, const int4* records, int n,
                                uint64_t step, float2* logit_out, const int32_t* step_pos,
                                double* stats_out, cudaStream_t st) {
  RBPR_CUDA(ctx, cudaMemsetAsync(ctx->stats, 0, RBPR_STATS_PER_STEP * sizeof(double), st));
  const int multi = (ctx->comm != nullptr && ctx->world > 1) ? 1 : 0;
  if (n > 0) {
    int lanes, nv, blocks = 1;
    rbpr_geometry(ctx->D, &lanes, &nv);
    int rc = pick_blocks(ctx, hp, n, lanes, nv, &blocks);
    if (rc) return rc;
    const int stride = blocks * (kPhaseAThreads / 32);
    rc = ensure_step_scratch(ctx, n, 1, stride);
    if (rc) return rc;
    rc = ensure_adam_table(ctx, hp, (int64_t)step + 1, (int64_t)step + 1, st);
    if (rc) return rc;
    TrainParams p;
    fill_train_params(ctx, p, 0, hp);
    p.n = n;
    p.step = step;
    p.logit_out = logit_out;
    p.step_pos = step_pos;
    int nb = 0;
    rc = run_phase_a(ctx, p, hp, records, reinterpret_cast<float4*>(ctx->partials[0]), blocks, st, &nb);
    if (rc) return rc;
    // data-parallel: users (owned by this rank) now, items after the exchange below
    rc = run_apply(ctx, step, hp, 0, multi ? 0 : 1, records, n, st, reinterpret_cast<const float4*>(ctx->partials[0]),
                   nb * (kPhaseAThreads / 32), ctx->stats);
    if (rc) return rc;
  }
  if (multi) {  // every rank takes part in the step's one exchange, with or without local triples
    int rc = ensure_adam_table(ctx, hp, (int64_t)step + 1, (int64_t)step + 1, st);
    if (rc) return rc;
    rc = rbpr_internal_exchange_apply(ctx, step, hp, st);
    if (rc) return rc;
    rc = rbpr_internal_fx_wait(ctx, st);
    if (rc) return rc;
  }
  if (stats_out)
    RBPR_CUDA(ctx, cudaMemcpyAsync(stats_out, ctx->stats, RBPR_STATS_PER_STEP * sizeof(double),
                                   cudaMemcpyDeviceToDevice, st));
  return 0;
}

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
  cudaFree(ctx->bloom);
  ctx->bloom = nullptr;
  RBPR_CUDA(ctx, cudaMalloc(&ctx->bloom, (size_t)num_users * 8 * sizeof(uint32_t)));
  RBPR_CUDA(ctx, cudaMemsetAsync(ctx->bloom, 0, (size_t)num_users * 8 * sizeof(uint32_t), st));
  const int64_t threads = num_users * 32;
  expand_rows<<<(int)((threads + 255) / 256), 256, 0, st>>>(indptr, num_users, nnz,
                                                            (uint32_t)ctx->I, ctx->coo_user,
                                                            indices, ctx->bloom, ctx->flag);
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
  if (n >= (1ll << 31)) RBPR_FAIL(ctx, RBPR_ERR_ARG, "sample: n must be < 2^31 per call");
  // the wave sampler with one "step" of n slots and no records: only neg_out is written
  p.triple_idx = triple_idx;
  p.batch = n;
  p.neg_out = neg_out;
  int rc2 = run_sample(ctx, p, nullptr, n, step, st);
  if (rc2) return rc2;
  return check_flag(ctx, st);
}

}  // extern "C"

// The training loop behind rbpr_train_steps / rbpr_train_steps_host.  With host sources the
// wave's slice of triple ids (and injected negatives) is copied host->device by the preparation
// stream right before that wave is counted and sampled, i.e. the PCIe transfer of wave w+1 overlaps
// the training of wave w instead of preceding the whole call.
static int train_steps_impl(rbpr_ctx* ctx, int64_t* triple_idx, int64_t n, int64_t batch,
                            uint64_t seed, uint64_t step0, const rbpr_hparams* hp, int64_t* neg_in,
                            int64_t* neg_out, double* stats_out, const int64_t* host_idx,
                            const int64_t* host_neg_in, void* stream) {
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
  const int64_t steps = (n + batch - 1) / batch;
  // A call is processed in WAVES of whole steps (<= kWaveTriples triples): wave w+1 is sorted and
  // sampled on the context's auxiliary stream while wave w trains on the caller's stream (the
  // static samplers depend on (seed, step, triple, CSR) only, never on the model).
  // The adaptive sampler reads the CURRENT user rows: it is sampled step by step on the caller's
  // stream, never ahead of the model.
  const bool adaptive = hp->sampler == RBPR_SAMPLER_ADAPTIVE;
  int64_t spw = adaptive ? 1 : kWaveTriples / batch;
  if (spw > kCounterEntries / ctx->U) spw = kCounterEntries / ctx->U;  // (steps, U) counter table
  if (spw > 65535) spw = 65535;                                        // gridDim.y of the prep kernels
  if (spw < 1) spw = 1;
  if (spw > steps) spw = steps;
  // the first wave is short (its preparation is the only one that is not overlapped)
  int64_t first = kFirstWaveTriples / batch;
  if (first < 1) first = 1;
  if (first > spw) first = spw;
  const int64_t nwaves = 1 + (steps - first + spw - 1) / spw;
  auto wave_step0 = [&](int64_t w) { return w == 0 ? (int64_t)0 : first + (w - 1) * spw; };
  auto wave_steps = [&](int64_t w) {
    const int64_t a0 = wave_step0(w), a1 = (w == 0) ? first : a0 + spw;
    return (a1 < steps ? a1 : steps) - a0;
  };
  const int64_t wave_cap = (spw * batch < n) ? spw * batch : n;
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  int blocks = 1;
  rc = pick_blocks(ctx, hp, batch < n ? batch : n, lanes, nv, &blocks);
  if (rc) return rc;
  const int stride = blocks * (kPhaseAThreads / 32);
  // scratch is sized for a full wave from the first call on, so later (longer) calls never allocate
  const int64_t alloc_cap = wave_cap > kWaveTriples ? wave_cap : kWaveTriples;
  int64_t spw_cap = kWaveTriples / batch;
  if (spw_cap > kCounterEntries / ctx->U) spw_cap = kCounterEntries / ctx->U;
  if (spw_cap > 65535) spw_cap = 65535;
  const int64_t alloc_spw = spw_cap > spw ? spw_cap : spw;
  rc = ensure_capacity(ctx, alloc_cap, steps, alloc_spw * ctx->U);
  if (rc) return rc;
  rc = ensure_step_scratch(ctx, alloc_cap, alloc_spw, stride);
  if (rc) return rc;
  rc = ensure_adam_table(ctx, hp, (int64_t)step0 + 1, (int64_t)step0 + steps, st);
  if (rc) return rc;
  TrainParams p;
  fill_train_params(ctx, p, seed, hp);
  const bool piped = nwaves > 1 && !adaptive;
  // small batches (the reference configs' own 256): a persistent cluster kernel runs whole waves
  const bool small = hp->optimizer == RBPR_OPT_SGD && !adaptive && !(ctx->comm != nullptr && ctx->world > 1) &&
                     rbpr_small_batch_eligible(ctx, batch) && getenv("RBPR_NO_SMALL_BATCH") == nullptr;
  const bool fx_split_users = getenv("RBPR_FX_SPLIT_USERS") != nullptr;
  const bool fx_pdl = !fx_split_users && (getenv("RBPR_FX_PDL") == nullptr || atoi(getenv("RBPR_FX_PDL")) != 0);
  cudaStream_t prep_st = piped ? ctx->aux : st;
  if (piped) {
    RBPR_CUDA(ctx, cudaEventRecord(ctx->ev_inputs, st));
    RBPR_CUDA(ctx, cudaStreamWaitEvent(ctx->aux, ctx->ev_inputs, 0));
  }
  auto prepare = [&](int64_t w) -> int {
    NvtxRange nvtx_prep("rbpr.prepare_wave (count users, sample negatives)");
    const int b = (int)(w & 1);
    const int64_t off = wave_step0(w) * batch;
    const int64_t nw = (n - off) < wave_steps(w) * batch ? (n - off) : wave_steps(w) * batch;
    if (piped && w >= 2) RBPR_CUDA(ctx, cudaStreamWaitEvent(ctx->aux, ctx->ev_free[b], 0));
    if (host_idx != nullptr)
      RBPR_CUDA(ctx, cudaMemcpyAsync(triple_idx + off, host_idx + off, nw * sizeof(int64_t),
                                     cudaMemcpyHostToDevice, prep_st));
    if (host_neg_in != nullptr)
      RBPR_CUDA(ctx, cudaMemcpyAsync(neg_in + off, host_neg_in + off, nw * sizeof(int64_t),
                                     cudaMemcpyHostToDevice, prep_st));
    int r = count_batches(ctx, triple_idx + off, nw, batch, prep_st);
    if (r) return r;
    TrainParams q = p;
    q.triple_idx = triple_idx + off;
    q.batch = batch;
    if (small) {  // per-step item occurrence counters: the sampler designates one applier per item and step
      const int64_t need = wave_steps(w) * ctx->I;
      if (need > ctx->icnt_cap) {
        cudaFree(ctx->icnt);
        ctx->icnt = nullptr;
        ctx->icnt_cap = 0;
        const int64_t cap = alloc_spw * ctx->I > need ? alloc_spw * ctx->I : need;
        RBPR_CUDA(ctx, cudaMalloc(&ctx->icnt, (size_t)cap * sizeof(uint32_t)));
        ctx->icnt_cap = cap;
      }
      RBPR_CUDA(ctx, cudaMemsetAsync(ctx->icnt, 0, (size_t)need * sizeof(uint32_t), prep_st));
      q.icnt = ctx->icnt;
    }
    q.neg_in = neg_in ? neg_in + off : nullptr;
    q.neg_out = neg_out ? neg_out + off : nullptr;
    if (adaptive) {
      const uint64_t s_glob = step0 + (uint64_t)wave_step0(w);
      r = rbpr_internal_sample_adaptive_csr(ctx, q, ctx->records[b], nw, s_glob,
                                            (double)hp->adaptive_prob, hp->optimizer, prep_st);
      if (r) return r;
      // AdaptiveSampler.sample refreshes its snapshot AFTER the draw of every N-th call
      // (neg_samplers.py:122-123), i.e. from the item table before this step's update
      if (hp->adaptive_every > 0 && (s_glob + 1) % (uint64_t)hp->adaptive_every == 0) {
        r = rbpr_adaptive_update_stats(ctx, prep_st);
        if (r) return r;
      }
    } else {
      if (!small) {  // bpr_apply's user half reads a compact list instead of scanning the records
        RBPR_CUDA(ctx, cudaMemsetAsync(ctx->mh_count[b], 0, (size_t)wave_steps(w) * sizeof(uint32_t), prep_st));
        q.mh_list = ctx->mh_list[b];
        q.mh_count = ctx->mh_count[b];
      }
      r = run_sample(ctx, q, ctx->records[b], nw, step0 + (uint64_t)wave_step0(w), prep_st);
      if (r) return r;
    }
    if (piped) RBPR_CUDA(ctx, cudaEventRecord(ctx->ev_ready[b], ctx->aux));
    return 0;
  };
  if (piped) {
    rc = prepare(0);
    if (rc) return rc;
  }
  for (int64_t w = 0; w < nwaves; ++w) {
    const int b = (int)(w & 1);
    if (!piped) {  // just in time, on the caller's stream
      rc = prepare(w);
      if (rc) return rc;
    } else if (w + 1 < nwaves) {
      rc = prepare(w + 1);
      if (rc) return rc;
    }
    const int64_t off = wave_step0(w) * batch;
    const int64_t nw = (n - off) < wave_steps(w) * batch ? (n - off) : wave_steps(w) * batch;
    const int64_t wsteps = wave_steps(w);
    if (piped) RBPR_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_ready[b], 0));
    if (small) {
      NvtxRange nvtx_small("rbpr.small_batch_wave (cluster kernel)");
      // all steps of the wave in ONE launch of one thread-block cluster (train_small.cu)
      p.batch = batch;
      p.step = step0 + (uint64_t)wave_step0(w);
      rc = rbpr_launch_small_steps(ctx, p, reinterpret_cast<const int4*>(ctx->records[b]), nw, (int)wsteps,
                                   ctx->stats + wave_step0(w) * RBPR_STATS_PER_STEP, st);
      if (rc) return rc;
    }
    for (int64_t s = 0; s < wsteps && !small; ++s) {
      const int64_t soff = s * batch;
      p.n = (int)((nw - soff) < batch ? (nw - soff) : batch);
      p.step = step0 + (uint64_t)(wave_step0(w) + s);
      p.item_grad = ctx->item_grad;  // alternates between two buffers under the fused exchange
      p.bias_grad = ctx->item_bias ? ctx->item_grad + ctx->I * ctx->D : nullptr;
      p.pdl = 0;
      if (ctx->fx_bound && ctx->fx_wait_epoch != 0) {
        p.xwait_flags = ctx->fx_flags_local + RBPR_MAX_PEERS;
        p.xwait_n = ctx->world;
        p.xwait_epoch = ctx->fx_wait_epoch;
        p.pdl = (fx_pdl && s > 0) ? 1 : 0;  // the kernel before this one in the stream is the exchange of step s-1
      }
      const int4* recs = reinterpret_cast<const int4*>(ctx->records[b]) + soff;
      const int32_t* mh_l = adaptive ? nullptr : ctx->mh_list[b] + soff;
      const uint32_t* mh_c = adaptive ? nullptr : ctx->mh_count[b] + s;
      float4* parts = reinterpret_cast<float4*>(ctx->partials[b]) + s * stride;
      int nb = 0;
      {
        NvtxRange nvtx_a("rbpr.phase_a");
        rc = run_phase_a(ctx, p, hp, recs, parts, blocks, st, &nb);
      }
      if (rc) return rc;
      const int multi = (ctx->comm != nullptr && ctx->world > 1) ? 1 : 0;
      double* step_stats = ctx->stats + (wave_step0(w) + s) * RBPR_STATS_PER_STEP;
      if (multi && ctx->fx_bound && !fx_split_users) {
        // The one exchange of the step, ONE kernel: user half of the apply + statistics (local),
        // then reduce / item update / publish across ranks over peer memory (exchange.cu).
        NvtxRange nvtx_x("rbpr.exchange (users + reduce + item update across ranks)");
        const ApplyParams ua = make_apply_params(ctx, p.step, hp, 0, 0, recs, p.n, parts, nb * (kPhaseAThreads / 32),
                                                 step_stats, mh_l, mh_c);
        rc = rbpr_internal_fused_exchange(ctx, p.step, hp, st, &ua);
        if (rc) return rc;
      } else if (multi) {
        // NCCL fallback (or RBPR_FX_SPLIT_USERS=1): dense item gradient summed over ranks on the
        // caller's stream; the user half of the apply (local rows only) and the statistics run on a
        // side stream meanwhile.
        RBPR_CUDA(ctx, cudaEventRecord(ctx->ev_phase_a, st));
        RBPR_CUDA(ctx, cudaStreamWaitEvent(ctx->aux2, ctx->ev_phase_a, 0));
        rc = run_apply(ctx, p.step, hp, 0, 0, recs, p.n, ctx->aux2, parts, nb * (kPhaseAThreads / 32),
                       step_stats, mh_l, mh_c);
        if (rc) return rc;
        RBPR_CUDA(ctx, cudaEventRecord(ctx->ev_users, ctx->aux2));
        NvtxRange nvtx_x("rbpr.exchange (reduce + item update across ranks)");
        rc = rbpr_internal_exchange_apply(ctx, p.step, hp, st);
        if (rc) return rc;
        RBPR_CUDA(ctx, cudaStreamWaitEvent(st, ctx->ev_users, 0));
      } else {
        NvtxRange nvtx_b("rbpr.apply");
        rc = run_apply(ctx, p.step, hp, 0, 1, recs, p.n, st, parts, nb * (kPhaseAThreads / 32), step_stats, mh_l, mh_c);
        if (rc) return rc;
      }
    }
    if (piped) RBPR_CUDA(ctx, cudaEventRecord(ctx->ev_free[b], st));
  }
  rc = rbpr_internal_fx_wait(ctx, st);  // data parallel: the item table is complete when the call's work is
  if (rc) return rc;
  rc = rbpr_internal_fx_trace_report(ctx, st);
  if (rc) return rc;
  if (stats_out)
    RBPR_CUDA(ctx, cudaMemcpyAsync(stats_out, ctx->stats,
                                   steps * RBPR_STATS_PER_STEP * sizeof(double),
                                   cudaMemcpyDeviceToDevice, st));
  return 0;
}

extern "C" {

int rbpr_train_steps(rbpr_ctx* ctx, const int64_t* triple_idx, int64_t n, int64_t batch,
                     uint64_t seed, uint64_t step0, const rbpr_hparams* hp, const int64_t* neg_in,
                     int64_t* neg_out, double* stats_out, void* stream) {
  return train_steps_impl(ctx, const_cast<int64_t*>(triple_idx), n, batch, seed, step0, hp,
                          const_cast<int64_t*>(neg_in), neg_out, stats_out, nullptr, nullptr, stream);
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
  const bool inj = hp->sampler == RBPR_SAMPLER_INJECTED;
  if (inj && !neg_in_host) RBPR_FAIL(ctx, RBPR_ERR_ARG, "train_host: injected sampler needs neg_in");
  // host->device copies happen wave by wave on the preparation stream (train_steps_impl);
  // neg_in and neg_out may alias the same staging buffer: each position is read before written
  rc = train_steps_impl(ctx, ctx->stage_idx, n, batch, seed, step0, hp,
                        inj ? ctx->stage_neg : nullptr, neg_out_host ? ctx->stage_neg : nullptr,
                        nullptr, triple_idx_host, inj ? neg_in_host : nullptr, st);
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
  if (ctx->fx_bound) RBPR_FAIL(ctx, RBPR_ERR_STATE, "grad_step: not available once the peer-memory exchange is bound");
  if (n < 0 || n >= (1ll << 31)) RBPR_FAIL(ctx, RBPR_ERR_ARG, "grad_step: bad n");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  rc = ensure_capacity(ctx, n > 0 ? n : 1, 1, ctx->U);
  if (rc) return rc;
  RBPR_CUDA(ctx, cudaMemsetAsync(ctx->stats, 0, RBPR_STATS_PER_STEP * sizeof(double), st));
  if (n > 0) {
    if (!triple_idx) RBPR_FAIL(ctx, RBPR_ERR_ARG, "grad_step: null triple_idx");
    if (hp->sampler == RBPR_SAMPLER_INJECTED && !neg_in)
      RBPR_FAIL(ctx, RBPR_ERR_ARG, "grad_step: injected sampler needs neg_in");
    rc = count_batches(ctx, triple_idx, n, n, st);
    if (rc) return rc;
    int lanes, nv;
    rbpr_geometry(ctx->D, &lanes, &nv);
    int blocks = 1;
    rc = pick_blocks(ctx, hp, n, lanes, nv, &blocks);
    if (rc) return rc;
    const int stride = blocks * (kPhaseAThreads / 32);
    rc = ensure_step_scratch(ctx, n, 1, stride);
    if (rc) return rc;
    rc = ensure_adam_table(ctx, hp, (int64_t)step + 1, (int64_t)step + 1, st);
    if (rc) return rc;
    TrainParams p;
    fill_train_params(ctx, p, seed, hp);
    p.triple_idx = triple_idx;
    p.batch = n;
    p.neg_in = neg_in;
    p.neg_out = neg_out;
    rc = run_sample(ctx, p, ctx->records[0], n, step, st);
    if (rc) return rc;
      p.n = (int)n;
    p.step = step;
    int nb = 0;
    rc = run_phase_a(ctx, p, hp, reinterpret_cast<const int4*>(ctx->records[0]),
                     reinterpret_cast<float4*>(ctx->partials[0]), blocks, st, &nb);
    if (rc) return rc;
    // users are owned by this rank: finish the multi-occurrence ones now (items wait for the
    // all-reduce, rbpr_apply_item_grads); block 0 also sums the step statistics
    rc = run_apply(ctx, step, hp, 0, 0, reinterpret_cast<const int4*>(ctx->records[0]), (int)n, st,
                   reinterpret_cast<const float4*>(ctx->partials[0]), nb * (kPhaseAThreads / 32),
                   ctx->stats);
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
  if (ctx->fx_bound) RBPR_FAIL(ctx, RBPR_ERR_STATE, "item_grad_buffer: not available once the peer-memory exchange is bound");
  if (ptr) *ptr = ctx->item_grad;
  if (numel) *numel = ctx->I * ctx->D + (ctx->item_bias ? ctx->I : 0);
  return 0;
}

int rbpr_apply_item_grads(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, void* stream) {
  int rc = check_ready(ctx, hp);
  if (rc) return rc;
  if (ctx->fx_bound) RBPR_FAIL(ctx, RBPR_ERR_STATE, "apply_item_grads: not available once the peer-memory exchange is bound");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  rc = ensure_adam_table(ctx, hp, (int64_t)step + 1, (int64_t)step + 1, (cudaStream_t)stream);
  if (rc) return rc;
  return run_apply(ctx, step, hp, 1, 1, nullptr, 0, (cudaStream_t)stream);
}

int rbpr_sync_check(rbpr_ctx* ctx, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  return check_flag(ctx, (cudaStream_t)stream);
}

int rbpr_flush_lazy(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!hp) RBPR_FAIL(ctx, RBPR_ERR_ARG, "hparams is NULL");
  if (hp->optimizer == RBPR_OPT_SGD) return 0;
  if (!ctx->user_m || !ctx->user_last || (hp->optimizer == RBPR_OPT_ADAM && !ctx->user_v))
    RBPR_FAIL(ctx, RBPR_ERR_STATE, "optimizer state not bound");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  int rc = ensure_adam_table(ctx, hp, (int64_t)step + 1, (int64_t)step + 1, (cudaStream_t)stream);
  if (rc) return rc;
  rc = hp->optimizer == RBPR_OPT_ADAM   ? rbpr_launch_flush_users_adam(ctx, (int64_t)step, hp, lanes, nv, (cudaStream_t)stream)
       : hp->optimizer == RBPR_OPT_SGDM ? rbpr_launch_flush_users_sgdm(ctx, (int64_t)step, hp, lanes, nv, (cudaStream_t)stream)
                                        : rbpr_launch_flush_users_rms(ctx, (int64_t)step, hp, lanes, nv, (cudaStream_t)stream);
  if (rc) return rc;
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

}  // extern "C"
