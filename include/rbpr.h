/*
 * rbpr.h — C ABI of the B200-native BPR hot path (librbpr.so).
 *
 * Plain C, no torch types: raw device/host pointers and sizes only.  Every entry
 * point returns 0 on success or a negative rbpr_status; the message for the last
 * failure of a context is available from rbpr_last_error().  No C++ exception
 * crosses this boundary.
 *
 * The reference (Nemexur/revisit-bpr) has no FFI: its hot path is a sequence of
 * PyTorch calls.  Each entry point below names the reference call sites it
 * replaces (paths relative to the reference repo root).
 *
 * Ownership: embedding tables, optimizer state and CSR arrays are OWNED BY THE
 * CALLER (PyTorch storages in the shipped host shell); the library borrows the
 * pointers, updates tables in place, and never frees or reallocates them.  The
 * context owns only scratch (per-wave records and user-occurrence counters, dense item- and
 * user-gradient accumulators, touched flags, step statistics, the adaptive sampler's snapshot) and
 * the NCCL communicator.
 *
 * Threading: a context is not thread-safe; one context per device per process.
 * All device work is enqueued on the cudaStream_t passed in (as void*).
 */
#ifndef RBPR_H_
#define RBPR_H_

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define RBPR_ABI_VERSION 9

typedef struct rbpr_ctx rbpr_ctx;

typedef enum {
  RBPR_OK = 0,
  RBPR_ERR_ARG = -1,   /* bad argument / shape mismatch / not bound */
  RBPR_ERR_CUDA = -2,  /* CUDA runtime error */
  RBPR_ERR_STATE = -3, /* call order violated */
  RBPR_ERR_DATA = -4,  /* invalid data (e.g. a user who has seen every item) */
  RBPR_ERR_COMM = -5   /* NCCL error / libnccl not loadable */
} rbpr_status;

typedef enum {
  RBPR_OPT_SGD = 0,     /* torch.optim.SGD(lr)                                              */
  RBPR_OPT_ADAM = 1,    /* torch.optim.Adam(lr, betas, eps)                                 */
  RBPR_OPT_SGDM = 2,    /* torch.optim.SGD(lr, momentum[, nesterov]), dampening 0           */
  RBPR_OPT_RMSPROP = 3  /* torch.optim.RMSprop(lr, alpha, eps), momentum 0, not centered    */
} rbpr_optimizer;

typedef enum {
  RBPR_SAMPLER_UNIFORM = 0,  /* uniform over {1..I-1} \ seen(u)                */
  RBPR_SAMPLER_WEIGHTED = 1, /* ∝ item_weight over {1..I-1} \ seen(u) (alias)  */
  RBPR_SAMPLER_INJECTED = 2, /* negatives supplied by the caller (neg_in)      */
  RBPR_SAMPLER_ADAPTIVE = 3  /* factor/rank adaptive (needs the current user rows: sampled
                                step by step, never ahead of the model)        */
} rbpr_sampler;

/* Hyper-parameters of one training call.
 * reg_*: replaces Model.regularization (revisit_bpr/models/bpr/model.py:70-93) —
 *        the host resolves `all`/`neg` defaulting before filling these.
 * optimizer fields: torch.optim.{SGD,Adam} param_groups as instantiated by
 *        experiments/bpr/exp.py:103-105. */
typedef struct {
  int32_t optimizer; /* rbpr_optimizer */
  int32_t sampler;   /* rbpr_sampler   */
  float lr;
  float beta1, beta2, eps; /* Adam: betas, eps.  SGDM: beta1 = momentum, beta2 != 0 selects nesterov.
                              RMSprop: beta2 = alpha, eps = eps.  (configs/RQ2/optimizers/*.yaml.j2) */
  float reg_user, reg_item, reg_neg;
  float adaptive_prob;    /* Geometric success probability (AdaptiveSampler sampling_prob)      */
  int32_t adaptive_every; /* refresh the item snapshot after every N-th sampled step (0: never) */
  int32_t reserved0;
} rbpr_hparams;

/* Per-step statistics written by the training entry points (doubles). */
#define RBPR_STATS_PER_STEP 4
/* [0] bpr_loss = Σ -logσ(x)      (model.py:65, loss.py:19-21)
 * [1] l2_reg   = Σ reg terms     (model.py:66)
 * [2] Σ|x|     (logits_diff numerator, experiments/bpr/exp.py:391)
 * [3] number of triples in the step */

int rbpr_abi_version(void);

/* Create / destroy a context on CUDA device `device`. */
int rbpr_create(int device, rbpr_ctx** out);
void rbpr_destroy(rbpr_ctx* ctx);
const char* rbpr_last_error(const rbpr_ctx* ctx);

/* Borrow the model tables.  Replaces nothing by itself; these are the storages of
 * MF._user_emb.weight (U,D), MF._item_emb.weight (I,D) and MF._item_bias (I,) or
 * NULL (revisit_bpr/models/bpr/model.py:96-129,147-153).  Row 0 of both tables is
 * the padding row.  D must be a multiple of 4, 4 <= D <= 1024; rows 16-byte aligned. */
int rbpr_bind_tables(rbpr_ctx* ctx, float* user_emb, int64_t num_users, float* item_emb,
                     int64_t num_items, int32_t dim, float* item_bias);

/* Borrow Adam state (same shapes as the tables; all zero-initialised by the caller).
 * user_last_step (U,) int32 holds, per user row, the number of optimizer steps already
 * applied to it (lazy dense-Adam catch-up).  Replaces the state that
 * torch.optim.Adam keeps (experiments/trainer.py:79). bias_* may be NULL iff no bias. */
int rbpr_bind_adam_state(rbpr_ctx* ctx, float* user_m, float* user_v, int32_t* user_last_step,
                         float* item_m, float* item_v, float* bias_m, float* bias_v);

/* State of the single-state optimizers (SGDM: momentum_buffer, RMSprop: square_avg), same shapes as
 * the tables, zero-initialised by the caller; user_last_step as for Adam (lazy dense semantics:
 * rows without gradient still move / decay, replayed when next touched or at rbpr_flush_lazy). */
int rbpr_bind_state1(rbpr_ctx* ctx, float* user_s, int32_t* user_last_step, float* item_s,
                     float* bias_s);

/* Borrow the training interaction matrix in CSR form (device pointers):
 * indptr (U+1,) int64, indices (nnz,) int32 item ids, sorted ascending within a row.
 * Triple t (0 <= t < nnz) is (user = row containing t, item = indices[t]): the COO
 * flattening of experiments/bpr/dataset.py:153-156; the row doubles as the user's
 * seen_items list (dataset.py:157-163) used to reject negatives.
 * Fails with RBPR_ERR_DATA if some user has seen every non-padding item (the reference's
 * torch.multinomial raises on such a row). Synchronises the stream once. */
int rbpr_bind_csr(rbpr_ctx* ctx, const int64_t* indptr, const int32_t* indices, int64_t num_users,
                  int64_t nnz, void* stream);

/* Optional: Walker alias table over items for the popularity-weighted static sampler
 * (experiments/bpr/exp.py:85-91,282-293: weights = count^alpha).  prob (I,) float in [0,1],
 * alias (I,) int32; entry 0 (padding) must have prob 0 and a non-zero alias. Device ptrs. */
int rbpr_bind_item_alias(rbpr_ctx* ctx, const float* prob, const int32_t* alias);

/* Adaptive sampler state (owned by the context): snapshot the item table, per-factor unbiased std
 * over items 1..I-1, every factor column sorted once (descending, ties by item id) + inverse.
 * Replaces AdaptiveSampler.update_stats (revisit_bpr/modules/neg_samplers.py:126-132) and
 * BPRExperiment._update_adaptive_stats (experiments/bpr/exp.py:344-354). */
int rbpr_adaptive_update_stats(rbpr_ctx* ctx, void* stream);
/* Borrow the state for inspection (device pointers: std (D), order (D,I), pos (D,I)). */
int rbpr_adaptive_stats(rbpr_ctx* ctx, const float** factor_std, const int32_t** order,
                        const int32_t** pos);

/* Adaptive negatives for a batch in the reference's layout: users (batch,) int64, seen
 * (batch,width) int64 0-padded; neg_out (batch,num) int64.  factor ~ |u_f|*std_f of the CURRENT
 * user row, rank ~ Geometric(sampling_prob) clamped to the number of unseen items, item = that
 * rank among the user's unseen items in the snapshot's factor order (top if u_f > 0 else bottom).
 * Replaces AdaptiveSampler.sample (neg_samplers.py:74-124) / _adaptive_sampling (exp.py:295-342):
 * no (B,I) scatter, no per-row argsort over I.  Counter-based draw of DESIGN.md §3.3.
 * hp / opt_step (hp may be NULL): with a stateful optimizer bound, user rows are updated lazily; the
 * draw then replays, in registers, the zero-gradient steps a row missed up to `opt_step` optimizer
 * steps, so it sees what dense torch.optim would hold (the reference reads model.get_features()). */
int rbpr_sample_adaptive_padded(rbpr_ctx* ctx, const int64_t* users, const int64_t* seen,
                                int64_t batch, int64_t width, int64_t num, double sampling_prob,
                                uint64_t seed, uint64_t step, int64_t* neg_out, uint64_t opt_step,
                                const rbpr_hparams* hp, void* stream);

/* Negative sampler alone: for each triple id t = triple_idx[k] draw
 * neg_out[k] = f(seed, step, t, CSR) — the counter-based specification in DESIGN.md §3
 * (Philox4x32-10, key = seed, subsequence = t, offset = (step<<8 | block)), rejecting
 * item 0 and the user's seen items.  Replaces UniformSampler.sample
 * (revisit_bpr/modules/neg_samplers.py:31-37) and BPRExperiment._static_sampling
 * (experiments/bpr/exp.py:290-293).  All pointers are device pointers. */
int rbpr_sample_negatives(rbpr_ctx* ctx, const int64_t* triple_idx, int64_t n, uint64_t seed,
                          uint64_t step, int32_t sampler, int64_t* neg_out, void* stream);

/* The fused training path.  Splits triple_idx[0..n) into ceil(n/batch) consecutive
 * minibatches; for each: sample negatives, gather (u,i+,i-), x = u·(i+ - i-) [+ bias],
 * loss = Σ softplus(-x) + L2, exact synchronous-minibatch gradients (duplicates summed),
 * optimizer update in place.  Step s uses global step number step0 + s for the sampler
 * and the Adam bias correction (step0 = optimizer steps taken so far).
 * Replaces, per step: BPRExperiment._train_batch (experiments/bpr/exp.py:356-367),
 * Model.forward train branch (model.py:48-68), MF.forward (model.py:131-145),
 * Loss.forward (loss.py:19-21), Model.regularization (model.py:70-93),
 * accelerator.backward + optimizer.step + zero_grad (experiments/trainer.py:76-81).
 *   triple_idx : device int64 (n,)  — a slice of the epoch permutation of [0,nnz)
 *   neg_in     : device int64 (n,) or NULL — required iff hp->sampler == INJECTED
 *   neg_out    : device int64 (n,) or NULL — negatives used, aligned with triple_idx
 *   stats_out  : device double (ceil(n/batch), RBPR_STATS_PER_STEP) or NULL */
int rbpr_train_steps(rbpr_ctx* ctx, const int64_t* triple_idx, int64_t n, int64_t batch,
                     uint64_t seed, uint64_t step0, const rbpr_hparams* hp, const int64_t* neg_in,
                     int64_t* neg_out, double* stats_out, void* stream);

/* Synchronise `stream` and report (then clear) any error raised on the device by earlier
 * asynchronous calls (sampler exhaustion, out-of-range triple index).  The *_host entry
 * points do this themselves. */
int rbpr_sync_check(rbpr_ctx* ctx, void* stream);

/* Same call with HOST buffers: triple_idx (and neg_in) are copied host→device, stats (and
 * neg_out) device→host, and the stream is synchronised before returning. */
int rbpr_train_steps_host(rbpr_ctx* ctx, const int64_t* triple_idx_host, int64_t n, int64_t batch,
                          uint64_t seed, uint64_t step0, const rbpr_hparams* hp,
                          const int64_t* neg_in_host, int64_t* neg_out_host,
                          double* stats_out_host, void* stream);

/* Data-parallel split of one step (experiments/launcher.py:35-73 + DDP allreduce inside
 * accelerator.backward, experiments/trainer.py:76).  Each rank calls rbpr_grad_step on its
 * local batch (users owned by the rank: user rows are updated locally), then all ranks
 * all-reduce(sum) the buffer returned by rbpr_item_grad_buffer, then each rank calls
 * rbpr_apply_item_grads (dense, identical on every rank). */
int rbpr_grad_step(rbpr_ctx* ctx, const int64_t* triple_idx, int64_t n, uint64_t seed,
                   uint64_t step, const rbpr_hparams* hp, const int64_t* neg_in, int64_t* neg_out,
                   double* stats_out, void* stream);
int rbpr_item_grad_buffer(rbpr_ctx* ctx, float** ptr, int64_t* numel);
int rbpr_apply_item_grads(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, void* stream);

/* ---- entry points shaped like the reference's Python call sites (explicit id batches) ---------- */

/* One training step on an explicit batch of (user, positive, negative) ids, as a DataLoader of the
 * reference delivers them (device int64, length n): forward + backward + optimizer update, exact
 * minibatch semantics, the three rows of every triple updated in place.  Row 0 of either table is
 * the padding row: it contributes to the logits but receives no embedding gradient
 * (nn.Embedding(padding_idx=0), example.py:311-321).
 * Replaces Model.forward train branch + Loss + regularization (model.py:48-93, loss.py:19-21),
 * accelerator.backward and optimizer.step/zero_grad (experiments/trainer.py:72-81,
 * example.py:175-178) for one batch.
 *   logits_out : device float (n,2) or NULL — (logits_pos, logits_neg) per triple, input order
 *   stats_out  : device double (RBPR_STATS_PER_STEP) or NULL
 *   step       : optimizer steps taken so far (Adam bias correction / lazy catch-up)            */
int rbpr_train_step_triples(rbpr_ctx* ctx, const int64_t* users, const int64_t* items,
                            const int64_t* negs, int64_t n, uint64_t step, const rbpr_hparams* hp,
                            float* logits_out, double* stats_out, void* stream);

/* logits[b,k] = <U[users[b]], V[items[b,k]]> + item_bias[items[b,k]] + user_bias[users[b]] for an
 * arbitrary (n_users, per_user) block of item ids; entries with mask == 0 become -1e13.
 * Replaces MF.forward (model.py:131-145) in eval mode and Model.forward's eval branch
 * (model.py:43-47) for any collator (AllItemsCollator, OnePosCollator, ManyPosCollator:
 * experiments/bpr/dataset.py:193-296).  mask, user_bias may be NULL.  Device pointers. */
int rbpr_pair_logits(rbpr_ctx* ctx, const int64_t* users, const int64_t* items, const float* mask,
                     int64_t n_users, int64_t per_user, const float* user_bias, float* out,
                     void* stream);

/* Negative sampling from the reference's batch layout: seen (batch,width) int64, 0-padded rows in
 * any order (batch["seen_items"], experiments/bpr/dataset.py:175-181).  neg_out (batch,num) int64.
 * Same counter-based draw as rbpr_sample_negatives with the slot (row*num + s) as subsequence;
 * draws of one row are independent.  Replaces _sampling_weights + torch.multinomial
 * (revisit_bpr/modules/neg_samplers.py:31-37,135-141; experiments/bpr/exp.py:282-293). */
int rbpr_sample_negatives_padded(rbpr_ctx* ctx, const int64_t* seen, int64_t batch, int64_t width,
                                 int64_t num_items, int64_t num, uint64_t seed, uint64_t step,
                                 int32_t sampler, int64_t* neg_out, void* stream);

/* Data-parallel communicator, one process per GPU.  Rank 0 obtains a 128-byte NCCL unique id and
 * hands it to the other ranks through any host channel (the shipped host code broadcasts it with
 * torch.distributed); every rank then joins.  Once a communicator with world > 1 exists,
 * rbpr_train_steps runs the data-parallel step by itself: phase A on the rank's own triples (users
 * sharded by owner), ONE ncclAllReduce(sum, fp32) of the dense item-gradient buffer on the
 * caller's stream, then the dense item update, identical on every rank — preparation waves still
 * overlap.  Replaces mp.spawn + init_process_group("nccl") + DDP (experiments/launcher.py:35-73,
 * experiments/bpr/exp.py:102, experiments/trainer.py:76). */
int rbpr_comm_unique_id(rbpr_ctx* ctx, void* out128);
int rbpr_comm_init(rbpr_ctx* ctx, int32_t world, int32_t rank, const void* id128);
/* The all-reduce alone, for callers driving rbpr_grad_step / rbpr_apply_item_grads themselves. */
int rbpr_comm_allreduce_item_grads(rbpr_ctx* ctx, void* stream);
int64_t rbpr_collective_count(const rbpr_ctx* ctx);

/* The exchange as ONE kernel over NVLink peer memory instead of ncclAllReduce + dense apply: every rank
 * owns a slice of the item rows, sums the ranks' gradients for it straight from their accumulators,
 * applies the optimizer and stores the updated rows into every replica (csrc/exchange.cu).  The host
 * shell gathers every rank's export blob (RBPR_IPC_BLOB_BYTES each, cudaIpc handles of the library's
 * gradient buffers and of the item table / bias storages) and hands all of them to every rank.
 * After a successful bind rbpr_train_steps / rbpr_train_step_triples use the fused exchange (the
 * communicator of rbpr_comm_init stays the fallback when this fails: no peer access, no IPC).
 * With a stateful optimizer the item table's optimizer state is SHARDED: only rows
 * [I*rank/world, I*(rank+1)/world) of item_m / item_v are kept current on a rank.
 * rbpr_grad_step / rbpr_item_grad_buffer / rbpr_apply_item_grads are not available once bound. */
#define RBPR_MAX_PEERS 8
#define RBPR_IPC_BLOB_BYTES 512
int rbpr_comm_ipc_export(rbpr_ctx* ctx, void* blob_out);
int rbpr_comm_ipc_bind(rbpr_ctx* ctx, const void* blobs, int32_t world, int32_t rank, void* stream);
/* The same exchange over a SYMMETRIC buffer the host shell allocates (same size on every rank, mapped
 * into every peer — torch.distributed._symmetric_memory, or cuMemCreate + cuMemMap + cuMulticastBindMem):
 * peer_bases[q] is this process's mapping of rank q's buffer (own one included), multicast_base the
 * multicast alias of the buffer or 0.  With a multicast alias the reduction runs INSIDE the NVSwitch
 * (multimem.ld_reduce: one 16-byte response per element instead of one load per rank) and the
 * updated rows leave the owner once (multimem.st), which roughly halves the NVLink bytes per GPU;
 * the order of the in-switch sum is the hardware's, so results may differ in the last bit from the
 * rank-order sum of the unicast path (replicas stay bit-identical: the owner computes, everybody
 * receives the same bits).  The buffer must hold rbpr_comm_symm_bytes(ctx) bytes, be 256-byte
 * aligned and ZERO before any rank binds.  The bind MOVES the item table and bias into the buffer
 * and returns their new device addresses: the caller re-points its tensors there (the old storage
 * is no longer read or written by the library).  Collective: every rank calls it. */
int64_t rbpr_comm_symm_bytes(const rbpr_ctx* ctx);
int rbpr_comm_symm_bind(rbpr_ctx* ctx, const uint64_t* peer_bases, uint64_t multicast_base, int32_t world,
                        int32_t rank, uint64_t* item_emb_out, uint64_t* item_bias_out, void* stream);
int32_t rbpr_fused_exchange_multicast(const rbpr_ctx* ctx);
int64_t rbpr_fused_exchange_count(const rbpr_ctx* ctx);

/* Bring every lazily-updated user row up to `step` optimizer steps (dense-Adam semantics
 * of torch.optim.Adam: rows with zero gradient still move).  Call before reading the user
 * table (eval, checkpoint).  No-op for SGD. */
int rbpr_flush_lazy(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, void* stream);

/* Full-catalog scoring + top-k + ranking metrics for a block of users.
 * Replaces eval Model.forward/MF.forward over AllItemsCollator batches
 * (model.py:43-47,131-145; experiments/bpr/dataset.py:274-296), _remove_seen_items
 * (experiments/bpr/exp.py:369-374), prepare_target (revisit_bpr/metrics/metric.py:110-113),
 * NDCG.compute (metrics/ndcg.py:8-13,69-78) and Recall.compute (metrics/recall.py:44-51).
 *   users            device int64 (n_users,)
 *   seen_indptr/idx  device CSR (row per entry of `users`, local numbering 0..n_users) of items
 *                    to mask to -1e13 (NULL,NULL = no masking); item 0 is always masked
 *   held_indptr/idx  device CSR (same local numbering) of held-out positives, sorted per row
 *   k_max            top-k depth kept (1..RBPR_MAX_TOPK)
 *   ks, n_ks         host array of cut-offs (each <= k_max)
 *   topk_items       device int32 (n_users,k_max) or NULL;  topk_scores device float or NULL
 *   ndcg_out, recall_out  device float (n_users, n_ks) or NULL                                  */
#define RBPR_MAX_TOPK 128
int rbpr_score_topk(rbpr_ctx* ctx, const int64_t* users, int64_t n_users,
                    const int64_t* seen_indptr, const int32_t* seen_indices,
                    const int64_t* held_indptr, const int32_t* held_indices, int32_t k_max,
                    const int32_t* ks, int32_t n_ks, int32_t* topk_items, float* topk_scores,
                    float* ndcg_out, float* recall_out, void* stream);

/* Every ranking metric the reference configs attach, for ALL cut-offs, from ONE scoring + ranking
 * pass over a block of users (what attach_metrics' update_handler, experiments/options.py:42-51,
 * obtains by calling up to 14 metric objects, each of which re-sorts the (B,I) logits:
 * revisit_bpr/metrics/{ndcg,recall,precision,map,fbeta}.py).  Arguments as rbpr_score_topk; every
 * output is device float (n_users, n_ks) or NULL:
 *   ndcg        gain 2^t-1, discount 1/log2(rank+2)  (ndcg.py:8-13,69-78)
 *   ndcg_linear gain t,     discount 1/(rank+1)      (ndcg.py:16-24)
 *   recall, precision                                 (recall.py:44-51, precision.py:44-51)
 *   map         average precision, denominator min(n_pos,k) if map_normalized else hits (map.py:45-64)
 *   topk_items  device int32 (n_users,k_max) or NULL */
typedef struct {
  float* ndcg;
  float* ndcg_linear;
  float* recall;
  float* precision;
  float* map;
  int32_t* topk_items;
  int32_t map_normalized;
  int32_t reserved0;
} rbpr_metric_outputs;
int rbpr_score_metrics(rbpr_ctx* ctx, const int64_t* users, int64_t n_users,
                       const int64_t* seen_indptr, const int32_t* seen_indices,
                       const int64_t* held_indptr, const int32_t* held_indices, int32_t k_max,
                       const int32_t* ks, int32_t n_ks, const rbpr_metric_outputs* out, void* stream);
/* Number of ranking passes so far (one per block of users per call). */
int64_t rbpr_topk_launch_count(const rbpr_ctx* ctx);
/* Which scoring path ran: blocks of users that went through the tensor-core candidate filter
 * (csrc/score_tc.cu), and users it handed back to the dense fp32 path (candidate overflow). */
int rbpr_score_path_counts(const rbpr_ctx* ctx, int64_t* tensor_passes, int64_t* overflow_users);

/* Dense scores for a block of users: out (n_users, I) float, masked like above.
 * The eval-mode Model.forward output `logits` (model.py:43-47) for drop-in callers that
 * want the full matrix (revisit_bpr.metrics on dense tensors). */
int rbpr_score_dense(rbpr_ctx* ctx, const int64_t* users, int64_t n_users,
                     const int64_t* seen_indptr, const int32_t* seen_indices, float* out,
                     void* stream);

/* Ranking metrics from DENSE tensors, the calling convention of revisit_bpr.metrics
 * (Metric.__call__/compute(output (B,I), target (B,I) multi-hot)): per row, top-k_max of `scores`
 * (ties -> lower column first), hits against `target`, NDCG@k / Recall@k / Precision@k for every
 * cut-off.  Replaces prepare_target (revisit_bpr/metrics/metric.py:110-113: a full argsort per
 * metric), NDCG.compute (metrics/ndcg.py:8-24,69-78; linear_gain selects gain_function="linear"),
 * Recall.compute (metrics/recall.py:44-51), Precision.compute (metrics/precision.py:44-51),
 * MAP.compute (metrics/map.py:45-64; map_normalized selects the min(n_pos,k) denominator).
 * A target value outside {0,1} raises RBPR_ERR_DATA at the next rbpr_sync_check
 * (validate_metric_inputs, metric.py:100-107).  Outputs (n_rows, n_ks) float or NULL;
 * topk_items (n_rows,k_max) int32 or NULL.  Device pointers, contiguous rows. */
int rbpr_topk_metrics_dense(rbpr_ctx* ctx, const float* scores, const float* target, int64_t n_rows,
                            int64_t n_cols, int32_t k_max, const int32_t* ks, int32_t n_ks,
                            int32_t linear_gain, float* ndcg_out, float* recall_out,
                            float* precision_out, float* map_out, int32_t map_normalized,
                            int32_t* topk_items, void* stream);

/* logits[b, seen[b,c]] = -1e13 and logits[b,0] = -1e13 for the reference's padded seen matrix
 * (batch,width) int64; logits (batch,n_cols) float, in place.  Replaces
 * BPRExperiment._remove_seen_items (experiments/bpr/exp.py:369-374). */
int rbpr_mask_seen_padded(rbpr_ctx* ctx, float* logits, const int64_t* seen, int64_t batch,
                          int64_t width, int64_t n_cols, void* stream);

/* Per-row ROC AUC over (positive, negative) pairs: positives target != 0, negatives target == 0 and
 * mask != 0 (mask may be NULL = all ones); 0/0 -> NaN.  Replaces RocAucManySlow.compute
 * (revisit_bpr/metrics/auc.py:149-166).  auc_out (n_rows,) float.  Device pointers. */
int rbpr_auc_dense(rbpr_ctx* ctx, const float* scores, const float* target, const float* mask,
                   int64_t n_rows, int64_t n_cols, float* auc_out, void* stream);

/* ---- item-neighbourhood logits models (reference revisit_bpr/models/bpr/model.py:156-251) --------
 * A seen entry (b,s) is "kept" unless seen[b,s] occurs among item[b,:] (model.py:184-190,230-235).
 * item (batch,n_items) int64, seen (batch,width) int64 (0-padded like the reference's collator; row
 * 0 is an ordinary row here, as in the reference, whose Parameter has no padding gradient mask),
 * keep (batch,width) uint8, logits / grad_logits (batch,n_items) float.  Ids outside
 * [0,num_items) raise the device flag (next rbpr_sync_check).  Device pointers.
 *
 * ItemKNN.forward (model.py:176-196): weights (num_items,hidden), bias (num_items) or NULL;
 * logits[b,i] = weights[item[b,i]] . SUM_{s kept} weights[seen[b,s]] + bias[item[b,i]].
 * keep_out and profile_out (batch,hidden: the kept-row sums) are what the backward needs. */
int rbpr_knn_forward(rbpr_ctx* ctx, const float* weights, int64_t num_items, int32_t hidden,
                     const float* bias, const int64_t* item, int64_t batch, int64_t n_items,
                     const int64_t* seen, int64_t width, uint8_t* keep_out, float* profile_out,
                     float* logits_out, void* stream);

/* Gradient of rbpr_knn_forward w.r.t. weights and bias (what autograd derives for model.py:176-196):
 * ACCUMULATES into grad_weights (num_items,hidden) and grad_bias (num_items, NULL = no bias). */
int rbpr_knn_backward(rbpr_ctx* ctx, const float* weights, int64_t num_items, int32_t hidden,
                      const int64_t* item, int64_t batch, int64_t n_items, const int64_t* seen,
                      int64_t width, const uint8_t* keep, const float* profile,
                      const float* grad_logits, float* grad_weights, float* grad_bias, void* stream);

/* FreeItemKNN.forward (model.py:222-248): weights (num_items,num_items);
 * logits[b,i] = SUM_{s kept} weights[item[b,i], seen[b,s]] + bias[item[b,i]]. */
int rbpr_freeknn_forward(rbpr_ctx* ctx, const float* weights, int64_t num_items, const float* bias,
                         const int64_t* item, int64_t batch, int64_t n_items, const int64_t* seen,
                         int64_t width, uint8_t* keep_out, float* logits_out, void* stream);

/* Gradient of rbpr_freeknn_forward: ACCUMULATES grad_logits[b,i] into
 * grad_weights[item[b,i], seen[b,s]] for every kept s, and into grad_bias[item[b,i]]. */
int rbpr_freeknn_backward(rbpr_ctx* ctx, int64_t num_items, const int64_t* item, int64_t batch,
                          int64_t n_items, const int64_t* seen, int64_t width, const uint8_t* keep,
                          const float* grad_logits, float* grad_weights, float* grad_bias,
                          void* stream);

/* ---- host-side JSONL ingest (no GPU involved; thread-safe; errors via rbpr_ingest_last_error) ---
 * One mmap'ed pass over the reference's on-disk files (bin/datasets/format-repro.sh:56-81); other
 * keys on a line are skipped (their values are stepped over structurally — strings, nested arrays
 * and objects — without being validated, and bytes after the closing brace of a line are ignored:
 * every file json.loads accepts yields the same values, some files it rejects are accepted).
 * Outputs are malloc'ed int64 arrays released with rbpr_ingest_free.
 * Replaces the per-line json.loads loops of experiments/bpr/dataset.py:16-24,183-190.
 *   pairs: {"<key_a>": int, "<key_b>": int}            -> a[n], b[n]
 *   lists: {"<key_a>": int, "<key_list>": [int, ...]}  -> a[rows], offsets[rows+1], values[...]   */
int rbpr_ingest_pairs(const char* path, const char* key_a, const char* key_b, int64_t** a_out,
                      int64_t** b_out, int64_t* n_out);
int rbpr_ingest_lists(const char* path, const char* key_a, const char* key_list, int64_t** a_out,
                      int64_t** offsets_out, int64_t** values_out, int64_t* n_rows_out);
void rbpr_ingest_free(void* p);
const char* rbpr_ingest_last_error(void);

/* Instrumentation: number of kernels this context has launched so far, and the device time
 * (ms, CUDA events on the launch stream) spent in the dominant training kernel since the
 * last reset, with the number of launches sampled.  Timing is off unless enabled; when on,
 * every 8th launch of the kernel is bracketed by an event pair. */
int64_t rbpr_launch_count(const rbpr_ctx* ctx);
int rbpr_kernel_timing(rbpr_ctx* ctx, int32_t enable);
int rbpr_kernel_time_ms(rbpr_ctx* ctx, double* ms_out, int64_t* launches_out);

#ifdef __cplusplus
}
#endif
#endif /* RBPR_H_ */
