// SGD-with-momentum instantiations of the fused training kernels (see train_kernels.cuh).
#include "train_kernels.cuh"

using namespace rbpr_dev;

int rbpr_launch_phase_a_sgdm(rbpr_ctx* ctx, const TrainParams& p, int lanes, int nv,
                              const int4* records, int blocks, cudaStream_t st) {
#define X(L, V)                                                                       \
  if (lanes == L && nv == V) {                                                        \
    launch_phase_a(bpr_phase_a<L, V, RBPR_OPT_SGDM>, blocks, st, p, records);    \
    return 0;                                                                         \
  }
  RBPR_FOR_EACH_GEOMETRY(X)
#undef X
  RBPR_FAIL(ctx, RBPR_ERR_ARG, "unsupported dim geometry lanes=%d nv=%d", lanes, nv);
}

// How many CTAs of the instantiation fit on one SM (the grid is sized as one resident wave).
int rbpr_phase_a_prepare_sgdm(rbpr_ctx* ctx, int dim, int lanes, int nv, int* blocks_per_sm) {
  (void)dim;
#define X(L, V)                                                                                \
  if (lanes == L && nv == V) {                                                                 \
    RBPR_CUDA(ctx, cudaOccupancyMaxActiveBlocksPerMultiprocessor(                              \
                       blocks_per_sm, bpr_phase_a<L, V, RBPR_OPT_SGDM>, kPhaseAThreads, 0));     \
    return 0;                                                                                  \
  }
  RBPR_FOR_EACH_GEOMETRY(X)
#undef X
  RBPR_FAIL(ctx, RBPR_ERR_ARG, "unsupported dim geometry lanes=%d nv=%d", lanes, nv);
}

int rbpr_launch_apply_sgdm(rbpr_ctx* ctx, const ApplyParams& p, int lanes, int nv,
                            cudaStream_t st) {
  const int groups_per_block = 256 / lanes;
  const int64_t work = (p.do_items ? p.I : 0) > (p.do_users ? (int64_t)p.n : 0) ? (p.do_items ? p.I : 0) : (p.do_users ? (int64_t)p.n : 0);
  const int64_t blocks64 = (work + groups_per_block - 1) / groups_per_block > 0 ? (work + groups_per_block - 1) / groups_per_block : 1;
  const int64_t maxb = (int64_t)ctx->sm_count * 8;
  const int blocks = (int)(blocks64 < maxb ? blocks64 : maxb);
#define X(L, V)                                                   \
  if (lanes == L && nv == V) {                                    \
    bpr_apply<L, V, RBPR_OPT_SGDM><<<blocks, 256, 0, st>>>(p);      \
    return 0;                                                     \
  }
  RBPR_FOR_EACH_GEOMETRY(X)
#undef X
  RBPR_FAIL(ctx, RBPR_ERR_ARG, "unsupported dim geometry lanes=%d nv=%d", lanes, nv);
}

int rbpr_launch_flush_users_sgdm(rbpr_ctx* ctx, int64_t step, const rbpr_hparams* hp, int lanes, int nv,
                                 cudaStream_t st) {
  const int blocks = ctx->sm_count * 8;
  const OptScalars h = {hp->lr, hp->beta1, hp->beta2, hp->eps, 0.f, 1.f};
#define X(L, V)                                                                                  \
  if (lanes == L && nv == V) {                                                                   \
    bpr_flush_users<L, V, RBPR_OPT_SGDM><<<blocks, 256, 0, st>>>(                                 \
        ctx->user_emb, ctx->user_m, ctx->user_v, ctx->user_last, ctx->U, ctx->D, step, ctx->adam_tab, h); \
    return 0;                                                                                    \
  }
  RBPR_FOR_EACH_GEOMETRY(X)
#undef X
  RBPR_FAIL(ctx, RBPR_ERR_ARG, "unsupported dim geometry lanes=%d nv=%d", lanes, nv);
}
