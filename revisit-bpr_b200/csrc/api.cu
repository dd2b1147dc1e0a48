// Context management and table binding for librbpr.so (C ABI in include/rbpr.h).
#include "common.cuh"

void rbpr_internal_comm_destroy(rbpr_ctx* ctx);  // comm.cu
void rbpr_internal_fx_destroy(rbpr_ctx* ctx);    // exchange.cu

extern "C" {

int rbpr_abi_version(void) { return RBPR_ABI_VERSION; }

int rbpr_create(int device, rbpr_ctx** out) {
  if (!out) return RBPR_ERR_ARG;
  *out = nullptr;
  int count = 0;
  if (cudaGetDeviceCount(&count) != cudaSuccess || device < 0 || device >= count)
    return RBPR_ERR_CUDA;
  if (cudaSetDevice(device) != cudaSuccess) return RBPR_ERR_CUDA;
  cudaDeviceProp prop;
  if (cudaGetDeviceProperties(&prop, device) != cudaSuccess) return RBPR_ERR_CUDA;
  if (prop.major != 10) return RBPR_ERR_CUDA;  // sm_100a cubin only: no fallback path exists
  rbpr_ctx* ctx = new (std::nothrow) rbpr_ctx();
  if (!ctx) return RBPR_ERR_CUDA;
  ctx->device = device;
  ctx->sm_count = prop.multiProcessorCount;
  if (cudaMalloc(&ctx->flag, sizeof(int32_t)) != cudaSuccess ||
      cudaMemset(ctx->flag, 0, sizeof(int32_t)) != cudaSuccess) {
    delete ctx;
    return RBPR_ERR_CUDA;
  }
  bool ok = cudaStreamCreateWithFlags(&ctx->aux, cudaStreamNonBlocking) == cudaSuccess &&
            cudaStreamCreateWithFlags(&ctx->aux2, cudaStreamNonBlocking) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ev_phase_a, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ev_users, cudaEventDisableTiming) == cudaSuccess &&
            cudaEventCreateWithFlags(&ctx->ev_inputs, cudaEventDisableTiming) == cudaSuccess;
  for (int b = 0; b < 2 && ok; ++b)
    ok = cudaEventCreateWithFlags(&ctx->ev_ready[b], cudaEventDisableTiming) == cudaSuccess &&
         cudaEventCreateWithFlags(&ctx->ev_free[b], cudaEventDisableTiming) == cudaSuccess;
  if (!ok) {
    rbpr_destroy(ctx);
    return RBPR_ERR_CUDA;
  }
  *out = ctx;
  return 0;
}

void rbpr_destroy(rbpr_ctx* ctx) {
  if (!ctx) return;
  cudaSetDevice(ctx->device);
  rbpr_internal_fx_destroy(ctx);
  rbpr_internal_comm_destroy(ctx);
  cudaFree(ctx->coo_user);
  cudaFree(ctx->bloom);
  cudaFree(ctx->item_grad);
  cudaFree(ctx->user_grad);
  cudaFree(ctx->touched);
  cudaFree(ctx->icnt);
  cudaFree(ctx->ord);
  cudaFree(ctx->cnt);
  cudaFree(ctx->stats);
  for (int b = 0; b < 2; ++b) {
    cudaFree(ctx->partials[b]);
    cudaFree(ctx->mh_list[b]);
    cudaFree(ctx->mh_count[b]);
    cudaFree(ctx->records[b]);
    if (ctx->ev_ready[b]) cudaEventDestroy(ctx->ev_ready[b]);
    if (ctx->ev_free[b]) cudaEventDestroy(ctx->ev_free[b]);
  }
  if (ctx->ev_inputs) cudaEventDestroy(ctx->ev_inputs);
  if (ctx->aux) cudaStreamDestroy(ctx->aux);
  if (ctx->aux2) cudaStreamDestroy(ctx->aux2);
  if (ctx->ev_phase_a) cudaEventDestroy(ctx->ev_phase_a);
  if (ctx->ev_users) cudaEventDestroy(ctx->ev_users);
  cudaFree(ctx->flag);
  cudaFree(ctx->stage_idx);
  cudaFree(ctx->stage_neg);
  cudaFree(ctx->score_buf);
  cudaFree(ctx->tc_items);
  cudaFree(ctx->tc_users);
  cudaFree(ctx->tc_gmax);
  cudaFree(ctx->tc_mask);
  cudaFree(ctx->tc_small);
  cudaFree(ctx->tc_cand);
  cudaFree(ctx->tc_ovf_users);
  cudaFree(ctx->adam_tab);
  cudaFree(ctx->ad_snap);
  cudaFree(ctx->ad_keys);
  cudaFree(ctx->ad_keys_sorted);
  cudaFree(ctx->ad_std);
  cudaFree(ctx->ad_ids);
  cudaFree(ctx->ad_order);
  cudaFree(ctx->ad_pos);
  cudaFree(ctx->ad_tmp);
  for (cudaEvent_t e : ctx->ev) cudaEventDestroy(e);
  delete ctx;
}

const char* rbpr_last_error(const rbpr_ctx* ctx) { return ctx ? ctx->err.c_str() : "null context"; }

int rbpr_bind_tables(rbpr_ctx* ctx, float* user_emb, int64_t num_users, float* item_emb,
                     int64_t num_items, int32_t dim, float* item_bias) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!user_emb || !item_emb) RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_tables: null table pointer");
  if (ctx->fx_bound) RBPR_FAIL(ctx, RBPR_ERR_STATE, "bind_tables: peer-memory exchange is bound to the current tables");
  if (num_users < 2 || num_items < 3)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_tables: need >=2 user rows and >=3 item rows (row 0 pads)");
  if (num_items >= (1ll << 31) || num_users >= (1ll << 31))
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_tables: ids must fit int32");
  if (dim < 4 || dim > 1024 || dim % 4 != 0)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_tables: dim=%d must be a multiple of 4 in [4,1024]", dim);
  if (((uintptr_t)user_emb | (uintptr_t)item_emb) & 15u)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_tables: tables must be 16-byte aligned");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  cudaFree(ctx->item_grad);
  cudaFree(ctx->user_grad);
  cudaFree(ctx->touched);
  cudaFree(ctx->icnt);
  ctx->icnt = nullptr;
  ctx->icnt_cap = 0;
  ctx->item_grad = nullptr;
  ctx->user_grad = nullptr;
  ctx->touched = nullptr;
  const size_t gbytes = ((size_t)num_items * dim + (size_t)num_items) * sizeof(float);
  RBPR_CUDA(ctx, cudaMalloc(&ctx->item_grad, gbytes));
  RBPR_CUDA(ctx, cudaMemset(ctx->item_grad, 0, gbytes));
  const size_t ubytes = (size_t)num_users * dim * sizeof(float);
  RBPR_CUDA(ctx, cudaMalloc(&ctx->user_grad, ubytes));
  RBPR_CUDA(ctx, cudaMemset(ctx->user_grad, 0, ubytes));
  RBPR_CUDA(ctx, cudaMalloc(&ctx->touched, num_items * sizeof(uint32_t)));
  RBPR_CUDA(ctx, cudaMemset(ctx->touched, 0, num_items * sizeof(uint32_t)));
  ctx->user_emb = user_emb;
  ctx->item_emb = item_emb;
  ctx->item_bias = item_bias;
  ctx->U = num_users;
  ctx->I = num_items;
  ctx->D = dim;
  for (int o = 0; o < 4; ++o) ctx->phase_a_blocks_per_sm[o] = 0;
  return 0;
}

int rbpr_bind_adam_state(rbpr_ctx* ctx, float* user_m, float* user_v, int32_t* user_last_step,
                         float* item_m, float* item_v, float* bias_m, float* bias_v) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!user_m || !user_v || !user_last_step || !item_m || !item_v)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_adam_state: null pointer");
  if (((uintptr_t)user_m | (uintptr_t)user_v | (uintptr_t)item_m | (uintptr_t)item_v) & 15u)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_adam_state: state must be 16-byte aligned");
  ctx->user_m = user_m;
  ctx->user_v = user_v;
  ctx->user_last = user_last_step;
  ctx->item_m = item_m;
  ctx->item_v = item_v;
  ctx->bias_m = bias_m;
  ctx->bias_v = bias_v;
  return 0;
}

int rbpr_bind_state1(rbpr_ctx* ctx, float* user_s, int32_t* user_last_step, float* item_s,
                     float* bias_s) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!user_s || !user_last_step || !item_s) RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_state1: null pointer");
  if (((uintptr_t)user_s | (uintptr_t)item_s) & 15u)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "bind_state1: state must be 16-byte aligned");
  ctx->user_m = user_s;
  ctx->user_v = nullptr;
  ctx->user_last = user_last_step;
  ctx->item_m = item_s;
  ctx->item_v = nullptr;
  ctx->bias_m = bias_s;
  ctx->bias_v = nullptr;
  return 0;
}

int64_t rbpr_launch_count(const rbpr_ctx* ctx) { return ctx ? ctx->launches : 0; }

int rbpr_kernel_timing(rbpr_ctx* ctx, int32_t enable) {
  if (!ctx) return RBPR_ERR_ARG;
  ctx->timing = enable != 0;
  if (ctx->timing) {  // pre-create events so that no cudaEventCreate lands inside a timed region
    RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
    while (ctx->ev.size() < 2048) {
      cudaEvent_t e;
      RBPR_CUDA(ctx, cudaEventCreate(&e));
      ctx->ev.push_back(e);
    }
  }
  return 0;
}

int rbpr_kernel_time_ms(rbpr_ctx* ctx, double* ms_out, int64_t* launches_out) {
  if (!ctx) return RBPR_ERR_ARG;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  for (size_t k = 0; k + 1 < ctx->ev_used; k += 2) {
    RBPR_CUDA(ctx, cudaEventSynchronize(ctx->ev[k + 1]));
    float ms = 0.f;
    RBPR_CUDA(ctx, cudaEventElapsedTime(&ms, ctx->ev[k], ctx->ev[k + 1]));
    ctx->timed_ms += ms;
    ctx->timed_launches++;
  }
  ctx->ev_used = 0;
  if (ms_out) *ms_out = ctx->timed_ms;
  if (launches_out) *launches_out = ctx->timed_launches;
  ctx->timed_ms = 0.0;
  ctx->timed_launches = 0;
  return 0;
}

}  // extern "C"
