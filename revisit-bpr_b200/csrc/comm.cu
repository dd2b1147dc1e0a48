// NCCL glue of the data-parallel step (SURVEY.md §8 e): ONE ncclAllReduce(sum, fp32) of the dense
// item-gradient buffer per step, on the caller's stream, between phase A and the (then dense) item
// update.  Replaces the DDP all-parameter gradient all-reduce of the reference
// (experiments/launcher.py:59-70, accelerator.backward at experiments/trainer.py:76).
// NCCL is bound at run time (dlopen of the libnccl.so.2 that PyTorch ships and has usually already
// loaded), so librbpr.so carries no link-time dependency on it.
#include <dlfcn.h>

#include "common.cuh"

namespace {

typedef struct { char internal[128]; } nccl_unique_id;  // ncclUniqueId (NCCL_UNIQUE_ID_BYTES = 128)
typedef void* nccl_comm_t;
typedef int (*fn_get_unique_id)(nccl_unique_id*);
typedef int (*fn_comm_init_rank)(nccl_comm_t*, int, nccl_unique_id, int);
typedef int (*fn_all_reduce)(const void*, void*, size_t, int, int, nccl_comm_t, cudaStream_t);
typedef int (*fn_comm_destroy)(nccl_comm_t);
typedef int (*fn_comm_abort)(nccl_comm_t);
typedef const char* (*fn_error_string)(int);
constexpr int kNcclFloat32 = 7, kNcclSum = 0;

struct NcclApi {
  void* handle = nullptr;
  fn_get_unique_id get_unique_id = nullptr;
  fn_comm_init_rank comm_init_rank = nullptr;
  fn_all_reduce all_reduce = nullptr;
  fn_comm_destroy comm_destroy = nullptr;
  fn_comm_abort comm_abort = nullptr;
  fn_error_string error_string = nullptr;
};

NcclApi* nccl_api(rbpr_ctx* ctx) {
  static NcclApi api;
  if (api.handle) return &api;
  const char* env = getenv("RBPR_NCCL_LIB");
  void* h = nullptr;
  if (env && *env) h = dlopen(env, RTLD_NOW | RTLD_GLOBAL);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_NOLOAD);
  if (!h) h = dlopen("libnccl.so.2", RTLD_NOW | RTLD_GLOBAL);
  if (!h) {
    ctx->err = std::string("cannot load libnccl.so.2 (set RBPR_NCCL_LIB): ") + dlerror();
    return nullptr;
  }
  api.get_unique_id = (fn_get_unique_id)dlsym(h, "ncclGetUniqueId");
  api.comm_init_rank = (fn_comm_init_rank)dlsym(h, "ncclCommInitRank");
  api.all_reduce = (fn_all_reduce)dlsym(h, "ncclAllReduce");
  api.comm_destroy = (fn_comm_destroy)dlsym(h, "ncclCommDestroy");
  api.comm_abort = (fn_comm_abort)dlsym(h, "ncclCommAbort");
  api.error_string = (fn_error_string)dlsym(h, "ncclGetErrorString");
  if (!api.get_unique_id || !api.comm_init_rank || !api.all_reduce || !api.comm_destroy) {
    ctx->err = "libnccl.so.2 lacks an expected symbol";
    return nullptr;
  }
  api.handle = h;
  return &api;
}

}  // namespace

#define RBPR_NCCL(ctx, api, expr)                                                              \
  do {                                                                                         \
    int _r = (expr);                                                                           \
    if (_r != 0)                                                                               \
      RBPR_FAIL(ctx, RBPR_ERR_COMM, "%s failed: %s", #expr,                                    \
                (api)->error_string ? (api)->error_string(_r) : "nccl error");                 \
  } while (0)

int rbpr_internal_allreduce_item_grads(rbpr_ctx* ctx, cudaStream_t st) {
  if (!ctx->comm || ctx->world <= 1) return 0;
  NcclApi* api = nccl_api(ctx);
  if (!api) return RBPR_ERR_COMM;
  const size_t numel = (size_t)ctx->I * ctx->D + (ctx->item_bias ? (size_t)ctx->I : 0);
  RBPR_NCCL(ctx, api, api->all_reduce(ctx->item_grad, ctx->item_grad, numel, kNcclFloat32, kNcclSum,
                                      (nccl_comm_t)ctx->comm, st));
  ctx->collectives++;
  return 0;
}

void rbpr_internal_comm_destroy(rbpr_ctx* ctx) {
  if (!ctx->comm) return;
  NcclApi* api = nccl_api(ctx);
  // ncclCommAbort, not ncclCommDestroy: contexts die when Python collects them, at a different moment
  // on every rank; a teardown that waits for the peers deadlocks against whatever collective the
  // other rank has moved on to (seen with two experiments run back to back in one process group)
  if (api) {
    if (api->comm_abort) api->comm_abort((nccl_comm_t)ctx->comm);
    else api->comm_destroy((nccl_comm_t)ctx->comm);
  }
  ctx->comm = nullptr;
}

extern "C" {

int rbpr_comm_unique_id(rbpr_ctx* ctx, void* out128) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!out128) RBPR_FAIL(ctx, RBPR_ERR_ARG, "comm_unique_id: null output");
  NcclApi* api = nccl_api(ctx);
  if (!api) return RBPR_ERR_COMM;
  nccl_unique_id id;
  RBPR_NCCL(ctx, api, api->get_unique_id(&id));
  memcpy(out128, &id, sizeof(id));
  return 0;
}

int rbpr_comm_init(rbpr_ctx* ctx, int32_t world, int32_t rank, const void* id128) {
  if (!ctx) return RBPR_ERR_ARG;
  if (world < 1 || rank < 0 || rank >= world || !id128)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "comm_init: bad world/rank/id");
  if (ctx->comm) RBPR_FAIL(ctx, RBPR_ERR_STATE, "comm_init: communicator already initialised");
  NcclApi* api = nccl_api(ctx);
  if (!api) return RBPR_ERR_COMM;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  nccl_unique_id id;
  memcpy(&id, id128, sizeof(id));
  nccl_comm_t comm = nullptr;
  RBPR_NCCL(ctx, api, api->comm_init_rank(&comm, world, id, rank));
  ctx->comm = comm;
  ctx->world = world;
  ctx->rank = rank;
  return 0;
}

int rbpr_comm_allreduce_item_grads(rbpr_ctx* ctx, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!ctx->item_grad) RBPR_FAIL(ctx, RBPR_ERR_STATE, "tables not bound");
  if (ctx->fx_bound) RBPR_FAIL(ctx, RBPR_ERR_STATE, "allreduce_item_grads: not available once the peer-memory exchange is bound");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  return rbpr_internal_allreduce_item_grads(ctx, (cudaStream_t)stream);
}

int64_t rbpr_collective_count(const rbpr_ctx* ctx) { return ctx ? ctx->collectives : 0; }

}  // extern "C"
