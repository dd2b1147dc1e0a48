// The data-parallel step's ONE exchange as ONE kernel over NVLink peer memory (sm_100a, NVSwitch):
// reduce the ranks' dense item gradients, apply the optimizer, and publish the updated item rows to
// every replica — tile by tile, no separate all-reduce pass, no dense re-read of the reduced buffer.
//
//   every rank r owns a contiguous slice of the item rows, [I*r/W, I*(r+1)/W):
//     g   = sum over ranks q = 0..W-1 of grad_q[row]      (128-bit loads from the peers' accumulators
//                                                           through NVLink; fixed order: deterministic)
//     row = optimizer(row, g [, m, v of the slice])        (SGD / Adam / SGD-momentum / RMSprop; the
//                                                           optimizer state of the item table is
//                                                           SHARDED: only the owner keeps its slice)
//     item_q[row] = row   for every rank q                 (128-bit stores into the peers' tables)
//   so replicas are bit-identical by construction, each rank sweeps 1/W of the table (the dense
//   Adam sweep of the 8-GPU MSD configuration shrinks 8x), and per-GPU NVLink traffic is one
//   gradient buffer in + one table slice out per step instead of an all-reduce plus a dense apply.
//
// Synchronisation: two gradient accumulators used alternately (step s accumulates into buf[s&1], the
// exchange of step s clears the local buf[(s+1)&1], which the peers finished reading one step ago),
// and two cross-rank barriers per step (flags in peer memory, st.release.sys / ld.acquire.sys):
// B1 "every rank's phase A has landed" before the reduce, B2 "every rank's rows have landed" before
// the next phase A.  Memory is shared with cudaIpc handles exchanged by the host shell
// (rbpr_comm_ipc_export / rbpr_comm_ipc_bind); NCCL stays the fallback when peer access or IPC is
// not available.  Replaces the DDP gradient all-reduce + dense optimizer step of the reference
// (experiments/launcher.py:59-70, experiments/trainer.py:76-79).
#include "train_kernels.cuh"

using namespace rbpr_dev;

namespace {

constexpr int kMaxWorld = RBPR_MAX_PEERS;

struct IpcBlob {  // what one rank tells the others (host-exchanged, fixed size: RBPR_IPC_BLOB_BYTES)
  cudaIpcMemHandle_t sym;    // library-owned: buf[0] | buf[1] | flags
  cudaIpcMemHandle_t item;   // base allocation holding the item table (PyTorch storage)
  cudaIpcMemHandle_t bias;   // base allocation holding the item bias (may equal `item`)
  uint64_t item_off, bias_off, gbytes;
  int32_t has_bias, device;
  int64_t I;
  int32_t D, pad;
};
static_assert(sizeof(IpcBlob) <= RBPR_IPC_BLOB_BYTES, "blob too large");

struct ExchangeParams {
  const float* gsrc[kMaxWorld];  // every rank's accumulator of this step's parity (own one included)
  float* idst[kMaxWorld];        // every rank's item table
  float* bdst[kMaxWorld];        // every rank's item bias (or null)
  float* gzero;                  // local accumulator of the OTHER parity: cleared here
  float* item_m;                 // local optimizer state of the item table (slice rows are current)
  float* item_v;
  float* bias_m;
  float* bias_v;
  // barriers folded into this kernel: B1 (signal + wait at the start), B2 (signal by the last CTA)
  uint32_t* const* flags;        // device array: flags[q] = rank q's flag words: [0,W) = B1, [W_MAX, W_MAX+W) = B2
  uint32_t epoch;                // this exchange's barrier epoch
  uint32_t* done;                // CTAs finished (self-resetting counter)
  int32_t* err;
  int world, rank;
  int64_t I, lo, hi;             // rows; [lo, hi) is this rank's slice
  int D;
  uint64_t step;
  float lr, beta1, beta2, eps;
  const float2* adam_tab;
};

__device__ __forceinline__ float4 ldcg4x(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }

template <int LANES, int NV, int OPT>
__global__ void __launch_bounds__(256) bpr_exchange_apply(const ExchangeParams p) {
  const Group<LANES> g;
  const int D = p.D;
  const int64_t groups = ((int64_t)gridDim.x * blockDim.x) / LANES;
  const int64_t gid = ((int64_t)blockIdx.x * blockDim.x + threadIdx.x) / LANES;
  OptScalars h = {p.lr, p.beta1, p.beta2, p.eps, 0.f, 1.f};
  if (OPT == RBPR_OPT_ADAM) {
    const float2 t = __ldg(p.adam_tab + (p.step + 1));
    h.step_size = t.x;
    h.bc2_sqrt = t.y;
  }
  const int W = p.world;
  // ---- B1: "every rank's phase A has landed".  This kernel runs after this rank's phase A (stream
  // order), so CTA 0 publishes the arrival; every CTA then waits for all peers' arrivals.
  if (blockIdx.x == 0 && threadIdx.x < W) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flags[threadIdx.x] + p.rank), "r"(p.epoch) : "memory");
  }
  wait_peer_flags(p.flags[p.rank], W, p.epoch, p.err);
  // ---- this rank's slice: reduce over ranks, update, publish -------------------------------------
  for (int64_t r = p.lo + gid; r < p.hi; r += groups) {
    float4 gr[NV];
#pragma unroll
    for (int v = 0; v < NV; ++v) gr[v] = f4zero();
    for (int q = 0; q < W; ++q) {  // fixed order: every run sums in the same order
      const float* src = p.gsrc[q] + r * D;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = 4 * (g.gl + LANES * v);
        if (c < D) gr[v] = add4(gr[v], ldcg4x(src + c));
      }
    }
    const float* prow = p.idst[p.rank] + r * D;
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      const int c = 4 * (g.gl + LANES * v);
      if (c >= D) continue;
      float4 pp = ldcg4x(prow + c);
      if (OPT == RBPR_OPT_SGD) {
        pp.x -= p.lr * gr[v].x;
        pp.y -= p.lr * gr[v].y;
        pp.z -= p.lr * gr[v].z;
        pp.w -= p.lr * gr[v].w;
      } else {
        float4 m = ld4(p.item_m + r * D + c);
        float4 vv = opt_has_s2(OPT) ? ld4(p.item_v + r * D + c) : f4zero();
        opt4<OPT>(pp, m, vv, gr[v], h);
        st4(p.item_m + r * D + c, m);
        if (opt_has_s2(OPT)) st4(p.item_v + r * D + c, vv);
      }
      for (int q = 0; q < W; ++q) st4(p.idst[q] + r * D + c, pp);
    }
    if (g.gl == 0 && p.bdst[p.rank] != nullptr) {
      float gb = 0.f;
      for (int q = 0; q < W; ++q) gb += __ldcg(p.gsrc[q] + p.I * D + r);
      float b = __ldcg(p.bdst[p.rank] + r);
      if (OPT == RBPR_OPT_SGD) {
        b -= p.lr * gb;
      } else {
        float m = p.bias_m[r], vv = opt_has_s2(OPT) ? p.bias_v[r] : 0.f;
        opt1<OPT>(b, m, vv, gb, h);
        p.bias_m[r] = m;
        if (opt_has_s2(OPT)) p.bias_v[r] = vv;
      }
      for (int q = 0; q < W; ++q) p.bdst[q][r] = b;
    }
  }
  // ---- clear the accumulator the NEXT step uses (local; the peers read it one step ago) -----------
  const int64_t vecs = (p.I * D + (p.bdst[p.rank] != nullptr ? p.I : 0) + 3) / 4;  // buffer is padded to 16 B
  float4* z = reinterpret_cast<float4*>(p.gzero);
  for (int64_t k = (int64_t)blockIdx.x * blockDim.x + threadIdx.x; k < vecs; k += (int64_t)gridDim.x * blockDim.x)
    z[k] = f4zero();
  // ---- B2: "this rank's rows have landed everywhere": the last CTA to finish publishes it; the next
  // kernel that reads item rows (phase A, or the wait at the end of the call) waits for all ranks
  __syncthreads();
  if (threadIdx.x == 0) {
    __threadfence_system();
    if (atomicAdd(p.done, 1u) == gridDim.x - 1) {
      *p.done = 0u;
      __threadfence_system();
      for (int q = 0; q < W; ++q)
        asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flags[q] + kMaxWorld + p.rank), "r"(p.epoch) : "memory");
    }
  }
}

// End of a call: everything that reads the item table next (scoring, torch) runs after every rank's
// rows of the last step have landed.
__global__ void xwait_kernel(const uint32_t* flags, int n, uint32_t epoch, int32_t* err) {
  wait_peer_flags(flags, n, epoch, err);
}

// Cross-rank barrier on the caller's stream: thread q publishes this rank's arrival (a growing
// epoch) into rank q's flag word for this rank and waits for rank q's arrival in its own.  The
// fence + release make everything earlier kernels of this stream wrote (local reds, peer stores)
// visible to a rank that has seen the epoch.  Bounded spin: ranks out of step raise an error
// instead of hanging the GPU.
__global__ void xrank_barrier(uint32_t* const* flags, int world, int rank, uint32_t epoch, int32_t* err) {
  const int q = threadIdx.x;
  if (q >= world) return;
  __threadfence_system();
  uint32_t* theirs = flags[q] + 2 * kMaxWorld + rank;  // words [2*W_MAX, 3*W_MAX): this standalone barrier
  asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(theirs), "r"(epoch) : "memory");
  const uint32_t* mine = flags[rank] + 2 * kMaxWorld + q;
  for (uint32_t spin = 0; spin < (1u << 24); ++spin) {
    uint32_t seen;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(seen) : "l"(mine) : "memory");
    if ((int32_t)(seen - epoch) >= 0) return;
    __nanosleep(64);
  }
  atomicExch(err, 10);
}

typedef int (*fn_cuMemGetAddressRange)(unsigned long long*, size_t*, unsigned long long);

int base_of(rbpr_ctx* ctx, const void* ptr, void** base, uint64_t* off) {
  static fn_cuMemGetAddressRange fn = nullptr;
  if (!fn) {
    void* sym = nullptr;
    cudaDriverEntryPointQueryResult q;
    RBPR_CUDA(ctx, cudaGetDriverEntryPoint("cuMemGetAddressRange", &sym, cudaEnableDefault, &q));
    if (!sym) RBPR_FAIL(ctx, RBPR_ERR_CUDA, "cuMemGetAddressRange not available");
    fn = (fn_cuMemGetAddressRange)sym;
  }
  unsigned long long b = 0;
  size_t sz = 0;
  if (fn(&b, &sz, (unsigned long long)(uintptr_t)ptr) != 0)
    RBPR_FAIL(ctx, RBPR_ERR_CUDA, "cuMemGetAddressRange failed for %p", ptr);
  *base = (void*)(uintptr_t)b;
  *off = (uint64_t)((uintptr_t)ptr - (uintptr_t)b);
  return 0;
}

size_t sym_gbytes(const rbpr_ctx* ctx) {
  const size_t n = (size_t)ctx->I * ctx->D + (ctx->item_bias ? (size_t)ctx->I : 0);
  return ((n * sizeof(float) + 255) / 256) * 256;
}

}  // namespace

int rbpr_internal_xrank_barrier(rbpr_ctx* ctx, cudaStream_t st) {
  ctx->fx_bar_epoch++;
  xrank_barrier<<<1, 32, 0, st>>>(ctx->fx_flags_dev, ctx->world, ctx->rank, ctx->fx_bar_epoch, ctx->flag);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

// The exchange of one step on stream st (fused path).  ctx->item_grad is the accumulator phase A
// used; on return it points at the (cleared) accumulator of the next step.
int rbpr_internal_fused_exchange(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, cudaStream_t st) {
  int rc = 0;
  const int par = ctx->fx_par;
  ctx->fx_epoch++;
  ExchangeParams p;
  memset(&p, 0, sizeof(p));
  for (int q = 0; q < ctx->world; ++q) {
    p.gsrc[q] = ctx->fx_grad[q][par];
    p.idst[q] = ctx->fx_item[q];
    p.bdst[q] = ctx->fx_bias[q];
  }
  p.gzero = ctx->fx_grad[ctx->rank][par ^ 1];
  p.item_m = ctx->item_m;
  p.item_v = ctx->item_v;
  p.bias_m = ctx->bias_m;
  p.bias_v = ctx->bias_v;
  p.flags = ctx->fx_flags_dev;
  p.epoch = ctx->fx_epoch;
  p.done = ctx->fx_done;
  p.err = ctx->flag;
  p.world = ctx->world;
  p.rank = ctx->rank;
  p.I = ctx->I;
  p.lo = ctx->I * ctx->rank / ctx->world;
  p.hi = ctx->I * (ctx->rank + 1) / ctx->world;
  p.D = ctx->D;
  p.step = step;
  p.lr = hp->lr;
  p.beta1 = hp->beta1;
  p.beta2 = hp->beta2;
  p.eps = hp->eps;
  p.adam_tab = ctx->adam_tab;
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  const int blocks = ctx->sm_count * 4;
#define X(L, V)                                                                                       \
  if (lanes == L && nv == V) {                                                                        \
    switch (hp->optimizer) {                                                                          \
      case RBPR_OPT_SGD: bpr_exchange_apply<L, V, RBPR_OPT_SGD><<<blocks, 256, 0, st>>>(p); break;     \
      case RBPR_OPT_ADAM: bpr_exchange_apply<L, V, RBPR_OPT_ADAM><<<blocks, 256, 0, st>>>(p); break;   \
      case RBPR_OPT_SGDM: bpr_exchange_apply<L, V, RBPR_OPT_SGDM><<<blocks, 256, 0, st>>>(p); break;   \
      default: bpr_exchange_apply<L, V, RBPR_OPT_RMSPROP><<<blocks, 256, 0, st>>>(p); break;           \
    }                                                                                                 \
  } else
  RBPR_FOR_EACH_GEOMETRY(X)
#undef X
  RBPR_FAIL(ctx, RBPR_ERR_ARG, "unsupported dim geometry lanes=%d nv=%d", lanes, nv);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  ctx->fx_par = par ^ 1;
  ctx->fx_wait_epoch = ctx->fx_epoch;  // the next reader of item rows waits for B2 of this exchange
  ctx->item_grad = ctx->fx_grad[ctx->rank][ctx->fx_par];
  ctx->fused_exchanges++;
  return 0;
}

// Block stream st until every rank's rows of the last exchange have landed (no-op when none is pending).
int rbpr_internal_fx_wait(rbpr_ctx* ctx, cudaStream_t st) {
  if (!ctx->fx_bound || ctx->fx_wait_epoch == 0) return 0;
  xwait_kernel<<<1, 32, 0, st>>>(ctx->fx_flags_local + kMaxWorld, ctx->world, ctx->fx_wait_epoch, ctx->flag);
  ctx->launches++;
  RBPR_CUDA(ctx, cudaGetLastError());
  return 0;
}

void rbpr_internal_fx_destroy(rbpr_ctx* ctx) {
  if (!ctx->fx_bound && !ctx->fx_sym) return;
  for (void* m : ctx->fx_opened) cudaIpcCloseMemHandle(m);
  ctx->fx_opened.clear();
  cudaFree(ctx->fx_flags_dev);
  ctx->fx_flags_dev = nullptr;
  cudaFree(ctx->fx_done);
  ctx->fx_done = nullptr;
  if (ctx->fx_bound) ctx->item_grad = ctx->fx_item_grad_owned;  // freed by rbpr_destroy
  cudaFree(ctx->fx_sym);
  ctx->fx_sym = nullptr;
  ctx->fx_bound = false;
}

extern "C" {

int rbpr_comm_ipc_export(rbpr_ctx* ctx, void* blob_out) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!blob_out) RBPR_FAIL(ctx, RBPR_ERR_ARG, "ipc_export: null output");
  if (!ctx->item_emb || !ctx->item_grad) RBPR_FAIL(ctx, RBPR_ERR_STATE, "ipc_export: bind tables first");
  if (ctx->fx_bound) RBPR_FAIL(ctx, RBPR_ERR_STATE, "ipc_export: peer memory already bound");
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t gb = sym_gbytes(ctx);
  if (!ctx->fx_sym) {  // buf[0] | buf[1] | flags (kMaxWorld words, own 256-byte line)
    RBPR_CUDA(ctx, cudaMalloc(&ctx->fx_sym, 2 * gb + 256));
    RBPR_CUDA(ctx, cudaMemset(ctx->fx_sym, 0, 2 * gb + 256));
  }
  IpcBlob b;
  memset(&b, 0, sizeof(b));
  RBPR_CUDA(ctx, cudaIpcGetMemHandle(&b.sym, ctx->fx_sym));
  void* base = nullptr;
  int rc = base_of(ctx, ctx->item_emb, &base, &b.item_off);
  if (rc) return rc;
  RBPR_CUDA(ctx, cudaIpcGetMemHandle(&b.item, base));
  if (ctx->item_bias) {
    rc = base_of(ctx, ctx->item_bias, &base, &b.bias_off);
    if (rc) return rc;
    RBPR_CUDA(ctx, cudaIpcGetMemHandle(&b.bias, base));
    b.has_bias = 1;
  }
  b.gbytes = gb;
  b.device = ctx->device;
  b.I = ctx->I;
  b.D = ctx->D;
  memset(blob_out, 0, RBPR_IPC_BLOB_BYTES);
  memcpy(blob_out, &b, sizeof(b));
  return 0;
}

int rbpr_comm_ipc_bind(rbpr_ctx* ctx, const void* blobs, int32_t world, int32_t rank, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!blobs || world < 2 || world > kMaxWorld || rank < 0 || rank >= world)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "ipc_bind: need 2..%d ranks and every rank's export blob", kMaxWorld);
  if (!ctx->fx_sym) RBPR_FAIL(ctx, RBPR_ERR_STATE, "ipc_bind: call rbpr_comm_ipc_export first");
  if (ctx->fx_bound) RBPR_FAIL(ctx, RBPR_ERR_STATE, "ipc_bind: already bound");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t gb = sym_gbytes(ctx);
  std::vector<std::pair<cudaIpcMemHandle_t, void*>> opened;  // one mapping per distinct allocation
  auto open = [&](const cudaIpcMemHandle_t& h, void** out) -> int {
    for (auto& o : opened)
      if (memcmp(&o.first, &h, sizeof(h)) == 0) {
        *out = o.second;
        return 0;
      }
    void* m = nullptr;
    RBPR_CUDA(ctx, cudaIpcOpenMemHandle(&m, h, cudaIpcMemLazyEnablePeerAccess));
    opened.push_back({h, m});
    ctx->fx_opened.push_back(m);
    *out = m;
    return 0;
  };
  uint32_t* flags_host[kMaxWorld];
  for (int q = 0; q < world; ++q) {
    IpcBlob b;
    memcpy(&b, (const char*)blobs + (size_t)q * RBPR_IPC_BLOB_BYTES, sizeof(b));
    if (b.I != ctx->I || b.D != ctx->D || b.gbytes != gb || b.has_bias != (ctx->item_bias ? 1 : 0))
      RBPR_FAIL(ctx, RBPR_ERR_ARG, "ipc_bind: rank %d holds different tables", q);
    char* sym = nullptr;
    if (q == rank) {
      sym = (char*)ctx->fx_sym;
      ctx->fx_item[q] = ctx->item_emb;
      ctx->fx_bias[q] = ctx->item_bias;
    } else {
      void* m = nullptr;
      int rc = open(b.sym, &m);
      if (rc) return rc;
      sym = (char*)m;
      rc = open(b.item, &m);
      if (rc) return rc;
      ctx->fx_item[q] = (float*)((char*)m + b.item_off);
      ctx->fx_bias[q] = nullptr;
      if (b.has_bias) {
        rc = open(b.bias, &m);
        if (rc) return rc;
        ctx->fx_bias[q] = (float*)((char*)m + b.bias_off);
      }
    }
    ctx->fx_grad[q][0] = (float*)sym;
    ctx->fx_grad[q][1] = (float*)(sym + gb);
    flags_host[q] = (uint32_t*)(sym + 2 * gb);
  }
  RBPR_CUDA(ctx, cudaMalloc(&ctx->fx_done, sizeof(uint32_t)));
  RBPR_CUDA(ctx, cudaMemsetAsync(ctx->fx_done, 0, sizeof(uint32_t), st));
  ctx->fx_flags_local = flags_host[rank];
  ctx->fx_wait_epoch = 0;
  RBPR_CUDA(ctx, cudaMalloc(&ctx->fx_flags_dev, kMaxWorld * sizeof(uint32_t*)));
  RBPR_CUDA(ctx, cudaMemcpyAsync(ctx->fx_flags_dev, flags_host, world * sizeof(uint32_t*), cudaMemcpyHostToDevice, st));
  RBPR_CUDA(ctx, cudaStreamSynchronize(st));
  ctx->world = world;
  ctx->rank = rank;
  ctx->fx_par = 0;
  ctx->fx_epoch = 0;
  ctx->fx_item_grad_owned = ctx->item_grad;
  ctx->item_grad = ctx->fx_grad[rank][0];
  ctx->fx_bound = true;
  // nobody accumulates before every rank has mapped and cleared its buffers
  return rbpr_internal_xrank_barrier(ctx, st);
}

int64_t rbpr_fused_exchange_count(const rbpr_ctx* ctx) { return ctx ? ctx->fused_exchanges : 0; }

}  // extern "C"
