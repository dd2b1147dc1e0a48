// The data-parallel step's ONE exchange as ONE kernel over NVLink / NVSwitch (sm_100a): reduce the
// ranks' dense item gradients, apply the optimizer, and publish the updated item rows to every
// replica — element by element, no separate all-reduce pass, no dense re-read of the reduced buffer —
// plus everything else a step does after phase A (user half of the apply, statistics, accumulator
// clear), so that a data-parallel step is two launches.
//
//   every rank r owns a contiguous slice of the item rows, [I*r/W, I*(r+1)/W):
//     g   = sum over ranks q = 0..W-1 of grad_q[element]   multicast binding: ONE multimem.ld_reduce —
//                                                           the NVSwitch reads the W copies and returns
//                                                           their sum; otherwise W independent 128-bit
//                                                           peer loads summed in rank order
//     x   = optimizer(x, g [, m, v of the slice])          SGD / Adam / SGD-momentum / RMSprop; the
//                                                           optimizer state of the item table is
//                                                           SHARDED: only the owner keeps its slice
//     item_q[element] = x   for every rank q               multicast binding: ONE multimem.st;
//                                                           otherwise W 128-bit peer stores
//   so replicas are bit-identical by construction and each rank sweeps 1/W of the table (the dense
//   Adam sweep of the 8-GPU MSD configuration shrinks 8x).
//
// Synchronisation: two gradient accumulators used alternately (step s accumulates into buf[s&1], the
// exchange of step s clears the local buf[(s+1)&1], which the peers finished reading one step ago),
// and two cross-rank barriers per step (flag words in peer memory, st.release.sys / ld.acquire.sys,
// bounded spins): B1 "every rank's phase A has landed" before the reduce, B2 "every rank's rows have
// landed" before the next phase A, which is launched with programmatic stream serialization and waits
// for the flags itself.
//
// Memory is shared in one of two ways (the host shell chooses; include/rbpr.h):
//   * rbpr_comm_symm_bind: ONE symmetric buffer per rank allocated by the host (torch symmetric
//     memory / cuMem + cuMulticast) holding accumulators, flags AND the item table / bias, mapped
//     into every peer, optionally aliased by an NVSwitch multicast address;
//   * rbpr_comm_ipc_export / rbpr_comm_ipc_bind: cudaIpc handles of the library's accumulators and
//     of the storages holding the item table / bias (unicast only).
// NCCL stays the fallback when neither is available.  Replaces the DDP gradient all-reduce + dense
// optimizer step of the reference (experiments/launcher.py:59-70, experiments/trainer.py:76-79).
// Measured anatomy of the kernel on 2 and 8 B200s: profiles/round2/z_exchange_trace.txt, DESIGN.md §6.
#include <algorithm>

#include "train_kernels.cuh"

using namespace rbpr_dev;

namespace {

constexpr int kMaxWorld = RBPR_MAX_PEERS;
constexpr int kTraceSlots = 512;  // RBPR_FX_TRACE ring: exchanges kept per report

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
  // barriers folded into this kernel: B1 (signal at the start, wait before the reduce), B2 (signal by the last CTA)
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
  unsigned long long* trace;     // RBPR_FX_TRACE: 8 globaltimer words of this exchange, or null
  // NVSwitch multicast (symmetric-memory binding only; null = unicast loads / stores through gsrc / idst):
  const float* mc_grad;          // multicast address of this parity's accumulators: multimem.ld_reduce sums
                                 // the W copies INSIDE the switch, one 16-byte response per element
  float* mc_item;                // multicast address of the item tables: one multimem.st reaches every replica
  float* mc_bias;
  int n_xchg;                    // CTAs [0, n_xchg) run the exchange, the others the local work
  int mc_chunks;                 // > 1: chunk-major pull / push pipeline of the multicast path
};

__device__ __forceinline__ float4 ldcg4x(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }
__device__ __forceinline__ float4 add4(float4 a, float4 b) { return make_float4(a.x + b.x, a.y + b.y, a.z + b.z, a.w + b.w); }
__device__ __forceinline__ float4 mc_ld_reduce4(const float* p) {
  float4 v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.v4.f32 {%0,%1,%2,%3}, [%4];"
               : "=f"(v.x), "=f"(v.y), "=f"(v.z), "=f"(v.w) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st4(float* p, float4 v) {
  asm volatile("multimem.st.relaxed.sys.global.v4.f32 [%0], {%1,%2,%3,%4};"
               :: "l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}
__device__ __forceinline__ float mc_ld_reduce1(const float* p) {
  float v;
  asm volatile("multimem.ld_reduce.relaxed.sys.global.add.f32 %0, [%1];" : "=f"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ void mc_st1(float* p, float v) {
  asm volatile("multimem.st.relaxed.sys.global.f32 [%0], %1;" :: "l"(p), "f"(v) : "memory");
}
__device__ __forceinline__ unsigned long long gtime() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}

// One launch per step and rank; the CTAs split into two roles that run side by side (they are bound
// by different resources: NVLink vs HBM):
//   exchange CTAs [0, n_xchg):
//     1. CTA 0 publishes "this rank's phase A has landed" (B1) to every peer; all wait for every peer's B1;
//     2. the slice, one 16-byte element per thread and iteration: the W accumulators summed — by the
//        NVSwitch (multimem.ld_reduce: one response per element) under the symmetric-memory binding,
//        else W independent 128-bit peer loads in flight at once, summed in rank order — optimizer,
//        then the updated element stored into every replica (multimem.st, else W 128-bit stores);
//   local CTAs [n_xchg, grid): work that needs no peer — the user half of the step's apply (users are
//        sharded by owner, their gradients never leave the rank), the step statistics, clearing the
//        accumulator of the next step;
//   the last CTA to finish publishes "this rank's rows have landed" (B2).
template <int LANES, int NV, int OPT>
__global__ void __launch_bounds__(256, 4) bpr_exchange_apply(const ExchangeParams p, const ApplyParams u) {
  const int D = p.D;
  OptScalars h = {p.lr, p.beta1, p.beta2, p.eps, 0.f, 1.f};
  if (OPT == RBPR_OPT_ADAM) {
    const float2 t = __ldg(p.adam_tab + (p.step + 1));
    h.step_size = t.x;
    h.bc2_sqrt = t.y;
  }
  // the next phase A may be placed as soon as SM resources free up (launch_phase_a, train_kernels.cuh)
  asm volatile("griddepcontrol.launch_dependents;" ::: "memory");
  const int W = p.world;
  const int G = (int)gridDim.x;
  const bool both = p.n_xchg >= G;  // no role split: every CTA does the local work, then the exchange
  const bool tracer = p.trace != nullptr && blockIdx.x == 0 && threadIdx.x == 0;
  if (tracer) p.trace[0] = gtime();
  // B1 signal: this kernel runs after this rank's phase A (stream order)
  if (blockIdx.x == 0 && threadIdx.x < W) {
    __threadfence_system();
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p.flags[threadIdx.x] + p.rank), "r"(p.epoch) : "memory");
  }
  if (both || (int)blockIdx.x >= p.n_xchg) {
    // ---- local CTAs
    const int lb = both ? (int)blockIdx.x : (int)blockIdx.x - p.n_xchg, ln = both ? G : G - p.n_xchg;
    const int64_t nthreads = (int64_t)ln * blockDim.x;
    const int64_t tid = (int64_t)lb * blockDim.x + threadIdx.x;
    if (u.do_users) apply_users<LANES, NV, OPT>(u, h, tid / LANES, nthreads / LANES);
    // the accumulator the NEXT step uses: the peers finished reading it before their B2 of the
    // previous exchange, which this rank's phase A (earlier in this stream) has waited for
    const int64_t vecs = (p.I * D + (p.bdst[p.rank] != nullptr ? p.I : 0) + 3) / 4;  // buffer is padded to 16 B
    float4* z = reinterpret_cast<float4*>(p.gzero);
    for (int64_t k = tid; k < vecs; k += nthreads) z[k] = f4zero();
    if (lb == ln - 1 && u.stats_out != nullptr) apply_stats(u);
    if (p.trace != nullptr && threadIdx.x == 0) atomicMax(p.trace + 1, gtime());
  }
  if (both || (int)blockIdx.x < p.n_xchg) {
    // ---- exchange CTAs
    const int64_t nthreads = (int64_t)(both ? G : p.n_xchg) * blockDim.x;
    const int64_t tid = (int64_t)blockIdx.x * blockDim.x + threadIdx.x;
    wait_peer_flags(p.flags[p.rank], W, p.epoch, p.err);
    if (tracer) p.trace[2] = gtime();
    const int64_t e_lo = p.lo * D / 4, e_hi = p.hi * D / 4;
    // optimizer on one 16-byte element of the slice, then the new value into every replica
    auto finish = [&](int64_t off, float4 gr) {
      float4 pp = ldcg4x(p.idst[p.rank] + off);
      if (OPT == RBPR_OPT_SGD) {
        pp.x -= p.lr * gr.x;
        pp.y -= p.lr * gr.y;
        pp.z -= p.lr * gr.z;
        pp.w -= p.lr * gr.w;
      } else {
        float4 m = ld4(p.item_m + off);
        float4 vv = opt_has_s2(OPT) ? ld4(p.item_v + off) : f4zero();
        opt4<OPT>(pp, m, vv, gr, h);
        st4(p.item_m + off, m);
        if (opt_has_s2(OPT)) st4(p.item_v + off, vv);
      }
      if (p.mc_item != nullptr) {
        mc_st4(p.mc_item + off, pp);
      } else {
#pragma unroll
        for (int j = 0; j < kMaxWorld; ++j)
          if (j < W) st4(p.idst[j] + off, pp);
      }
    };
    if (p.mc_grad != nullptr && p.mc_chunks > 1) {
      // EXPERIMENTAL (RBPR_FX_MC_CHUNKS > 1; default off, see the launcher).
      // Multicast: the pull loads every GPU's OUT links (the switch reads all W copies), the push its
      // IN links (every replica receives every slice), so the two can overlap — if pushes start while
      // pulls are still queued.  The slice is cut into kChunks consecutive chunks; a thread issues its
      // element of every chunk back to back (requests reach the switch chunk-major) and finishes them
      // in that order: chunk c is being published while chunks > c are still being reduced.
      constexpr int kChunks = 4;
      const int64_t n = e_hi - e_lo;
      for (int64_t base = 0; base < n; base += nthreads * kChunks) {
        const int64_t span = (n - base) < nthreads * kChunks ? (n - base) : nthreads * kChunks;
        const int64_t T = (span + kChunks - 1) / kChunks;
        if (tid >= T) continue;
        float4 g4[kChunks];
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int64_t idx = c * T + tid;
          if (idx < span) g4[c] = mc_ld_reduce4(p.mc_grad + 4 * (e_lo + base + idx));
        }
#pragma unroll
        for (int c = 0; c < kChunks; ++c) {
          const int64_t idx = c * T + tid;
          if (idx < span) finish(4 * (e_lo + base + idx), g4[c]);
        }
      }
    } else {
      for (int64_t e = e_lo + tid; e < e_hi; e += nthreads) {
        const int64_t off = 4 * e;
        float4 gr;
        if (p.mc_grad != nullptr) {
          gr = mc_ld_reduce4(p.mc_grad + off);
        } else {
          float4 t[kMaxWorld];
#pragma unroll
          for (int j = 0; j < kMaxWorld; ++j)  // all W loads are independent: issued back to back
            if (j < W) t[j] = ldcg4x(p.gsrc[j] + off);
          gr = t[0];
#pragma unroll
          for (int j = 1; j < kMaxWorld; ++j)  // fixed order (rank 0, 1, ...): every run sums the same way
            if (j < W) gr = add4(gr, t[j]);
        }
        finish(off, gr);
      }
    }
    if (p.bdst[p.rank] != nullptr) {
      for (int64_t r = p.lo + tid; r < p.hi; r += nthreads) {
        float gb = 0.f;
        if (p.mc_grad != nullptr) {
          gb = mc_ld_reduce1(p.mc_grad + p.I * D + r);
        } else {
          for (int q = 0; q < W; ++q) gb += __ldcg(p.gsrc[q] + p.I * D + r);
        }
        float b = __ldcg(p.bdst[p.rank] + r);
        if (OPT == RBPR_OPT_SGD) {
          b -= p.lr * gb;
        } else {
          float m = p.bias_m[r], vv = opt_has_s2(OPT) ? p.bias_v[r] : 0.f;
          opt1<OPT>(b, m, vv, gb, h);
          p.bias_m[r] = m;
          if (opt_has_s2(OPT)) p.bias_v[r] = vv;
        }
        if (p.mc_bias != nullptr) {
          mc_st1(p.mc_bias + r, b);
        } else {
          for (int q = 0; q < W; ++q) p.bdst[q][r] = b;
        }
      }
    }
    if (p.trace != nullptr && threadIdx.x == 0) atomicMax(p.trace + 3, gtime());
  }
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
      if (p.trace != nullptr) p.trace[4] = gtime();
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
// used; on return it points at the (cleared) accumulator of the next step.  `users` (optional): the
// user half of the step's apply and its statistics, run inside the same kernel ahead of the B1 wait.
int rbpr_internal_fused_exchange(rbpr_ctx* ctx, uint64_t step, const rbpr_hparams* hp, cudaStream_t st,
                                 const ApplyParams* users) {
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
  if (ctx->fx_trace != nullptr) p.trace = ctx->fx_trace + 8 * (size_t)(ctx->fx_trace_n++ % kTraceSlots);
  if (ctx->fx_mc_grad[par] != nullptr) {
    // RBPR_FX_MC_PUSH=0: in-switch reduction for the pull, unicast stores for the push
    static const bool mc_push = [] {
      const char* e = getenv("RBPR_FX_MC_PUSH");
      return e == nullptr || atoi(e) != 0;
    }();
    p.mc_grad = ctx->fx_mc_grad[par];
    if (mc_push) {
      p.mc_item = ctx->fx_mc_item;
      p.mc_bias = ctx->fx_mc_bias;
    }
  }
  ApplyParams u;
  if (users != nullptr) {
    u = *users;
  } else {
    memset(&u, 0, sizeof(u));
  }
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  // one wave of 4 CTAs per SM (<= 64 registers).  RBPR_FX_XCHG_CTAS = how many of the 4 run the
  // exchange role (the rest the local role); 4 = no split: every CTA does the local work first
  static const int xchg_per_sm = [] {
    const char* e = getenv("RBPR_FX_XCHG_CTAS");
    const int v = e ? atoi(e) : 4;
    return v < 1 ? 1 : (v > 4 ? 4 : v);
  }();
  static const int mc_chunks = [] {
    const char* e = getenv("RBPR_FX_MC_CHUNKS");
    return e ? atoi(e) : 1;  // off by default: measured slower at N=2 (85.6 vs 82.7 us/step), unmeasured at N=8
  }();
  const int blocks = ctx->sm_count * 4;
  p.n_xchg = ctx->sm_count * xchg_per_sm;
  p.mc_chunks = mc_chunks;
#define X(L, V)                                                                                       \
  if (lanes == L && nv == V) {                                                                        \
    switch (hp->optimizer) {                                                                          \
      case RBPR_OPT_SGD: bpr_exchange_apply<L, V, RBPR_OPT_SGD><<<blocks, 256, 0, st>>>(p, u); break;     \
      case RBPR_OPT_ADAM: bpr_exchange_apply<L, V, RBPR_OPT_ADAM><<<blocks, 256, 0, st>>>(p, u); break;   \
      case RBPR_OPT_SGDM: bpr_exchange_apply<L, V, RBPR_OPT_SGDM><<<blocks, 256, 0, st>>>(p, u); break;   \
      default: bpr_exchange_apply<L, V, RBPR_OPT_RMSPROP><<<blocks, 256, 0, st>>>(p, u); break;           \
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

// RBPR_FX_TRACE=1: where the time of the exchange kernel goes (globaltimer stamps written by the
// kernel), printed per call to stderr.  Diagnostic only: synchronises the stream.
int rbpr_internal_fx_trace_report(rbpr_ctx* ctx, cudaStream_t st) {
  if (ctx->fx_trace == nullptr || ctx->fx_trace_n == 0) return 0;
  const int n = ctx->fx_trace_n < kTraceSlots ? (int)ctx->fx_trace_n : kTraceSlots;
  std::vector<unsigned long long> t((size_t)8 * kTraceSlots);
  RBPR_CUDA(ctx, cudaMemcpyAsync(t.data(), ctx->fx_trace, t.size() * sizeof(unsigned long long),
                                 cudaMemcpyDeviceToHost, st));
  RBPR_CUDA(ctx, cudaStreamSynchronize(st));
  // medians: the first exchange of a call waits for the slowest HOST to get going, which would own the mean
  std::vector<double> f[6];
  for (int i = 0; i < n; ++i) {
    const unsigned long long* w = t.data() + 8 * i;
    f[0].push_back((double)(w[1] - w[0]));  // local CTAs done (they run beside / before the exchange CTAs)
    f[1].push_back((double)(w[2] - w[0]));
    f[2].push_back((double)(w[3] - w[2]));
    f[3].push_back((double)(w[4] - (w[3] > w[1] ? w[3] : w[1])));
    f[4].push_back((double)(w[4] - w[0]));
    if (i + 1 < n) f[5].push_back((double)(t[8 * (i + 1)] - w[4]));
  }
  auto med = [](std::vector<double>& v) {
    if (v.empty()) return 0.0;
    std::sort(v.begin(), v.end());
    return v[v.size() / 2] * 1e-3;
  };
  fprintf(stderr,
          "[rbpr fx trace] rank %d: %d exchanges, median us: local work done at %.2f | B1 signal+wait %.2f | slice reduce+publish %.2f | "
          "tail(fence+B2) %.2f | kernel %.2f | B2 -> next exchange start (phase A etc.) %.2f\n",
          ctx->rank, n, med(f[0]), med(f[1]), med(f[2]), med(f[3]), med(f[4]), med(f[5]));
  RBPR_CUDA(ctx, cudaMemsetAsync(ctx->fx_trace, 0, t.size() * sizeof(unsigned long long), st));
  ctx->fx_trace_n = 0;
  return 0;
}

void rbpr_internal_fx_destroy(rbpr_ctx* ctx) {
  if (!ctx->fx_bound && !ctx->fx_sym) return;
  ctx->fx_mc_grad[0] = ctx->fx_mc_grad[1] = nullptr;
  ctx->fx_mc_item = ctx->fx_mc_bias = nullptr;
  cudaFree(ctx->fx_trace);
  ctx->fx_trace = nullptr;
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
  if (getenv("RBPR_FX_TRACE") != nullptr && ctx->fx_trace == nullptr) {
    RBPR_CUDA(ctx, cudaMalloc(&ctx->fx_trace, (size_t)8 * kTraceSlots * sizeof(unsigned long long)));
    RBPR_CUDA(ctx, cudaMemsetAsync(ctx->fx_trace, 0, (size_t)8 * kTraceSlots * sizeof(unsigned long long), st));
    ctx->fx_trace_n = 0;
  }
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

// ---- symmetric-memory binding ---------------------------------------------------------------------
// Layout of the host-allocated symmetric buffer (every offset 256-byte aligned):
//   [ accumulator 0 | accumulator 1 | flag words | item table | item bias ]
static size_t symm_item_off(const rbpr_ctx* ctx) { return 2 * sym_gbytes(ctx) + 256; }
static size_t symm_bias_off(const rbpr_ctx* ctx) {
  return symm_item_off(ctx) + (((size_t)ctx->I * ctx->D * sizeof(float) + 255) / 256) * 256;
}

int64_t rbpr_comm_symm_bytes(const rbpr_ctx* ctx) {
  if (!ctx || !ctx->item_emb) return 0;
  return (int64_t)(symm_bias_off(ctx) + (ctx->item_bias ? (((size_t)ctx->I * sizeof(float) + 255) / 256) * 256 : 0));
}

int rbpr_comm_symm_bind(rbpr_ctx* ctx, const uint64_t* peer_bases, uint64_t multicast_base, int32_t world,
                        int32_t rank, uint64_t* item_emb_out, uint64_t* item_bias_out, void* stream) {
  if (!ctx) return RBPR_ERR_ARG;
  if (!peer_bases || world < 2 || world > kMaxWorld || rank < 0 || rank >= world)
    RBPR_FAIL(ctx, RBPR_ERR_ARG, "symm_bind: need 2..%d ranks and every rank's mapping of the symmetric buffer", kMaxWorld);
  if (!ctx->item_emb || !ctx->item_grad) RBPR_FAIL(ctx, RBPR_ERR_STATE, "symm_bind: bind tables first");
  if (ctx->fx_bound) RBPR_FAIL(ctx, RBPR_ERR_STATE, "symm_bind: peer memory already bound");
  for (int q = 0; q < world; ++q)
    if (peer_bases[q] == 0 || (peer_bases[q] & 255u)) RBPR_FAIL(ctx, RBPR_ERR_ARG, "symm_bind: mapping %d is null or not 256-byte aligned", q);
  if (multicast_base & 255u) RBPR_FAIL(ctx, RBPR_ERR_ARG, "symm_bind: multicast address not 256-byte aligned");
  cudaStream_t st = (cudaStream_t)stream;
  RBPR_CUDA(ctx, cudaSetDevice(ctx->device));
  const size_t gb = sym_gbytes(ctx), ioff = symm_item_off(ctx), boff = symm_bias_off(ctx);
  // the tables move into the symmetric buffer (the caller re-points its tensors at *_out)
  char* mine = (char*)(uintptr_t)peer_bases[rank];
  RBPR_CUDA(ctx, cudaMemcpyAsync(mine + ioff, ctx->item_emb, (size_t)ctx->I * ctx->D * sizeof(float),
                                 cudaMemcpyDeviceToDevice, st));
  if (ctx->item_bias)
    RBPR_CUDA(ctx, cudaMemcpyAsync(mine + boff, ctx->item_bias, (size_t)ctx->I * sizeof(float), cudaMemcpyDeviceToDevice, st));
  uint32_t* flags_host[kMaxWorld];
  for (int q = 0; q < world; ++q) {
    char* sym = (char*)(uintptr_t)peer_bases[q];
    ctx->fx_grad[q][0] = (float*)sym;
    ctx->fx_grad[q][1] = (float*)(sym + gb);
    flags_host[q] = (uint32_t*)(sym + 2 * gb);
    ctx->fx_item[q] = (float*)(sym + ioff);
    ctx->fx_bias[q] = ctx->item_bias ? (float*)(sym + boff) : nullptr;
  }
  if (multicast_base != 0) {
    char* mc = (char*)(uintptr_t)multicast_base;
    ctx->fx_mc_grad[0] = (const float*)mc;
    ctx->fx_mc_grad[1] = (const float*)(mc + gb);
    ctx->fx_mc_item = (float*)(mc + ioff);
    ctx->fx_mc_bias = ctx->item_bias ? (float*)(mc + boff) : nullptr;
  }
  ctx->fx_item_prev = ctx->item_emb;
  ctx->fx_bias_prev = ctx->item_bias;
  ctx->item_emb = ctx->fx_item[rank];
  if (ctx->item_bias) ctx->item_bias = ctx->fx_bias[rank];
  if (item_emb_out) *item_emb_out = (uint64_t)(uintptr_t)ctx->item_emb;
  if (item_bias_out) *item_bias_out = (uint64_t)(uintptr_t)ctx->item_bias;
  ctx->fx_symm_host = true;
  RBPR_CUDA(ctx, cudaMalloc(&ctx->fx_done, sizeof(uint32_t)));
  RBPR_CUDA(ctx, cudaMemsetAsync(ctx->fx_done, 0, sizeof(uint32_t), st));
  if (getenv("RBPR_FX_TRACE") != nullptr && ctx->fx_trace == nullptr) {
    RBPR_CUDA(ctx, cudaMalloc(&ctx->fx_trace, (size_t)8 * kTraceSlots * sizeof(unsigned long long)));
    RBPR_CUDA(ctx, cudaMemsetAsync(ctx->fx_trace, 0, (size_t)8 * kTraceSlots * sizeof(unsigned long long), st));
    ctx->fx_trace_n = 0;
  }
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
  // nobody accumulates (or reads a replica) before every rank has placed its tables
  return rbpr_internal_xrank_barrier(ctx, st);
}

int32_t rbpr_fused_exchange_multicast(const rbpr_ctx* ctx) { return (ctx && ctx->fx_bound && ctx->fx_mc_item) ? 1 : 0; }

int64_t rbpr_fused_exchange_count(const rbpr_ctx* ctx) { return ctx ? ctx->fused_exchanges : 0; }

}  // extern "C"
