// Small-batch training path for sm_100a: MANY consecutive steps in ONE launch of ONE thread-block
// cluster (8 CTAs x 512 threads), hardware cluster barriers between the halves of a step.
//
// Why: at the reference configs' own batch size (train_batch_size 256, README.md:305) a step moves
// 0.8 MB — a fraction of a microsecond of memory time — and the two dependent launches of the
// large-batch path (bpr_phase_a, bpr_apply) cost 10.6 us per step.  A grid-wide cooperative kernel
// was tried in round 1 and lost (two grid.sync per step cost more than two launches:
// profiles/r02d_coop_experiment.txt).  A cluster barrier is a hardware barrier among <= 8 SMs
// (~0.2 us), and 4096 threads are exactly one 16-lane group per triple at B=256, D=128.
//
// One step (exact synchronous-minibatch semantics, same arithmetic as train_kernels.cuh):
//   A  every lane group takes triples of the step: the three rows are read with ld.global.cg
//      (L2: another CTA of the cluster may have written them in the previous step, L1 is not
//      coherent), dot / softplus / gradients as in bpr_phase_a; item gradients go to the dense
//      accumulator with red.global.add.v4.f32, a user occurring once is updated in place, a user
//      occurring several times accumulates into the user-gradient buffer;
//   -- barrier.cluster (release / acquire at cluster scope) --
//   B  the same groups walk the same records again: the slot the preparation kernel designated for an
//      item row (one per touched item and step, flags kRecApplyPos / kRecApplyNeg baked into the
//      record one wave ahead — no atomic on the critical path) applies the accumulated gradient and
//      clears it; the designated triple of a repeated user applies the user row.  No scan over the
//      catalogue, no touched-flag pass: the work of a step is proportional to its batch.
//   -- barrier.cluster --
// Plain SGD only (a stateful optimizer moves every item row every step — dense torch.optim
// semantics — which is a sweep over the table, not a small-batch operation); everything else takes
// the large-batch path.  Reference call sites replaced: as rbpr_train_steps (include/rbpr.h).
#include "train_kernels.cuh"

using namespace rbpr_dev;

namespace {

// 4096 threads either as 8 CTAs x 512 (portable cluster size) or 16 CTAs x 256 (non-portable: two warps
// per scheduler instead of four, the same chain of round trips with half the issue contention)
constexpr int kSmallTotalThreads = 4096;

struct SmallParams {
  TrainParams t;        // tables, accumulators, hyper-parameters (t.batch = triples per step)
  float* item_emb_w;    // the item table, writable
  float* item_bias_w;   // or null
  const int4* records;  // the wave's records {u, i+, i-, flags}
  int64_t n;            // triples in the wave
  int n_steps;
  double* stats;        // (n_steps, RBPR_STATS_PER_STEP), zeroed by the caller
};

__device__ __forceinline__ float4 ldcg4(const float* p) { return __ldcg(reinterpret_cast<const float4*>(p)); }

// Release / acquire at cluster scope orders every earlier global access of the arriving threads (reds,
// stores) before every later access of the threads that have waited: all readers and writers of a
// row sit in this one cluster, so no device-scope fence is needed on top.
__device__ __forceinline__ void cluster_barrier() {
  asm volatile("barrier.cluster.arrive.release.aligned;\n" ::: "memory");
  asm volatile("barrier.cluster.wait.acquire.aligned;\n" ::: "memory");
}

template <int LANES, int NV, int kSmallThreads>
__global__ void __launch_bounds__(kSmallThreads, 1) bpr_small_steps(const SmallParams sp) {
  const TrainParams& p = sp.t;
  const Group<LANES> g;
  const int D = p.D;
  __shared__ float4 s_part[kSmallThreads / 32];
  const uint32_t groups_total = (gridDim.x * kSmallThreads) / LANES;
  const uint32_t gid = (blockIdx.x * kSmallThreads + threadIdx.x) / LANES;
  const float lr = p.lr;
  bool colok[NV];
#pragma unroll
  for (int v = 0; v < NV; ++v) colok[v] = 4 * (g.gl + LANES * v) < D;

  // Records are static, so the group's first record of step s+1 is fetched during step s and the rows
  // it names are prefetched into L2 one step ahead (the user table does not fit L2: without this the
  // gather of every step waits for DRAM).  A prefetch only warms L2; the load that follows the
  // barrier still reads the current value.
  auto first_record = [&](int step) -> int4 {
    const int64_t o = (int64_t)step * p.batch;
    const int64_t l = sp.n - o;
    const int64_t nn = l < p.batch ? l : p.batch;
    return (step < sp.n_steps && (int64_t)gid < nn) ? __ldg(sp.records + o + gid) : make_int4(0, 0, 0, 0);
  };
  auto prefetch_rows = [&](const int4& r) {
    const int c = 4 * g.gl;  // one 16-byte column per lane: LANES lanes cover the first 16*LANES bytes, NV strides the rest
#pragma unroll
    for (int v = 0; v < NV; ++v) {
      if (!colok[v]) continue;
      const int cc = c + 4 * LANES * v;
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.user_emb + (size_t)r.x * D + cc));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.item_emb + (size_t)r.y * D + cc));
      asm volatile("prefetch.global.L2 [%0];" ::"l"(p.item_emb + (size_t)r.z * D + cc));
    }
  };
  int4 rec_first = first_record(0), rec_ahead = first_record(1);
  for (int s = 0; s < sp.n_steps; ++s) {
    const int64_t off = (int64_t)s * p.batch;
    const int64_t left = sp.n - off;
    const uint32_t n = (uint32_t)(left < p.batch ? left : p.batch);
    const int4* recs = sp.records + off;
    float loss_acc = 0.f, absx_acc = 0.f, l2_acc = 0.f, cnt_acc = 0.f;
    const int4 rec0 = rec_first;
    if (s + 1 < sp.n_steps) prefetch_rows(rec_ahead);  // rows of step s+1 (its record arrived a step ago)

    // ---- A: gather, loss, gradients ---------------------------------------------------------------
    for (uint32_t k = gid; k < n; k += groups_total) {
      const int4 rec = (k == gid) ? rec0 : __ldg(recs + k);
      const int32_t uu = rec.x, i = rec.y, j = rec.z;
      const bool single = (rec.w & kRecSingle) != 0;
      const float* urow = p.user_emb + (size_t)uu * D;
      const float* irow = p.item_emb + (size_t)i * D;
      const float* jrow = p.item_emb + (size_t)j * D;
      float4 u[NV], vi[NV], vj[NV];
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        const int c = 4 * (g.gl + LANES * v);
        u[v] = colok[v] ? ldcg4(urow + c) : f4zero();
        vi[v] = colok[v] ? ldcg4(irow + c) : f4zero();
        vj[v] = colok[v] ? ldcg4(jrow + c) : f4zero();
      }
      float pp = 0.f, pn = 0.f, sq = 0.f, usq = 0.f;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        pp += dot4(u[v], vi[v]);
        pn += dot4(u[v], vj[v]);
        sq += p.reg_item * dot4(vi[v], vi[v]) + p.reg_neg * dot4(vj[v], vj[v]);
        usq += dot4(u[v], u[v]);
      }
      float x = g.sum(pp - pn);
      if (p.item_bias != nullptr) x += __ldcg(p.item_bias + i) - __ldcg(p.item_bias + j);
      const float e = __expf(-fabsf(x));
      const float spl = fmaxf(-x, 0.f) + __logf(1.0f + e);
      const float inv = __fdividef(1.0f, 1.0f + e);
      const float c = (x >= 0.f) ? e * inv : inv;
      l2_acc += 0.5f * (sq + p.reg_user * usq);
      if (g.gl == 0) {
        loss_acc += spl;
        absx_acc += fabsf(x);
        cnt_acc += 1.f;
        if (p.bias_grad != nullptr) {
          atomicAdd(p.bias_grad + i, -c);
          atomicAdd(p.bias_grad + j, c);
        }
      }
      float* gi = (i != 0) ? p.item_grad + (size_t)i * D : nullptr;
      float* gj = (j != 0) ? p.item_grad + (size_t)j * D : nullptr;
      float* gurow = p.user_grad + (size_t)uu * D;
      float* uout = p.user_emb + (size_t)uu * D;
#pragma unroll
      for (int v = 0; v < NV; ++v) {
        if (!colok[v]) continue;
        const int cidx = 4 * (g.gl + LANES * v);
        const float4 cu = make_float4(c * u[v].x, c * u[v].y, c * u[v].z, c * u[v].w);
        float4 a, b, gu;
        a.x = p.reg_item * vi[v].x - cu.x;
        a.y = p.reg_item * vi[v].y - cu.y;
        a.z = p.reg_item * vi[v].z - cu.z;
        a.w = p.reg_item * vi[v].w - cu.w;
        b.x = p.reg_neg * vj[v].x + cu.x;
        b.y = p.reg_neg * vj[v].y + cu.y;
        b.z = p.reg_neg * vj[v].z + cu.z;
        b.w = p.reg_neg * vj[v].w + cu.w;
        if (gi != nullptr) red4(gi + cidx, a);
        if (gj != nullptr) red4(gj + cidx, b);
        gu.x = p.reg_user * u[v].x - c * (vi[v].x - vj[v].x);
        gu.y = p.reg_user * u[v].y - c * (vi[v].y - vj[v].y);
        gu.z = p.reg_user * u[v].z - c * (vi[v].z - vj[v].z);
        gu.w = p.reg_user * u[v].w - c * (vi[v].w - vj[v].w);
        if (uu == 0) continue;
        if (!single) {
          red4(gurow + cidx, gu);
        } else {
          float4 o;
          o.x = u[v].x - lr * gu.x;
          o.y = u[v].y - lr * gu.y;
          o.z = u[v].z - lr * gu.z;
          o.w = u[v].w - lr * gu.w;
          st4(uout + cidx, o);
        }
      }
    }
    {  // per-warp statistics partial of the step
      float a = loss_acc, b = l2_acc, cabs = absx_acc, d = cnt_acc;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        cabs += __shfl_xor_sync(0xffffffffu, cabs, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
      }
      if ((threadIdx.x & 31) == 0) s_part[threadIdx.x >> 5] = make_float4(a, b, cabs, d);
    }
    cluster_barrier();

    if (threadIdx.x < 32 && sp.stats != nullptr) {  // one atomic per statistic per CTA
      const float4 v = (threadIdx.x < kSmallThreads / 32) ? s_part[threadIdx.x] : f4zero();
      double a = v.x, b = v.y, c2 = v.z, d = v.w;
#pragma unroll
      for (int o = 16; o > 0; o >>= 1) {
        a += __shfl_xor_sync(0xffffffffu, a, o);
        b += __shfl_xor_sync(0xffffffffu, b, o);
        c2 += __shfl_xor_sync(0xffffffffu, c2, o);
        d += __shfl_xor_sync(0xffffffffu, d, o);
      }
      if (threadIdx.x == 0) {
        double* out = sp.stats + (size_t)s * RBPR_STATS_PER_STEP;
        atomicAdd(out + 0, a);
        atomicAdd(out + 1, b);
        atomicAdd(out + 2, c2);
        atomicAdd(out + 3, d);
      }
    }

    // ---- B: apply the step's gradients, driven by the step's own records ----------------------------
    // up to three rows per slot (positive, negative, repeated user): all loads are issued before any
    // dependent arithmetic, so the phase costs ONE round trip to L2
    for (uint32_t k = gid; k < n; k += groups_total) {
      const int4 rec = (k == gid) ? rec0 : __ldg(recs + k);
      float* grow[3];
      float* prow[3];
      bool on[3];
      on[0] = (rec.w & kRecApplyPos) != 0;
      on[1] = (rec.w & kRecApplyNeg) != 0;
      on[2] = (rec.w & kRecMultiHead) != 0 && rec.x != 0;
      grow[0] = p.item_grad + (size_t)rec.y * D;
      prow[0] = sp.item_emb_w + (size_t)rec.y * D;
      grow[1] = p.item_grad + (size_t)rec.z * D;
      prow[1] = sp.item_emb_w + (size_t)rec.z * D;
      grow[2] = p.user_grad + (size_t)rec.x * D;
      prow[2] = p.user_emb + (size_t)rec.x * D;
      float4 gr[3][NV], pv[3][NV];
#pragma unroll
      for (int w = 0; w < 3; ++w)
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          const int c = 4 * (g.gl + LANES * v);
          const bool ld = on[w] && colok[v];
          gr[w][v] = ld ? ldcg4(grow[w] + c) : f4zero();
          pv[w][v] = ld ? ldcg4(prow[w] + c) : f4zero();
        }
      float gb[2] = {0.f, 0.f}, bv[2] = {0.f, 0.f};
      if (g.gl == 0 && p.bias_grad != nullptr) {
        if (on[0]) { gb[0] = __ldcg(p.bias_grad + rec.y); bv[0] = __ldcg(sp.item_bias_w + rec.y); }
        if (on[1]) { gb[1] = __ldcg(p.bias_grad + rec.z); bv[1] = __ldcg(sp.item_bias_w + rec.z); }
      }
#pragma unroll
      for (int w = 0; w < 3; ++w) {
        if (!on[w]) continue;
#pragma unroll
        for (int v = 0; v < NV; ++v) {
          if (!colok[v]) continue;
          const int c = 4 * (g.gl + LANES * v);
          float4 o = pv[w][v];
          o.x -= lr * gr[w][v].x;
          o.y -= lr * gr[w][v].y;
          o.z -= lr * gr[w][v].z;
          o.w -= lr * gr[w][v].w;
          st4(prow[w] + c, o);
          st4(grow[w] + c, f4zero());
        }
      }
      if (g.gl == 0 && p.bias_grad != nullptr) {
        if (on[0]) { sp.item_bias_w[rec.y] = bv[0] - lr * gb[0]; p.bias_grad[rec.y] = 0.f; }
        if (on[1]) { sp.item_bias_w[rec.z] = bv[1] - lr * gb[1]; p.bias_grad[rec.z] = 0.f; }
      }
    }
    rec_first = rec_ahead;
    rec_ahead = first_record(s + 2);  // consumed two steps from now: its latency is never waited for
    cluster_barrier();
  }
}

}  // namespace

// Can this call take the small-batch path?  (plain SGD, single GPU, static sampler — checked by the
// caller — and at most 4 triples per lane group per step)
bool rbpr_small_batch_eligible(const rbpr_ctx* ctx, int64_t batch) {
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  const int64_t groups = (int64_t)kSmallTotalThreads / lanes;
  return batch <= 4 * groups;
}

namespace {
template <int L, int V, int T>
int launch_small(rbpr_ctx* ctx, const SmallParams& sp, int cluster, cudaStream_t st) {
  static bool allowed = false;
  if (cluster > 8 && !allowed) {
    RBPR_CUDA(ctx, cudaFuncSetAttribute(bpr_small_steps<L, V, T>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1));
    allowed = true;
  }
  cudaLaunchConfig_t cfg;
  memset(&cfg, 0, sizeof(cfg));
  cfg.gridDim = dim3(cluster);
  cfg.blockDim = dim3(T);
  cfg.stream = st;
  cudaLaunchAttribute attr[1];
  attr[0].id = cudaLaunchAttributeClusterDimension;
  attr[0].val.clusterDim.x = cluster;
  attr[0].val.clusterDim.y = 1;
  attr[0].val.clusterDim.z = 1;
  cfg.attrs = attr;
  cfg.numAttrs = 1;
  RBPR_CUDA(ctx, cudaLaunchKernelEx(&cfg, bpr_small_steps<L, V, T>, sp));
  return 0;
}
}  // namespace

// `n_steps` consecutive steps over prepared records in one cluster launch on stream st.
int rbpr_launch_small_steps(rbpr_ctx* ctx, const TrainParams& p, const int4* records, int64_t n,
                            int n_steps, double* stats, cudaStream_t st) {
  SmallParams sp;
  sp.t = p;
  sp.item_emb_w = ctx->item_emb;
  sp.item_bias_w = ctx->item_bias;
  sp.records = records;
  sp.n = n;
  sp.n_steps = n_steps;
  sp.stats = stats;
  if (stats)
    RBPR_CUDA(ctx, cudaMemsetAsync(stats, 0, (size_t)n_steps * RBPR_STATS_PER_STEP * sizeof(double), st));
  int lanes, nv;
  rbpr_geometry(ctx->D, &lanes, &nv);
  // 16 CTAs x 256 threads when the device can co-schedule such a cluster (measured: 4.2 us per step
  // against 5.5 us for 8 x 512 at B=256, D=128), else the portable 8 x 512; RBPR_SMALL_CLUSTER overrides
  static int cluster = 0;
  if (cluster == 0) {
    const char* e = getenv("RBPR_SMALL_CLUSTER");
    cluster = (e && atoi(e) == 8) ? 8 : 16;
    if (cluster == 16) {
      cudaLaunchConfig_t probe;
      memset(&probe, 0, sizeof(probe));
      probe.gridDim = dim3(16);
      probe.blockDim = dim3(256);
      cudaLaunchAttribute pa[1];
      pa[0].id = cudaLaunchAttributeClusterDimension;
      pa[0].val.clusterDim.x = 16;
      pa[0].val.clusterDim.y = 1;
      pa[0].val.clusterDim.z = 1;
      probe.attrs = pa;
      probe.numAttrs = 1;
      int n_clusters = 0;
      cudaError_t e1 = cudaFuncSetAttribute(bpr_small_steps<16, 2, 256>, cudaFuncAttributeNonPortableClusterSizeAllowed, 1);
      cudaError_t e2 = e1 == cudaSuccess ? cudaOccupancyMaxActiveClusters(&n_clusters, bpr_small_steps<16, 2, 256>, &probe)
                                         : e1;
      if (e2 != cudaSuccess || n_clusters < 1) {
        cudaGetLastError();  // clear
        cluster = 8;
      }
    }
  }
#define X(L, V)                                                                              \
  if (lanes == L && nv == V) {                                                               \
    int rc = cluster == 16 ? launch_small<L, V, 256>(ctx, sp, 16, st) : launch_small<L, V, 512>(ctx, sp, 8, st); \
    if (rc) return rc;                                                                       \
    ctx->launches++;                                                                         \
    ctx->small_launches++;                                                                   \
    return 0;                                                                                \
  }
  RBPR_FOR_EACH_GEOMETRY(X)
#undef X
  RBPR_FAIL(ctx, RBPR_ERR_ARG, "unsupported dim geometry lanes=%d nv=%d", lanes, nv);
}
