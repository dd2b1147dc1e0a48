// Memory-system floor of the BPR access pattern on one B200 (diagnostic, not product code):
// per "triple" read 3 random rows of D floats, optionally write one and reduce-add two.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o scripts/micro/gather_floor scripts/micro/gather_floor.cu
#include <cuda_runtime.h>
#include <cstdio>
#include <cstdlib>
#include <vector>
#include <random>
#include <algorithm>

#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); exit(1);} } while (0)

__device__ __forceinline__ void red4(float* p, float4 v) {
  asm volatile("red.global.add.v4.f32 [%0], {%1,%2,%3,%4};" ::"l"(p), "f"(v.x), "f"(v.y), "f"(v.z), "f"(v.w) : "memory");
}

// one warp per triple, D = 128 (one float4 per lane); mode bits: 1 = write user row, 2 = red item rows
template <int MODE>
__global__ void k(const int4* __restrict__ rec, int n, float* __restrict__ ue, const float* __restrict__ ie,
                  float* __restrict__ ig, float* __restrict__ sink) {
  const int lane = threadIdx.x & 31;
  const int warps = (gridDim.x * blockDim.x) >> 5;
  float acc = 0.f;
  for (int t = (blockIdx.x * blockDim.x + threadIdx.x) >> 5; t < n; t += warps) {
    const int4 r = __ldg(rec + t);
    const float4 u = *reinterpret_cast<const float4*>(ue + (size_t)r.x * 128 + lane * 4);
    const float4 a = *reinterpret_cast<const float4*>(ie + (size_t)r.y * 128 + lane * 4);
    const float4 b = *reinterpret_cast<const float4*>(ie + (size_t)r.z * 128 + lane * 4);
    const float d = u.x * (a.x - b.x) + u.y * (a.y - b.y) + u.z * (a.z - b.z) + u.w * (a.w - b.w);
    acc += d;
    if (MODE & 1) {
      float4 o = make_float4(u.x + 1e-6f * a.x, u.y + 1e-6f * a.y, u.z + 1e-6f * b.z, u.w + 1e-6f * b.w);
      *reinterpret_cast<float4*>(ue + (size_t)r.x * 128 + lane * 4) = o;
    }
    if (MODE & 2) {
      red4(ig + (size_t)r.y * 128 + lane * 4, make_float4(d, u.y, u.z, u.w));
      red4(ig + (size_t)r.z * 128 + lane * 4, make_float4(-d, -u.y, -u.z, -u.w));
    }
  }
  if (acc == 1234.5f) sink[0] = acc;
}

int main(int argc, char** argv) {
  const int U = 136678, I = 20109, D = 128;
  const int B = argc > 1 ? atoi(argv[1]) : 65536;
  const int steps = 64;
  float *ue, *ie, *ig, *sink;
  CK(cudaMalloc(&ue, (size_t)U * D * 4)); CK(cudaMalloc(&ie, (size_t)I * D * 4));
  CK(cudaMalloc(&ig, (size_t)I * D * 4)); CK(cudaMalloc(&sink, 4));
  CK(cudaMemset(ue, 0, (size_t)U * D * 4)); CK(cudaMemset(ie, 0, (size_t)I * D * 4)); CK(cudaMemset(ig, 0, (size_t)I * D * 4));
  std::mt19937 g(13);
  std::vector<int4> h((size_t)B * steps);
  for (auto& r : h) { r.x = 1 + g() % (U - 1); r.y = 1 + g() % (I - 1); r.z = 1 + g() % (I - 1); r.w = 1; }
  // sort each step's records by user (like the product path)
  for (int s = 0; s < steps; ++s) std::sort(h.begin() + (size_t)s * B, h.begin() + (size_t)(s + 1) * B, [](const int4& a, const int4& b) { return a.x < b.x; });
  int4* rec; CK(cudaMalloc(&rec, h.size() * sizeof(int4)));
  CK(cudaMemcpy(rec, h.data(), h.size() * sizeof(int4), cudaMemcpyHostToDevice));
  cudaEvent_t e0, e1; cudaEventCreate(&e0); cudaEventCreate(&e1);
  for (int blocks_per_sm : {4, 8, 16}) {
    for (int mode = 0; mode < 4; ++mode) {
      const int blocks = 148 * blocks_per_sm;
      auto launch = [&](int s) {
        const int4* r = rec + (size_t)s * B;
        switch (mode) {
          case 0: k<0><<<blocks, 128>>>(r, B, ue, ie, ig, sink); break;
          case 1: k<1><<<blocks, 128>>>(r, B, ue, ie, ig, sink); break;
          case 2: k<2><<<blocks, 128>>>(r, B, ue, ie, ig, sink); break;
          default: k<3><<<blocks, 128>>>(r, B, ue, ie, ig, sink); break;
        }
      };
      for (int s = 0; s < 8; ++s) launch(s);
      CK(cudaDeviceSynchronize());
      cudaEventRecord(e0);
      for (int s = 0; s < steps; ++s) launch(s);
      cudaEventRecord(e1);
      CK(cudaDeviceSynchronize());
      float ms; cudaEventElapsedTime(&ms, e0, e1);
      const double us = ms * 1e3 / steps;
      const double bytes = (double)B * 128 * 4 * (3 + ((mode & 1) ? 1 : 0) + ((mode & 2) ? 2 : 0));
      printf("B=%d blocks/SM=%2d mode=%d (%s%s): %7.2f us/launch  %7.1f GB/s moved  (%.1f GB/s at 24*D B/triple)\n", B, blocks_per_sm, mode,
             (mode & 1) ? "+write-user " : "", (mode & 2) ? "+red-2-item-rows" : "read-only", us, bytes / us / 1e3,
             (double)B * 24 * 128 / us / 1e3);
    }
  }
  return 0;
}
