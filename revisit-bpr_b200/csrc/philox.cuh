// Philox4x32-10 counter-based generator (Salmon et al., SC'11) — the same function cuRAND's
// curandStatePhilox4_32_10_t evaluates: with key = (seed_lo, seed_hi) and counter
// (c0,c1,c2,c3) = (offset_lo, offset_hi, subsequence_lo, subsequence_hi) one call equals
// curand_init(seed, subsequence, 4*offset, &s); curand4(&s).  Written out by hand so the
// training kernel carries no generator state; oracle/philox.py (numpy)
// restates it for the CPU.
#pragma once
#include <stdint.h>

#ifdef __CUDACC__
#define RBPR_HD __host__ __device__ __forceinline__
#else
#define RBPR_HD static inline
#endif

struct philox4 {
  uint32_t x, y, z, w;
};

RBPR_HD void philox_mulhilo(uint32_t a, uint32_t b, uint32_t& hi, uint32_t& lo) {
#ifdef __CUDA_ARCH__
  hi = __umulhi(a, b);
  lo = a * b;
#else
  uint64_t p = (uint64_t)a * (uint64_t)b;
  hi = (uint32_t)(p >> 32);
  lo = (uint32_t)p;
#endif
}

RBPR_HD philox4 philox4x32_10(uint32_t c0, uint32_t c1, uint32_t c2, uint32_t c3, uint32_t k0,
                              uint32_t k1) {
  const uint32_t M0 = 0xD2511F53u, M1 = 0xCD9E8D57u, W0 = 0x9E3779B9u, W1 = 0xBB67AE85u;
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    uint32_t hi0, lo0, hi1, lo1;
    philox_mulhilo(M0, c0, hi0, lo0);
    philox_mulhilo(M1, c2, hi1, lo1);
    uint32_t n0 = hi1 ^ c1 ^ k0;
    uint32_t n1 = lo1;
    uint32_t n2 = hi0 ^ c3 ^ k1;
    uint32_t n3 = lo0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += W0; k1 += W1;
  }
  philox4 o = {c0, c1, c2, c3};
  return o;
}
