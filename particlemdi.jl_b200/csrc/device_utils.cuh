// Device helpers for libpmdi_cuda.so (sm_100a): RNG addressing, warp/block reductions,
// log-factorial lookup, cache-bypassing loads for data other CTAs write, the grid barrier.
#pragma once
#ifndef __CUDACC_RTC__
#include <cuda_runtime.h>
#include <stdint.h>
#else
#define INFINITY (__longlong_as_double(0x7ff0000000000000LL))
#endif

#include "pmdi_internal.h"

#define FULL 0xffffffffu

// Philox4x32-10 (Salmon et al., SC'11), addressed per draw by (seed, iter, kind, step, k, index).
// Replaces Julia's rand() stream (src/pmdi.jl:253, src/misc.jl:28,43, src/pmdi.jl:350,367).
__host__ __device__ __forceinline__ double pmdi_philox_uniform(unsigned long long seed, unsigned iter,
                                                               unsigned kind, unsigned step,
                                                               unsigned k, unsigned index) {
  unsigned c0 = index, c1 = step, c2 = (kind << 16) | (k & 0xFFFFu), c3 = iter;
  unsigned k0 = (unsigned)seed, k1 = (unsigned)(seed >> 32);
#pragma unroll
  for (int r = 0; r < 10; ++r) {
    const unsigned long long p0 = (unsigned long long)0xD2511F53u * c0;
    const unsigned long long p1 = (unsigned long long)0xCD9E8D57u * c2;
    const unsigned n0 = (unsigned)(p1 >> 32) ^ c1 ^ k0;
    const unsigned n1 = (unsigned)p1;
    const unsigned n2 = (unsigned)(p0 >> 32) ^ c3 ^ k1;
    const unsigned n3 = (unsigned)p0;
    c0 = n0; c1 = n1; c2 = n2; c3 = n3;
    k0 += 0x9E3779B9u; k1 += 0xBB67AE85u;
  }
  const unsigned long long x = ((unsigned long long)c0 << 32) | c1;
  return (double)(x >> 11) * (1.0 / 9007199254740992.0);
}

// One copy each of the long FP64 routines inside the persistent kernel: the sweep's per-step
// code has to stay resident in the instruction cache (the math library inlines ~2-4 KB per call
// site).  Same routines, same bits as the inlined forms.
__device__ __noinline__ double pm_log(double x) { return log(x); }
__device__ __noinline__ double pm_exp(double x) { return exp(x); }
__device__ __noinline__ double pm_div(double a, double b) { return __ddiv_rn(a, b); }
__device__ __noinline__ double pm_uniform(unsigned long long seed, unsigned iter, unsigned kind, unsigned step,
                                          unsigned k, unsigned index) {
  return pmdi_philox_uniform(seed, iter, kind, step, k, index);
}
__device__ __noinline__ double pm_lfact_stirling(long long k) {
  const double z = (double)k + 1.0;
  const double zi = 1.0 / z, zi2 = zi * zi;
  return (z - 0.5) * log(z) - z + 0.91893853320467274178 +
         zi * (1.0 / 12.0 - zi2 * (1.0 / 360.0 - zi2 * (1.0 / 1260.0)));
}

// Butterfly reductions: every lane ends with the same bits (a+b == b+a, identical tree shape).
__device__ __forceinline__ double warp_sum(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(FULL, v, o);
  return v;
}
__device__ __forceinline__ double warp_max(double v) {
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) v = fmax(v, __shfl_xor_sync(FULL, v, o));
  return v;
}

// L2-only loads for anything another CTA may have written during this kernel (L1 is not coherent).
__device__ __forceinline__ double ldcg_f64(const double* p) { return __ldcg(p); }
__device__ __forceinline__ int ldcg_i32(const int* p) { return __ldcg(p); }
__device__ __forceinline__ double2 ldcg_f64x2(const double* p) { return __ldcg((const double2*)p); }
__device__ __forceinline__ longlong2 ldcg_i64x2(const long long* p) { return __ldcg((const longlong2*)p); }
__device__ __forceinline__ unsigned ldcg_u32(const unsigned* p) { return __ldcg(p); }
__device__ __forceinline__ uint8_t ldcg_u8(const uint8_t* p) { return __ldcg(p); }

// lgamma(k+1) for integer k >= 0: table (shared or global) below T, Stirling series above
// (z >= 256: the first omitted term 1/(1680 z^7) is < 1e-20).
__device__ __forceinline__ double lfact(long long k, const double* tab, int T) {
  if (k < (long long)T) return tab[k];
  return pm_lfact_stirling(k);
}

__device__ __forceinline__ unsigned ld_acquire_u32(const unsigned* p) {
  unsigned v;
  asm volatile("ld.acquire.gpu.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
  return v;
}
__device__ __forceinline__ unsigned long long globaltimer_ns() {
  unsigned long long t;
  asm volatile("mov.u64 %0, %globaltimer;" : "=l"(t));
  return t;
}
