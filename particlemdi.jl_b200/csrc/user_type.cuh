// User-defined cluster types as device functors (the reference's plugin contract, README.md:48-88:
// calc_logprob / cluster_add! / calc_logmarginal for a user's cluster struct).  This header is only
// compiled at run time (NVRTC), together with the user's source, into a private copy of the sweep
// kernels: pmdi_register_cluster_type hands over a struct
//
//   struct MyType {
//     static constexpr int WORDS = ...;                                        // doubles of state per feature (<= 8)
//     __device__ static void   init(double* st);                               // the empty cluster
//     __device__ static double logprob(const double* st, int n, double x);     // this feature's term of calc_logprob
//                                                                              //   (n = cluster size, x = the observation)
//     __device__ static void   add(double* st, int n, double x);               // cluster_add!; n = size AFTER the add
//     __device__ static double logmarginal(const double* st, int n);           // this feature's calc_logmarginal
//   };
//
// The cluster types of the reference are all of this form: feature-wise sufficient statistics, the
// observation's log-probability a sum over features.  State lives as ust[row][word][feature].
#pragma once
#include "cluster_types.cuh"

#ifdef PMDI_USER_STRUCT
typedef PMDI_USER_STRUCT PmdiUser;
static_assert(PmdiUser::WORDS >= 1 && PmdiUser::WORDS <= 8, "a user cluster type keeps 1..8 doubles per feature");

__device__ __forceinline__ double user_x(const DsDev& ds, unsigned xs_addr, int f, int x_is_int) {
  // staged observation: doubles, or 32-bit integers for integer-valued data
  if (x_is_int) { int v; asm("ld.shared.s32 %0, [%1];" : "=r"(v) : "r"(xs_addr + f * 4u)); return (double)v; }
  double v; asm("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(xs_addr + f * 8u)); return v;
}

// One block of one row, by one warp: mode 0 plain evaluation, 1 add + evaluation, 2 add into the row
// d_off doubles further on + evaluation of both (pool_types.cuh).  s_st at the row's word 0, this lane's
// first feature; xp / xc: shared-memory addresses of the block's first feature of this lane.
__device__ __noinline__ double user_block(const DsDev& ds, const double* s_st, long long d_off, const uint8_t* flag_p,
                                          int nit, int mode, int n, unsigned xp, unsigned xc, double* v_src) {
  const int W = PmdiUser::WORDS, Dp = ds.Dp, xi = ds.uW < 0;
  double ed = 0.0, es = 0.0;
#pragma unroll 1
  for (int it = 0; it < nit; ++it) {
#pragma unroll 1
    for (int h = 0; h < 2; ++h) {
      const int f = it * PMDI_WF + h;
      double st[PmdiUser::WORDS];
#pragma unroll
      for (int w = 0; w < W; ++w) st[w] = ldcg_f64(s_st + (long long)w * Dp + f);
      if (flag_p[f]) {
        const double y = user_x(ds, xc, f, xi);
        if (mode != 1) es += PmdiUser::logprob(st, n, y);
        if (mode) {
          PmdiUser::add(st, n + 1, user_x(ds, xp, f, xi));
          ed += PmdiUser::logprob(st, n + 1, y);
        }
      }
      if (mode) {
#pragma unroll
        for (int w = 0; w < W; ++w) const_cast<double*>(s_st)[d_off + (long long)w * Dp + f] = st[w];
      }
    }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ed += __shfl_xor_sync(FULL, ed, o);
    es += __shfl_xor_sync(FULL, es, o);
  }
  if (mode == 0) return es;
  *v_src = es;
  return ed;
}

// sequential cluster_add! of a member list into one feature of a row (prefix build, single-cluster evaluation)
__device__ __forceinline__ void user_build_feature(const DsDev& ds, long long row, int q, const int* mem, int cnt, bool on) {
  double st[PmdiUser::WORDS];
  PmdiUser::init(st);
  if (on)
    for (int t = 0; t < cnt; ++t) {
      const double x = ds.uW < 0 ? (double)((const int*)ds.x)[(size_t)mem[t] * ds.Dp + q] : ((const double*)ds.x)[(size_t)mem[t] * ds.Dp + q];
      PmdiUser::add(st, t + 1, x);
    }
  for (int w = 0; w < PmdiUser::WORDS; ++w) ds.ust[((long long)row * PmdiUser::WORDS + w) * ds.Dp + q] = st[w];
}
#endif
