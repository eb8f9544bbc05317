// Copy-on-write block operators of the pool engine (pool_kernel.cuh): cluster_add! of the previous
// observation applied to a SOURCE row and written to a DESTINATION row (the same row when every
// particle that refers to it chose it: in place; a fresh row when only some did: the split of
// src/pmdi.jl:288-309), fused with calc_logprob of the current observation for the updated row
// and, on a split, for the unchanged source row as well - the source is read once.
// Arithmetic: the reference's operation order for the add (gaussian_cluster.jl:57-63,
// categorical_cluster.jl:43-51, negbinom_cluster.jl:43-51), logs of short products for the
// predictive (same forms as cluster_types.cuh).
#pragma once
#include "cluster_types.cuh"

// ------------------------------------------------------------------------------------------
// Gaussian.  Pointers are at this lane's first feature of the 256-feature block; n = size AFTER
// the add.  Returns the updated row's predictive partial (aux - (n/2+1) * sum log(1 + d^2 lamn));
// with SPLIT, *e_src = sum over the block of log(1 + (y - mu_src)^2 lamn_src) of the source row.
// ------------------------------------------------------------------------------------------
template <bool SPLIT>
__device__ __noinline__ double gauss_cow_block(const double* s_sum, const double* s_beta, const double* s_mu,
                                               const double* s_lamn, double* d_sum, double* d_beta, double* d_mu,
                                               double* d_lamn, double* d_aux, const uint8_t* flag_p, int nit, int n,
                                               const double* xp, const double* xc, double* e_src) {
  const double nn = (double)n;
  const double c1 = __dadd_rn(__dadd_rn(nn, -1.0), 0.001);
  const double c2 = __dmul_rn(2.0, __dadd_rn(nn, 0.001));
  const double c3 = __dadd_rn(nn, 0.001);
  const double c4 = __dmul_rn(__dadd_rn(__dmul_rn(0.5, nn), 0.5), c3);
  const double c5 = __dadd_rn(nn, 1.001);
  const double c6 = __dadd_rn(nn, 1.0);
  const double r2 = __drcp_rn(c2), r3 = __drcp_rn(c3), r6 = __drcp_rn(c6);
  double prodl = 1.0, prode = 1.0, prods = 1.0;
#pragma unroll 1
  for (int h = 0; h < 4; h += 2) {
    double2 sm[2], bt[2], mu[2], ln[2];
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (h + i < nit) {
        const int o = (h + i) * PMDI_WF;
        sm[i] = ldcg_f64x2(s_sum + o); bt[i] = ldcg_f64x2(s_beta + o);
        mu[i] = ldcg_f64x2(s_mu + o); ln[i] = ldcg_f64x2(s_lamn + o);
      }
#pragma unroll
    for (int i = 0; i < 2; ++i)
      if (h + i < nit) {
        const int o = (h + i) * PMDI_WF;
        const double2 x = *(const double2*)(xp + o);
        const double2 y = *(const double2*)(xc + o);
        const uchar2 fl = *(const uchar2*)(flag_p + o);
        if (fl.x) {
          if (SPLIT) { const double ds = y.x - mu[i].x; prods *= fma(ds * ds, ln[i].x, 1.0); }
          sm[i].x = __dadd_rn(sm[i].x, x.x);
          const double dd = __dadd_rn(x.x, -mu[i].x);
          bt[i].x = __dadd_rn(bt[i].x, div_const(__dmul_rn(c1, __dmul_rn(dd, dd)), c2, r2));
          mu[i].x = div_const(sm[i].x, c3, r3);
          ln[i].x = div_const(__ddiv_rn(c4, __dmul_rn(bt[i].x, c5)), c6, r6);
          prodl *= ln[i].x;
          const double d = y.x - mu[i].x;
          prode *= fma(d * d, ln[i].x, 1.0);
        }
        if (fl.y) {
          if (SPLIT) { const double ds = y.y - mu[i].y; prods *= fma(ds * ds, ln[i].y, 1.0); }
          sm[i].y = __dadd_rn(sm[i].y, x.y);
          const double dd = __dadd_rn(x.y, -mu[i].y);
          bt[i].y = __dadd_rn(bt[i].y, div_const(__dmul_rn(c1, __dmul_rn(dd, dd)), c2, r2));
          mu[i].y = div_const(sm[i].y, c3, r3);
          ln[i].y = div_const(__ddiv_rn(c4, __dmul_rn(bt[i].y, c5)), c6, r6);
          prodl *= ln[i].y;
          const double d = y.y - mu[i].y;
          prode *= fma(d * d, ln[i].y, 1.0);
        }
        *(double2*)(d_sum + o) = sm[i];
        *(double2*)(d_beta + o) = bt[i];
        *(double2*)(d_mu + o) = mu[i];
        *(double2*)(d_lamn + o) = ln[i];
      }
  }
  double a = 0.5 * pm_log(prodl), e = pm_log(prode), es = SPLIT ? pm_log(prods) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    a += __shfl_xor_sync(FULL, a, o);
    e += __shfl_xor_sync(FULL, e, o);
    if (SPLIT) es += __shfl_xor_sync(FULL, es, o);
  }
  if ((threadIdx.x & 31) == 0) *d_aux = a;
  if (SPLIT) *e_src = es;
  return a - (0.5 * nn + 1.0) * e;
}

// ------------------------------------------------------------------------------------------
// NegBinom.  S pointers at this lane's first feature; observations by shared-memory address.
// Returns the updated row's partial (new aux + per-observation terms); with SPLIT, *v_src = the
// per-observation terms of the source row (its aux is added by the caller).
// ------------------------------------------------------------------------------------------
template <bool SPLIT>
__device__ __noinline__ double nb_cow_block(const long long* s_S, long long* d_S, double* d_aux, int nit, int n,
                                            unsigned xp, unsigned xc, unsigned lf_s, int T, double* v_src) {
  double aux = 0.0, ed = 0.0, es = 0.0;
  const long long nd2 = n + 2, ns2 = n + 1;
#pragma unroll 1
  for (int it = 0; it < nit; ++it) {
    longlong2 s = ldcg_i64x2(s_S + it * PMDI_WF);
    const int2 x = lds_i32x2(xp + it * PMDI_WF * 4);
    const int2 y = lds_i32x2(xc + it * PMDI_WF * 4);
    if (SPLIT) {
      if (y.x >= 0) { const long long b = s.x + y.x; es += lfact_s(b, lf_s, T) - lfact_s(b + ns2, lf_s, T); }
      if (y.y >= 0) { const long long b = s.y + y.y; es += lfact_s(b, lf_s, T) - lfact_s(b + ns2, lf_s, T); }
    }
    if (x.x >= 0) { s.x += x.x; aux += lfact_s(s.x + n + 1, lf_s, T) - lfact_s(s.x, lf_s, T); }
    if (x.y >= 0) { s.y += x.y; aux += lfact_s(s.y + n + 1, lf_s, T) - lfact_s(s.y, lf_s, T); }
    *(longlong2*)(d_S + it * PMDI_WF) = s;
    if (y.x >= 0) { const long long b = s.x + y.x; ed += lfact_s(b, lf_s, T) - lfact_s(b + nd2, lf_s, T); }
    if (y.y >= 0) { const long long b = s.y + y.y; ed += lfact_s(b, lf_s, T) - lfact_s(b + nd2, lf_s, T); }
  }
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    aux += __shfl_xor_sync(FULL, aux, o);
    ed += __shfl_xor_sync(FULL, ed, o);
    if (SPLIT) es += __shfl_xor_sync(FULL, es, o);
  }
  if ((threadIdx.x & 31) == 0) *d_aux = aux;
  if (SPLIT) *v_src = es;
  return aux + ed;
}

// ------------------------------------------------------------------------------------------
// Categorical, packed counts: a feature's level counts are `fpw` fields of 64/fpw bits in each of
// `wpf` consecutive 64-bit words (cw[row][q][wpf]); level l (1-based) is field (l-1) % fpw of
// word (l-1) / fpw.  With <= 4 levels and n_obs < 65536 a feature is ONE word: 8 bytes per
// predictive term, the reference's own width (an Int64 count), and one 128-bit load per lane
// serves two features.
// ------------------------------------------------------------------------------------------
__device__ __forceinline__ unsigned long long ldcg_u64(const unsigned long long* p) { return __ldcg(p); }
__device__ __forceinline__ ulonglong2 ldcg_u64x2(const unsigned long long* p) { return __ldcg((const ulonglong2*)p); }

__device__ __forceinline__ double cat_field(unsigned long long w, int f, int fpw) {
  const int fw = 64 / fpw;
  const unsigned long long mask = fw == 64 ? ~0ull : ((1ull << fw) - 1ull);
  return 0.5 + (double)((w >> (f * fw)) & mask);
}

// calc_logprob partial of nits 64-feature iterations; cw at this lane's first feature.
__device__ __noinline__ double cat_eval_pool(const unsigned long long* cw, int wpf, int fpw, unsigned xs, int nits) {
  double acc = 0.0;
  const int fw = 64 / fpw;
#pragma unroll 1
  for (int i0 = 0; i0 < nits; i0 += 4) {
    double prod = 1.0;
    if (wpf == 1) {
      ulonglong2 w[4];
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (i0 + it < nits) w[it] = ldcg_u64x2(cw + (size_t)(i0 + it) * PMDI_WF);
#pragma unroll
      for (int it = 0; it < 4; ++it)
        if (i0 + it < nits) {
          const int2 lv = lds_i32x2(xs + (i0 + it) * PMDI_WF * 4);
          if (lv.x) prod *= cat_field(w[it].x, lv.x - 1, fpw);
          if (lv.y) prod *= cat_field(w[it].y, lv.y - 1, fpw);
        }
    } else {
#pragma unroll 1
      for (int it = 0; it < 4; ++it)
        if (i0 + it < nits) {
          const int2 lv = lds_i32x2(xs + (i0 + it) * PMDI_WF * 4);
          const unsigned long long* f0 = cw + (size_t)(i0 + it) * PMDI_WF * wpf;
          if (lv.x) prod *= cat_field(ldcg_u64(f0 + (lv.x - 1) / fpw), (lv.x - 1) % fpw, fpw);
          if (lv.y) prod *= cat_field(ldcg_u64(f0 + wpf + (lv.y - 1) / fpw), (lv.y - 1) % fpw, fpw);
        }
    }
    acc += pm_log(prod);
  }
  (void)fw;
  return warp_sum(acc);
}

// add of the previous observation (src -> dst) + predictive of the current one, one 256-block.
template <bool SPLIT>
__device__ __noinline__ double cat_cow_block(const unsigned long long* s_cw, unsigned long long* d_cw, int wpf,
                                             int fpw, int nit, unsigned xp, unsigned xc, double* v_src) {
  const int fw = 64 / fpw;
  double prodd = 1.0, prods = 1.0;
#pragma unroll 1
  for (int it = 0; it < nit; ++it) {
    const int2 x = lds_i32x2(xp + it * PMDI_WF * 4);
    const int2 y = lds_i32x2(xc + it * PMDI_WF * 4);
    if (wpf == 1) {
      ulonglong2 w = ldcg_u64x2(s_cw + (size_t)it * PMDI_WF);
      if (SPLIT) {
        if (y.x) prods *= cat_field(w.x, y.x - 1, fpw);
        if (y.y) prods *= cat_field(w.y, y.y - 1, fpw);
      }
      if (x.x) w.x += 1ull << ((x.x - 1) * fw);
      if (x.y) w.y += 1ull << ((x.y - 1) * fw);
      *(ulonglong2*)(d_cw + (size_t)it * PMDI_WF) = w;
      if (y.x) prodd *= cat_field(w.x, y.x - 1, fpw);
      if (y.y) prodd *= cat_field(w.y, y.y - 1, fpw);
    } else {
#pragma unroll 1
      for (int h = 0; h < 2; ++h) {
        const int xv = h ? x.y : x.x, yv = h ? y.y : y.x;
        const unsigned long long* sf = s_cw + ((size_t)it * PMDI_WF + h) * wpf;
        unsigned long long* df = d_cw + ((size_t)it * PMDI_WF + h) * wpf;
#pragma unroll 1
        for (int wd = 0; wd < wpf; ++wd) {
          unsigned long long w = ldcg_u64(sf + wd);
          if (SPLIT && yv && (yv - 1) / fpw == wd) prods *= cat_field(w, (yv - 1) % fpw, fpw);
          if (xv && (xv - 1) / fpw == wd) w += 1ull << (((xv - 1) % fpw) * fw);
          df[wd] = w;
          if (yv && (yv - 1) / fpw == wd) prodd *= cat_field(w, (yv - 1) % fpw, fpw);
        }
      }
    }
  }
  double ed = pm_log(prodd), es = SPLIT ? pm_log(prods) : 0.0;
#pragma unroll
  for (int o = 16; o > 0; o >>= 1) {
    ed += __shfl_xor_sync(FULL, ed, o);
    if (SPLIT) es += __shfl_xor_sync(FULL, es, o);
  }
  if (SPLIT) *v_src = es;
  return ed;
}
