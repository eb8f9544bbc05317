// The conditional-SMC sweep as ONE persistent cooperative kernel: the per-observation loop of
// the reference (src/pmdi.jl:209-342) runs on the device with no host involvement.  One CTA per
// SM.  Every (dataset, particle-slot) unit of statistics is owned by one CTA for the whole sweep,
// so predictive, proposal and add of a unit are CTA-local; only the particle weights need the
// whole grid, and that exchange is SPLIT-PHASE: a CTA arrives at the grid barrier of step t when
// its proposals are out, starts the predictive pass of step t+1 at once, and only looks at the
// other CTAs' increments (weights + ESS of step t) when the first of its units is ready to
// propose for step t+1.  Barrier latency and step-to-step imbalance between CTAs hide behind the
// next step's streaming; the rare resampling step (ESS <= P/2) discards that pass and redoes it
// after the particles have moved.
//
// Per observation step t a CTA runs:
//   top       wait for the staged observation, prefetch the next one (cp.async, 3-deep ring: the
//             previous row is still needed by the fused add), build this step's item list;
//   queue     ONE dynamically scheduled list of work items handed to warps through a
//             shared-memory counter, longest first:
//               fused item  = the row the particle chose in step t-1, one 256-feature block:
//                             cluster_add! of x[t-1] (src/pmdi.jl:300) and calc_logprob of x[t]
//                             in the same pass over the row;
//               plain item  = calc_logprob of `qb` blocks of any other occupied row
//                             (src/pmdi.jl:218-220);
//             the warp that finishes the last item of a unit runs the unit's proposal
//             (src/pmdi.jl:223-265: softmax-cdf, draw, weight increment) -> lab/inc[t&1][k][slot]
//             - provided step t-1 is RESOLVED: its grid barrier has completed and one warp has
//             folded every particle's increments and the Phi coupling (src/misc.jl:50-59) into
//             the CTA's private log-weights and evaluated calc_ESS (src/misc.jl:15-25; identical
//             bits in every CTA, so all CTAs take the same branch).  Units that finish earlier
//             are deferred to the end of the queue;
//   arrive    at the grid barrier of step t (no wait).
// Resampling (draw_partstar src/misc.jl:27-47 by CTA 0, then all CTAs move the duplicated
// particles' rows) costs two blocking grid barriers and happens on a few steps per sweep.
#pragma once
#include "cluster_types.cuh"

struct SweepSmem {
  double red[3 * 32];
  double lp_empty[PMDI_MAX_K];  // predictive of the empty cluster, this step
  double res_mx;                // max log-weight of the last resolved step
  int item_ctr, total_items, n_fused, n_defer;
  int res_step;   // last step whose weights / ESS this CTA has folded in
  int res_claim;  // step a warp has claimed to resolve
  int res_flag;   // ESS <= P/2 at res_step: resample before going on
  int fail;
  unsigned rows_eval[PMDI_MAX_K];
  unsigned long long tacc[8];
  unsigned long long t_prev;
};

// ---- block-wide helpers (fixed shapes -> identical bits in every CTA) -------------------------
__device__ __forceinline__ double block_max(double v, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  v = warp_max(v);
  __syncthreads();
  if (lane == 0) red[w] = v;
  __syncthreads();
  double r = red[lane < (PMDI_NT / 32) ? lane : 0];
  return warp_max(r);
}
__device__ __forceinline__ void block_sum2(double& a, double& b, double* red) {
  const int lane = threadIdx.x & 31, w = threadIdx.x >> 5;
  a = warp_sum(a);
  b = warp_sum(b);
  __syncthreads();
  if (lane == 0) { red[w] = a; red[32 + w] = b; }
  __syncthreads();
  double ra = lane < (PMDI_NT / 32) ? red[lane] : 0.0;
  double rb = lane < (PMDI_NT / 32) ? red[32 + lane] : 0.0;
  a = warp_sum(ra);
  b = warp_sum(rb);
}

// exclusive scan of in[0..P) (global scratch, CTA-local use) into out; returns the total.
__device__ int block_excl_scan(const int* in, int* out, int P, int* s_tmp /* PMDI_NT+1 ints */) {
  const int t = threadIdx.x;
  const int seg = (P + PMDI_NT - 1) / PMDI_NT;
  const int b = min(P, t * seg), e = min(P, b + seg);
  int s = 0;
  for (int i = b; i < e; ++i) s += in[i];
  __syncthreads();
  s_tmp[t] = s;
  __syncthreads();
  if (t == 0) {
    int acc = 0;
    for (int i = 0; i < PMDI_NT; ++i) { const int v = s_tmp[i]; s_tmp[i] = acc; acc += v; }
    s_tmp[PMDI_NT] = acc;
  }
  __syncthreads();
  int acc = s_tmp[t];
  for (int i = b; i < e; ++i) { const int v = in[i]; out[i] = acc; acc += v; }
  __syncthreads();
  return s_tmp[PMDI_NT];
}

// draw_partstar (src/misc.jl:27-47) + the slot plan, by CTA 0.  Output: anc_log[ev][P] (1-based),
// slot_of/logical_of[next][P], copies[], plan_out[0] = number of copies.
// The Fisher-Yates shuffle followed by partstar[1]=1 and sort! only decides WHICH element of the
// sorted systematic sample is replaced by the reference particle: the one the shuffle moves to
// position 1.  That index is traced through the swaps without moving anything.
__device__ __noinline__ void resample_plan(const SweepParams& sp, int step, int ev, double mx, const double* lw,
                              int* s_tmp) {
  const int P = sp.P, t = threadIdx.x;
  const int* slot_cur = sp.slot_of + (ev & 1) * P;
  int* slot_nxt = sp.slot_of + ((ev + 1) & 1) * P;
  int* logi_nxt = sp.logical_of + ((ev + 1) & 1) * P;
  for (int p = t; p < P; p += PMDI_NT) {
    sp.sc_w[p] = pm_exp(__ldcg(lw + p) - mx);
    // Fisher-Yates pick for position pos = p+1 (entry index p), pos >= 2
    const double us = sp.tape_shuffle ? sp.tape_shuffle[(size_t)step * P + p]
                                      : pm_uniform(sp.seed, sp.iter, DRAW_SHUFFLE, step, 0, p);
    int jj = 1 + (int)floor(us * (double)(p + 1));
    if (jj > p + 1) jj = p + 1;
    sp.sc_j[p] = jj;
    sp.sc_b[p] = 0;  // has_child
  }
  __syncthreads();
  if (t == 0) {  // pprob = cumsum(exp.(logweight .- max)), sequential (misc.jl:29)
    double acc = 0.0;
    for (int p = 0; p < P; ++p) { acc += sp.sc_w[p]; sp.sc_pp[p] = acc; }
  } else if (t == 32) {  // u, u + 1/P, ... by repeated addition (misc.jl:28,35)
    const double r = sp.tape_resamp ? sp.tape_resamp[step]
                                    : pm_uniform(sp.seed, sp.iter, DRAW_RESAMP, step, 0, 0);
    double u = r / (double)P;
    for (int i = 0; i < P; ++i) { sp.sc_u[i] = u; u += 1.0 / (double)P; }
  } else if (t == 64) {  // index of the pre-shuffle element that ends at position 1
    int tt = 0;
    for (int pos = 2; pos <= P; ++pos)
      if (sp.sc_j[pos - 1] - 1 == tt) tt = pos - 1;
    s_tmp[PMDI_NT + 1] = tt;
  }
  __syncthreads();
  const double tot = sp.sc_pp[P - 1];
  for (int i = t; i < P; i += PMDI_NT) {  // first p with pprob[p]/last >= u_i (misc.jl:33-38)
    const double ui = sp.sc_u[i];
    int lo = 0, hi = P;  // hi == P means none
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (pm_div(sp.sc_pp[mid], tot) >= ui) hi = mid; else lo = mid + 1;
    }
    sp.sc_anc0[i] = (lo < P) ? lo + 1 : P;
  }
  __syncthreads();
  const int drop = s_tmp[PMDI_NT + 1];
  int* anc = sp.anc_log + (size_t)ev * P;
  for (int i = t; i < P; i += PMDI_NT) {
    const int a = (i == 0) ? 1 : ((i - 1 < drop) ? sp.sc_anc0[i - 1] : sp.sc_anc0[i]);
    anc[i] = a;
    if (sp.dbg_anc) sp.dbg_anc[(size_t)step * P + i] = a;
  }
  __syncthreads();
  // first child of every ancestor keeps the ancestor's slot; the others take dead slots
  for (int i = t; i < P; i += PMDI_NT) {
    const int first = (i == 0) || (anc[i] != anc[i - 1]);
    sp.sc_a[i] = first ? 0 : 1;  // extra child
    if (first) sp.sc_b[anc[i] - 1] = 1;
  }
  __syncthreads();
  for (int p = t; p < P; p += PMDI_NT) sp.sc_b[p] = sp.sc_b[p] ? 0 : 1;  // dead
  __syncthreads();
  const int n_extra = block_excl_scan(sp.sc_a, sp.sc_c, P, s_tmp);  // sc_c = extra rank
  block_excl_scan(sp.sc_b, sp.sc_d, P, s_tmp);                      // sc_d = dead rank
  for (int p = t; p < P; p += PMDI_NT)
    if (sp.sc_b[p]) sp.sc_j[sp.sc_d[p]] = ldcg_i32(slot_cur + p);  // dead slot list
  __syncthreads();
  for (int i = t; i < P; i += PMDI_NT) {
    const int src = ldcg_i32(slot_cur + anc[i] - 1);
    int dst = src;
    if (sp.sc_a[i]) {
      dst = sp.sc_j[sp.sc_c[i]];
      sp.copies[sp.sc_c[i]] = make_int2(src, dst);
    }
    slot_nxt[i] = dst;
    logi_nxt[dst] = i;
  }
  if (t == 0) {
    sp.plan_out[0] = n_extra;
    sp.ev_of_step[step] = ev;
  }
  __syncthreads();
}

__device__ __forceinline__ void cp_async16(void* smem_dst, const void* gsrc) {
  const unsigned d = (unsigned)__cvta_generic_to_shared(smem_dst);
  asm volatile("cp.async.cg.shared.global [%0], [%1], 16;" ::"r"(d), "l"(gsrc) : "memory");
}
__device__ __forceinline__ void cp_async_commit_wait_all() {
  asm volatile("cp.async.commit_group;\ncp.async.wait_group 0;" ::: "memory");
}

// optional per-warp event trace of one CTA and one step (PMDI_TRACE_STEP): tag << 48 | clock64
__device__ __noinline__ void trace_mark(const SweepParams& sp, int step, unsigned tag) {
  __shared__ int tr_cnt[PMDI_NT / 32];
  if (sp.trace && (int)blockIdx.x == sp.trace_cta && step == sp.trace_step && (threadIdx.x & 31) == 0) {
    const int w = threadIdx.x >> 5;
    unsigned long long* t = sp.trace + w * 128;
    const int n = (t[0] == 0 ? 0 : tr_cnt[w]) + 1;
    if (n < 128) {
      t[n] = ((unsigned long long)tag << 48) | (clock64() & 0xFFFFFFFFFFFFull);
      t[0] = n;
      tr_cnt[w] = n;
    }
  }
}

// per-CTA views into dynamic shared memory
struct CtaTables {
  unsigned* urow;   // [max_units][N]   occupied rows of a unit: label | n << 8
  int* ucount;      // [max_units]      number of occupied rows
  int* foff;        // [max_units + 1]  first fused item of a unit in this step's list
  int* poff;        // [max_units + 1]  first plain item of a unit (after all fused items)
  int* pbase;       // [max_units]      first partial-sum slot of a unit (rows x J slots)
  int* uinfo;       // [max_units]      k << 24 | slot
  int* ulog;        // [max_units]      logical particle of the unit's slot (RNG address, log index)
  int* pend;        // [max_units]      pending add: label | n_after << 8, or -1
  int* pe;          // [max_units]      position of the pending row in urow
  int* remaining;   // [max_units]      items of the unit not yet finished this step
  int* defer;       // [max_units]      units whose proposal waits for the previous step's weights
  unsigned* items;  // [item_cap]       fused << 31 | u << 13 | e << 5 | j0
  double* part;     // [item_cap]       predictive partial sums of this step
  double* lp_s;     // [NW][Npad]       per-warp proposal scratch
  double* Pi_s;     // [K][N]
};

#define PMDI_ITEM_FUSED 0x80000000u

// cluster_add! of the pending row of every unit against observation buffer xb (only after the
// last observation: every other pending add is applied by the next step's fused items).
__device__ __noinline__ void flush_adds(const SweepParams& sp, const CtaTables& T, int nu, const unsigned char* xb,
                                        const double* lf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = PMDI_NT / 32, N = sp.N;
#pragma unroll 1
  for (int it = warp; it < nu * sp.Jmax; it += NW) {
    const int u = it / sp.Jmax, j = it - u * sp.Jmax;
    const int k = T.uinfo[u] >> 24, slot = T.uinfo[u] & 0xFFFFFF;
    const DsDev& ds = sp.ds[k];
    if (j >= ds.J || T.pend[u] < 0) continue;
    const int label = T.pend[u] & 0xFF, n = T.pend[u] >> 8;
    const long long row = (long long)slot * N + label;
    if (ds.type == T_GAUSSIAN) gauss_add_block(ds, row, j, n, (const double*)(xb + ds.x_off), lane);
    else if (ds.type == T_CATEGORICAL) cat_add_block(ds, row, j, (const int*)(xb + ds.x_off), lane);
    else nb_add_block(ds, row, j, n, (const int*)(xb + ds.x_off), lane, lf, sp.lf_T);
  }
  __syncthreads();
#pragma unroll 1
  for (int u = threadIdx.x; u < nu; u += PMDI_NT) T.pend[u] = -1;
}

// Proposal for one unit (dataset k, particle slot), by one warp: src/pmdi.jl:223-265.
// Written for code size (it runs once per unit per step between long streaming items and must
// not fall out of the instruction cache): label-indexed scratch in shared memory, plain loops.
__device__ __noinline__ void propose_unit(const SweepParams& sp, const CtaTables& T, int u, int step,
                                          const double* lp_empty_s, unsigned* rows_eval_s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = sp.K, N = sp.N, P = sp.P, par = step & 1;
  const int Npad = (N + 31) & ~31;
  const int k = T.uinfo[u] >> 24, slot = T.uinfo[u] & 0xFFFFFF;
  const DsDev& ds = sp.ds[k];
  const int p = T.ulog[u];
  double* lps = T.lp_s + (size_t)warp * Npad;  // lp[label], then f[label], then cumsum[label]
  const double lpe = lp_empty_s[k];
  const int cnt = T.ucount[u];
  const int fe = T.pend[u] >= 0 ? T.pe[u] : -1;  // the row that ran as fused items: one partial per block
  const double* part = T.part + T.pbase[u];
  const unsigned* urow = T.urow + (size_t)u * N;
  const int J = ds.J, qb = sp.qb;
#pragma unroll 1
  for (int m = lane; m < N; m += 32) lps[m] = lpe;
  __syncwarp();
#pragma unroll 1
  for (int e = lane; e < cnt; e += 32) {
    const unsigned ent = urow[e];
    double a = __ldg(ds.rc + (ent >> 8));
    const int stride = (e == fe) ? 1 : qb;
#pragma unroll 1
    for (int j = 0; j < J; j += stride) a += part[e * J + j];
    lps[ent & 0xFF] = a;
  }
  __syncwarp();
  double mx = -INFINITY;
#pragma unroll 1
  for (int m = lane; m < N; m += 32) {
    const double v = lps[m];
    mx = fmax(mx, v);
    if (sp.dbg_lp) sp.dbg_lp[(((size_t)step * K + k) * P + p) * N + m] = v;
  }
  mx = warp_max(mx);
  // f = exp(lp - max) * Pi ; sequential cumsum over labels (src/pmdi.jl:236-241)
#pragma unroll 1
  for (int m = lane; m < N; m += 32) lps[m] = pm_exp(lps[m] - mx) * T.Pi_s[k * N + m];
  __syncwarp();
  if (lane == 0) {
    double run = 0.0;
#pragma unroll 1
    for (int m = 0; m < N; ++m) { run += lps[m]; lps[m] = run; }
  }
  __syncwarp();
  const double tot = lps[N - 1];
  const double inc = pm_log(tot) + mx;
  int label;
  if (p == 0) {
    label = (int)sp.s_in[(size_t)k * sp.n_obs + sp.order[sp.n1 - 1 + step]] - 1;  // reference trajectory (:262)
  } else {
    const double uu = sp.tape_alloc ? sp.tape_alloc[((size_t)step * K + k) * P + p]
                                    : pm_uniform(sp.seed, sp.iter, DRAW_ALLOC, step, k, p);
    label = N - 1;
#pragma unroll 1
    for (int m0 = 0; m0 < N - 1; m0 += 32) {
      const int m = m0 + lane;
      const bool hit = (m < N - 1) && (pm_div(lps[m < N ? m : 0], tot) > uu);  // strict '>' (:255)
      const unsigned b = __ballot_sync(FULL, hit);
      if (b) { label = m0 + __ffs(b) - 1; break; }
    }
  }
  // bookkeeping of the chosen row: size, occupied-row list, pending add
  int pos = -1;
#pragma unroll 1
  for (int e0 = 0; e0 < cnt; e0 += 32) {
    const int e = e0 + lane;
    const bool hit = (e < cnt) && ((int)(urow[e] & 0xFF) == label);
    const unsigned b = __ballot_sync(FULL, hit);
    if (b) { pos = e0 + __ffs(b) - 1; break; }
  }
  if (lane == 0) {
    int n_new = 1;
    unsigned* urw = T.urow + (size_t)u * N;
    if (pos >= 0) {
      const unsigned ent = urw[pos] + (1u << 8);
      urw[pos] = ent;
      n_new = (int)(ent >> 8);
    } else {
      urw[cnt] = (unsigned)label | (1u << 8);
      T.ucount[u] = cnt + 1;
    }
    ds.n[(long long)slot * N + label] = n_new;
    T.pend[u] = label | (n_new << 8);
    T.pe[u] = pos >= 0 ? pos : cnt;
    sp.lab[((size_t)par * K + k) * P + p] = (uint8_t)label;
    sp.inc[((size_t)par * K + k) * P + p] = inc;
    sp.alloc_log[((size_t)step * K + k) * P + p] = (uint8_t)label;
    if (sp.dbg_alloc) sp.dbg_alloc[((size_t)step * K + k) * P + p] = label + 1;
    atomicAdd(&rows_eval_s[k], (unsigned)cnt);
  }
  __syncwarp();
}

// Weights of step `st` by ONE warp: fold every particle's K increments (dataset order, as
// src/pmdi.jl:210,233) and the Phi coupling (Phi_upweight!, src/misc.jl:50-59) into this CTA's
// private log-weights, then calc_ESS (src/misc.jl:15-25).  Fixed reduction shape -> identical
// bits in every CTA.  lab/inc are indexed by LOGICAL particle, so every load of a dataset is
// independent: 8 particles per lane are in flight per L2 round trip.
// Returns ESS <= P/2 (src/pmdi.jl:317); *mx_out = max log-weight.
#define PMDI_RCH 8
__device__ __noinline__ bool resolve_weights(const SweepParams& sp, int st, double* lw, int cta, double* mx_out) {
  const int lane = threadIdx.x & 31, K = sp.K, P = sp.P, par = st & 1;
  const uint8_t* lab_g = sp.lab + (size_t)par * K * P;
  const double* inc_g = sp.inc + (size_t)par * K * P;
  double mx = -INFINITY;
  double w[PMDI_RCH];
#pragma unroll 1
  for (int base = 0; base < P; base += 32 * PMDI_RCH) {
    unsigned long long lb[PMDI_RCH];
#pragma unroll
    for (int i = 0; i < PMDI_RCH; ++i) {
      const int p = base + lane + 32 * i;
      w[i] = (p < P) ? __ldcg(lw + p) : -INFINITY;
      lb[i] = 0ull;
    }
#pragma unroll 1
    for (int k = 0; k < K; ++k) {
      double t[PMDI_RCH];
      unsigned l8[PMDI_RCH];
#pragma unroll
      for (int i = 0; i < PMDI_RCH; ++i) {
        const int p = base + lane + 32 * i;
        t[i] = (p < P) ? ldcg_f64(inc_g + (size_t)k * P + p) : 0.0;
        l8[i] = (p < P) ? (unsigned)ldcg_u8(lab_g + (size_t)k * P + p) : 0u;
      }
#pragma unroll
      for (int i = 0; i < PMDI_RCH; ++i) {
        w[i] += t[i];
        lb[i] |= (unsigned long long)l8[i] << (8 * k);
      }
    }
    int idx = 0;
#pragma unroll 1
    for (int k1 = 0; k1 < K - 1; ++k1)
#pragma unroll 1
      for (int k2 = k1 + 1; k2 < K; ++k2) {
        const double phil = sp.l1phi[idx++];
#pragma unroll
        for (int i = 0; i < PMDI_RCH; ++i)
          w[i] += (((lb[i] >> (8 * k1)) & 0xFF) == ((lb[i] >> (8 * k2)) & 0xFF)) ? phil : 0.0;
      }
#pragma unroll
    for (int i = 0; i < PMDI_RCH; ++i) {
      const int p = base + lane + 32 * i;
      if (p < P) {
        __stcg(lw + p, w[i]);
        mx = fmax(mx, w[i]);
        if (cta == 0 && sp.dbg_lw) sp.dbg_lw[(size_t)st * P + p] = w[i];
      }
    }
  }
  mx = warp_max(mx);
  double num = 0.0, den = 0.0;
#pragma unroll 1
  for (int base = 0; base < P; base += 32 * PMDI_RCH) {
    if (P > 32 * PMDI_RCH) {  // otherwise the weights are still in registers
#pragma unroll
      for (int i = 0; i < PMDI_RCH; ++i) {
        const int p = base + lane + 32 * i;
        w[i] = (p < P) ? __ldcg(lw + p) : -INFINITY;
      }
    }
#pragma unroll
    for (int i = 0; i < PMDI_RCH; ++i) {
      const double e = pm_exp(w[i] - mx);  // exp(-inf) = 0 for the padding
      num += e;
      den += e * e;
    }
  }
  num = warp_sum(num);
  den = warp_sum(den);
  *mx_out = mx;
  const bool do_res = (num * num) / den <= 0.5 * (double)P;
  if (!do_res && cta == 0 && lane == 0) sp.ev_of_step[st] = -1;
  return do_res;
}

__device__ __forceinline__ unsigned long long ld_acquire_u64(const unsigned long long* p) {
  unsigned long long v;
  asm volatile("ld.acquire.gpu.global.u64 %0, [%1];" : "=l"(v) : "l"(p) : "memory");
  return v;
}

// ---- cold or once-per-step pieces as separate functions (code size, see pm_log) ----------------

// spin until the arrival counter reaches `target` (thread 0 of the caller); false = watchdog / error
__device__ __noinline__ bool bar_wait(const unsigned long long* bar, unsigned long long target, int* err) {
  const unsigned long long t0 = globaltimer_ns();
  unsigned spins = 0;
  while (ld_acquire_u64(bar) < target) {
    if (((++spins) & 0x3ffu) == 0) {
      if (__ldcg(err) != 0) return false;
      if (globaltimer_ns() - t0 > 4000000000ull) { atomicExch(err, 77); return false; }
    }
  }
  __threadfence();
  return true;
}

// blocking full grid barrier (resampling only)
__device__ __noinline__ bool grid_sync(unsigned long long* bar, unsigned long long& epoch, int G, int* err, int* s_fail) {
  __syncthreads();
  epoch += G;
  if (threadIdx.x == 0) {
    __threadfence();
    atomicAdd(bar, 1ull);
    if (!bar_wait(bar, epoch, err)) *s_fail = 1;
  }
  __syncthreads();
  return *s_fail == 0;
}

// occupied rows (and the logical particle) of every owned unit, from the state in HBM
__device__ __noinline__ void rebuild_rows(const SweepParams& sp, const CtaTables& T, int nu, int ev) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, N = sp.N;
#pragma unroll 1
  for (int u = warp; u < nu; u += PMDI_NT / 32) {
    const int k = T.uinfo[u] >> 24, slot = T.uinfo[u] & 0xFFFFFF;
    int cnt = 0;
#pragma unroll 1
    for (int m0 = 0; m0 < N; m0 += 32) {
      const int m = m0 + lane;
      const int nm = (m < N) ? ldcg_i32(sp.ds[k].n + (long long)slot * N + m) : 0;
      const unsigned b = __ballot_sync(FULL, nm > 0);
      if (nm > 0) T.urow[(size_t)u * N + cnt + __popc(b & ((1u << lane) - 1))] = (unsigned)m | ((unsigned)nm << 8);
      cnt += __popc(b);
    }
    if (lane == 0) {
      T.ucount[u] = cnt;
      T.ulog[u] = ldcg_i32(sp.logical_of + (ev & 1) * sp.P + slot);
    }
  }
}

// stage (asynchronously) the observation of one step into a ring buffer
__device__ __noinline__ void prefetch_obs(const SweepParams& sp, int step, unsigned char* xb) {
  if (step >= sp.steps) return;
  const int obs = sp.order[sp.n1 - 1 + step];
#pragma unroll 1
  for (int k = 0; k < sp.K; ++k) {
    const DsDev& ds = sp.ds[k];
    const int bytes = ds.Dp * (ds.type == T_GAUSSIAN ? 8 : 4);
    const unsigned char* src = (const unsigned char*)ds.xstage + (size_t)obs * bytes;
    unsigned char* dst = xb + ds.x_off;
#pragma unroll 1
    for (int o = threadIdx.x * 16; o < bytes; o += PMDI_NT * 16) cp_async16(dst + o, src + o);
  }
}

// This step's item list, by warp 0: [fused items of all units][plain items of all units], the
// partial-sum slots of every unit and its item count.  Returns through sm.
__device__ __noinline__ void build_item_offsets(const SweepParams& sp, const CtaTables& T, int nu, SweepSmem& sm) {
  const int lane = threadIdx.x & 31, qb = sp.qb;
  int runF = 0, runP = 0, runB = 0;
#pragma unroll 1
  for (int ub = 0; ub < nu; ub += 32) {
    const int u = ub + lane;
    int cF = 0, cP = 0, cB = 0;
    if (u < nu) {
      const int J = sp.ds[T.uinfo[u] >> 24].J, cnt = T.ucount[u], hp = T.pend[u] >= 0 ? 1 : 0;
      cF = hp ? J : 0;
      cP = (cnt - hp) * ((J + qb - 1) / qb);
      cB = cnt * J;
    }
    int iF = cF, iP = cP, iB = cB;
#pragma unroll
    for (int o = 1; o < 32; o <<= 1) {
      const int vF = __shfl_up_sync(FULL, iF, o), vP = __shfl_up_sync(FULL, iP, o), vB = __shfl_up_sync(FULL, iB, o);
      if (lane >= o) { iF += vF; iP += vP; iB += vB; }
    }
    if (u < nu) {
      T.foff[u] = runF + iF - cF; T.poff[u] = runP + iP - cP; T.pbase[u] = runB + iB - cB;
      T.remaining[u] = cF + cP;
    }
    runF += __shfl_sync(FULL, iF, 31); runP += __shfl_sync(FULL, iP, 31); runB += __shfl_sync(FULL, iB, 31);
  }
  if (lane == 0) {
    T.foff[nu] = runF; T.poff[nu] = runP;
    sm.n_fused = runF;
    sm.total_items = runF + runP;
    sm.item_ctr = 0;
    sm.n_defer = 0;
    if (runB > sp.item_cap) { atomicExch(sp.err, 78); sm.fail = 1; }
  }
}

// decode table: item -> (unit, row entry, first block), all threads
__device__ __noinline__ void build_item_codes(const SweepParams& sp, const CtaTables& T, int nu, int total, int nF) {
  const int qb = sp.qb;
#pragma unroll 1
  for (int it = threadIdx.x; it < total; it += PMDI_NT) {
    const bool fz = it < nF;
    const int* off = fz ? T.foff : T.poff;
    const int r0 = fz ? it : it - nF;
    int lo = 0, hi = nu - 1;  // last u with off[u] <= r0
#pragma unroll 1
    while (lo < hi) {
      const int mid = (lo + hi + 1) >> 1;
      if (off[mid] <= r0) lo = mid; else hi = mid - 1;
    }
    const int r = r0 - off[lo];
    unsigned code;
    if (fz) {
      code = PMDI_ITEM_FUSED | ((unsigned)lo << 13) | ((unsigned)T.pe[lo] << 5) | (unsigned)r;
    } else {
      const int JQ = (sp.ds[T.uinfo[lo] >> 24].J + qb - 1) / qb;
      int e = r / JQ;
      const int q = r - e * JQ;
      if (T.pend[lo] >= 0 && e >= T.pe[lo]) ++e;  // skip the fused row
      code = ((unsigned)lo << 13) | ((unsigned)e << 5) | (unsigned)(q * qb);
    }
    T.items[it] = code;
  }
}

// One work item by one warp; returns the predictive partial sum (identical in all lanes).
__device__ __noinline__ double run_item(const SweepParams& sp, const CtaTables& T, unsigned code,
                                        const unsigned char* xs_cur, const unsigned char* xs_prev,
                                        const double* lf) {
  const int lane = threadIdx.x & 31, N = sp.N, qb = sp.qb, lfT = sp.lf_T;
  const bool fused = (code & PMDI_ITEM_FUSED) != 0;
  const int u = (code >> 13) & 0x3FFFF, e = (code >> 5) & 0xFF, j0 = code & 31;
  const int k = T.uinfo[u] >> 24, slot = T.uinfo[u] & 0xFFFFFF;
  const DsDev& ds = sp.ds[k];
  const unsigned ent = T.urow[(size_t)u * N + e];
  const int m = ent & 0xFF, n = ent >> 8;  // for the fused row n is the size AFTER the pending add
  const long long row = (long long)slot * N + m;
  const int j1 = fused ? j0 + 1 : min(ds.J, j0 + qb);
  if (ds.type == T_GAUSSIAN) {
    if (fused) return gauss_fused_block(ds, row, j0, n, (const double*)(xs_prev + ds.x_off),
                                        (const double*)(xs_cur + ds.x_off), lane);
    return gauss_eval_item(ds, row, j0, j1, n, (const double*)(xs_cur + ds.x_off), lane);
  }
  if (ds.type == T_CATEGORICAL) {
    if (fused) {
      cat_add_block(ds, row, j0, (const int*)(xs_prev + ds.x_off), lane);
      __syncwarp();
    }
    return cat_eval_item(ds, row, j0, j1, (const int*)(xs_cur + ds.x_off), lane);
  }
  if (fused) {
    nb_add_block(ds, row, j0, n, (const int*)(xs_prev + ds.x_off), lane, lf, lfT);
    __syncwarp();
  }
  return nb_eval_item(ds, row, j0, j1, n, (const int*)(xs_cur + ds.x_off), lane, lf, lfT);
}

// Resampling after step `st` (draw_partstar src/misc.jl:27-47 by CTA 0, then every CTA moves the
// duplicated particles' rows, src/pmdi.jl:318-341 in dense form): two blocking grid barriers.
__device__ __noinline__ bool do_resample(const SweepParams& sp, const CtaTables& T, int nu, int st, int& ev,
                                         unsigned long long& epoch, double mx, double* lw, int* s_tmp, int* s_fail) {
  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5, NW = PMDI_NT / 32;
  const int cta = blockIdx.x, G = sp.G, K = sp.K, N = sp.N, P = sp.P;
  unsigned long long* bar = (unsigned long long*)sp.bar;
#pragma unroll 1
  for (int u = tid; u < nu; u += PMDI_NT) T.pend[u] = -1;
  if (cta == 0) resample_plan(sp, st, ev, mx, lw, s_tmp);
  if (!grid_sync(bar, epoch, G, sp.err, s_fail)) return false;
  const int ncopy = ldcg_i32(sp.plan_out);
  const int gw = cta * NW + warp, GW = G * NW;
#pragma unroll 1
  for (int idx = gw; idx < ncopy * K * N; idx += GW) {
    const int c = idx / (K * N), rem = idx - c * (K * N);
    const int k = rem / N, m = rem - k * N;
    const int2 cp = __ldcg(sp.copies + c);
    row_copy(sp.ds[k], (long long)cp.x * N + m, (long long)cp.y * N + m, lane);
  }
#pragma unroll 1
  for (int p = tid; p < P; p += PMDI_NT) __stcg(lw + p, 1.0);  // logweight .= 1.0 (src/pmdi.jl:319)
  if (cta == 0 && tid == 0) { sp.counters[0] += 1; sp.counters[1] += ncopy; }
  ++ev;
  if (!grid_sync(bar, epoch, G, sp.err, s_fail)) return false;
  rebuild_rows(sp, T, nu, ev);
  __syncthreads();
  return true;
}

extern "C" __global__ void __launch_bounds__(PMDI_NT, 1) k_sweep(const __grid_constant__ SweepParams sp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ SweepSmem sm;
  __shared__ int s_tmp[PMDI_NT + 2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NW = PMDI_NT / 32;
  const int cta = blockIdx.x, G = sp.G;
  const int K = sp.K, N = sp.N, P = sp.P;
  const int Npad = (N + 31) & ~31;

  // dynamic shared memory: [3 x observation][lf table][lp scratch][Pi][part][items][unit tables]
  unsigned char* xbuf[3] = {smem_raw, smem_raw + sp.sm_x_bytes, smem_raw + 2 * (size_t)sp.sm_x_bytes};
  double* lf = (double*)(smem_raw + 3 * (size_t)sp.sm_x_bytes);
  CtaTables T;
  T.lp_s = lf + sp.lf_T;
  T.Pi_s = T.lp_s + (size_t)NW * Npad;
  T.part = T.Pi_s + (size_t)K * N;
  T.items = (unsigned*)(T.part + sp.item_cap);
  T.urow = T.items + sp.item_cap;
  const int u0 = sp.cta_off[cta], nu = sp.cta_off[cta + 1] - u0;
  const int MU = sp.max_units;
  T.ucount = (int*)(T.urow + (size_t)MU * N);
  T.foff = T.ucount + MU;
  T.poff = T.foff + MU + 1;
  T.pbase = T.poff + MU + 1;
  T.uinfo = T.pbase + MU;
  T.ulog = T.uinfo + MU;
  T.pend = T.ulog + MU;
  T.pe = T.pend + MU;
  T.remaining = T.pe + MU;
  T.defer = T.remaining + MU;
  const int lfT = sp.lf_T;
  for (int i = tid; i < lfT; i += PMDI_NT) lf[i] = sp.lf_glob[i];
  for (int i = tid; i < K * N; i += PMDI_NT) T.Pi_s[i] = sp.Pi[i];
  if (tid < PMDI_MAX_K) sm.rows_eval[tid] = 0;
  for (int u = tid; u < nu; u += PMDI_NT) { T.uinfo[u] = sp.cta_units[u0 + u]; T.pend[u] = -1; T.pe[u] = 0; }
  if (tid == 0) { sm.res_step = -1; sm.res_claim = -1; sm.res_flag = 0; sm.n_defer = 0; sm.fail = 0; }

  double* lw = sp.lw + (size_t)cta * P;  // this CTA's private copy of the log-weights (L2)
  for (int p = tid; p < P; p += PMDI_NT) __stcg(lw + p, sp.lw_init);

  unsigned long long* bar = (unsigned long long*)sp.bar;
  unsigned long long epoch = 0;  // arrivals this CTA has made, times G
  int ev = 0;
  const bool timing = sp.phase_ns != nullptr && tid == 0;
  if (timing) {
    for (int i = 0; i < 8; ++i) sm.tacc[i] = 0;
    sm.t_prev = globaltimer_ns();
  }
#define PHASE_MARK(i_)                                   \
  if (timing) {                                          \
    const unsigned long long now_ = globaltimer_ns();    \
    sm.tacc[i_] += now_ - sm.t_prev;                     \
    sm.t_prev = now_;                                    \
  }
#define TRACE(tag_) if (sp.trace) trace_mark(sp, step, (tag_));
  __syncthreads();  // uinfo is visible to every warp
  rebuild_rows(sp, T, nu, ev);
  prefetch_obs(sp, 0, xbuf[0]);

  unsigned long long ep_prev = 0;  // epoch that completes the barrier of the previous step
  for (int step = 0; step < sp.steps; ++step) {
    const unsigned char* xs_cur = xbuf[step % 3];
    const unsigned char* xs_prev = xbuf[(step + 2) % 3];
    bool first_pass = true;
  redo_step:
    TRACE(1)
    if (first_pass) {
      cp_async_commit_wait_all();   // this step's observation has landed (own copies)
      __syncthreads();              // ... everybody's; last step's proposals are complete
      prefetch_obs(sp, step + 1, xbuf[(step + 1) % 3]);
    }
    if (tid < K) sm.lp_empty[tid] = sp.lp_empty[(size_t)step * K + tid];
    if (warp == 0) build_item_offsets(sp, T, nu, sm);
    __syncthreads();
    if (sm.fail) return;  // the other CTAs leave through the barrier's error check
    const int total = sm.total_items;
    build_item_codes(sp, T, nu, total, sm.n_fused);
    __syncthreads();
    PHASE_MARK(0)
    TRACE(2)
    // ------------------------------------------------------------------ the item queue
    for (;;) {
      int it = 0;
      if (lane == 0) it = atomicAdd(&sm.item_ctr, 1);
      it = __shfl_sync(FULL, it, 0);
      if (it >= total) break;
      const unsigned code = T.items[it];
      TRACE(0x100 | (code & 31) | (((code >> 5) & 7) << 5) | ((T.uinfo[(code >> 13) & 0x3FFFF] >> 24) << 12) | ((code >> 31) << 11))
      const double v = run_item(sp, T, code, xs_cur, xs_prev, lf);
      TRACE(3)
      const int u = (code >> 13) & 0x3FFFF;
      int act = 0;  // 0 nothing, 1 propose, 2 defer, 3 resolve the previous step first
      if (lane == 0) {
        T.part[T.pbase[u] + ((code >> 5) & 0xFF) * sp.ds[T.uinfo[u] >> 24].J + (code & 31)] = v;
        __threadfence_block();
        if (atomicSub(&T.remaining[u], 1) == 1) {  // this warp finished the unit
          const int rs = *(volatile int*)&sm.res_step;
          if (rs >= step - 1) act = (*(volatile int*)&sm.res_flag) ? 0 : 1;
          else if (rs == step - 2 && ld_acquire_u64(bar) >= ep_prev &&
                   atomicCAS(&sm.res_claim, step - 2, step - 1) == step - 2) act = 3;
          else act = 2;
        }
      }
      act = __shfl_sync(FULL, act, 0);
      if (act == 3) {
        __threadfence();
        TRACE(9)
        double mxv;
        const bool r = resolve_weights(sp, step - 1, lw, cta, &mxv);
        if (lane == 0) {
          sm.res_mx = mxv;
          sm.res_flag = r ? 1 : 0;
          __threadfence_block();
          *(volatile int*)&sm.res_step = step - 1;
        }
        __syncwarp();
        TRACE(10)
        act = r ? 0 : 1;
      }
      if (act == 1) {
        __threadfence_block();
        TRACE(4)
        propose_unit(sp, T, u, step, sm.lp_empty, sm.rows_eval);
        TRACE(5)
      } else if (act == 2) {
        if (lane == 0) T.defer[atomicAdd(&sm.n_defer, 1)] = u;
      }
    }
    TRACE(6)
    __syncthreads();
    PHASE_MARK(1)
    // ------------------------------------------------------------------ previous step resolved?
    if (step > 0 && sm.res_step < step - 1) {
      if (warp == 0) {
        if (lane == 0 && !bar_wait(bar, ep_prev, sp.err)) sm.fail = 1;
        __syncwarp();
        PHASE_MARK(4)
        TRACE(9)
        double mxv;
        const bool r = resolve_weights(sp, step - 1, lw, cta, &mxv);
        if (lane == 0) {
          sm.res_mx = mxv;
          sm.res_flag = r ? 1 : 0;
          sm.res_claim = step - 1;
          sm.res_step = step - 1;
        }
        TRACE(10)
      }
      __syncthreads();
      if (sm.fail) return;
      PHASE_MARK(5)
    }
    if (sm.res_flag) {
      // resampling after step-1: this step's first pass applied every pending add; its predictive
      // sums are discarded and the step is redone on the moved particles
      if (!do_resample(sp, T, nu, step - 1, ev, epoch, sm.res_mx, lw, s_tmp, &sm.fail)) return;
      if (tid == 0) sm.res_flag = 0;
      PHASE_MARK(6)
      first_pass = false;
      goto redo_step;
    }
    // ------------------------------------------------------------------ deferred proposals, arrive
    {
      const int nd = sm.n_defer;
      for (int i = warp; i < nd; i += NW) {
        TRACE(4)
        propose_unit(sp, T, T.defer[i], step, sm.lp_empty, sm.rows_eval);
        TRACE(5)
      }
      for (int u = warp; u < nu; u += NW)  // units with no occupied row at all
        if (T.foff[u + 1] == T.foff[u] && T.poff[u + 1] == T.poff[u])
          propose_unit(sp, T, u, step, sm.lp_empty, sm.rows_eval);
    }
    __syncthreads();
    epoch += G;
    ep_prev = epoch;
    if (tid == 0) {
      __threadfence();
      atomicAdd(bar, 1ull);
    }
    PHASE_MARK(2)
    TRACE(7)
  }
  // ------------------------------------------------------------------ weights of the last step
  {
    const int st = sp.steps - 1;
    if (warp == 0) {
      if (lane == 0 && !bar_wait(bar, ep_prev, sp.err)) sm.fail = 1;
      __syncwarp();
      double mxv;
      const bool r = resolve_weights(sp, st, lw, cta, &mxv);
      if (lane == 0) { sm.res_mx = mxv; sm.res_flag = r ? 1 : 0; }
    }
    __syncthreads();
    if (sm.fail) return;
    PHASE_MARK(4)
    if (sm.res_flag) {
      flush_adds(sp, T, nu, xbuf[st % 3], lf);
      if (!do_resample(sp, T, nu, st, ev, epoch, sm.res_mx, lw, s_tmp, &sm.fail)) return;
      PHASE_MARK(6)
    }
  }
  __syncthreads();
  if (tid < K) atomicAdd(sp.rows_eval + tid, (unsigned long long)sm.rows_eval[tid]);
  if (cta == 0)
    for (int p = tid; p < P; p += PMDI_NT) sp.lw_out[p] = __ldcg(lw + p);
  if (timing)
    for (int i = 0; i < 8; ++i) sp.phase_ns[(size_t)cta * 8 + i] = sm.tacc[i];
  if (cta == 0 && tid == 0) sp.counters[2] = ev;
#undef PHASE_MARK
#undef TRACE
}
