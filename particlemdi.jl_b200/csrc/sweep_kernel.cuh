// The conditional-SMC sweep as ONE persistent cooperative kernel: the per-observation loop of
// the reference (src/pmdi.jl:209-342) runs on the device with two grid barriers per step and no
// host involvement.  One CTA per SM; every (dataset, particle-slot) unit of statistics is owned
// by one CTA for the whole sweep, so the add of step t and the predictive of step t+1 need no
// grid-wide ordering.
//
//   phase A  predictive      calc_logprob for every occupied cluster row of the CTA's units
//                            (src/pmdi.jl:218-220) -> part[row][block]
//   -- grid barrier --
//   phase B  proposal        per particle and dataset: sum partials, softmax-cdf, draw, weight
//                            increment (src/pmdi.jl:223-265); Phi coupling (src/misc.jl:50-59)
//   -- grid barrier --
//   phase E  ESS             every CTA evaluates calc_ESS (src/misc.jl:15-25) redundantly
//   phase C  cluster_add!    chosen row of every owned unit (src/pmdi.jl:275-310, dense form)
//   [resampling steps only]  CTA 0: draw_partstar (src/misc.jl:27-47) + copy plan; barrier;
//                            all CTAs move the duplicated particles' rows; barrier
#pragma once
#include "cluster_types.cuh"

struct SweepSmem {
  double red[3 * 32];
  double lp_empty[PMDI_MAX_K];
  double inc[PMDI_NT / 32];
  int lab[PMDI_NT / 32];
  int n_entries;
  int item_ctr;
  int flag;
  unsigned rows_eval[PMDI_MAX_K];
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
// slot_of[next][P], copies[], plan_out[0] = number of copies.
// The Fisher-Yates shuffle followed by partstar[1]=1 and sort! only decides WHICH element of the
// sorted systematic sample is replaced by the reference particle: the one the shuffle moves to
// position 1.  That index is traced through the swaps without moving anything.
__device__ void resample_plan(const SweepParams& sp, int step, int ev, double mx, int* s_tmp) {
  const int P = sp.P, t = threadIdx.x;
  const int* slot_cur = sp.slot_of + (ev & 1) * P;
  int* slot_nxt = sp.slot_of + ((ev + 1) & 1) * P;
  for (int p = t; p < P; p += PMDI_NT) {
    sp.sc_w[p] = exp(ldcg_f64(sp.lw + p) - mx);
    // Fisher-Yates pick for position pos = p+1 (entry index p), pos >= 2
    const double us = sp.tape_shuffle ? sp.tape_shuffle[(size_t)step * P + p]
                                      : pmdi_philox_uniform(sp.seed, sp.iter, DRAW_SHUFFLE, step, 0, p);
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
                                    : pmdi_philox_uniform(sp.seed, sp.iter, DRAW_RESAMP, step, 0, 0);
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
    int lo = 0, hi = P;  // answer in [lo, hi]; hi == P means none
    while (lo < hi) {
      const int mid = (lo + hi) >> 1;
      if (sp.sc_pp[mid] / tot >= ui) hi = mid; else lo = mid + 1;
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
    if (sp.sc_a[i]) {
      const int dst = sp.sc_j[sp.sc_c[i]];
      slot_nxt[i] = dst;
      sp.copies[sp.sc_c[i]] = make_int2(src, dst);
    } else {
      slot_nxt[i] = src;
    }
  }
  if (t == 0) {
    sp.plan_out[0] = n_extra;
    sp.ev_of_step[step] = ev;
  }
  __syncthreads();
}

extern "C" __global__ void __launch_bounds__(PMDI_NT, 1) k_sweep(const SweepParams sp) {
  extern __shared__ __align__(16) unsigned char smem_raw[];
  __shared__ SweepSmem sm;
  __shared__ int s_tmp[PMDI_NT + 2];

  const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
  const int NW = PMDI_NT / 32;
  const int cta = blockIdx.x, G = sp.G;
  const int K = sp.K, N = sp.N, P = sp.P;

  unsigned char* xs_raw = smem_raw;
  double* lf = (double*)(smem_raw + sp.sm_x_bytes);
  unsigned* rowlist = (unsigned*)(smem_raw + sp.sm_x_bytes + (size_t)sp.lf_T * 8);
  const int lfT = sp.lf_T;
  for (int i = tid; i < lfT; i += PMDI_NT) lf[i] = sp.lf_glob[i];
  if (tid < PMDI_MAX_K) sm.rows_eval[tid] = 0;

  const int u0 = sp.cta_off[cta], u1 = sp.cta_off[cta + 1];
  const int empty_slot = P + 1;
  const long long empty_row = (long long)empty_slot * N;

  unsigned epoch = 0;
  int ev = 0;
  const bool timing = sp.phase_ns != nullptr && cta == 0 && tid == 0;
  unsigned long long tacc[6] = {0, 0, 0, 0, 0, 0}, t_prev = 0;
  if (timing) t_prev = globaltimer_ns();
#define PHASE_MARK(i_)                                   \
  if (timing) {                                          \
    const unsigned long long now_ = globaltimer_ns();    \
    tacc[i_] += now_ - t_prev;                           \
    t_prev = now_;                                       \
  }

  // B-phase geometry: warp (pl, k) of a CTA round handles dataset k of one particle
  const int ppb = NW / K;
  const int b_rounds = (P + ppb * G - 1) / (ppb * G);

  for (int step = 0; step < sp.steps; ++step) {
    const int obs = sp.order[sp.n1 - 1 + step];
    const int* slot_cur = sp.slot_of + (ev & 1) * P;

    // ------------------------------------------------------------------ stage the observation
    for (int k = 0; k < K; ++k) {
      const DsDev& ds = sp.ds[k];
      if (ds.type == T_GAUSSIAN) {
        const double* src = (const double*)ds.x + (size_t)obs * ds.Dp;
        double* dst = (double*)(xs_raw + ds.x_off);
        for (int q = tid; q < ds.Dp; q += PMDI_NT) dst[q] = src[q];
      } else {
        const int* src = (const int*)ds.x + (size_t)obs * ds.Dp;
        int* dst = (int*)(xs_raw + ds.x_off);
        const int skip = ds.type == T_CATEGORICAL ? 0 : -1;
        for (int q = tid; q < ds.Dp; q += PMDI_NT) dst[q] = ds.flag[q] ? src[q] : skip;
      }
    }
    if (tid == 0) { sm.n_entries = 0; sm.item_ctr = 0; }
    __syncthreads();
    // ------------------------------------------------------------------ occupied rows of my units
    for (int idx = tid; idx < (u1 - u0) * N; idx += PMDI_NT) {
      const int ul = idx / N, m = idx - ul * N;
      const int unit = sp.cta_units[u0 + ul];
      const int k = unit >> 24, slot = unit & 0xFFFFFF;
      bool occ;
      if (slot == empty_slot) occ = (m == 0);
      else occ = ldcg_i32(sp.ds[k].n + (long long)slot * N + m) > 0;
      if (occ) {
        const int e = atomicAdd(&sm.n_entries, 1);
        rowlist[e] = ((unsigned)ul << 8) | (unsigned)m;
        atomicAdd(&sm.rows_eval[k], 1u);
      }
    }
    __syncthreads();
    PHASE_MARK(0)
    // ------------------------------------------------------------------ phase A: predictive
    {
      const int n_entries = sm.n_entries;
      const int JM = sp.Jmax;  // item index space per entry
      // items are (entry, block) pairs handed out dynamically to warps
      for (;;) {
        int it = 0;
        if (lane == 0) it = atomicAdd(&sm.item_ctr, 1);
        it = __shfl_sync(FULL, it, 0);
        // decode: walk entries in order; entry e contributes J_k(e) items
        // (n_entries * J is small; a flat decode by division on the max J wastes few grabs)
        const int e = it / JM, j = it - e * JM;
        if (e >= n_entries) break;
        const unsigned ent = rowlist[e];
        const int ul = ent >> 8, m = ent & 0xFF;
        const int unit = sp.cta_units[u0 + ul];
        const int k = unit >> 24, slot = unit & 0xFFFFFF;
        const DsDev& ds = sp.ds[k];
        if (j >= ds.J) continue;
        const long long row = (long long)slot * N + m;
        double v;
        if (ds.type == T_GAUSSIAN) {
          const int n = (slot == empty_slot) ? 0 : ldcg_i32(ds.n + row);
          v = gauss_eval_block(ds, row, j, n, (const double*)(xs_raw + ds.x_off), lane);
        } else if (ds.type == T_CATEGORICAL) {
          v = cat_eval_block(ds, row, j, (const int*)(xs_raw + ds.x_off), lane);
        } else {
          const int n = (slot == empty_slot) ? 0 : ldcg_i32(ds.n + row);
          v = nb_eval_block(ds, row, j, n, (const int*)(xs_raw + ds.x_off), lane, lf, lfT);
        }
        if (lane == 0) ds.part[row * ds.J + j] = v;
      }
    }
    PHASE_MARK(1)
    if (!grid_barrier(sp.bar, epoch, G, sp.err)) return;
    PHASE_MARK(2)

    // ------------------------------------------------------------------ phase B: proposal
    if (warp < K) {
      const DsDev& ds = sp.ds[warp];
      double acc = ds.rc[0];
      for (int j = 0; j < ds.J; ++j) acc += ldcg_f64(ds.part + empty_row * ds.J + j);
      if (lane == 0) sm.lp_empty[warp] = acc;
    }
    __syncthreads();
    for (int r = 0; r < b_rounds; ++r) {
      const int pl = warp / K, k = warp - pl * K;
      const int p = (r * G + cta) * ppb + pl;  // logical particle
      const bool active = (pl < ppb) && (p < P);
      int label = 0;
      double inc = 0.0;
      if (active) {
        const DsDev& ds = sp.ds[k];
        const int slot = ldcg_i32(slot_cur + p);
        const long long row0 = (long long)slot * N;
        double lpv[PMDI_MAX_N / 32];
        double mx = -INFINITY;
#pragma unroll
        for (int c = 0; c < PMDI_MAX_N / 32; ++c) {
          const int m = c * 32 + lane;
          lpv[c] = -INFINITY;
          if (m < N) {
            const int nm = ldcg_i32(ds.n + row0 + m);
            double a;
            if (nm > 0) {
              a = ds.rc[nm];
              const double* pp = ds.part + (row0 + m) * ds.J;
              for (int j = 0; j < ds.J; ++j) a += ldcg_f64(pp + j);
            } else {
              a = sm.lp_empty[k];
            }
            lpv[c] = a;
            mx = fmax(mx, a);
            if (sp.dbg_lp) sp.dbg_lp[(((size_t)step * K + k) * P + p) * N + m] = a;
          }
        }
        mx = warp_max(mx);
        // f = exp(lp - max) * Pi ; sequential cumsum over labels (src/pmdi.jl:236-241)
        double cv[PMDI_MAX_N / 32];
        double run = 0.0;
#pragma unroll
        for (int c = 0; c < PMDI_MAX_N / 32; ++c) {
          const int m = c * 32 + lane;
          double f = 0.0;
          if (m < N) f = exp(lpv[c] - mx) * sp.Pi[k * N + m];
          cv[c] = 0.0;
          if (c * 32 < N) {
            const int lim = min(32, N - c * 32);
#pragma unroll 1
            for (int l = 0; l < lim; ++l) {
              run += __shfl_sync(FULL, f, l);
              if (lane == l) cv[c] = run;
            }
          }
        }
        const double tot = run;
        inc = log(tot) + mx;
        if (p == 0) {
          label = (int)sp.s_in[(size_t)k * sp.n_obs + obs] - 1;  // reference trajectory (:262)
        } else {
          const double u = sp.tape_alloc ? sp.tape_alloc[((size_t)step * K + k) * P + p]
                                         : pmdi_philox_uniform(sp.seed, sp.iter, DRAW_ALLOC, step, k, p);
          label = N - 1;
          bool found = false;
#pragma unroll
          for (int c = 0; c < PMDI_MAX_N / 32; ++c) {
            if (c * 32 < N && !found) {
              const int m = c * 32 + lane;
              const bool hit = (m < N - 1) && (cv[c] / tot > u);  // strict '>' (:255)
              const unsigned b = __ballot_sync(FULL, hit);
              if (b) { label = c * 32 + __ffs(b) - 1; found = true; }
            }
          }
        }
        if (lane == 0) {
          sp.lab[(size_t)k * P + slot] = (uint8_t)label;
          sp.alloc_log[((size_t)step * K + k) * P + p] = (uint8_t)label;
          ds.n[row0 + label] = ldcg_i32(ds.n + row0 + label) + 1;
          if (sp.dbg_alloc) sp.dbg_alloc[((size_t)step * K + k) * P + p] = label + 1;
          sm.inc[warp] = inc;
          sm.lab[warp] = label;
        }
      }
      __syncthreads();
      if (active && k == 0 && lane == 0) {  // weight update in dataset order, then Phi coupling
        double w = ldcg_f64(sp.lw + p);
        for (int kk = 0; kk < K; ++kk) w += sm.inc[pl * K + kk];
        int idx = 0;
        for (int k1 = 0; k1 < K - 1; ++k1)
          for (int k2 = k1 + 1; k2 < K; ++k2) {
            w += (sm.lab[pl * K + k1] == sm.lab[pl * K + k2]) ? sp.l1phi[idx] : 0.0;
            ++idx;
          }
        sp.lw[p] = w;
        if (sp.dbg_lw) sp.dbg_lw[(size_t)step * P + p] = w;
      }
      __syncthreads();
    }
    PHASE_MARK(3)
    if (!grid_barrier(sp.bar, epoch, G, sp.err)) return;
    PHASE_MARK(2)

    // ------------------------------------------------------------------ phase E: ESS (redundant)
    double mx = -INFINITY;
    for (int p = tid; p < P; p += PMDI_NT) mx = fmax(mx, ldcg_f64(sp.lw + p));
    mx = block_max(mx, sm.red);
    double num = 0.0, den = 0.0;
    for (int p = tid; p < P; p += PMDI_NT) {
      const double w = exp(ldcg_f64(sp.lw + p) - mx);
      num += w;
      den += w * w;
    }
    block_sum2(num, den, sm.red);
    const bool do_res = (num * num) / den <= 0.5 * (double)P;
    if (!do_res && cta == 0 && tid == 0) sp.ev_of_step[step] = -1;

    if (do_res && cta == 0) resample_plan(sp, step, ev, mx, s_tmp);

    // ------------------------------------------------------------------ phase C: cluster_add!
    for (int it = warp; it < (u1 - u0) * sp.Jmax; it += NW) {
      const int ul = it / sp.Jmax, j = it - ul * sp.Jmax;
      const int unit = sp.cta_units[u0 + ul];
      const int k = unit >> 24, slot = unit & 0xFFFFFF;
      const DsDev& ds = sp.ds[k];
      if (slot == empty_slot || j >= ds.J) continue;
      const int label = ldcg_u8(sp.lab + (size_t)k * P + slot);
      const long long row = (long long)slot * N + label;
      const int n = ldcg_i32(ds.n + row);
      if (ds.type == T_GAUSSIAN) gauss_add_block(ds, row, j, n, (const double*)(xs_raw + ds.x_off), lane);
      else if (ds.type == T_CATEGORICAL) cat_add_block(ds, row, j, (const int*)(xs_raw + ds.x_off), lane);
      else nb_add_block(ds, row, j, n, (const int*)(xs_raw + ds.x_off), lane, lf, lfT);
    }
    PHASE_MARK(4)

    if (do_res) {
      if (!grid_barrier(sp.bar, epoch, G, sp.err)) return;
      const int ncopy = ldcg_i32(sp.plan_out);
      const int gw = cta * NW + warp, GW = G * NW;
      for (int idx = gw; idx < ncopy * K * N; idx += GW) {
        const int c = idx / (K * N), rem = idx - c * (K * N);
        const int k = rem / N, m = rem - k * N;
        const int2 cp = __ldcg(sp.copies + c);
        row_copy(sp.ds[k], (long long)cp.x * N + m, (long long)cp.y * N + m, lane);
      }
      if (cta == 0) {
        for (int p = tid; p < P; p += PMDI_NT) sp.lw[p] = 1.0;  // logweight .= 1.0 (:319)
        if (tid == 0) { sp.counters[0] += 1; sp.counters[1] += ncopy; }
      }
      ++ev;
      if (!grid_barrier(sp.bar, epoch, G, sp.err)) return;
      PHASE_MARK(5)
    } else {
      __syncthreads();  // adds of this step are visible to the whole CTA before the next predictive
    }
  }
  if (tid < K) atomicAdd(sp.rows_eval + tid, (unsigned long long)sm.rows_eval[tid]);
  if (timing)
    for (int i = 0; i < 6; ++i) sp.phase_ns[i] = tacc[i];
  if (cta == 0 && tid == 0) sp.counters[2] = ev;
#undef PHASE_MARK
}
