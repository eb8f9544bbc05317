// The conditional-SMC sweep as ONE persistent cooperative kernel: the per-observation loop of
// the reference (src/pmdi.jl:209-342) runs on the device with ONE grid barrier per observation
// and no host involvement.  One CTA per SM.  Every (dataset, particle-slot) unit of statistics
// is owned by one CTA for the whole sweep, so predictive, proposal and add of a unit are
// CTA-local; only the particle weights need the whole grid.
//
// Per observation step a CTA runs ONE dynamically scheduled queue of work items
// (unit, occupied row, 256-feature block), handed to warps through a shared-memory counter:
//   * item        calc_logprob of the row block against the staged observation
//                 (src/pmdi.jl:218-220); for the row the particle chose in the previous step the
//                 item first applies the pending cluster_add! (src/pmdi.jl:300) in the same pass;
//   * proposal    the warp that finishes the last item of a unit sums the unit's partials,
//                 builds the softmax-cdf, draws the label and the weight increment
//                 (src/pmdi.jl:223-265) -> lab/inc[step parity][k][slot], pending add;
//   -- grid barrier --
//   * weights     every CTA folds all increments and the Phi coupling (src/misc.jl:50-59) into
//                 its private copy of the log-weights and evaluates calc_ESS (src/misc.jl:15-25):
//                 identical bits everywhere, no second barrier;
//   * resampling steps only: pending adds are flushed, CTA 0 runs draw_partstar
//                 (src/misc.jl:27-47) + the copy plan; barrier; all CTAs move the duplicated
//                 particles' rows; barrier.
// Observation rows are prefetched one step ahead with cp.async into a 3-deep ring (the previous
// row is still needed by the fused add).
#pragma once
#include "cluster_types.cuh"

struct SweepSmem {
  double red[3 * 32];
  int item_ctr;
  int total_items;
  int plan_drop;
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
    sp.sc_w[p] = exp(lw[p] - mx);
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
    int lo = 0, hi = P;  // hi == P means none
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

// per-CTA views into dynamic shared memory
struct CtaTables {
  unsigned* urow;   // [max_units][N]   occupied rows of a unit: label | n << 8
  int* ucount;      // [max_units]      number of occupied rows
  int* uoff;        // [max_units + 1]  first item of a unit in this step's queue
  int* uinfo;       // [max_units]      k << 24 | slot
  int* pend;        // [max_units]      pending add: label | n_after << 8, or -1
  int* remaining;   // [max_units]      items of the unit not yet finished this step
  unsigned* items;  // [item_cap]       u << 13 | e << 5 | j
  double* part;     // [item_cap]       predictive partial sums of this step
  double* lp_s;     // [NW][Npad]       per-warp proposal scratch
};

// cluster_add! of the pending row of every unit against observation buffer xb (resampling steps:
// the copies need the statistics up to date).
__device__ __noinline__ void flush_adds(const SweepParams& sp, const CtaTables& T, int nu, const unsigned char* xb,
                                        const double* lf) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, NW = PMDI_NT / 32, N = sp.N;
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
  for (int u = threadIdx.x; u < nu; u += PMDI_NT) T.pend[u] = -1;
}

// Proposal for one unit (dataset k, particle slot), by one warp: src/pmdi.jl:223-265.
__device__ __noinline__ void propose_unit(const SweepParams& sp, const CtaTables& T, int u, int step, int ev,
                                          unsigned* rows_eval_s) {
  const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
  const int K = sp.K, N = sp.N, P = sp.P, par = step & 1;
  const int Npad = (N + 31) & ~31;
  const int k = T.uinfo[u] >> 24, slot = T.uinfo[u] & 0xFFFFFF;
  const DsDev& ds = sp.ds[k];
  const int p = ldcg_i32(sp.logical_of + (ev & 1) * P + slot);  // logical particle: RNG address, log index
  const long long row0 = (long long)slot * N;
  double* lps = T.lp_s + (size_t)warp * Npad;
  const double lpe = sp.lp_empty[(size_t)step * K + k];
  for (int m = lane; m < Npad; m += 32) lps[m] = lpe;
  __syncwarp();
  const int cnt = T.ucount[u];
  const double* part = T.part + T.uoff[u];
  for (int e = lane; e < cnt; e += 32) {
    const unsigned ent = T.urow[(size_t)u * N + e];
    double a = ds.rc[ent >> 8];
    const int JQ = (ds.J + PMDI_QB - 1) / PMDI_QB;
    for (int q = 0; q < JQ; ++q) a += part[e * JQ + q];
    lps[ent & 0xFF] = a;
  }
  __syncwarp();
  double lpv[PMDI_MAX_N / 32];
  double mx = -INFINITY;
#pragma unroll
  for (int c = 0; c < PMDI_MAX_N / 32; ++c) {
    lpv[c] = -INFINITY;
    if (c * 32 < N) {
      const int m = c * 32 + lane;
      if (m < N) {
        lpv[c] = lps[m];
        mx = fmax(mx, lpv[c]);
        if (sp.dbg_lp) sp.dbg_lp[(((size_t)step * K + k) * P + p) * N + m] = lpv[c];
      }
    }
  }
  mx = warp_max(mx);
  // f = exp(lp - max) * Pi ; sequential cumsum over labels (src/pmdi.jl:236-241)
  double cv[PMDI_MAX_N / 32];
  double run = 0.0;
#pragma unroll
  for (int c = 0; c < PMDI_MAX_N / 32; ++c) {
    cv[c] = 0.0;
    if (c * 32 < N) {
      const int m = c * 32 + lane;
      double f = 0.0;
      if (m < N) f = exp(lpv[c] - mx) * sp.Pi[k * N + m];
      const int lim = min(32, N - c * 32);
#pragma unroll 1
      for (int l = 0; l < lim; ++l) {
        run += __shfl_sync(FULL, f, l);
        if (lane == l) cv[c] = run;
      }
    }
  }
  const double tot = run;
  const double inc = log(tot) + mx;
  int label;
  if (p == 0) {
    label = (int)sp.s_in[(size_t)k * sp.n_obs + sp.order[sp.n1 - 1 + step]] - 1;  // reference trajectory (:262)
  } else {
    const double uu = sp.tape_alloc ? sp.tape_alloc[((size_t)step * K + k) * P + p]
                                    : pmdi_philox_uniform(sp.seed, sp.iter, DRAW_ALLOC, step, k, p);
    label = N - 1;
    bool found = false;
#pragma unroll
    for (int c = 0; c < PMDI_MAX_N / 32; ++c) {
      if (c * 32 < N && !found) {
        const int m = c * 32 + lane;
        const bool hit = (m < N - 1) && (cv[c] / tot > uu);  // strict '>' (:255)
        const unsigned b = __ballot_sync(FULL, hit);
        if (b) { label = c * 32 + __ffs(b) - 1; found = true; }
      }
    }
  }
  // bookkeeping of the chosen row: size, occupied-row list, pending add
  int pos = -1;
  for (int e0 = 0; e0 < cnt; e0 += 32) {
    const int e = e0 + lane;
    const bool hit = (e < cnt) && ((int)(T.urow[(size_t)u * N + e] & 0xFF) == label);
    const unsigned b = __ballot_sync(FULL, hit);
    if (b) { pos = e0 + __ffs(b) - 1; break; }
  }
  if (lane == 0) {
    int n_new = 1;
    if (pos >= 0) {
      const unsigned ent = T.urow[(size_t)u * N + pos] + (1u << 8);
      T.urow[(size_t)u * N + pos] = ent;
      n_new = (int)(ent >> 8);
    } else {
      T.urow[(size_t)u * N + cnt] = (unsigned)label | (1u << 8);
      T.ucount[u] = cnt + 1;
    }
    ds.n[row0 + label] = n_new;
    T.pend[u] = label | (n_new << 8);
    sp.lab[((size_t)par * K + k) * P + slot] = (uint8_t)label;
    sp.inc[((size_t)par * K + k) * P + slot] = inc;
    sp.alloc_log[((size_t)step * K + k) * P + p] = (uint8_t)label;
    if (sp.dbg_alloc) sp.dbg_alloc[((size_t)step * K + k) * P + p] = label + 1;
    atomicAdd(&rows_eval_s[k], (unsigned)cnt);
  }
  __syncwarp();
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

  // dynamic shared memory: [3 x observation][lf table][lp scratch][part][items][unit tables]
  unsigned char* xbuf[3] = {smem_raw, smem_raw + sp.sm_x_bytes, smem_raw + 2 * (size_t)sp.sm_x_bytes};
  double* lf = (double*)(smem_raw + 3 * (size_t)sp.sm_x_bytes);
  CtaTables T;
  T.lp_s = lf + sp.lf_T;
  T.part = T.lp_s + (size_t)NW * Npad;
  T.items = (unsigned*)(T.part + sp.item_cap);
  T.urow = T.items + sp.item_cap;
  const int u0 = sp.cta_off[cta], nu = sp.cta_off[cta + 1] - u0;
  T.ucount = (int*)(T.urow + (size_t)sp.max_units * N);
  T.uoff = T.ucount + sp.max_units;
  T.uinfo = T.uoff + sp.max_units + 1;
  T.pend = T.uinfo + sp.max_units;
  T.remaining = T.pend + sp.max_units;
  const int lfT = sp.lf_T;
  for (int i = tid; i < lfT; i += PMDI_NT) lf[i] = sp.lf_glob[i];
  if (tid < PMDI_MAX_K) sm.rows_eval[tid] = 0;
  for (int u = tid; u < nu; u += PMDI_NT) { T.uinfo[u] = sp.cta_units[u0 + u]; T.pend[u] = -1; }

  double* lw = sp.lw + (size_t)cta * P;  // private copy, thread t owns p = t, t + NT, ...
  for (int p = tid; p < P; p += PMDI_NT) lw[p] = sp.lw_init;

  unsigned epoch = 0;
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

  int tr_n = 0;
#define TRACE(tag_)                                                                              \
  if (sp.trace && cta == sp.trace_cta && step == sp.trace_step && lane == 0 && tr_n < 127) {     \
    sp.trace[warp * 128 + (++tr_n)] = ((unsigned long long)(tag_) << 48) | (clock64() & 0xFFFFFFFFFFFFull); \
    sp.trace[warp * 128] = tr_n;                                                                 \
  }
  // occupied rows of every owned unit, from the statistics in HBM
  auto rebuild_rows = [&]() {
    for (int u = warp; u < nu; u += NW) {
      const int k = T.uinfo[u] >> 24, slot = T.uinfo[u] & 0xFFFFFF;
      int cnt = 0;
      for (int m0 = 0; m0 < N; m0 += 32) {
        const int m = m0 + lane;
        const int nm = (m < N) ? ldcg_i32(sp.ds[k].n + (long long)slot * N + m) : 0;
        const unsigned b = __ballot_sync(FULL, nm > 0);
        if (nm > 0) T.urow[(size_t)u * N + cnt + __popc(b & ((1u << lane) - 1))] = (unsigned)m | ((unsigned)nm << 8);
        cnt += __popc(b);
      }
      if (lane == 0) T.ucount[u] = cnt;
    }
  };
  // stage (asynchronously) the observation of one step into ring buffer b
  auto prefetch_obs = [&](int step, int b) {
    if (step >= sp.steps) return;
    const int obs = sp.order[sp.n1 - 1 + step];
    for (int k = 0; k < K; ++k) {
      const DsDev& ds = sp.ds[k];
      const int bytes = ds.Dp * (ds.type == T_GAUSSIAN ? 8 : 4);
      const unsigned char* src = (const unsigned char*)ds.xstage + (size_t)obs * bytes;
      unsigned char* dst = xbuf[b] + ds.x_off;
      for (int o = tid * 16; o < bytes; o += PMDI_NT * 16) cp_async16(dst + o, src + o);
    }
  };
  __syncthreads();  // uinfo is visible to every warp
  rebuild_rows();
  prefetch_obs(0, 0);

  for (int step = 0; step < sp.steps; ++step) {
    const int par = step & 1;
    const unsigned char* xs_cur = xbuf[step % 3];
    const unsigned char* xs_prev = xbuf[(step + 2) % 3];
    const int* slot_cur = sp.slot_of + (ev & 1) * P;
    const uint8_t* lab_g = sp.lab + (size_t)par * K * P;
    const double* inc_g = sp.inc + (size_t)par * K * P;

    TRACE(1)
    cp_async_commit_wait_all();   // this step's observation has landed (own copies)
    __syncthreads();              // ... everybody's; last step's row lists are complete
    prefetch_obs(step + 1, (step + 1) % 3);
    if (warp == 0) {  // item offsets of the units: uoff[u] = sum_{v<u} ucount[v] * J_v
      int run = 0;
      for (int ub = 0; ub < nu; ub += 32) {
        const int u = ub + lane;
        int c = 0;
        if (u < nu) c = T.ucount[u] * ((sp.ds[T.uinfo[u] >> 24].J + PMDI_QB - 1) / PMDI_QB);
        int inc = c;
#pragma unroll
        for (int o = 1; o < 32; o <<= 1) {
          const int v = __shfl_up_sync(FULL, inc, o);
          if (lane >= o) inc += v;
        }
        if (u < nu) { T.uoff[u] = run + inc - c; T.remaining[u] = c; }
        run += __shfl_sync(FULL, inc, 31);
      }
      if (lane == 0) {
        T.uoff[nu] = run;
        sm.total_items = run;
        sm.item_ctr = 0;
        if (run > sp.item_cap) atomicExch(sp.err, 78);
      }
    }
    __syncthreads();
    const int total = sm.total_items;
    if (total > sp.item_cap) return;  // every CTA sees err through the barrier watchdog
    for (int it = tid; it < total; it += PMDI_NT) {  // decode table: item -> (unit, row entry, block)
      int lo = 0, hi = nu - 1;  // last u with uoff[u] <= it
      while (lo < hi) {
        const int mid = (lo + hi + 1) >> 1;
        if (T.uoff[mid] <= it) lo = mid; else hi = mid - 1;
      }
      const int JQ = (sp.ds[T.uinfo[lo] >> 24].J + PMDI_QB - 1) / PMDI_QB;
      const int r = it - T.uoff[lo];
      const int e = r / JQ;
      T.items[it] = ((unsigned)lo << 13) | ((unsigned)e << 5) | (unsigned)(r - e * JQ);
    }
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
      TRACE(0x100 | (code & 0xFF) | ((T.uinfo[code >> 13] >> 24) << 12))
      const int u = code >> 13, e = (code >> 5) & 0xFF, qd = code & 31;
      const int k = T.uinfo[u] >> 24, slot = T.uinfo[u] & 0xFFFFFF;
      const DsDev& ds = sp.ds[k];
      const unsigned ent = T.urow[(size_t)u * N + e];
      const int m = ent & 0xFF, n = ent >> 8;
      const long long row = (long long)slot * N + m;
      const int j0 = qd * PMDI_QB, j1 = min(ds.J, j0 + PMDI_QB);
      const int pd = T.pend[u];
      const bool fused = (pd >= 0) && ((pd & 0xFF) == m);  // pending add of the previous step (n == pd >> 8)
      double v;
      if (ds.type == T_GAUSSIAN) {
        if (fused) {
          v = 0.0;
          for (int j = j0; j < j1; ++j)
            v += gauss_fused_block(ds, row, j, n, (const double*)(xs_prev + ds.x_off),
                                   (const double*)(xs_cur + ds.x_off), lane);
        } else {
          v = gauss_eval_item(ds, row, j0, j1, n, (const double*)(xs_cur + ds.x_off), lane);
        }
      } else if (ds.type == T_CATEGORICAL) {
        if (fused) {
          for (int j = j0; j < j1; ++j) cat_add_block(ds, row, j, (const int*)(xs_prev + ds.x_off), lane);
          __syncwarp();
        }
        v = cat_eval_item(ds, row, j0, j1, (const int*)(xs_cur + ds.x_off), lane);
      } else {
        if (fused) {
          for (int j = j0; j < j1; ++j) nb_add_block(ds, row, j, n, (const int*)(xs_prev + ds.x_off), lane, lf, lfT);
          __syncwarp();
        }
        v = nb_eval_item(ds, row, j0, j1, n, (const int*)(xs_cur + ds.x_off), lane, lf, lfT);
      }
      TRACE(3)
      int last = 0;
      if (lane == 0) {
        T.part[it] = v;
        __threadfence_block();
        last = (atomicSub(&T.remaining[u], 1) == 1);
      }
      last = __shfl_sync(FULL, last, 0);
      if (last) {  // this warp finished the unit: run its proposal now
        __threadfence_block();
        TRACE(4)
        propose_unit(sp, T, u, step, ev, sm.rows_eval);
        TRACE(5)
      }
    }
    TRACE(6)
    for (int u = warp; u < nu; u += NW)  // units with no occupied row at all
      if (T.uoff[u + 1] == T.uoff[u]) propose_unit(sp, T, u, step, ev, sm.rows_eval);
    PHASE_MARK(1)
    if (!grid_barrier(sp.bar, epoch, G, sp.err)) return;
    PHASE_MARK(4)
    TRACE(7)

    // ------------------------------------------------------------------ weights + ESS
    double mx = -INFINITY;
    for (int p = tid; p < P; p += PMDI_NT) {
      const int slot = ldcg_i32(slot_cur + p);
      double w = lw[p];
      int labs[PMDI_MAX_K];
#pragma unroll
      for (int k = 0; k < PMDI_MAX_K; ++k)
        if (k < K) {
          w += ldcg_f64(inc_g + (size_t)k * P + slot);  // dataset order, as src/pmdi.jl:210,233
          labs[k] = ldcg_u8(lab_g + (size_t)k * P + slot);
        }
      int idx = 0;
#pragma unroll
      for (int k1 = 0; k1 < PMDI_MAX_K - 1; ++k1)
#pragma unroll
        for (int k2 = k1 + 1; k2 < PMDI_MAX_K; ++k2)
          if (k2 < K) {
            w += (labs[k1] == labs[k2]) ? sp.l1phi[idx] : 0.0;  // Phi_upweight! (misc.jl:50-59)
            ++idx;
          }
      lw[p] = w;
      mx = fmax(mx, w);
      if (cta == 0 && sp.dbg_lw) sp.dbg_lw[(size_t)step * P + p] = w;
    }
    mx = block_max(mx, sm.red);
    double num = 0.0, den = 0.0;
    for (int p = tid; p < P; p += PMDI_NT) {
      const double w = exp(lw[p] - mx);
      num += w;
      den += w * w;
    }
    block_sum2(num, den, sm.red);
    const bool do_res = (num * num) / den <= 0.5 * (double)P;  // src/pmdi.jl:317
    if (!do_res && cta == 0 && tid == 0) sp.ev_of_step[step] = -1;
    PHASE_MARK(5)
    TRACE(8)

    if (do_res) {
      flush_adds(sp, T, nu, xs_cur, lf);
      if (cta == 0) resample_plan(sp, step, ev, mx, lw, s_tmp);
      if (!grid_barrier(sp.bar, epoch, G, sp.err)) return;
      const int ncopy = ldcg_i32(sp.plan_out);
      const int gw = cta * NW + warp, GW = G * NW;
      for (int idx = gw; idx < ncopy * K * N; idx += GW) {
        const int c = idx / (K * N), rem = idx - c * (K * N);
        const int k = rem / N, m = rem - k * N;
        const int2 cp = __ldcg(sp.copies + c);
        row_copy(sp.ds[k], (long long)cp.x * N + m, (long long)cp.y * N + m, lane);
      }
      for (int p = tid; p < P; p += PMDI_NT) lw[p] = 1.0;  // logweight .= 1.0 (src/pmdi.jl:319)
      if (cta == 0 && tid == 0) { sp.counters[0] += 1; sp.counters[1] += ncopy; }
      ++ev;
      if (!grid_barrier(sp.bar, epoch, G, sp.err)) return;
      rebuild_rows();
      PHASE_MARK(6)
    }
  }
  __syncthreads();
  if (tid < K) atomicAdd(sp.rows_eval + tid, (unsigned long long)sm.rows_eval[tid]);
  if (cta == 0)
    for (int p = tid; p < P; p += PMDI_NT) sp.lw_out[p] = lw[p];
  if (timing)
    for (int i = 0; i < 8; ++i) sp.phase_ns[(size_t)cta * 8 + i] = sm.tacc[i];
  if (cta == 0 && tid == 0) sp.counters[2] = ev;
#undef PHASE_MARK
}
