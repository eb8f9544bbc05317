// libpmdi_cuda.so — host side of the C-ABI declared in include/pmdi_cuda.h.
// Owns device memory, builds the x-independent tables, assigns (dataset, particle) units to CTAs
// and launches the kernels.  No CPU fallback: every compute entry point needs a CUDA device.
#include "../../include/pmdi_cuda.h"

#include <cuda_runtime.h>

#include <algorithm>
#include <cmath>
#include <cstdio>
#include <cstring>
#include <cstdlib>
#include <map>
#include <queue>
#include <string>
#include <vector>

#include "aux_kernels.cuh"
#include "spec_kernel.cuh"

namespace {

thread_local std::string g_err;

int fail(int code, const std::string& msg) {
  g_err = msg;
  return code;
}

#define CK(call)                                                                            \
  do {                                                                                      \
    cudaError_t e_ = (call);                                                                \
    if (e_ != cudaSuccess)                                                                  \
      return fail(100 + (int)e_, std::string(#call) + ": " + cudaGetErrorString(e_));       \
  } while (0)

template <class T>
struct DevBuf {
  T* p = nullptr;
  size_t cap = 0;
  bool owned = true;
  void view(void* ptr, size_t n) {  // a slice of the context's shared arena (not freed here)
    if (p && owned) cudaFree(p);
    p = (T*)ptr; cap = n; owned = false;
  }
  cudaError_t ensure(size_t n) {
    if (n <= cap && p) return cudaSuccess;
    if (p && owned) cudaFree(p);
    owned = true;
    p = nullptr;
    cap = 0;
    cudaError_t e = cudaMalloc((void**)&p, std::max<size_t>(n, 1) * sizeof(T));
    if (e == cudaSuccess) cap = n;
    return e;
  }
  void release() {
    if (p && owned) cudaFree(p);
    p = nullptr;
    cap = 0;
    owned = true;
  }
};

struct Dataset {
  bool bound = false;
  int type = -1, D = 0, Dp = 0, J = 0, Lmax = 0, FB = PMDI_FB;
  long long max_arg = 0;               // NegBinom: largest lgamma argument the sweep can form
  std::vector<uint8_t> flag;           // [Dp]
  std::vector<double> nlevels;         // categorical [D]
  bool rc_dirty = true;
  DevBuf<unsigned char> x, xq;
  bool xq_dirty = true;
  DevBuf<uint8_t> d_flag;
  DevBuf<double> rc, d_nlevels;
  DevBuf<double> mu, lamn, sum, beta, part, aux;
  DevBuf<uint32_t> cnt;
  DevBuf<long long> S;
  DevBuf<int> n;
  // pool engine (PoolDev)
  int cap = 0, wpf = 1, fpw = 4;
  DevBuf<unsigned long long> cw;
  DevBuf<int> p_refcnt, p_chosen, p_dst, p_neval, p_live, p_free, p_rowmap, p_ctr;
  DevBuf<double> p_lp;
  DevBuf<RowInfo> p_info;
  DevBuf<int> p_mark;
  void release() {
    x.release(); xq.release(); d_flag.release(); rc.release(); d_nlevels.release();
    mu.release(); lamn.release(); sum.release(); beta.release(); part.release(); aux.release();
    cnt.release(); S.release(); n.release(); cw.release();
    p_refcnt.release(); p_chosen.release(); p_dst.release(); p_neval.release(); p_live.release();
    p_free.release(); p_rowmap.release(); p_ctr.release(); p_lp.release();
    p_info.release(); p_mark.release();
  }
};

}  // namespace

struct pmdi_ctx {
  int K = 0, N = 0, P = 0, device = 0;
  long long n = 0;
  cudaStream_t stream = nullptr;
  bool own_stream = false;
  int n_sm = 0, G = 0, GP = 0;  // CTAs of the sweep kernel; spec engine: the first GP propose, the others evaluate
  std::vector<Dataset> ds;
  bool layout_dirty = true;
  // everything another rank has to reach (statistics, grid counter, ESS partials, log-weights, allocation
  // log) lives in ONE allocation with the same layout on every rank: one IPC handle, peer address =
  // local address + delta
  unsigned char* arena = nullptr;
  size_t arena_bytes = 0;
  int rank = 0, R = 1, Ps = 0;  // this rank, ranks, particle slots held here (P / R)
  int engine = 1;               // 1 pool (copy-on-write rows, pool_kernel.cuh), 0 dense (sweep_kernel.cuh)
  int obs_ring = PMDI_OBS_RING;
  unsigned long long wd_ns = 10000000000ull;
  long long peer_delta[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  void* peer_base[8] = {nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr, nullptr};
  // static layout
  std::vector<int> cta_off, cta_units;
  int max_units = 0, sm_x_bytes = 0, lf_T = 0, lf_want = 0, item_cap = 0, Jmax = 0;
  bool assigned = false;
  size_t dyn_smem = 0;
  std::vector<double> lf_host;
  DevBuf<double> lf_dev;
  DevBuf<int> d_cta_off, d_cta_units;
  // per-sweep inputs
  DevBuf<double> Pi, l1phi, tape_alloc, tape_resamp, tape_shuffle, tape_select;
  DevBuf<long long> s_in, s_out, d_pstar, cluster_n, counters;
  DevBuf<int> order, slot_of, anc_log, ev_of_step, members, mem_off, cur_at, plan_out, err;
  DevBuf<int> sc_j, sc_anc0, sc_a, sc_b, sc_c, sc_d;
  DevBuf<double> lw, lw_out, sc_w, sc_pp, sc_u, inc, lp_empty, ess_part;
  DevBuf<int> logical_of;
  DevBuf<uint8_t> lab, alloc_log;
  DevBuf<int2> copies;
  DevBuf<unsigned> bar;
  DevBuf<unsigned long long> rows_eval, rows_ref, phase_ns, trace;
  DevBuf<long long> label_counts, pair_agree, contingency;
  DevBuf<unsigned> psm;
  DevBuf<long long> psm_s;
  long long psm_samples = 0;
  DevBuf<double> rank_part;
  DevBuf<int4> pull_jobs;
  DevBuf<unsigned char> dec;
  DevBuf<int4> glist, elist;
  DevBuf<int> gcnt;
  DevBuf<unsigned long long> rows_spec, rows_add;
  DevBuf<double> dbg_lp, dbg_lw, scratch_d;
  DevBuf<int> dbg_alloc, dbg_anc, scratch_i, wd_state;
  DevBuf<uint8_t> scratch_u8;
  SweepParams sp;
  bool uploaded = false, ran = false;
  unsigned long long sweep_seq = 1;  // pool / spec engines: counters and step tags run on from sweep to sweep
  DevBuf<unsigned long long> bar_state;
  bool bar_state_init = false;
  unsigned sweep_flags = 0;
  cudaEvent_t ev0 = nullptr, ev1 = nullptr, ev2 = nullptr, ev3 = nullptr;
};

namespace {

int round_up(int v, int m) { return (v + m - 1) / m * m; }

void fill_dsdev(pmdi_ctx* c, int k, DsDev& d) {
  Dataset& s = c->ds[k];
  std::memset(&d, 0, sizeof(d));
  d.type = s.type; d.D = s.D; d.Dp = s.Dp; d.J = s.J; d.Lmax = s.Lmax; d.FB = s.FB;
  int nflag = 0;
  for (int q = 0; q < s.D; ++q) nflag += s.flag[q] ? 1 : 0;
  d.nflag = nflag;
  d.all_on = (nflag == s.D && s.D == s.Dp) ? 1 : 0;
  // padded Gaussian features evaluate to a factor of exactly 1 without the flag test
  if (s.type == T_GAUSSIAN && nflag == s.D) d.all_on = 1;
  d.x = s.x.p; d.flag = s.d_flag.p; d.rc = s.rc.p;
  d.xstage = (s.type != T_GAUSSIAN && nflag != s.D) ? (const void*)s.xq.p : (const void*)s.x.p;
  d.mu = s.mu.p; d.lamn = s.lamn.p; d.sum = s.sum.p; d.beta = s.beta.p;
  d.cnt = s.cnt.p; d.S = s.S.p; d.part = c->engine ? nullptr : s.part.p; d.aux = s.aux.p; d.n = s.n.p;
}

void fill_pooldev(pmdi_ctx* c, int k, PoolDev& d) {
  Dataset& s = c->ds[k];
  std::memset(&d, 0, sizeof(d));
  d.cap = s.cap; d.wpf = s.wpf; d.fpw = s.fpw;
  d.refcnt = s.p_refcnt.p; d.chosen = s.p_chosen.p; d.dst = s.p_dst.p; d.n_eval = s.p_neval.p;
  d.live = s.p_live.p; d.freelist = s.p_free.p; d.rowmap = s.p_rowmap.p; d.ctr = s.p_ctr.p; d.cw = s.cw.p;
  d.lp = s.p_lp.p; d.info = s.p_info.p; d.mark = s.p_mark.p;
}

// x-independent row constants by cluster size n:
//  Gaussian    nflag * (log(1/sqrt(pi)) + lgamma(n/2+1) - lgamma(n/2+1/2))   gaussian_cluster.jl:38-40
//  Categorical -sum_q flag_q log(nlevels_q + n)                              categorical_cluster.jl:30
//  NegBinom    nflag * (lgamma(n+2) - lgamma(n+1)) = nflag log(n+1)          negbinom_cluster.jl:33,37
int build_rc(pmdi_ctx* c, int k) {
  Dataset& s = c->ds[k];
  const long long n = c->n;
  std::vector<double> rc(n + 1);
  int nflag = 0;
  for (int q = 0; q < s.D; ++q) nflag += s.flag[q] ? 1 : 0;
  if (s.type == T_GAUSSIAN) {
    for (long long i = 0; i <= n; ++i) {
      const double nn = (double)i;
      rc[i] = nflag * (std::log(1.0 / std::sqrt(M_PI)) + std::lgamma(0.5 * nn + 1.0) -
                       std::lgamma(0.5 * nn + 0.5));
    }
  } else if (s.type == T_CATEGORICAL) {
    std::map<double, long long> hist;
    for (int q = 0; q < s.D; ++q)
      if (s.flag[q]) hist[s.nlevels[q]] += 1;
    for (long long i = 0; i <= n; ++i) {
      double acc = 0.0;
      for (auto& kv : hist) acc += (double)kv.second * std::log(kv.first + (double)i);
      rc[i] = -acc;
    }
  } else {
    for (long long i = 0; i <= n; ++i) rc[i] = nflag * std::log((double)i + 1.0);
  }
  CK(s.rc.ensure(n + 1));
  CK(cudaMemcpyAsync(s.rc.p, rc.data(), sizeof(double) * (n + 1), cudaMemcpyHostToDevice, c->stream));
  if (s.type != T_GAUSSIAN && nflag != s.D) {  // staging source with the flags folded in
    CK(s.xq.ensure((size_t)n * s.Dp * 4));
    k_mark_x<<<c->n_sm * 4, 256, 0, c->stream>>>((const int*)s.x.p, s.d_flag.p, (int*)s.xq.p, n, s.Dp,
                                                  s.type == T_CATEGORICAL ? 0 : -1);
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(c->stream));
  s.rc_dirty = false;
  return 0;
}

// Static layout (per bound data): staging offsets, log-factorial table.
int build_layout(pmdi_ctx* c) {
  const int K = c->K;
  for (int k = 0; k < K; ++k)
    if (!c->ds[k].bound) return fail(3, "pmdi: dataset " + std::to_string(k) + " is not bound");
  if (c->P % c->R != 0) return fail(1, "pmdi: particles must be a multiple of the number of ranks");
  c->Ps = c->P / c->R;
  // engine: the copy-on-write pool unless PMDI_ENGINE=dense asks for the dense form (every particle owns its
  // N rows)
  // N rows); default: the speculative single-barrier form of the pool on one GPU, the two-barrier pool when
  // the particles are sharded over several GPUs
  const char* eng = getenv("PMDI_ENGINE");
  // (spec deals one row task per evaluation CTA and step: with many datasets a step has more tasks than
  // evaluation CTAs and the two-barrier form, which uses every SM for the evaluations, measured faster - cfg3)
  c->engine = c->K > 4 ? 1 : 2;
  if (eng && std::string(eng) == "dense") c->engine = 0;
  if (eng && std::string(eng) == "pool") c->engine = 1;
  if (eng && std::string(eng) == "spec") c->engine = 2;
  // spec engine: role split of the grid (proposal CTAs own particle slots, evaluation CTAs own rows)
  {
    // proposal CTAs: one (dataset, particle) unit per warp where the SMs allow it; the rest evaluate rows
    const int ge0 = std::max(8, c->n_sm / 8);
    const int want = ((long long)c->Ps * c->K + 14) / 15;
    c->GP = std::max(1, std::min(std::min(c->Ps, want), c->n_sm - ge0));
    const int ge = std::max(1, std::min(c->n_sm - c->GP - 1, 64));
    c->G = c->engine == 2 ? c->GP + ge + 1 : std::min(c->n_sm, c->Ps);  // + the CTA that decides on the resamplings
  }
  int off = 0, Jmax = 1;
  long long nb_max_arg = -1;
  for (int k = 0; k < K; ++k) {
    Dataset& s = c->ds[k];
    // a warp's share of a row: 256 features; the spec engine spreads a row over twice as many warps (latency)
    s.FB = c->engine == 2 ? 128 : PMDI_FB;
    s.J = (s.Dp + s.FB - 1) / s.FB;
    off = round_up(off, 16);
    off += s.Dp * (s.type == T_GAUSSIAN ? 8 : 4);
    Jmax = std::max(Jmax, s.J);
    if (s.type == T_NEGBINOM) nb_max_arg = std::max(nb_max_arg, s.max_arg);
  }
  c->sm_x_bytes = round_up(off, 16);
  c->Jmax = Jmax;
  c->lf_want = 0;
  if (nb_max_arg >= 0) {
    c->lf_want = (int)std::min<long long>(nb_max_arg + 2, 28000);
    if (c->lf_want < 256) c->lf_want = 256;
    c->lf_host.resize(c->lf_want);
    for (int i = 0; i < c->lf_want; ++i) c->lf_host[i] = std::lgamma((double)i + 1.0);
    CK(c->lf_dev.ensure(c->lf_want));
    CK(cudaMemcpy(c->lf_dev.p, c->lf_host.data(), sizeof(double) * c->lf_want, cudaMemcpyHostToDevice));
  }
  // ---- the shared arena: [grid counter | ESS partials | log-weights | allocation log | statistics]
  if (c->P % c->R != 0) return fail(1, "pmdi: particles must be a multiple of the number of ranks");
  c->Ps = c->P / c->R;
  if (c->R > 1 && c->peer_base[c->rank] != nullptr)
    return fail(1, "pmdi: datasets cannot be re-bound after pmdi_ipc_export (the peers hold this layout)");
  if (getenv("PMDI_WATCHDOG_S")) c->wd_ns = (unsigned long long)(atof(getenv("PMDI_WATCHDOG_S")) * 1e9);
  // dense: particle slots, prototypes, the shared empty row.  pool: at most Ps*N live rows, one reservation per
  // chosen row in flight (<= Ps per dataset), the N prefix rows, the empty cluster; with several ranks a
  // resampling pulls the rows of remote ancestors before the dead local rows are freed (another Ps*N at most)
  // spec: a live row holds its child and the two ids handed out for the next step (x4)
  const long long rows = c->engine == 2 ? (c->R > 1 ? 8ll : 4ll) * c->Ps * c->N + c->N + 4096
                       : c->engine ? (long long)(c->R > 1 ? 2 : 1) * c->Ps * c->N + c->Ps + c->N + 2
                                   : (long long)(c->Ps + 2) * c->N;
  if (rows > 0x7fffff00ll) return fail(1, "pmdi: particles x N too large");
  const int Gmax = std::min(c->n_sm, c->Ps);
  size_t off_b = 0;
  auto take = [&](size_t bytes) { const size_t o = off_b; off_b = (off_b + bytes + 255) / 256 * 256; return o; };
  const size_t o_bar = take(64);
  const size_t o_ess = take(sizeof(double) * 6 * (size_t)c->R * Gmax);
  const size_t o_lw = take(sizeof(double) * (size_t)c->P);
  const size_t o_log = take(2 * (size_t)c->n * K * c->P);  // at most n_obs observation steps; two sweeps' worth (below)
  const size_t o_rankp = take(sizeof(double) * 2 * 8 * 4);
  struct Off { size_t mu, lamn, sum, beta, cnt, S, part, aux, n, cw, refcnt, chosen, dst, neval, live, free_, rowmap, ctr, lp, info, mark; };
  std::vector<Off> offs(K);
  for (int k = 0; k < K; ++k) {
    Dataset& s = c->ds[k];
    Off& o = offs[k];
    std::memset(&o, 0, sizeof(o));
    if (s.type == T_GAUSSIAN) {
      o.mu = take(8 * rows * s.Dp); o.lamn = take(8 * rows * s.Dp);
      o.sum = take(8 * rows * s.Dp); o.beta = take(8 * rows * s.Dp);
    } else if (s.type == T_CATEGORICAL) {
      if (c->engine) {
        s.fpw = c->n < 65536 ? 4 : 2;  // count fields per 64-bit word: 16 bits hold any count when n_obs < 65536
        s.wpf = (s.Lmax + s.fpw - 1) / s.fpw;
        o.cw = take(8 * rows * (size_t)s.Dp * s.wpf);
      } else {
        o.cnt = take(4 * rows * (size_t)s.Lmax * s.Dp);
      }
    } else {
      o.S = take(8 * rows * s.Dp);
    }
    o.part = take(c->engine ? 8 : 8 * rows * s.J); o.aux = take(8 * rows * s.J); o.n = take(4 * rows);
    if (c->engine) {
      s.cap = (int)rows;
      o.refcnt = take(8 * rows); o.chosen = take(12 * rows); o.dst = take(8 * rows); o.neval = take(4 * rows);
      o.live = take(4 * rows); o.free_ = take(4 * rows); o.rowmap = take(8 * (size_t)c->Ps * c->N); o.ctr = take(64);
      o.lp = take(8 * rows);
      if (c->engine == 2) {
        o.info = take(2 * sizeof(RowInfo) * rows); o.mark = take(4 * rows);
      }
    }
  }
  if (c->arena) { cudaFree(c->arena); c->arena = nullptr; }
  CK(cudaMalloc((void**)&c->arena, off_b));
  CK(cudaMemset(c->arena, 0, off_b));
  c->arena_bytes = off_b;
  unsigned char* A = c->arena;
  c->bar.view(A + o_bar, 16);
  c->ess_part.view(A + o_ess, 6 * (size_t)c->R * Gmax);
  c->lw.view(A + o_lw, c->P);
  c->alloc_log.view(A + o_log, 2 * (size_t)c->n * K * c->P);
  c->rank_part.view(A + o_rankp, 2 * 8 * 4);
  for (int k = 0; k < K; ++k) {
    Dataset& s = c->ds[k];
    const Off& o = offs[k];
    if (c->engine) {
      if (s.type == T_CATEGORICAL) s.cw.view(A + o.cw, rows * (size_t)s.Dp * s.wpf);
      s.p_refcnt.view(A + o.refcnt, 2 * rows); s.p_chosen.view(A + o.chosen, 3 * rows); s.p_dst.view(A + o.dst, 2 * rows);
      s.p_neval.view(A + o.neval, rows); s.p_live.view(A + o.live, rows); s.p_free.view(A + o.free_, rows);
      s.p_rowmap.view(A + o.rowmap, 2 * (size_t)c->Ps * c->N); s.p_ctr.view(A + o.ctr, 16);
      s.p_lp.view(A + o.lp, rows);
      if (c->engine == 2) {
        s.p_info.view(A + o.info, 2 * rows); s.p_mark.view(A + o.mark, rows);
      }
    }
    if (s.type == T_GAUSSIAN) {
      s.mu.view(A + o.mu, rows * s.Dp); s.lamn.view(A + o.lamn, rows * s.Dp);
      s.sum.view(A + o.sum, rows * s.Dp); s.beta.view(A + o.beta, rows * s.Dp);
    } else if (s.type == T_CATEGORICAL) {
      if (!c->engine) s.cnt.view(A + o.cnt, rows * (size_t)s.Lmax * s.Dp);
    } else {
      s.S.view(A + o.S, rows * s.Dp);
    }
    s.part.view(A + o.part, c->engine ? 1 : rows * s.J); s.aux.view(A + o.aux, rows * s.J); s.n.view(A + o.n, rows);
    DsDev d;
    fill_dsdev(c, k, d);
    if (!(c->engine && s.type == T_CATEGORICAL)) k_init_rows<<<c->n_sm * 4, 256, 0, c->stream>>>(d, rows);
    else {  // the arena is zeroed: packed counts, aux, part and n of every row start empty
    }
    CK(cudaGetLastError());
  }
  CK(cudaStreamSynchronize(c->stream));
  c->layout_dirty = false;
  c->assigned = false;
  return 0;
}

// Particle slots -> CTAs (all K datasets of a slot stay together: the particle's weight update is
// CTA-local), round-robin so that the counts differ by at most one; then the shared-memory
// budget of the sweep kernel.
int assign_units(pmdi_ctx* c) {
  const int K = c->K, P = c->Ps, N = c->N;  // the particle slots this rank holds
  if (c->engine != 2) c->G = std::min(c->n_sm, P);  // one persistent CTA per SM; never a CTA without a particle
  const int G = c->engine == 2 ? c->GP : c->G;      // CTAs that own particle slots
  c->cta_off.assign(G + 1, 0);
  c->cta_units.clear();
  int max_slots = 0;
  for (int g = 0; g < G; ++g) {
    c->cta_off[g] = (int)c->cta_units.size();
    int ns = 0;
    for (int slot = g; slot < P; slot += G, ++ns)
      for (int k = 0; k < K; ++k) c->cta_units.push_back((k << 24) | slot);
    max_slots = std::max(max_slots, ns);
  }
  c->cta_off[G] = (int)c->cta_units.size();
  c->max_units = max_slots * K;
  const int Npad = (N + 31) & ~31;
  int dev_smem = 0;
  CK(cudaDeviceGetAttribute(&dev_smem, cudaDevAttrMaxSharedMemoryPerBlockOptin, c->device));
  const long long MU = c->max_units, MS = max_slots;
  if (c->engine) {
    // pool engine: observation ring (2..4 deep) | lf | proposal scratch | Pi | lw | inc | unit tables
    long long tables = (long long)(PMDI_NT / 32) * Npad * 8 + (long long)K * N * 8 + MS * 8 + MU * 8 +
                       (MU * 11 + MS + MU * N) * 4 + 64;
    if (c->engine == 2) {  // the two roles overlay one region: proposal tables | entry-list and free-id caches
      const long long pt = (long long)(PMDI_NT / 32) * Npad * 12 + (long long)K * N * 8 + MS * 8 + MU * 8 +
                           (MU * 8 + MS + MU * N) * 4 + 64;
      const long long et = 2ll * SPEC_EC * 16 + (long long)K * SPEC_FC * 4 + K * 4 + 64;
      tables = std::max(pt, et);
    }
    const long long budget = (long long)dev_smem - 8192 - tables;  // static shared memory: the parameter block, PoolSmem
    int ring = PMDI_OBS_RING;
    while (ring > 2 && (long long)ring * c->sm_x_bytes + std::min<long long>((long long)c->lf_want * 8, 64 * 1024) > budget) --ring;
    if ((long long)ring * c->sm_x_bytes + (c->lf_want > 0 ? 2048 : 0) > budget)
      return fail(4, "pmdi: one observation over all datasets is " + std::to_string(c->sm_x_bytes) +
                         " bytes; two of them must fit in shared memory (" + std::to_string(budget / 2) +
                         " bytes each with this N and particle count)");
    c->obs_ring = ring;
    const long long lf_b = std::min<long long>((long long)c->lf_want * 8, budget - (long long)ring * c->sm_x_bytes);
    c->lf_T = (int)(lf_b / 8);
    c->item_cap = 0;
    c->dyn_smem = (size_t)((long long)ring * c->sm_x_bytes + (long long)c->lf_T * 8 + 8 + tables);
    CK(cudaFuncSetAttribute(k_sweep_pool, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->dyn_smem));
    CK(cudaFuncSetAttribute(k_sweep_pool_dbg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->dyn_smem));
    CK(cudaFuncSetAttribute(k_sweep_spec, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->dyn_smem));
    CK(cudaFuncSetAttribute(k_sweep_spec_dbg, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->dyn_smem));
  } else {
    // dense engine: 4 observation buffers | lf | proposal scratch | Pi | lw | inc | part x2 | items x2 | unit tables
    const long long fixed = 4LL * c->sm_x_bytes + (long long)(PMDI_NT / 32) * Npad * 8 + (long long)K * N * 8 +
                            MS * 8 + 2 * MU * 8 + MU * N * 4 + (12 * MU + 2 * MS) * 4 + 64;
    const long long avail = (long long)dev_smem - 6144 - fixed;
    const long long want_items = MU * N * c->Jmax;  // every label of every unit occupied
    if (avail < 24LL * MU * c->Jmax * 2)
      return fail(4, "pmdi: datasets too wide / too many particles per SM for the shared-memory work queue of the "
                     "dense engine (4 x " + std::to_string(c->sm_x_bytes) + " bytes of staged observations)");
    long long items_b = std::min(want_items * 24, avail / 2);
    const long long lf_b = std::min<long long>((long long)c->lf_want * 8, avail - items_b);
    items_b = std::min(want_items * 24, avail - lf_b);
    c->item_cap = (int)(items_b / 24);
    c->lf_T = (int)(lf_b / 8);
    if (c->lf_want > 0 && c->lf_T < 256) return fail(4, "pmdi: no shared memory left for the log-factorial table");
    c->dyn_smem = (size_t)fixed + (size_t)c->lf_T * 8 + (size_t)c->item_cap * 24;
    CK(cudaFuncSetAttribute(k_sweep, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)c->dyn_smem));
  }
  if (c->sm_x_bytes > 48 * 1024)
    CK(cudaFuncSetAttribute(k_empty_lp, cudaFuncAttributeMaxDynamicSharedMemorySize, c->sm_x_bytes));
  CK(c->d_cta_off.ensure(c->cta_off.size()));
  CK(c->d_cta_units.ensure(c->cta_units.size()));
  CK(cudaMemcpyAsync(c->d_cta_off.p, c->cta_off.data(), sizeof(int) * c->cta_off.size(), cudaMemcpyHostToDevice, c->stream));
  CK(cudaMemcpyAsync(c->d_cta_units.p, c->cta_units.data(), sizeof(int) * c->cta_units.size(), cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  c->assigned = true;
  return 0;
}

int fill_params(pmdi_ctx* c) {
  SweepParams& sp = c->sp;
  int off = 0;
  for (int k = 0; k < c->K; ++k) {
    fill_dsdev(c, k, sp.ds[k]);
    off = round_up(off, 16);
    sp.ds[k].x_off = off;
    off += c->ds[k].Dp * (c->ds[k].type == T_GAUSSIAN ? 8 : 4);
  }
  for (int k = 0; k < c->K; ++k) fill_pooldev(c, k, sp.pd[k]);
  sp.engine = c->engine; sp.obs_ring = c->obs_ring; sp.wd_ns = c->wd_ns;
  sp.proto_base = c->engine ? 0 : (long long)c->Ps * c->N;
  sp.rank_part = c->rank_part.p;
  sp.K = c->K; sp.N = c->N; sp.P = c->P; sp.n_obs = (int)c->n; sp.G = c->G; sp.GP = c->GP;
  sp.R = c->R; sp.rank = c->rank; sp.Ps = c->Ps; sp.slot0 = c->rank * c->Ps;
  for (int r = 0; r < 8; ++r) sp.peer_delta[r] = c->peer_delta[r];
  sp.Jmax = c->Jmax;
  sp.cta_off = c->d_cta_off.p; sp.cta_units = c->d_cta_units.p;
  sp.max_units = c->max_units; sp.sm_x_bytes = c->sm_x_bytes; sp.lf_T = c->lf_T;
  sp.item_cap = c->item_cap; sp.lf_glob = c->lf_dev.p; sp.lf_glob_T = c->lf_want;
  // 256-feature blocks per plain work item.  With few units per CTA rows have to be split so that the 16
  // warps have work (cfg2, 6 units: 1-2 blocks best, 3 already +3 %); with many units a whole row per item
  // is best (cfg4, 14 units: 8 blocks 9 % faster than 2) - profiles/r01_k_sweep_cfg4.md.  PMDI_QB overrides.
  sp.qb = getenv("PMDI_QB") ? std::max(1, atoi(getenv("PMDI_QB"))) : (c->max_units >= 11 ? std::max(2, c->Jmax) : 2);
  sp.jq = 1;
  while (sp.jq < c->Jmax && sp.jq < PMDI_NT / 32) sp.jq *= 2;
  return 0;
}

}  // namespace

extern "C" {

const char* pmdi_last_error(void) { return g_err.c_str(); }
int pmdi_version(void) { return 100; }

int pmdi_device_count(void) {
  int n = 0;
  if (cudaGetDeviceCount(&n) != cudaSuccess) {
    cudaGetLastError();
    return 0;
  }
  return n;
}

double pmdi_uniform(uint64_t seed, uint32_t iter, uint32_t kind, uint32_t step, uint32_t k,
                    uint32_t index) {
  return pmdi_philox_uniform(seed, iter, kind, step, k, index);
}

int pmdi_ctx_create(pmdi_ctx** out, int32_t K, int64_t n_obs, int32_t N, int32_t particles,
                    int32_t device) {
  if (!out) return fail(1, "pmdi_ctx_create: out is NULL");
  *out = nullptr;
  // pre-conditions of pmdi() (src/pmdi.jl:50-55)
  if (K < 1 || K > PMDI_MAX_K) return fail(1, "pmdi_ctx_create: need 1 <= K <= 8 datasets");
  if (!(N > 1 && (int64_t)N <= n_obs)) return fail(1, "pmdi_ctx_create: need 1 < N <= n_obs");
  if (N > PMDI_MAX_N) return fail(1, "pmdi_ctx_create: N > 256 is not supported");
  if (particles <= 1) return fail(1, "pmdi_ctx_create: need particles > 1");
  if (n_obs > 0x7fffffff / 2) return fail(1, "pmdi_ctx_create: n_obs too large");
  if (pmdi_device_count() <= device || device < 0)
    return fail(2, "pmdi_ctx_create: no CUDA device " + std::to_string(device) +
                       " (libpmdi_cuda has no CPU fallback)");
  CK(cudaSetDevice(device));
  cudaDeviceProp prop;
  CK(cudaGetDeviceProperties(&prop, device));
  if (prop.major < 10)
    return fail(2, "pmdi_ctx_create: built for sm_100a (B200); found compute capability " +
                       std::to_string(prop.major) + "." + std::to_string(prop.minor));
  int coop = 0;
  CK(cudaDeviceGetAttribute(&coop, cudaDevAttrCooperativeLaunch, device));
  if (!coop) return fail(2, "pmdi_ctx_create: device lacks cooperative launch");
  pmdi_ctx* c = new pmdi_ctx();
  c->K = K; c->n = n_obs; c->N = N; c->P = particles; c->device = device;
  c->n_sm = prop.multiProcessorCount;
  c->G = c->n_sm;  // one persistent CTA per SM
  c->ds.resize(K);
  std::memset(&c->sp, 0, sizeof(c->sp));
  cudaError_t e = cudaStreamCreateWithFlags(&c->stream, cudaStreamNonBlocking);
  if (e != cudaSuccess) { delete c; return fail(100 + (int)e, "cudaStreamCreate failed"); }
  c->own_stream = true;
  cudaEventCreate(&c->ev0); cudaEventCreate(&c->ev1); cudaEventCreate(&c->ev2); cudaEventCreate(&c->ev3);
  *out = c;
  return 0;
}

int pmdi_ctx_destroy(pmdi_ctx* c) {
  if (!c) return 0;
  cudaSetDevice(c->device);
  cudaStreamSynchronize(c->stream);
  for (auto& d : c->ds) d.release();
  DevBuf<double>* dd[] = {&c->lf_dev, &c->Pi, &c->l1phi, &c->tape_alloc, &c->tape_resamp, &c->tape_shuffle,
                          &c->tape_select, &c->lw, &c->lw_out, &c->sc_w, &c->sc_pp, &c->sc_u, &c->dbg_lp,
                          &c->dbg_lw, &c->scratch_d, &c->inc, &c->lp_empty, &c->ess_part};
  for (auto* b : dd) b->release();
  DevBuf<int>* di[] = {&c->d_cta_off, &c->d_cta_units, &c->order, &c->slot_of, &c->anc_log, &c->ev_of_step,
                       &c->members, &c->mem_off, &c->cur_at, &c->plan_out, &c->err, &c->sc_j, &c->sc_anc0,
                       &c->sc_a, &c->sc_b, &c->sc_c, &c->sc_d, &c->dbg_alloc, &c->dbg_anc, &c->scratch_i, &c->logical_of};
  for (auto* b : di) b->release();
  DevBuf<long long>* dl[] = {&c->s_in, &c->s_out, &c->d_pstar, &c->cluster_n, &c->counters};
  for (auto* b : dl) b->release();
  c->lab.release(); c->alloc_log.release(); c->copies.release(); c->bar.release();
  c->rows_eval.release(); c->phase_ns.release(); c->scratch_u8.release();
  c->dec.release(); c->rows_spec.release(); c->rows_add.release(); c->bar_state.release(); c->glist.release(); c->elist.release(); c->gcnt.release();
  c->rows_ref.release(); c->trace.release(); c->wd_state.release(); c->pull_jobs.release(); c->rank_part.release();
  c->label_counts.release(); c->pair_agree.release(); c->contingency.release(); c->psm.release(); c->psm_s.release();
  for (int r = 0; r < c->R; ++r)
    if (r != c->rank && c->peer_base[r]) cudaIpcCloseMemHandle(c->peer_base[r]);
  if (c->arena) cudaFree(c->arena);
  if (c->ev0) { cudaEventDestroy(c->ev0); cudaEventDestroy(c->ev1); cudaEventDestroy(c->ev2); cudaEventDestroy(c->ev3); }
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  delete c;
  return 0;
}

int pmdi_ctx_set_stream(pmdi_ctx* c, void* s) {
  if (!c) return fail(1, "NULL context");
  if (c->own_stream && c->stream) cudaStreamDestroy(c->stream);
  c->stream = (cudaStream_t)s;
  c->own_stream = false;
  return 0;
}
void* pmdi_ctx_get_stream(pmdi_ctx* c) { return c ? (void*)c->stream : nullptr; }

int pmdi_set_dataset(pmdi_ctx* c, int32_t k, int32_t type_tag, int32_t elem_kind, const void* data,
                     int64_t n_obs, int64_t D, int64_t ld) {
  if (!c) return fail(1, "NULL context");
  if (k < 0 || k >= c->K) return fail(1, "pmdi_set_dataset: k out of range");
  if (n_obs != c->n) return fail(1, "pmdi_set_dataset: datasets must have n_obs rows (src/pmdi.jl:52)");
  if (D < 1 || D > 8192) return fail(1, "pmdi_set_dataset: need 1 <= D <= 8192 features per dataset");
  if (ld < n_obs) return fail(1, "pmdi_set_dataset: ld < n_obs");
  if (type_tag < 0 || type_tag > 2) return fail(1, "pmdi_set_dataset: unknown cluster type tag");
  if ((type_tag == PMDI_GAUSSIAN) != (elem_kind == PMDI_F64))
    return fail(1, "pmdi_set_dataset: Gaussian needs PMDI_F64 data, Categorical/NegBinom need PMDI_I64");
  CK(cudaSetDevice(c->device));
  Dataset& s = c->ds[k];
  s.release();
  s.type = type_tag; s.D = (int)D; s.Dp = round_up((int)D, PMDI_WF);
  s.J = (s.Dp + PMDI_FB - 1) / PMDI_FB;
  s.Lmax = 0; s.max_arg = 0;
  const long long n = c->n;
  const int Dp = s.Dp;
  s.flag.assign(Dp, 0);
  for (int q = 0; q < s.D; ++q) s.flag[q] = 1;
  if (type_tag == PMDI_GAUSSIAN) {
    std::vector<double> h((size_t)n * Dp, 0.0);
    const double* x = (const double*)data;
    for (int q = 0; q < s.D; ++q)
      for (long long i = 0; i < n; ++i) h[(size_t)i * Dp + q] = x[(size_t)i + (size_t)ld * q];
    CK(s.x.ensure(h.size() * 8));
    CK(cudaMemcpy(s.x.p, h.data(), h.size() * 8, cudaMemcpyHostToDevice));
  } else {
    std::vector<int> h((size_t)n * Dp, type_tag == PMDI_NEGBINOM ? -1 : 0);
    const int64_t* x = (const int64_t*)data;
    s.nlevels.assign(s.D, 0.0);
    long long gmax = 0, max_colsum = 0;
    for (int q = 0; q < s.D; ++q) {
      long long cmax = x[(size_t)ld * q], csum = 0;
      for (long long i = 0; i < n; ++i) {
        const int64_t v = x[(size_t)i + (size_t)ld * q];
        if (type_tag == PMDI_CATEGORICAL && v < 1)
          return fail(1, "pmdi_set_dataset: categorical levels must be >= 1 (use coerce_categorical)");
        if (type_tag == PMDI_NEGBINOM && v < 0) return fail(1, "pmdi_set_dataset: counts must be >= 0");
        if (v > 0x3fffffff) return fail(1, "pmdi_set_dataset: value too large");
        h[(size_t)i * Dp + q] = (int)v;
        cmax = std::max<long long>(cmax, v);
        csum += v;
      }
      s.nlevels[q] = 0.5 * (double)cmax;  // categorical_cluster.jl:10
      gmax = std::max(gmax, cmax);
      max_colsum = std::max(max_colsum, csum);
    }
    if (type_tag == PMDI_CATEGORICAL) {
      if (gmax > 4096) return fail(1, "pmdi_set_dataset: more than 4096 categorical levels");
      s.Lmax = (int)gmax;  // categorical_cluster.jl:8
      CK(s.d_nlevels.ensure(s.D));
      CK(cudaMemcpy(s.d_nlevels.p, s.nlevels.data(), sizeof(double) * s.D, cudaMemcpyHostToDevice));
    } else {
      s.max_arg = max_colsum + n + 3;
    }
    CK(s.x.ensure(h.size() * 4));
    CK(cudaMemcpy(s.x.p, h.data(), h.size() * 4, cudaMemcpyHostToDevice));
  }
  CK(s.d_flag.ensure(Dp));
  CK(cudaMemcpy(s.d_flag.p, s.flag.data(), Dp, cudaMemcpyHostToDevice));
  s.bound = true;
  s.rc_dirty = true;
  c->layout_dirty = true;  // the statistics are (re)allocated in the shared arena by build_layout()
  return 0;
}

int pmdi_ctx_set_ranks(pmdi_ctx* c, int32_t rank, int32_t n_ranks) {
  if (!c) return fail(1, "NULL context");
  if (n_ranks < 1 || n_ranks > 8 || rank < 0 || rank >= n_ranks) return fail(1, "pmdi_ctx_set_ranks: need 0 <= rank < n_ranks <= 8");
  if (c->P % n_ranks != 0) return fail(1, "pmdi_ctx_set_ranks: particles must be a multiple of the number of ranks");
  if (c->P / n_ranks < 1) return fail(1, "pmdi_ctx_set_ranks: fewer particles than ranks");
  if (c->arena) return fail(1, "pmdi_ctx_set_ranks: call it before the first sweep / export");
  c->rank = rank; c->R = n_ranks; c->Ps = c->P / n_ranks;
  c->layout_dirty = true;
  return 0;
}

static int prepare(pmdi_ctx* c);

int pmdi_ipc_export(pmdi_ctx* c, void* handle_out, int64_t* arena_bytes) {
  if (!c || !handle_out) return fail(1, "pmdi_ipc_export: NULL argument");
  int rc = prepare(c);  // builds the shared arena (all datasets must be bound)
  if (rc) return rc;
  static_assert(sizeof(cudaIpcMemHandle_t) == PMDI_IPC_HANDLE_BYTES, "IPC handle size");
  cudaIpcMemHandle_t h;
  CK(cudaIpcGetMemHandle(&h, c->arena));
  std::memcpy(handle_out, &h, sizeof(h));
  if (arena_bytes) *arena_bytes = (int64_t)c->arena_bytes;
  return 0;
}

int pmdi_ipc_import(pmdi_ctx* c, const void* handles, const int64_t* arena_bytes) {
  if (!c || !handles) return fail(1, "pmdi_ipc_import: NULL argument");
  if (!c->arena) return fail(1, "pmdi_ipc_import: call pmdi_ipc_export first");
  CK(cudaSetDevice(c->device));
  for (int r = 0; r < c->R; ++r) {
    if (arena_bytes && arena_bytes[r] != (int64_t)c->arena_bytes)
      return fail(1, "pmdi_ipc_import: rank " + std::to_string(r) + " has a different arena layout (same datasets, N, particles on every rank?)");
    if (r == c->rank) { c->peer_base[r] = c->arena; c->peer_delta[r] = 0; continue; }
    cudaIpcMemHandle_t h;
    std::memcpy(&h, (const unsigned char*)handles + (size_t)r * sizeof(h), sizeof(h));
    void* p = nullptr;
    CK(cudaIpcOpenMemHandle(&p, h, cudaIpcMemLazyEnablePeerAccess));
    c->peer_base[r] = p;
    c->peer_delta[r] = (long long)((unsigned char*)p - c->arena);
  }
  return fill_params(c);
}

int pmdi_set_feature_flags(pmdi_ctx* c, int32_t k, const uint8_t* flags) {
  if (!c) return fail(1, "NULL context");
  if (k < 0 || k >= c->K || !c->ds[k].bound) return fail(1, "pmdi_set_feature_flags: dataset not bound");
  CK(cudaSetDevice(c->device));
  Dataset& s = c->ds[k];
  for (int q = 0; q < s.D; ++q) s.flag[q] = flags[q] ? 1 : 0;
  CK(cudaMemcpyAsync(s.d_flag.p, s.flag.data(), s.Dp, cudaMemcpyHostToDevice, c->stream));
  CK(cudaStreamSynchronize(c->stream));
  s.rc_dirty = true;
  return 0;
}

static int prepare(pmdi_ctx* c) {
  CK(cudaSetDevice(c->device));
  if (c->Ps == 0) c->Ps = c->P / c->R;
  if (c->layout_dirty) {
    int rc = build_layout(c);
    if (rc) return rc;
  }
  for (int k = 0; k < c->K; ++k)
    if (c->ds[k].rc_dirty) {
      int rc = build_rc(c, k);
      if (rc) return rc;
    }
  return fill_params(c);
}

int pmdi_sweep_upload(pmdi_ctx* c, const pmdi_sweep_args* a) {
  if (!c || !a) return fail(1, "pmdi_sweep_upload: NULL argument");
  int rc = prepare(c);
  if (rc) return rc;
  const int K = c->K, N = c->N, P = c->P;
  const long long n = c->n;
  if (a->n1 < 1 || a->n1 > n) return fail(1, "pmdi_sweep: need 1 <= n1 <= n_obs (n1 = floor(rho*n_obs), src/pmdi.jl:161)");
  if (!a->s || !a->order_obs || !a->Pi) return fail(1, "pmdi_sweep: s, order_obs and Pi are required");
  if (K > 1 && !a->phi) return fail(1, "pmdi_sweep: phi is required when K > 1");
  const int steps = (int)(n - a->n1 + 1);
  SweepParams& sp = c->sp;
  if (!c->assigned) {
    rc = assign_units(c);
    if (rc) return rc;
    fill_params(c);
  }
  // validate + convert
  std::vector<int> order(n);
  std::vector<uint8_t> seen(n, 0);
  for (long long i = 0; i < n; ++i) {
    const int64_t o = a->order_obs[i];
    if (o < 1 || o > n || seen[o - 1]) return fail(1, "pmdi_sweep: order_obs is not a permutation of 1..n_obs");
    seen[o - 1] = 1;
    order[i] = (int)(o - 1);
  }
  for (long long i = 0; i < n * K; ++i)
    if (a->s[i] < 1 || a->s[i] > N) return fail(1, "pmdi_sweep: allocation label outside 1..N");
  const int npairs = K * (K - 1) / 2;
  std::vector<double> l1phi(std::max(npairs, 1), 0.0);
  for (int i = 0; i < npairs; ++i) l1phi[i] = std::log(1 + a->phi[i]);  // src/misc.jl:53
  cudaStream_t st = c->stream;
  CK(c->s_in.ensure(n * K)); CK(c->s_out.ensure(n * K)); CK(c->order.ensure(n));
  CK(c->Pi.ensure((size_t)N * K)); CK(c->l1phi.ensure(l1phi.size()));
  CK(cudaMemcpyAsync(c->s_in.p, a->s, sizeof(int64_t) * n * K, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->order.p, order.data(), sizeof(int) * n, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->Pi.p, a->Pi, sizeof(double) * N * K, cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(c->l1phi.p, l1phi.data(), sizeof(double) * l1phi.size(), cudaMemcpyHostToDevice, st));
  sp.tape_alloc = sp.tape_resamp = sp.tape_shuffle = sp.tape_select = nullptr;
  if (a->tape_alloc) {
    CK(c->tape_alloc.ensure((size_t)steps * K * P));
    CK(cudaMemcpyAsync(c->tape_alloc.p, a->tape_alloc, sizeof(double) * steps * K * P, cudaMemcpyHostToDevice, st));
    sp.tape_alloc = c->tape_alloc.p;
  }
  if (a->tape_resamp) {
    CK(c->tape_resamp.ensure(steps));
    CK(cudaMemcpyAsync(c->tape_resamp.p, a->tape_resamp, sizeof(double) * steps, cudaMemcpyHostToDevice, st));
    sp.tape_resamp = c->tape_resamp.p;
  }
  if (a->tape_shuffle) {
    CK(c->tape_shuffle.ensure((size_t)steps * P));
    CK(cudaMemcpyAsync(c->tape_shuffle.p, a->tape_shuffle, sizeof(double) * steps * P, cudaMemcpyHostToDevice, st));
    sp.tape_shuffle = c->tape_shuffle.p;
  }
  if (a->tape_select) {
    CK(c->tape_select.ensure(1));
    CK(cudaMemcpyAsync(c->tape_select.p, a->tape_select, sizeof(double), cudaMemcpyHostToDevice, st));
    sp.tape_select = c->tape_select.p;
  }
  // state buffers
  CK(c->lw_out.ensure(P)); CK(c->slot_of.ensure(2 * (size_t)P));
  CK(c->logical_of.ensure(2 * (size_t)P)); CK(c->inc.ensure(2 * (size_t)K * P)); CK(c->lp_empty.ensure((size_t)steps * K));
  CK(c->lab.ensure(2 * (size_t)K * P));
  CK(c->anc_log.ensure((size_t)steps * P)); CK(c->ev_of_step.ensure(steps));
  CK(c->sc_w.ensure(P)); CK(c->sc_pp.ensure(P)); CK(c->sc_u.ensure(P));
  CK(c->sc_j.ensure(P)); CK(c->sc_anc0.ensure(P)); CK(c->sc_a.ensure(P)); CK(c->sc_b.ensure(P));
  CK(c->sc_c.ensure(P)); CK(c->sc_d.ensure(P)); CK(c->copies.ensure(P)); CK(c->plan_out.ensure(4));
  CK(c->err.ensure(4)); CK(c->rows_eval.ensure(PMDI_MAX_K)); CK(c->counters.ensure(8));
  CK(c->phase_ns.ensure(8 * (size_t)c->G)); CK(c->d_pstar.ensure(1)); CK(c->cluster_n.ensure((size_t)K * P * N));
  CK(c->members.ensure((size_t)K * std::max<long long>(a->n1 - 1, 1))); CK(c->mem_off.ensure((size_t)K * (N + 1)));
  CK(c->cur_at.ensure(steps));
  CK(c->rows_ref.ensure(PMDI_MAX_K)); CK(c->label_counts.ensure((size_t)N * K)); CK(c->pair_agree.ensure(std::max(npairs, 1)));
  CK(c->contingency.ensure((size_t)std::max(npairs, 1) * N * N));
  sp.rows_ref = c->rows_ref.p;
  CK(c->dec.ensure(steps)); CK(c->rows_spec.ensure(PMDI_MAX_K)); CK(c->rows_add.ensure(PMDI_MAX_K));
  CK(cudaMemsetAsync(c->rows_spec.p, 0, 8 * PMDI_MAX_K, st));
  CK(cudaMemsetAsync(c->rows_add.p, 0, 8 * PMDI_MAX_K, st));
  sp.dec = c->dec.p; sp.rows_spec = c->rows_spec.p; sp.rows_add = c->rows_add.p;
  if (c->engine == 2) {  // the live-row list: at most one entry per (particle slot, label) and dataset
    const long long lcap = (long long)K * ((long long)c->Ps * N + N + 2);
    CK(c->glist.ensure(lcap)); CK(c->gcnt.ensure(4));
    CK(c->elist.ensure(2 * (size_t)(c->G - c->GP - 1) * lcap));
    CK(cudaMemsetAsync(c->gcnt.p, 0, 16, st));
    sp.glist = c->glist.p; sp.gcnt = c->gcnt.p; sp.elist = c->elist.p; sp.lcap = lcap;
    // cached free ids per E-CTA and dataset: never more than a quarter of the pool over all E-CTAs
    const long long cap = c->ds[0].cap, ge = c->G - c->GP - 1;
    sp.fc_target = (int)std::max<long long>(8, std::min<long long>(SPEC_FC / 2, cap / (4 * ge)));
  }
  sp.pull_jobs = nullptr;
  if (c->engine && c->R > 1) {
    CK(c->pull_jobs.ensure(2 * (size_t)K * c->Ps * N));
    sp.pull_jobs = c->pull_jobs.p;
  }
  sp.n1 = (int)a->n1; sp.steps = steps; sp.flags = (int)a->flags;
  sp.Pi = c->Pi.p; sp.l1phi = c->l1phi.p; sp.s_in = c->s_in.p; sp.order = c->order.p;
  sp.lw_init = a->logweight_init; sp.seed = a->seed; sp.iter = a->iter;
  sp.lw = c->lw.p; sp.ess_part = c->ess_part.p; sp.lw_out = c->lw_out.p; sp.slot_of = c->slot_of.p; sp.logical_of = c->logical_of.p;
  sp.inc = c->inc.p; sp.lp_empty = c->lp_empty.p; sp.lab = c->lab.p;
  // a peer that is one sweep ahead writes its allocations into the other half while this rank's finish kernel reads
  sp.alloc_log = c->alloc_log.p + (c->sweep_seq & 1) * (size_t)n * K * P;
  sp.tag_base = c->sweep_seq << 32;
  CK(c->bar_state.ensure(2));
  if (!c->bar_state_init) { CK(cudaMemsetAsync(c->bar_state.p, 0, 16, st)); c->bar_state_init = true; }
  sp.bar_state = c->bar_state.p;
  sp.anc_log = c->anc_log.p; sp.ev_of_step = c->ev_of_step.p;
  sp.sc_w = c->sc_w.p; sp.sc_pp = c->sc_pp.p; sp.sc_u = c->sc_u.p; sp.sc_j = c->sc_j.p;
  sp.sc_anc0 = c->sc_anc0.p; sp.sc_a = c->sc_a.p; sp.sc_b = c->sc_b.p; sp.sc_c = c->sc_c.p;
  sp.sc_d = c->sc_d.p; sp.copies = c->copies.p; sp.plan_out = c->plan_out.p;
  sp.bar = c->bar.p; sp.err = c->err.p; sp.rows_eval = c->rows_eval.p; sp.counters = c->counters.p;
  sp.phase_ns = (a->flags & PMDI_SWEEP_TIME_PHASES) ? c->phase_ns.p : nullptr;
  sp.dbg_lp = nullptr; sp.dbg_lw = nullptr; sp.dbg_alloc = nullptr; sp.dbg_anc = nullptr;
  if (a->flags & PMDI_SWEEP_DEBUG) {
    CK(c->dbg_lp.ensure((size_t)steps * K * P * N)); CK(c->dbg_lw.ensure((size_t)steps * P));
    CK(c->dbg_alloc.ensure((size_t)steps * K * P)); CK(c->dbg_anc.ensure((size_t)steps * P));
    CK(cudaMemsetAsync(c->dbg_anc.p, 0, sizeof(int) * (size_t)steps * P, st));
    // with several ranks each entry of these is written by the rank that holds the particle: zero the rest
    CK(cudaMemsetAsync(c->dbg_lp.p, 0, sizeof(double) * (size_t)steps * K * P * N, st));
    CK(cudaMemsetAsync(c->dbg_lw.p, 0, sizeof(double) * (size_t)steps * P, st));
    CK(cudaMemsetAsync(c->dbg_alloc.p, 0, sizeof(int) * (size_t)steps * K * P, st));
    sp.dbg_lp = c->dbg_lp.p; sp.dbg_lw = c->dbg_lw.p; sp.dbg_alloc = c->dbg_alloc.p; sp.dbg_anc = c->dbg_anc.p;
  }
  if (c->R > 1 && c->peer_base[c->rank] == nullptr)
    return fail(1, "pmdi_sweep: call pmdi_ipc_export / pmdi_ipc_import on every rank first");
  if (!c->engine) {  // dense engine: counters restart; with R > 1 the caller barriers all ranks between upload and run
    CK(cudaMemsetAsync(c->bar.p, 0, 16, st));
  }
  CK(cudaMemsetAsync(c->phase_ns.p, 0, 64 * (size_t)c->G, st));
  CK(c->wd_state.ensure((size_t)c->G * 16 * 16));
  CK(cudaMemsetAsync(c->wd_state.p, 0xff, sizeof(int) * (size_t)c->G * 16 * 16, st));
  sp.wd_state = c->wd_state.p;
  sp.trace = nullptr;
  if (getenv("PMDI_TRACE_STEP")) {
    CK(c->trace.ensure(16 * 128));
    CK(cudaMemsetAsync(c->trace.p, 0, 16 * 128 * 8, st));
    sp.trace = c->trace.p;
    sp.trace_step = atoi(getenv("PMDI_TRACE_STEP"));
    sp.trace_cta = getenv("PMDI_TRACE_CTA") ? atoi(getenv("PMDI_TRACE_CTA")) : 0;
  }
  c->sweep_flags = a->flags;
  c->uploaded = true;
  c->ran = false;
  // pageable host sources: make the call safe to return from
  CK(cudaStreamSynchronize(st));
  return 0;
}

int pmdi_sweep_run(pmdi_ctx* c) {
  if (!c || !c->uploaded) return fail(1, "pmdi_sweep_run: call pmdi_sweep_upload first");
  CK(cudaSetDevice(c->device));
  const SweepParams& sp = c->sp;
  cudaStream_t st = c->stream;
  const int K = c->K, N = c->N, P = c->P;
  CK(cudaEventRecord(c->ev0, st));
  k_sweep_init<<<(P + 255) / 256, 256, 0, st>>>(sp);
  if (sp.n1 > 1) {
    k_prefix_lists<<<K, PMDI_MAX_N + 32, 0, st>>>(sp, c->members.p, c->mem_off.p);
  } else {
    CK(cudaMemsetAsync(c->mem_off.p, 0, sizeof(int) * K * (N + 1), st));
  }
  int maxDp = 0;
  for (int k = 0; k < K; ++k) maxDp = std::max(maxDp, c->ds[k].Dp);
  k_prefix_build<<<dim3((maxDp + 127) / 128, N, K), 128, 0, st>>>(sp, c->members.p, c->mem_off.p);
  k_proto_aux<<<dim3(N, K), 256, 0, st>>>(sp);
  void* args[] = {(void*)&c->sp};
  if (c->engine == 2) {
    k_spec_init<<<K, 1024, 0, st>>>(sp);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, st));
    const bool dbg = (c->sweep_flags & (PMDI_SWEEP_DEBUG | PMDI_SWEEP_TIME_PHASES)) || sp.trace;
    CK(cudaLaunchCooperativeKernel(dbg ? (const void*)k_sweep_spec_dbg : (const void*)k_sweep_spec, dim3(c->G), dim3(PMDI_NT), args,
                                   c->dyn_smem, st));
  } else if (c->engine) {
    k_pool_init<<<K, 1024, 0, st>>>(sp);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, st));
    const bool dbg = (c->sweep_flags & (PMDI_SWEEP_DEBUG | PMDI_SWEEP_TIME_PHASES)) || sp.trace;
    CK(cudaLaunchCooperativeKernel(dbg ? (const void*)k_sweep_pool_dbg : (const void*)k_sweep_pool, dim3(c->G), dim3(PMDI_NT), args,
                                   c->dyn_smem, st));
  } else {
    k_broadcast<<<c->n_sm * 8, 256, 0, st>>>(sp);
    k_empty_lp<<<sp.steps, 256, c->sm_x_bytes, st>>>(sp, c->lp_empty.p);
    CK(cudaGetLastError());
    CK(cudaEventRecord(c->ev1, st));
    CK(cudaLaunchCooperativeKernel((const void*)k_sweep, dim3(c->G), dim3(PMDI_NT), args, c->dyn_smem, st));
  }
  CK(cudaEventRecord(c->ev2, st));
  k_finish_pool<<<1, 256, 0, st>>>(sp, (c->sweep_flags & PMDI_SWEEP_SSTAR_COMPAT) ? 1 : 0, c->s_out.p, c->d_pstar.p,
                                   c->cluster_n.p, c->cur_at.p, c->label_counts.p, c->pair_agree.p, c->contingency.p);
  CK(cudaGetLastError());
  CK(cudaEventRecord(c->ev3, st));
  c->sweep_seq += 1;
  c->ran = true;
  return 0;
}

int pmdi_sweep_download(pmdi_ctx* c, pmdi_sweep_out* o) {
  if (!c || !o || !c->ran) return fail(1, "pmdi_sweep_download: call pmdi_sweep_run first");
  CK(cudaSetDevice(c->device));
  cudaStream_t st = c->stream;
  const int K = c->K, N = c->N, P = c->P;
  const long long n = c->n;
  const int steps = c->sp.steps;
  int err = 0;
  long long counters[8] = {0, 0, 0, 0, 0, 0, 0, 0};
  unsigned long long rows[PMDI_MAX_K], rref[PMDI_MAX_K], rspec[PMDI_MAX_K], radd[PMDI_MAX_K];
  std::vector<unsigned long long> phase(8 * (size_t)c->G, 0ull);
  long long pstar = 0;
  CK(cudaMemcpyAsync(&err, c->err.p, sizeof(int), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(counters, c->counters.p, sizeof(counters), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(rows, c->rows_eval.p, sizeof(rows), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(rref, c->rows_ref.p, sizeof(rref), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(rspec, c->rows_spec.p, sizeof(rspec), cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(radd, c->rows_add.p, sizeof(radd), cudaMemcpyDeviceToHost, st));
  if (o->label_counts)
    CK(cudaMemcpyAsync(o->label_counts, c->label_counts.p, sizeof(int64_t) * N * K, cudaMemcpyDeviceToHost, st));
  if (o->pair_agree && K > 1)
    CK(cudaMemcpyAsync(o->pair_agree, c->pair_agree.p, sizeof(int64_t) * (K * (K - 1) / 2), cudaMemcpyDeviceToHost, st));
  if (o->contingency && K > 1)
    CK(cudaMemcpyAsync(o->contingency, c->contingency.p, sizeof(int64_t) * (K * (K - 1) / 2) * N * N, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(phase.data(), c->phase_ns.p, 8 * sizeof(unsigned long long) * c->G, cudaMemcpyDeviceToHost, st));
  CK(cudaMemcpyAsync(&pstar, c->d_pstar.p, sizeof(long long), cudaMemcpyDeviceToHost, st));
  if (o->s) CK(cudaMemcpyAsync(o->s, c->s_out.p, sizeof(int64_t) * n * K, cudaMemcpyDeviceToHost, st));
  if (o->logweight) CK(cudaMemcpyAsync(o->logweight, c->lw_out.p, sizeof(double) * P, cudaMemcpyDeviceToHost, st));
  if (o->cluster_n)
    CK(cudaMemcpyAsync(o->cluster_n, c->cluster_n.p, sizeof(int64_t) * K * P * N, cudaMemcpyDeviceToHost, st));
  if (c->sweep_flags & PMDI_SWEEP_DEBUG) {
    if (o->dbg_lp) CK(cudaMemcpyAsync(o->dbg_lp, c->dbg_lp.p, sizeof(double) * (size_t)steps * K * P * N, cudaMemcpyDeviceToHost, st));
    if (o->dbg_lw) CK(cudaMemcpyAsync(o->dbg_lw, c->dbg_lw.p, sizeof(double) * (size_t)steps * P, cudaMemcpyDeviceToHost, st));
    if (o->dbg_alloc) CK(cudaMemcpyAsync(o->dbg_alloc, c->dbg_alloc.p, sizeof(int) * (size_t)steps * K * P, cudaMemcpyDeviceToHost, st));
    if (o->dbg_anc) CK(cudaMemcpyAsync(o->dbg_anc, c->dbg_anc.p, sizeof(int) * (size_t)steps * P, cudaMemcpyDeviceToHost, st));
  }
  CK(cudaStreamSynchronize(st));
  if (c->sp.trace && getenv("PMDI_TRACE_FILE")) {
    std::vector<unsigned long long> tr(16 * 128);
    CK(cudaMemcpy(tr.data(), c->trace.p, tr.size() * 8, cudaMemcpyDeviceToHost));
    FILE* f = fopen(getenv("PMDI_TRACE_FILE"), "w");
    if (f) {
      for (int w = 0; w < 16; ++w)
        for (unsigned long long i = 1; i <= tr[w * 128] && i < 128; ++i)
          fprintf(f, "%d %llu %llu\n", w, tr[w * 128 + i] >> 48, tr[w * 128 + i] & 0xFFFFFFFFFFFFull);
      fclose(f);
    }
  }
  if (err == 79 && getenv("PMDI_WD_DUMP")) {
    std::vector<int> w((size_t)c->G * 16 * 16);
    CK(cudaMemcpy(w.data(), c->wd_state.p, w.size() * 4, cudaMemcpyDeviceToHost));
    for (int g = 0; g < c->G; ++g)
      for (int wp = 0; wp < 16; ++wp) {
        const int* v = &w[((size_t)g * 16 + wp) * 16];
        if (v[0] == -1 && v[11] == -1) continue;
        fprintf(stderr, "cta %d warp %d: t=%d h=%d gen=%d head=%d tail=%d units_in=%d left=%d res_step=%d res_flag=%d "
                        "arrived=%d claim=%d nu=%d parked=%d gen'=%d units_in'=%d pdone=%d\n", g, wp, v[0], v[1], v[2],
                v[3], v[4], v[5], v[6], v[7], v[8], v[9], v[10], v[11], v[12], v[13], v[14], v[15]);
      }
  }
  if (err != 0)
    return fail(50 + err, err == 77 ? "pmdi_sweep: grid barrier watchdog fired (a CTA did not arrive)"
                          : err == 79 ? "pmdi_sweep: work-queue watchdog fired (a warp waited 4 s for work)"
                          : err == 78 ? "pmdi_sweep: shared-memory work queue overflow (too many occupied clusters per SM)"
                          : err == 80 ? "pmdi_sweep: cluster pool exhausted"
                                    : "pmdi_sweep: device-side error " + std::to_string(err));
  if (o->p_star) *o->p_star = pstar;
  o->n_resamples = counters[0];
  o->n_copies = counters[1];
  o->n_remote_rows = counters[3];
  o->rows_evaluated_ahead = counters[4];
  long long ev = 0, dense = 0;
  for (int k = 0; k < K; ++k) {
    if (!c->engine) {
      rref[k] = rows[k];                     // dense engine: every referenced row is evaluated
      rows[k] += (unsigned long long)steps;  // the shared empty cluster: one evaluation per step
    }
    ev += (long long)rows[k] * c->ds[k].D;
    dense += (long long)steps * P * N * c->ds[k].D;
    o->rows_evaluated[k] = (int64_t)rows[k];
    o->rows_referenced[k] = (int64_t)rref[k];
    o->rows_computed[k] = c->engine == 2 ? (int64_t)rspec[k] : (int64_t)rows[k];
    o->rows_added[k] = c->engine ? (int64_t)radd[k] : (int64_t)steps * c->Ps;  // dense: every particle's chosen cluster
  }
  for (int k = K; k < 8; ++k) { o->rows_evaluated[k] = 0; o->rows_referenced[k] = 0; o->rows_computed[k] = 0; o->rows_added[k] = 0; }
  o->engine = c->engine;
  o->n_evals = ev;
  o->n_evals_dense = dense;
  float ms = 0.f;
  CK(cudaEventElapsedTime(&ms, c->ev0, c->ev3));
  o->device_ms = ms;
  CK(cudaEventElapsedTime(&ms, c->ev1, c->ev2));
  o->sweep_kernel_ms = ms;
  for (int i = 0; i < 8; ++i) {
    double mean = 0.0, mx = 0.0;
    for (int g = 0; g < c->G; ++g) {
      const double v = (double)phase[(size_t)g * 8 + i] * 1e-6;
      mean += v / c->G;
      mx = std::max(mx, v);
    }
    o->phase_ms[i] = mean;
    o->phase_ms_max[i] = mx;
  }
  return 0;
}

int pmdi_sweep(pmdi_ctx* c, const pmdi_sweep_args* a, pmdi_sweep_out* o) {
  int rc = pmdi_sweep_upload(c, a);
  if (rc) return rc;
  rc = pmdi_sweep_run(c);
  if (rc) return rc;
  return pmdi_sweep_download(c, o);
}

// ---------------------------------------------------------------------------------------------
// feature selection and single-cluster evaluation
// ---------------------------------------------------------------------------------------------
static double gauss_marginal_cst(double n) {  // gaussian_cluster.jl:71-82
  const double a_n = n / 2 + 0.5, a_0 = 0.5, b_0 = 0.5, k_0 = 0.001, k_n = n + k_0;
  return (a_0 * std::log(b_0)) + std::lgamma(a_n) - std::lgamma(a_0) +
         0.5 * (std::log(k_0) - std::log(k_n)) - (n * 0.5) * std::log(2 * M_PI);
}

static int run_logmarginal(pmdi_ctx* c, int k, const std::vector<int>& off, const std::vector<int>& mem,
                           int use_flags, const double* base_host, double base_scale, double out_scale,
                           double* out_host, uint8_t* flags_host, const double* tape_f, uint64_t seed,
                           uint32_t iter) {
  Dataset& s = c->ds[k];
  const int nc = (int)off.size() - 1, D = s.D;
  std::vector<double> cst(std::max(nc, 1));
  for (int i = 0; i < nc; ++i) cst[i] = gauss_marginal_cst((double)(off[i + 1] - off[i]));
  cudaStream_t st = c->stream;
  CK(c->scratch_i.ensure(off.size() + mem.size() + 1));
  CK(c->scratch_d.ensure((size_t)nc + 1 + 3 * (size_t)D));
  CK(c->scratch_u8.ensure(D));
  int* d_off = c->scratch_i.p;
  int* d_mem = c->scratch_i.p + off.size();
  double* d_cst = c->scratch_d.p;
  double* d_base = d_cst + nc + 1;
  double* d_out = d_base + D;
  double* d_tape = d_out + D;
  CK(cudaMemcpyAsync(d_off, off.data(), sizeof(int) * off.size(), cudaMemcpyHostToDevice, st));
  if (!mem.empty()) CK(cudaMemcpyAsync(d_mem, mem.data(), sizeof(int) * mem.size(), cudaMemcpyHostToDevice, st));
  CK(cudaMemcpyAsync(d_cst, cst.data(), sizeof(double) * cst.size(), cudaMemcpyHostToDevice, st));
  if (base_host) CK(cudaMemcpyAsync(d_base, base_host, sizeof(double) * D, cudaMemcpyHostToDevice, st));
  if (tape_f) CK(cudaMemcpyAsync(d_tape, tape_f, sizeof(double) * D, cudaMemcpyHostToDevice, st));
  DsDev d;
  fill_dsdev(c, k, d);
  k_logmarginal<<<(D + 127) / 128, 128, 0, st>>>(d, nc, d_off, d_mem, d_cst, s.d_nlevels.p, use_flags,
                                                 base_host ? d_base : nullptr, base_scale, out_scale, d_out,
                                                 flags_host ? c->scratch_u8.p : nullptr,
                                                 tape_f ? d_tape : nullptr, seed, iter, k);
  CK(cudaGetLastError());
  CK(cudaMemcpyAsync(out_host, d_out, sizeof(double) * D, cudaMemcpyDeviceToHost, st));
  if (flags_host) CK(cudaMemcpyAsync(flags_host, c->scratch_u8.p, D, cudaMemcpyDeviceToHost, st));
  CK(cudaStreamSynchronize(st));
  return 0;
}

int pmdi_feature_null(pmdi_ctx* c, int32_t k, double* out_D) {
  if (!c || k < 0 || k >= c->K || !c->ds[k].bound) return fail(1, "pmdi_feature_null: dataset not bound");
  CK(cudaSetDevice(c->device));
  std::vector<int> off = {0, (int)c->n}, mem(c->n);
  for (long long i = 0; i < c->n; ++i) mem[i] = (int)i;
  return run_logmarginal(c, k, off, mem, 0, nullptr, 0.0, -1.0, out_D, nullptr, nullptr, 0, 0);
}

int pmdi_feature_select(pmdi_ctx* c, int32_t k, const int64_t* labels, const double* feature_null,
                        uint64_t seed, uint32_t iter, const double* tape_f, double* out_prob_D,
                        uint8_t* out_flags_D) {
  if (!c || k < 0 || k >= c->K || !c->ds[k].bound) return fail(1, "pmdi_feature_select: dataset not bound");
  CK(cudaSetDevice(c->device));
  const long long n = c->n;
  // occupied clusters in first-appearance order, members in index order (src/pmdi.jl:358-364)
  std::vector<int> order_of(c->N + 1, -1), occ;
  for (long long i = 0; i < n; ++i) {
    const int64_t l = labels[i];
    if (l < 1 || l > c->N) return fail(1, "pmdi_feature_select: label outside 1..N");
    if (order_of[l] < 0) { order_of[l] = (int)occ.size(); occ.push_back((int)l); }
  }
  std::vector<int> off(occ.size() + 1, 0), mem(n);
  for (long long i = 0; i < n; ++i) off[order_of[labels[i]] + 1] += 1;
  for (size_t i = 0; i < occ.size(); ++i) off[i + 1] += off[i];
  std::vector<int> w(off.begin(), off.end() - 1);
  for (long long i = 0; i < n; ++i) mem[w[order_of[labels[i]]]++] = (int)i;
  return run_logmarginal(c, k, off, mem, 0, feature_null, 1.0, 1.0, out_prob_D, out_flags_D, tape_f, seed, iter);
}

// ---------------------------------------------------------------------------------------------
// posterior similarity matrices (src/output_analysis/consensus_map.jl:31-65)
// ---------------------------------------------------------------------------------------------
__global__ void k_psm_add(const long long* s, unsigned* psm, long long n, int K) {
  const int k = blockIdx.z;
  const long long i = (long long)blockIdx.y * blockDim.y + threadIdx.y;
  const long long j = (long long)blockIdx.x * blockDim.x + threadIdx.x;
  if (i >= n || j >= n) return;
  const long long* sk = s + (size_t)k * n;
  if (sk[i] == sk[j]) psm[((size_t)k * n + i) * n + j] += 1u;
}
__global__ void k_psm_get(const unsigned* psm, double* out, size_t total, double inv) {
  const size_t stride = (size_t)gridDim.x * blockDim.x;
  for (size_t i = (size_t)blockIdx.x * blockDim.x + threadIdx.x; i < total; i += stride) out[i] = (double)psm[i] * inv;
}

int pmdi_psm_begin(pmdi_ctx* c) {
  if (!c) return fail(1, "NULL context");
  CK(cudaSetDevice(c->device));
  const size_t total = (size_t)c->K * c->n * c->n;
  CK(c->psm.ensure(total));
  CK(c->psm_s.ensure((size_t)c->K * c->n));
  CK(cudaMemsetAsync(c->psm.p, 0, sizeof(unsigned) * total, c->stream));
  c->psm_samples = 0;
  return 0;
}

int pmdi_psm_add(pmdi_ctx* c, const int64_t* s) {
  if (!c || !s) return fail(1, "pmdi_psm_add: NULL argument");
  if (!c->psm.p) return fail(1, "pmdi_psm_add: call pmdi_psm_begin first");
  CK(cudaSetDevice(c->device));
  for (long long i = 0; i < c->n * c->K; ++i)
    if (s[i] < 1 || s[i] > c->N) return fail(1, "pmdi_psm_add: allocation label outside 1..N");
  CK(cudaMemcpyAsync(c->psm_s.p, s, sizeof(int64_t) * c->n * c->K, cudaMemcpyHostToDevice, c->stream));
  const dim3 blk(32, 8), grd((unsigned)((c->n + 31) / 32), (unsigned)((c->n + 7) / 8), (unsigned)c->K);
  k_psm_add<<<grd, blk, 0, c->stream>>>(c->psm_s.p, c->psm.p, c->n, c->K);
  CK(cudaGetLastError());
  CK(cudaStreamSynchronize(c->stream));  // the host buffer may be reused by the caller
  c->psm_samples += 1;
  return 0;
}

int pmdi_psm_get(pmdi_ctx* c, double* out) {
  if (!c || !out) return fail(1, "pmdi_psm_get: NULL argument");
  if (!c->psm.p || c->psm_samples < 1) return fail(1, "pmdi_psm_get: nothing accumulated");
  CK(cudaSetDevice(c->device));
  const size_t total = (size_t)c->K * c->n * c->n;
  double* d = nullptr;
  CK(cudaMalloc((void**)&d, sizeof(double) * total));
  k_psm_get<<<c->n_sm * 8, 256, 0, c->stream>>>(c->psm.p, d, total, 1.0 / (double)c->psm_samples);
  cudaError_t e = cudaMemcpyAsync(out, d, sizeof(double) * total, cudaMemcpyDeviceToHost, c->stream);
  if (e == cudaSuccess) e = cudaStreamSynchronize(c->stream);
  cudaFree(d);
  if (e != cudaSuccess) return fail(100 + (int)e, std::string("pmdi_psm_get: ") + cudaGetErrorString(e));
  return 0;
}

int pmdi_cluster_eval(pmdi_ctx* c, int32_t k, const int64_t* rows, int64_t m, int64_t obs,
                      double* out_logprob, double* out_logmarg_D) {
  if (!c || k < 0 || k >= c->K || !c->ds[k].bound) return fail(1, "pmdi_cluster_eval: dataset not bound");
  int rc = prepare(c);
  if (rc) return rc;
  const long long n = c->n;
  if (m < 0 || m > n) return fail(1, "pmdi_cluster_eval: bad member count");
  std::vector<int> mem(m);
  for (int64_t i = 0; i < m; ++i) {
    if (rows[i] < 1 || rows[i] > n) return fail(1, "pmdi_cluster_eval: row outside 1..n_obs");
    mem[i] = (int)(rows[i] - 1);
  }
  Dataset& s = c->ds[k];
  cudaStream_t st = c->stream;
  if (out_logprob) {
    if (obs < 1 || obs > n) return fail(1, "pmdi_cluster_eval: obs outside 1..n_obs");
    CK(c->members.ensure(std::max<size_t>((size_t)c->K * n, 1)));
    if (m) CK(cudaMemcpyAsync(c->members.p, mem.data(), sizeof(int) * m, cudaMemcpyHostToDevice, st));
    CK(c->scratch_d.ensure(8));
    k_build_one<<<(s.Dp + 127) / 128, 128, 0, st>>>(c->sp, k, c->members.p, (int)m);
    k_aux_one<<<1, 256, 0, st>>>(c->sp, k);
    if (s.Dp * 8 > 48 * 1024)
      CK(cudaFuncSetAttribute(k_eval_row, cudaFuncAttributeMaxDynamicSharedMemorySize, s.Dp * 8));
    k_eval_row<<<1, 32, s.Dp * 8, st>>>(c->sp, k, (int)(obs - 1), c->scratch_d.p);
    CK(cudaGetLastError());
    CK(cudaMemcpyAsync(out_logprob, c->scratch_d.p, sizeof(double), cudaMemcpyDeviceToHost, st));
    CK(cudaStreamSynchronize(st));
    // leave the prototype row empty again (rows with n == 0 are always in the empty state)
    k_build_one<<<(s.Dp + 127) / 128, 128, 0, st>>>(c->sp, k, c->members.p, 0);
    k_aux_one<<<1, 256, 0, st>>>(c->sp, k);
    CK(cudaGetLastError());
    CK(cudaStreamSynchronize(st));
  }
  if (out_logmarg_D) {
    std::vector<int> off = {0, (int)m};
    rc = run_logmarginal(c, k, off, mem, 1, nullptr, 0.0, 1.0, out_logmarg_D, nullptr, nullptr, 0, 0);
    if (rc) return rc;
  }
  return 0;
}

}  // extern "C"
