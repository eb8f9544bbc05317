/*
 * ORACLE — TEST INFRASTRUCTURE ONLY (see pmdi_oracle.h for the parity statement).
 *
 * Restates, in C++17 and single-threaded like the reference:
 *   - the three cluster plugins: src/datatypes/gaussian_cluster.jl:11-83,
 *     categorical_cluster.jl:2-66, negbinom_cluster.jl:6-60
 *   - the sweep helpers:         src/misc.jl:1-59
 *   - the sweep itself:          src/__pmdi.jl:132-336 (correct resampling semantics),
 *                                cross-read with src/pmdi.jl:164-375
 * in two modes that must agree bit-for-bit once the F4 cache defect is corrected:
 *   DEDUP — the reference's own copy-on-write cluster pool and fprob cache;
 *   DENSE — every particle owns its N clusters, which is what the GPU does.
 * Compile with -ffp-contract=off so that no FMA contraction changes the literal arithmetic.
 */
#include "pmdi_oracle.h"
#include "philox.h"

#include <algorithm>
#include <cmath>
#include <cstdlib>
#include <cstring>
#include <map>
#include <vector>

namespace {

struct Dataset {
  int type = -1;
  int D = 0;
  int Lmax = 0;                 // categorical: maximum(Int, dataFile) (categorical_cluster.jl:8)
  std::vector<double> xf;       // row-major n x D (Gaussian)
  std::vector<int64_t> xi;      // row-major n x D (Categorical / NegBinom)
  std::vector<double> nlevels;  // categorical: 0.5 * column max (categorical_cluster.jl:10)
  std::vector<uint8_t> flag;    // featureFlag[k] (src/pmdi.jl:106-110)
};

}  // namespace

struct or_cluster {
  int64_t n = 0;
  std::vector<double> mu, sum, lam, beta;  // gaussian_cluster.jl:11-22
  std::vector<int64_t> counts;             // categorical: [r + Lmax*q] (categorical_cluster.jl:4)
  std::vector<int64_t> isum;               // negbinom_cluster.jl:8
};

struct or_ctx {
  int K, n, N, P;
  std::vector<Dataset> ds;
};

namespace {

typedef or_cluster Cluster;

// constructors: gaussian_cluster.jl:17-21, categorical_cluster.jl:6-10, negbinom_cluster.jl:9-10
Cluster make_empty(const Dataset& d) {
  Cluster c;
  c.n = 0;
  if (d.type == OR_GAUSSIAN) {
    c.mu.assign(d.D, 0.0);
    c.sum.assign(d.D, 0.0);
    c.lam.assign(d.D, 1.0);
    c.beta.assign(d.D, 0.5);
  } else if (d.type == OR_CATEGORICAL) {
    c.counts.assign((size_t)d.Lmax * d.D, 0);
  } else {
    c.isum.assign(d.D, 0);
  }
  return c;
}

// calc_logprob: gaussian_cluster.jl:37-52, categorical_cluster.jl:29-41, negbinom_cluster.jl:22-41
double calc_logprob(const Dataset& d, int64_t i, const Cluster& cl, const uint8_t* flag) {
  const int D = d.D;
  if (d.type == OR_GAUSSIAN) {
    const double* obs = &d.xf[(size_t)i * D];
    int nflag = 0;
    for (int q = 0; q < D; ++q) nflag += flag[q] ? 1 : 0;
    const double n = (double)cl.n;
    double out = nflag * (std::log(1.0 / std::sqrt(M_PI)) + std::lgamma(0.5 * n + 1.0) -
                          std::lgamma(0.5 * n + 0.5));
    for (int q = 0; q < D; ++q) {
      if (flag[q]) {
        out += 0.5 * std::log(cl.lam[q] / (n + 1.0));
        const double dd = obs[q] - cl.mu[q];
        out -= (0.5 * n + 1.0) * std::log(1.0 + (1.0 / (n + 1.0)) * (dd * dd) * cl.lam[q]);
      }
    }
    return out;
  } else if (d.type == OR_CATEGORICAL) {
    const int64_t* obs = &d.xi[(size_t)i * D];
    double s = 0.0;
    for (int q = 0; q < D; ++q)
      if (flag[q]) s += std::log(d.nlevels[q] + (double)cl.n);
    double out = -s;
    for (int q = 0; q < D; ++q) {
      if (flag[q]) {
        if (cl.n == 0) out += std::log(0.5);
        else out += std::log(0.5 + (double)cl.counts[(size_t)(obs[q] - 1) + (size_t)d.Lmax * q]);
      }
    }
    return out;
  } else {
    const int64_t* obs = &d.xi[(size_t)i * D];
    double out = 0.0;
    const double n = (double)cl.n;
    for (int q = 0; q < D; ++q) {
      if (flag[q]) {
        const double x = (double)obs[q], S = (double)cl.isum[q];
        out += std::lgamma(1 + n + 1) + std::lgamma(1 + x + S) + std::lgamma(1 + n + 1 + S) -
               std::lgamma(1 + n + 1 + 1 + x + S) - std::lgamma(1 + n) - std::lgamma(1 + S);
      }
    }
    return out;
  }
}

// cluster_add!: gaussian_cluster.jl:54-66, categorical_cluster.jl:43-51, negbinom_cluster.jl:43-51
void cluster_add(const Dataset& d, int64_t i, Cluster& cl, const uint8_t* flag) {
  const int D = d.D;
  cl.n += 1;
  if (d.type == OR_GAUSSIAN) {
    const double* obs = &d.xf[(size_t)i * D];
    const double n = (double)cl.n;
    for (int q = 0; q < D; ++q) {
      if (flag[q]) {
        cl.sum[q] += obs[q];
        const double dd = obs[q] - cl.mu[q];
        cl.beta[q] += (n - 1 + 0.001) * (dd * dd) / (2 * (n + 0.001));
        cl.mu[q] = cl.sum[q] / (n + 0.001);
        cl.lam[q] = ((0.5 * n + 0.5) * (n + 0.001)) / (cl.beta[q] * (n + 1.001));
      }
    }
  } else if (d.type == OR_CATEGORICAL) {
    const int64_t* obs = &d.xi[(size_t)i * D];
    for (int q = 0; q < D; ++q)
      if (flag[q]) cl.counts[(size_t)(obs[q] - 1) + (size_t)d.Lmax * q] += 1;
  } else {
    const int64_t* obs = &d.xi[(size_t)i * D];
    for (int q = 0; q < D; ++q)
      if (flag[q]) cl.isum[q] += obs[q];
  }
}

// calc_logmarginal: gaussian_cluster.jl:68-83, categorical_cluster.jl:53-66, negbinom_cluster.jl:53-60
void calc_logmarginal(const Dataset& d, const Cluster& cl, double* lm) {
  const int D = d.D;
  const double n = (double)cl.n;
  if (d.type == OR_GAUSSIAN) {
    const double a_n = n / 2 + 0.5, a_0 = 0.5, b_0 = 0.5, k_0 = 0.001, k_n = n + k_0;
    const double cst = (a_0 * std::log(b_0)) + std::lgamma(a_n) - std::lgamma(a_0) +
                       0.5 * (std::log(k_0) - std::log(k_n)) - (n * 0.5) * std::log(2 * M_PI);
    for (int q = 0; q < D; ++q) lm[q] = -a_n * std::log(cl.beta[q]) + cst;
  } else if (d.type == OR_CATEGORICAL) {
    for (int q = 0; q < D; ++q) {
      double v = 0.0;
      v += std::lgamma(d.nlevels[q] * 2) - std::lgamma(d.nlevels[q] * 2 + n);
      const int R = (int)(2 * d.nlevels[q]);
      for (int r = 0; r < R; ++r) v += std::lgamma((double)cl.counts[(size_t)r + (size_t)d.Lmax * q] + 0.5);
      lm[q] = v;
    }
  } else {
    for (int q = 0; q < D; ++q) {
      const double S = (double)cl.isum[q];
      lm[q] = std::lgamma(S + 1) - std::lgamma(S + (n + 1 + 1)) + std::lgamma(1 + n);
    }
  }
}

// calc_ESS: src/misc.jl:15-25
double calc_ess(const double* lw, int P) {
  double num = 0.0, den = 0.0, mx = lw[0];
  for (int p = 1; p < P; ++p) mx = std::max(mx, lw[p]);
  for (int p = 0; p < P; ++p) {
    const double w = std::exp(lw[p] - mx);
    num += w;
    den += w * w;
  }
  return (num * num) / den;
}

// draw_partstar: src/misc.jl:27-47.  `r` replaces rand() (:28); shuffle_u[i-1] is the uniform
// behind the Fisher-Yates pick for position i = P..2 (Random.shuffle!, :43).  Output is 1-based.
// Guards (documented deviations for measure-zero rounding cases where the Julia code would
// index out of bounds or leave zeros): at most P ancestors are written; missing ones get P.
void draw_partstar(const double* lw, int P, double r, const double* shuffle_u, int64_t* ps) {
  double u = r / P;
  double mx = lw[0];
  for (int p = 1; p < P; ++p) mx = std::max(mx, lw[p]);
  std::vector<double> pp(P);
  double acc = 0.0;
  for (int p = 0; p < P; ++p) { acc += std::exp(lw[p] - mx); pp[p] = acc; }
  int i = 0;
  for (int p = 0; p < P; ++p) {
    while (i < P && pp[p] / pp[P - 1] >= u) {
      u += 1.0 / P;
      ps[i++] = p + 1;
    }
  }
  while (i < P) ps[i++] = P;
  for (int pos = P; pos >= 2; --pos) {
    int j = 1 + (int)std::floor(shuffle_u[pos - 1] * pos);
    if (j > pos) j = pos;
    std::swap(ps[pos - 1], ps[j - 1]);
  }
  ps[0] = 1;
  std::sort(ps, ps + P);
}

// StatsBase.sample(1:P, Weights(w)) v0.33.0 (src/pmdi.jl:345-350): t = rand()*sum(w), linear scan.
int sample_weighted(const double* lw, int P, double u) {
  double mx = lw[0];
  for (int p = 1; p < P; ++p) mx = std::max(mx, lw[p]);
  std::vector<double> w(P);
  double tot = 0.0;
  for (int p = 0; p < P; ++p) { w[p] = std::exp(lw[p] - mx); tot += w[p]; }
  const double t = u * tot;
  int i = 0;
  double cw = w[0];
  while (cw < t && i < P - 1) { ++i; cw += w[i]; }
  return i;  // 0-based
}

struct Draws {
  const or_sweep_args* a;
  int K, P;
  double alloc(int step, int k, int p) const {  // p 0-based, p >= 1
    if (a->tape_alloc) return a->tape_alloc[((size_t)step * K + k) * P + p];
    return or_uniform(a->seed, a->iter, OR_DRAW_ALLOC, step, k, p);
  }
  double resamp(int step) const {
    if (a->tape_resamp) return a->tape_resamp[step];
    return or_uniform(a->seed, a->iter, OR_DRAW_RESAMP, step, 0, 0);
  }
  void shuffle(int step, std::vector<double>& out) const {
    out.resize(P);
    for (int i = 0; i < P; ++i)
      out[i] = a->tape_shuffle ? a->tape_shuffle[(size_t)step * P + i]
                               : or_uniform(a->seed, a->iter, OR_DRAW_SHUFFLE, step, 0, i);
  }
  double select() const {
    if (a->tape_select) return a->tape_select[0];
    return or_uniform(a->seed, a->iter, OR_DRAW_SELECT, 0, 0, 0);
  }
};

// Allocation proposal for one particle (src/pmdi.jl:231-247): in: lp[N]; out: fprob = cdf[N],
// returns the incremental log-weight.
double propose(const double* lp, const double* Pi_k, int N, double* fprob) {
  double mx = lp[0];
  for (int m = 1; m < N; ++m) mx = std::max(mx, lp[m]);
  for (int m = 0; m < N; ++m) {
    double f = lp[m] - mx;
    f = std::exp(f);
    f *= Pi_k[m];
    fprob[m] = f;
  }
  for (int m = 1; m < N; ++m) fprob[m] += fprob[m - 1];  // cumsum!
  const double inc = std::log(fprob[N - 1]) + mx;
  const double tot = fprob[N - 1];
  for (int m = 0; m < N; ++m) fprob[m] = fprob[m] / tot;
  return inc;
}

// inverse-CDF draw with strict '>' and a cap at N (src/pmdi.jl:252-260); returns 1-based label
int draw_label(const double* fprob, int N, double u) {
  int new_s = 1;
  for (int it = 1; it <= N - 1; ++it) {
    if (fprob[new_s - 1] > u) break;
    new_s += 1;
  }
  return new_s;
}

// Phi_upweight!: src/misc.jl:50-59
void phi_upweight(double* lw, const std::vector<int>& sstar_i /*[p*K+k]*/, int K, const double* phi,
                  int P) {
  int idx = 0;
  for (int k1 = 0; k1 < K - 1; ++k1)
    for (int k2 = k1 + 1; k2 < K; ++k2) {
      const double phil = std::log(1 + phi[idx]);
      for (int p = 0; p < P; ++p)
        lw[p] += (sstar_i[(size_t)p * K + k1] == sstar_i[(size_t)p * K + k2]) * phil;
      ++idx;
    }
}

// ---------------------------------------------------------------------------------------
// DENSE sweep: every particle owns N clusters per dataset.
// ---------------------------------------------------------------------------------------
int sweep_dense(or_ctx* c, or_sweep_args* a) {
  const int K = c->K, n = c->n, N = c->N, P = c->P;
  const int n1 = (int)a->n1;
  const int steps = n - n1 + 1;
  Draws dr{a, K, P};
  // cl[k][p*N + m]
  std::vector<std::vector<Cluster>> cl(K);
  std::vector<int32_t> sstar((size_t)P * n * K, 0);  // [ (p*n + i)*K + k ], labels 1-based
  auto SS = [&](int p, int i, int k) -> int32_t& { return sstar[((size_t)p * n + i) * K + k]; };
  std::vector<double> lw(P, a->logweight_init);
  int64_t n_ops = 0, n_res = 0;

  // prefix (src/pmdi.jl:188-207)
  for (int k = 0; k < K; ++k) {
    const Dataset& d = c->ds[k];
    std::vector<Cluster> proto(N, make_empty(d));
    for (int t = 0; t < n1 - 1; ++t) {
      const int i = (int)a->order_obs[t] - 1;
      const int lab = (int)a->s_in[(size_t)i + (size_t)n * k];
      cluster_add(d, i, proto[lab - 1], d.flag.data());
      for (int p = 0; p < P; ++p) SS(p, i, k) = lab;
    }
    cl[k].resize((size_t)P * N);
    for (int p = 0; p < P; ++p)
      for (int m = 0; m < N; ++m) cl[k][(size_t)p * N + m] = proto[m];
  }

  std::vector<double> lp(N), fprob(N);
  std::vector<int> chosen((size_t)P * K);
  std::vector<int64_t> anc(P);
  std::vector<double> shuf;
  for (int step = 0; step < steps; ++step) {
    const int i = (int)a->order_obs[n1 - 1 + step] - 1;
    for (int k = 0; k < K; ++k) {
      const Dataset& d = c->ds[k];
      const double* Pi_k = a->Pi + (size_t)N * k;
      for (int p = 0; p < P; ++p) {
        for (int m = 0; m < N; ++m) {
          lp[m] = calc_logprob(d, i, cl[k][(size_t)p * N + m], d.flag.data());
          ++n_ops;
        }
        if (a->dbg_lp)
          std::memcpy(a->dbg_lp + (((size_t)step * K + k) * P + p) * N, lp.data(), sizeof(double) * N);
        lw[p] += propose(lp.data(), Pi_k, N, fprob.data());
        int lab;
        if (p != 0) lab = draw_label(fprob.data(), N, dr.alloc(step, k, p));
        else lab = (int)a->s_in[(size_t)i + (size_t)n * k];  // reference trajectory (src/pmdi.jl:262)
        SS(p, i, k) = lab;
        chosen[(size_t)p * K + k] = lab;
        if (a->dbg_alloc) a->dbg_alloc[((size_t)step * K + k) * P + p] = lab;
        cluster_add(d, i, cl[k][(size_t)p * N + (lab - 1)], d.flag.data());
      }
    }
    if (K > 1) phi_upweight(lw.data(), chosen, K, a->phi, P);
    if (a->dbg_lw) std::memcpy(a->dbg_lw + (size_t)step * P, lw.data(), sizeof(double) * P);
    if (a->dbg_anc) std::memset(a->dbg_anc + (size_t)step * P, 0, sizeof(int32_t) * P);
    // resampling (src/pmdi.jl:317-341, src/__pmdi.jl:279-302)
    if (calc_ess(lw.data(), P) <= 0.5 * P) {
      ++n_res;
      dr.shuffle(step, shuf);
      draw_partstar(lw.data(), P, dr.resamp(step), shuf.data(), anc.data());
      std::fill(lw.begin(), lw.end(), 1.0);
      if (a->dbg_anc)
        for (int p = 0; p < P; ++p) a->dbg_anc[(size_t)step * P + p] = (int32_t)anc[p];
      for (int k = 0; k < K; ++k) {
        std::vector<Cluster> nc((size_t)P * N);
        for (int p = 0; p < P; ++p)
          for (int m = 0; m < N; ++m) nc[(size_t)p * N + m] = cl[k][(size_t)(anc[p] - 1) * N + m];
        cl[k].swap(nc);
      }
      if (!(a->mode & OR_MODE_SSTAR_COMPAT)) {
        std::vector<int32_t> ns(sstar.size());
        for (int p = 0; p < P; ++p)
          std::memcpy(&ns[(size_t)p * n * K], &sstar[(size_t)(anc[p] - 1) * n * K], sizeof(int32_t) * n * K);
        sstar.swap(ns);
      }
    }
  }
  if (a->logweight) std::memcpy(a->logweight, lw.data(), sizeof(double) * P);
  const int ps = sample_weighted(lw.data(), P, dr.select());
  if (a->p_star) *a->p_star = ps + 1;
  if (a->s_out)
    for (int k = 0; k < K; ++k)
      for (int i = 0; i < n; ++i) a->s_out[(size_t)i + (size_t)n * k] = SS(ps, i, k);
  if (a->n_ops) *a->n_ops = n_ops;
  if (a->n_resamples) *a->n_resamples = n_res;
  if (a->cluster_n)
    for (int k = 0; k < K; ++k)
      for (int p = 0; p < P; ++p)
        for (int m = 0; m < N; ++m) a->cluster_n[((size_t)k * P + p) * N + m] = cl[k][(size_t)p * N + m].n;
  return 0;
}

// ---------------------------------------------------------------------------------------
// DEDUP sweep: the reference's own data structures (src/pmdi.jl:131-146), 1-based ids with
// pool id 1 = the shared empty cluster.
// ---------------------------------------------------------------------------------------
int sweep_dedup(or_ctx* c, or_sweep_args* a) {
  const int K = c->K, n = c->n, N = c->N, P = c->P;
  const int n1 = (int)a->n1;
  const int steps = n - n1 + 1;
  const bool literal = (a->mode & OR_MODE_LITERAL_NEWID) != 0;
  Draws dr{a, K, P};
  const int NP1 = N * P + 1;
  std::vector<int> particle((size_t)N * P * K, 1);  // [m + N*(p + P*k)] -> pool id
  auto PART = [&](int m, int p, int k) -> int& { return particle[(size_t)m + (size_t)N * (p + (size_t)P * k)]; };
  std::vector<int> particle_id((size_t)P * K, 1);
  std::vector<int> new_id((size_t)N * (P + 1) * K, 0);  // [m + N*(id + (P+1)*k)], id 1..P
  auto NEWID = [&](int m, int id, int k) -> int& { return new_id[(size_t)m + (size_t)N * (id + (size_t)(P + 1) * k)]; };
  std::vector<double> fprob_dict((size_t)(N + 1) * (P + 1));
  std::vector<uint8_t> fprob_done(P + 1);
  std::vector<uint8_t> cluster_update(NP1 + 1);
  std::vector<double> logprob(NP1 + 1);
  std::vector<std::vector<Cluster>> clusters(K);
  std::vector<int64_t> counts((size_t)(NP1 + 1) * K, 0);
  auto CNT = [&](int id, int k) -> int64_t& { return counts[(size_t)id + (size_t)(NP1 + 1) * k]; };
  std::vector<int> sstar_id((size_t)P * K);
  std::vector<int32_t> sstar((size_t)P * n * K, 0);
  auto SS = [&](int p, int i, int k) -> int32_t& { return sstar[((size_t)p * n + i) * K + k]; };
  std::vector<double> lw(P, a->logweight_init);
  int64_t n_ops = 0, n_res = 0;

  for (int k = 0; k < K; ++k) {
    clusters[k].resize(NP1 + 1);
    CNT(1, k) = (int64_t)P * N;
  }
  // prefix (src/pmdi.jl:188-207)
  for (int k = 0; k < K; ++k) {
    const Dataset& d = c->ds[k];
    clusters[k][1] = make_empty(d);
    std::map<int, int> clust_ids;
    std::vector<int> us;  // unique, first-appearance order
    for (int t = 0; t < n1 - 1; ++t) {
      const int i = (int)a->order_obs[t] - 1;
      const int lab = (int)a->s_in[(size_t)i + (size_t)n * k];
      if (!clust_ids.count(lab)) { clust_ids[lab] = 0; us.push_back(lab); }
    }
    int id = 2;
    for (int u : us) {
      clusters[k][id] = make_empty(d);
      CNT(id, k) = P;
      CNT(1, k) -= P;
      clust_ids[u] = id;
      for (int p = 0; p < P; ++p) PART(u - 1, p, k) = id;
      ++id;
    }
    for (int t = 0; t < n1 - 1; ++t) {
      const int i = (int)a->order_obs[t] - 1;
      const int lab = (int)a->s_in[(size_t)i + (size_t)n * k];
      for (int p = 0; p < P; ++p) SS(p, i, k) = lab;
      cluster_add(d, i, clusters[k][clust_ids[lab]], d.flag.data());
    }
  }

  std::vector<double> fprob(N);
  std::vector<int> chosen((size_t)P * K);
  std::vector<int64_t> anc(P);
  std::vector<double> shuf;
  for (int step = 0; step < steps; ++step) {
    const int i = (int)a->order_obs[n1 - 1 + step] - 1;
    for (int k = 0; k < K; ++k) {
      const Dataset& d = c->ds[k];
      const double* Pi_k = a->Pi + (size_t)N * k;
      std::fill(cluster_update.begin(), cluster_update.end(), 0);
      std::fill(fprob_done.begin(), fprob_done.end(), 0);
      if (!literal)  // F4 correction: the (label, old id) -> new id map is per (i, k)
        std::fill(new_id.begin() + (size_t)N * (P + 1) * k, new_id.begin() + (size_t)N * (P + 1) * (k + 1), 0);
      int maxid = 1;
      for (int p = 0; p < P; ++p)
        for (int m = 0; m < N; ++m) maxid = std::max(maxid, PART(m, p, k));
      for (int id = 1; id <= maxid; ++id) {  // src/pmdi.jl:218-220
        logprob[id] = calc_logprob(d, i, clusters[k][id], d.flag.data());
        ++n_ops;
      }
      if (a->dbg_lp)
        for (int p = 0; p < P; ++p)
          for (int m = 0; m < N; ++m)
            a->dbg_lp[(((size_t)step * K + k) * P + p) * N + m] = logprob[PART(m, p, k)];
      int curr_id = 0;
      for (int p = 0; p < P; ++p) {  // src/pmdi.jl:223-273
        const int id = particle_id[(size_t)p + (size_t)P * k];
        double* dict = &fprob_dict[(size_t)(N + 1) * id];
        if (fprob_done[id]) {
          for (int m = 0; m < N; ++m) fprob[m] = dict[m];
          lw[p] += dict[N];
        } else {
          std::vector<double> lp(N);
          for (int m = 0; m < N; ++m) lp[m] = logprob[PART(m, p, k)];
          const double inc = propose(lp.data(), Pi_k, N, fprob.data());
          dict[N] = inc;
          lw[p] += inc;
          for (int m = 0; m < N; ++m) dict[m] = fprob[m];
          fprob_done[id] = 1;
        }
        int new_s;
        if (p != 0) new_s = draw_label(fprob.data(), N, dr.alloc(step, k, p));
        else new_s = (int)a->s_in[(size_t)i + (size_t)n * k];
        sstar_id[(size_t)p + (size_t)P * k] = PART(new_s - 1, p, k);
        SS(p, i, k) = new_s;
        chosen[(size_t)p * K + k] = new_s;
        if (a->dbg_alloc) a->dbg_alloc[((size_t)step * K + k) * P + p] = new_s;
        if (NEWID(new_s - 1, id, k) == 0) {
          ++curr_id;
          NEWID(new_s - 1, id, k) = curr_id;
          particle_id[(size_t)p + (size_t)P * k] = curr_id;
        } else {
          particle_id[(size_t)p + (size_t)P * k] = NEWID(new_s - 1, id, k);
        }
      }
      // copy-on-write cluster update (src/pmdi.jl:275-310)
      int max_k = maxid;
      for (int pp = 0; pp < P; ++pp) {
        const int pid = sstar_id[(size_t)pp + (size_t)P * k];
        if (cluster_update[pid]) continue;
        cluster_update[pid] = 1;
        int64_t ncopies = 0;
        for (int q = 0; q < P; ++q) ncopies += (sstar_id[(size_t)q + (size_t)P * k] == pid);
        int id;
        if (ncopies == CNT(pid, k)) {
          id = pid;
        } else {
          id = max_k + 1;
          CNT(pid, k) -= ncopies;
          CNT(id, k) = ncopies;
          clusters[k][id] = clusters[k][pid];  // deepcopy (:297)
          max_k += 1;
        }
        cluster_add(d, i, clusters[k][id], d.flag.data());
        if (id != pid) {
          for (int part = 0; part < P; ++part) {
            const int s_id = SS(part, i, k);
            if (PART(s_id - 1, part, k) == pid) PART(s_id - 1, part, k) = id;
          }
        }
      }
    }
    if (K > 1) phi_upweight(lw.data(), chosen, K, a->phi, P);
    if (a->dbg_lw) std::memcpy(a->dbg_lw + (size_t)step * P, lw.data(), sizeof(double) * P);
    if (a->dbg_anc) std::memset(a->dbg_anc + (size_t)step * P, 0, sizeof(int32_t) * P);
    if (calc_ess(lw.data(), P) <= 0.5 * P) {  // src/__pmdi.jl:279-302
      ++n_res;
      dr.shuffle(step, shuf);
      draw_partstar(lw.data(), P, dr.resamp(step), shuf.data(), anc.data());
      std::fill(lw.begin(), lw.end(), 1.0);
      if (a->dbg_anc)
        for (int p = 0; p < P; ++p) a->dbg_anc[(size_t)step * P + p] = (int32_t)anc[p];
      if (!(a->mode & OR_MODE_SSTAR_COMPAT)) {
        std::vector<int32_t> ns(sstar.size());
        for (int p = 0; p < P; ++p)
          std::memcpy(&ns[(size_t)p * n * K], &sstar[(size_t)(anc[p] - 1) * n * K], sizeof(int32_t) * n * K);
        sstar.swap(ns);
      }
      for (int k = 0; k < K; ++k) {
        std::vector<int> np((size_t)N * P), npid(P);
        for (int p = 0; p < P; ++p) {
          for (int m = 0; m < N; ++m) np[(size_t)m + (size_t)N * p] = PART(m, (int)anc[p] - 1, k);
          npid[p] = particle_id[(size_t)(anc[p] - 1) + (size_t)P * k];
        }
        for (int p = 0; p < P; ++p) {
          for (int m = 0; m < N; ++m) PART(m, p, k) = np[(size_t)m + (size_t)N * p];
          particle_id[(size_t)p + (size_t)P * k] = npid[p];
        }
        for (int id = 0; id <= NP1; ++id) CNT(id, k) = 0;
        // renumber pool ids to 1..U ascending (src/pmdi.jl:329-339); equivalent to the
        // reference's repeated full scans because the sorted map is monotone (new id <= old id)
        std::vector<int> uniq(np.begin(), np.end());
        std::sort(uniq.begin(), uniq.end());
        uniq.erase(std::unique(uniq.begin(), uniq.end()), uniq.end());
        std::vector<int> remap(NP1 + 1, 0);
        for (size_t j = 0; j < uniq.size(); ++j) {
          const int id = uniq[j], ni = (int)j + 1;
          remap[id] = ni;
          if (id != ni) clusters[k][ni] = clusters[k][id];
        }
        for (int p = 0; p < P; ++p)
          for (int m = 0; m < N; ++m) {
            int& v = PART(m, p, k);
            v = remap[v];
            CNT(v, k) += 1;
          }
      }
    }
  }
  if (a->logweight) std::memcpy(a->logweight, lw.data(), sizeof(double) * P);
  const int ps = sample_weighted(lw.data(), P, dr.select());
  if (a->p_star) *a->p_star = ps + 1;
  if (a->s_out)
    for (int k = 0; k < K; ++k)
      for (int i = 0; i < n; ++i) a->s_out[(size_t)i + (size_t)n * k] = SS(ps, i, k);
  if (a->n_ops) *a->n_ops = n_ops;
  if (a->n_resamples) *a->n_resamples = n_res;
  if (a->cluster_n) {
    for (int k = 0; k < K; ++k)
      for (int p = 0; p < P; ++p)
        for (int m = 0; m < N; ++m)
          a->cluster_n[((size_t)k * P + p) * N + m] = clusters[k][PART(m, p, k)].n;
    // the reference's second invariant (test/runtests.jl:149-153): ref-counts match the map
    for (int k = 0; k < K; ++k) {
      std::vector<int64_t> chk(NP1 + 1, 0);
      for (int p = 0; p < P; ++p)
        for (int m = 0; m < N; ++m) chk[PART(m, p, k)] += 1;
      for (int id = 1; id <= NP1; ++id)
        if (chk[id] != CNT(id, k)) return 100 + k;
    }
  }
  return 0;
}

}  // namespace

extern "C" {

or_ctx* or_create(int K, int n_obs, int N, int P) {
  or_ctx* c = new or_ctx();
  c->K = K; c->n = n_obs; c->N = N; c->P = P;
  c->ds.resize(K);
  return c;
}
void or_destroy(or_ctx* c) { delete c; }

int or_set_dataset(or_ctx* c, int k, int type, const void* data, int D) {
  if (k < 0 || k >= c->K) return 1;
  Dataset& d = c->ds[k];
  d.type = type; d.D = D;
  const int n = c->n;
  d.flag.assign(D, 1);
  if (type == OR_GAUSSIAN) {
    const double* x = (const double*)data;
    d.xf.resize((size_t)n * D);
    for (int i = 0; i < n; ++i)
      for (int q = 0; q < D; ++q) d.xf[(size_t)i * D + q] = x[(size_t)i + (size_t)n * q];
  } else {
    const int64_t* x = (const int64_t*)data;
    d.xi.resize((size_t)n * D);
    int64_t gmax = 0;
    d.nlevels.assign(D, 0.0);
    for (int q = 0; q < D; ++q) {
      int64_t cmax = x[(size_t)n * q];
      for (int i = 0; i < n; ++i) {
        const int64_t v = x[(size_t)i + (size_t)n * q];
        d.xi[(size_t)i * D + q] = v;
        cmax = std::max(cmax, v);
      }
      d.nlevels[q] = 0.5 * (double)cmax;
      gmax = std::max(gmax, cmax);
    }
    d.Lmax = (int)gmax;
  }
  return 0;
}

int or_set_flags(or_ctx* c, int k, const uint8_t* flags) {
  if (k < 0 || k >= c->K) return 1;
  Dataset& d = c->ds[k];
  for (int q = 0; q < d.D; ++q) d.flag[q] = flags[q] ? 1 : 0;
  return 0;
}

int or_sweep(or_ctx* c, or_sweep_args* a) {
  if (a->n1 < 1 || a->n1 > c->n) return 2;
  if (a->mode & OR_MODE_DEDUP) return sweep_dedup(c, a);
  return sweep_dense(c, a);
}

int or_feature_null(or_ctx* c, int k, double* out) {
  const Dataset& d = c->ds[k];
  Cluster cl = make_empty(d);
  std::vector<uint8_t> ones(d.D, 1);
  for (int i = 0; i < c->n; ++i) cluster_add(d, i, cl, ones.data());
  calc_logmarginal(d, cl, out);
  for (int q = 0; q < d.D; ++q) out[q] = -out[q];
  return 0;
}

// src/pmdi.jl:354-370: occupied clusters in first-appearance order, members in index order
int or_feature_select(or_ctx* c, int k, const int64_t* labels, const double* feature_null,
                      uint64_t seed, uint32_t iter, const double* tape_f, double* prob,
                      uint8_t* flags) {
  const Dataset& d = c->ds[k];
  const int n = c->n, D = d.D;
  std::vector<uint8_t> ones(D, 1);
  for (int q = 0; q < D; ++q) prob[q] = feature_null[q] + 0;
  std::vector<int64_t> occ;
  for (int i = 0; i < n; ++i)
    if (std::find(occ.begin(), occ.end(), labels[i]) == occ.end()) occ.push_back(labels[i]);
  std::vector<double> lm(D);
  for (int64_t lab : occ) {
    Cluster cl = make_empty(d);
    for (int i = 0; i < n; ++i)
      if (labels[i] == lab) cluster_add(d, i, cl, ones.data());
    calc_logmarginal(d, cl, lm.data());
    for (int q = 0; q < D; ++q) prob[q] += lm[q];
  }
  for (int q = 0; q < D; ++q) {
    const double u = tape_f ? tape_f[q] : or_uniform(seed, iter, OR_DRAW_FEATURE, 0, k, q);
    flags[q] = ((1 - 1 / (std::exp(prob[q] + 1))) > u) ? 1 : 0;
  }
  return 0;
}

or_cluster* or_cl_new(or_ctx* c, int k) { return new or_cluster(make_empty(c->ds[k])); }
void or_cl_free(or_cluster* cl) { delete cl; }
void or_cl_add(or_ctx* c, int k, or_cluster* cl, int64_t obs) {
  cluster_add(c->ds[k], obs - 1, *cl, c->ds[k].flag.data());
}
double or_cl_logprob(or_ctx* c, int k, or_cluster* cl, int64_t obs) {
  return calc_logprob(c->ds[k], obs - 1, *cl, c->ds[k].flag.data());
}
void or_cl_logmarginal(or_ctx* c, int k, or_cluster* cl, double* out) {
  calc_logmarginal(c->ds[k], *cl, out);
}
int64_t or_cl_n(or_cluster* cl) { return cl->n; }
int or_cl_get(or_cluster* cl, int field, double* out) {
  const std::vector<double>* v = nullptr;
  switch (field) {
    case 0: v = &cl->mu; break;
    case 1: v = &cl->sum; break;
    case 2: v = &cl->lam; break;
    case 3: v = &cl->beta; break;
    case 4: for (size_t j = 0; j < cl->counts.size(); ++j) out[j] = (double)cl->counts[j]; return 0;
    case 5: for (size_t j = 0; j < cl->isum.size(); ++j) out[j] = (double)cl->isum[j]; return 0;
    default: return 1;
  }
  std::memcpy(out, v->data(), sizeof(double) * v->size());
  return 0;
}

double or_calc_ess(const double* lw, int P) { return calc_ess(lw, P); }
void or_draw_partstar(const double* lw, int P, double r, const double* shuffle_u, int64_t* out) {
  draw_partstar(lw, P, r, shuffle_u, out);
}
double or_uniform_c(uint64_t seed, uint32_t iter, uint32_t kind, uint32_t step, uint32_t k,
                    uint32_t index) {
  return or_uniform(seed, iter, kind, step, k, index);
}

}  // extern "C"
