"""A SECOND, independent restatement of the reference sweep, in plain Python, used to cross-check the
C++ oracle (tests/test_oracle_vs_pyref.py).  Test infrastructure only; small cases only.

It follows src/__pmdi.jl:132-318 with the reference's OWN data structures - the copy-on-write cluster
pool ``clusters[k][id]`` with ``particle[n, p, k] -> id`` and ref-counts ``clusters_counts``, the
history classes ``particle_id`` / ``new_id`` with the ``fprob_dict`` cache, the renumbering after
resampling - and the cluster maths of src/datatypes/*.jl and the helpers of src/misc.jl:15-59.
Random numbers come from the same tapes the oracle takes.  Indices are 0-based here; every Julia
line it follows is cited.

Where Julia's own floating-point evaluation order cannot be known without Julia (``@fastmath`` at
gaussian_cluster.jl:38, the ``@simd`` loop inside ``sum`` at categorical_cluster.jl:30, ``^ 2.0`` at
gaussian_cluster.jl:47) this takes the strict left-to-right reading of the source text, as the
oracle does; agreement there shows the two restatements read the text the same way, nothing more.
"""
import ctypes
import math

_libm = ctypes.CDLL("libm.so.6")
_libm.lgamma.restype = ctypes.c_double
_libm.lgamma.argtypes = [ctypes.c_double]


def lgamma(x):  # glibc's, as the C++ oracle (CPython's math.lgamma is a different implementation)
    return _libm.lgamma(float(x))


GAUSSIAN, CATEGORICAL, NEGBINOM = 0, 1, 2


# ------------------------------------------------------------------ cluster types
class Gaussian:  # src/datatypes/gaussian_cluster.jl:11-22
    def __init__(self, X):
        D = len(X[0])
        self.n, self.mu, self.sum, self.lam, self.beta = 0, [0.0] * D, [0.0] * D, [1.0] * D, [0.5] * D

    def copy(self):
        c = Gaussian.__new__(Gaussian)
        c.n, c.mu, c.sum, c.lam, c.beta = self.n, self.mu[:], self.sum[:], self.lam[:], self.beta[:]
        return c

    def logprob(self, obs, flag):  # :37-52
        out = sum(flag) * (math.log(1 / math.sqrt(math.pi)) + lgamma(0.5 * self.n + 1.0) - lgamma(0.5 * self.n + 0.5))
        for q in range(len(obs)):
            if flag[q]:
                out += 0.5 * (math.log(self.lam[q] / (self.n + 1.0)))
                d = obs[q] - self.mu[q]
                out -= (0.5 * self.n + 1.0) * math.log(1.0 + (1.0 / (self.n + 1.0)) * (d * d) * self.lam[q])
        return out

    def add(self, obs, flag):  # :54-66
        self.n += 1
        for q in range(len(obs)):
            if flag[q]:
                self.sum[q] += obs[q]
                d = obs[q] - self.mu[q]
                self.beta[q] += (self.n - 1 + 0.001) * (d * d) / (2 * (self.n + 0.001))
                self.mu[q] = self.sum[q] / (self.n + 0.001)
                self.lam[q] = ((0.5 * self.n + 0.5) * (self.n + 0.001)) / (self.beta[q] * (self.n + 1.001))


class Categorical:  # src/datatypes/categorical_cluster.jl:2-11
    def __init__(self, X):
        D = len(X[0])
        self.Lmax = max(max(r) for r in X)
        self.n = 0
        self.counts = [[0] * D for _ in range(self.Lmax)]           # [level][feature]
        self.nlevels = [0.5 * max(r[q] for r in X) for q in range(D)]

    def copy(self):
        c = Categorical.__new__(Categorical)
        c.Lmax, c.n, c.nlevels = self.Lmax, self.n, self.nlevels
        c.counts = [row[:] for row in self.counts]
        return c

    def logprob(self, obs, flag):  # :29-41
        acc = 0.0
        for q in range(len(obs)):
            if flag[q]:
                acc += math.log(self.nlevels[q] + self.n)
        out = -acc
        for q in range(len(obs)):
            if flag[q]:
                if self.n == 0:
                    out += math.log(0.5)
                else:
                    out += math.log(0.5 + self.counts[obs[q] - 1][q])
        return out

    def add(self, obs, flag):  # :43-51
        self.n += 1
        for q in range(len(obs)):
            if flag[q]:
                self.counts[obs[q] - 1][q] += 1


class NegBinom:  # src/datatypes/negbinom_cluster.jl:6-11
    def __init__(self, X):
        self.n, self.S = 0, [0] * len(X[0])

    def copy(self):
        c = NegBinom.__new__(NegBinom)
        c.n, c.S = self.n, self.S[:]
        return c

    def logprob(self, obs, flag):  # :22-41
        out = 0.0
        n = self.n
        for q in range(len(obs)):
            if flag[q]:
                x, S = obs[q], self.S[q]
                out += (lgamma(1 + n + 1) + lgamma(1 + x + S) + lgamma(1 + n + 1 + S)
                        - lgamma(1 + n + 1 + 1 + x + S) - lgamma(1 + n) - lgamma(1 + S))
        return out

    def add(self, obs, flag):  # :43-51
        self.n += 1
        for q in range(len(obs)):
            if flag[q]:
                self.S[q] += obs[q]


TYPES = {GAUSSIAN: Gaussian, CATEGORICAL: Categorical, NEGBINOM: NegBinom}


# ------------------------------------------------------------------ src/misc.jl
def calc_ess(lw):  # :15-25
    num = den = 0.0
    mx = max(lw)
    for l in lw:
        w = math.exp(l - mx)
        num += w
        den += w * w
    return (num * num) / den


def draw_partstar(lw, P, r, shuffle_u):  # :27-47
    u = r / P
    mx = max(lw)
    pprob, acc = [], 0.0
    for l in lw:
        acc += math.exp(l - mx)
        pprob.append(acc)
    ps, i = [0] * P, 0
    for p in range(P):
        while i < P and pprob[p] / pprob[-1] >= u:
            u += 1 / P
            ps[i] = p + 1
            i += 1
    while i < P:  # (the oracle's guard for a rounding case in which Julia would leave zeros)
        ps[i] = P
        i += 1
    for pos in range(P, 1, -1):  # Random.shuffle!: i = n..2, j = rand(1:i)
        j = min(pos, 1 + int(math.floor(shuffle_u[pos - 1] * pos)))
        ps[pos - 1], ps[j - 1] = ps[j - 1], ps[pos - 1]
    ps[0] = 1
    ps.sort()
    return ps  # 1-based ancestors


def phi_upweight(lw, sstar_i, K, phi, P):  # :50-59 ; sstar_i[p][k]
    idx = 0
    for k1 in range(K - 1):
        for k2 in range(k1 + 1, K):
            phil = math.log(1 + phi[idx])
            for p in range(P):
                lw[p] += (sstar_i[p][k1] == sstar_i[p][k2]) * phil
            idx += 1


# ------------------------------------------------------------------ one sweep of __pmdi
def sweep(data, types, flags, N, P, s, order_obs, n1, Pi, phi, tapes, lw_init, literal_new_id):
    """data[k]: list of rows; s: list of rows of 1-based labels; order_obs 1-based; Pi[n][k].
    literal_new_id=True keeps ``new_id`` for the whole sweep as the reference does (:135); False
    clears it for every (observation, dataset) - the corrected cache key."""
    K, n_obs = len(data), len(data[0])
    steps = n_obs - n1 + 1
    T = N * P + 1
    logweight = [lw_init] * P
    particle = [[[0] * K for _ in range(P)] for _ in range(N)]       # [n][p][k] -> pool id (0 = empty cluster)
    particle_id = [[0] * K for _ in range(P)]
    new_id = [[[-1] * K for _ in range(P)] for _ in range(N)]         # [new_s][old id][k]
    fprob_dict = [[0.0] * P for _ in range(N + 1)]
    clusters = [[None] * T for _ in range(K)]
    counts = [[0] * K for _ in range(T)]
    for k in range(K):
        counts[0][k] = P * N                                          # :133
    sstar = [[[0] * K for _ in range(n_obs)] for _ in range(P)]      # [p][i][k]
    sstar_id = [[0] * K for _ in range(P)]
    logprob = [[0.0] * K for _ in range(T)]
    n_ops = n_res = 0
    out = dict(alloc=[[[0] * P for _ in range(K)] for _ in range(steps)], anc=[[0] * P for _ in range(steps)],
               lw=[None] * steps, lp=[[[None] * P for _ in range(K)] for _ in range(steps)])

    # prefix (:154-173)
    for k in range(K):
        clusters[k][0] = TYPES[types[k]](data[k])
        clust_ids, idn, us = {}, 1, []
        for i in order_obs[:n1 - 1]:
            if s[i - 1][k] not in us:
                us.append(s[i - 1][k])                                # unique(), first appearance
        for u in us:
            clusters[k][idn] = TYPES[types[k]](data[k])
            counts[idn][k] = P
            counts[0][k] -= P
            clust_ids[u] = idn
            for p in range(P):
                particle[u - 1][p][k] = idn
            idn += 1
        for i in order_obs[:n1 - 1]:
            idn = clust_ids[s[i - 1][k]]
            for p in range(P):
                sstar[p][i - 1][k] = s[i - 1][k]
            clusters[k][idn].add(data[k][i - 1], flags[k])

    for step, i in enumerate(order_obs[n1 - 1:]):                    # :175
        for k in range(K):
            obs = data[k][i - 1]
            max_k = max(particle[n][p][k] for n in range(N) for p in range(P))
            for idn in range(max_k + 1):                              # :184-187
                logprob[idn][k] = clusters[k][idn].logprob(obs, flags[k])
                n_ops += 1
            if not literal_new_id:
                for n in range(N):
                    for p in range(P):
                        new_id[n][p][k] = -1
            fprob_done = [False] * P
            curr_id = -1
            for p in range(P):                                        # :190-240
                idn = particle_id[p][k]
                if fprob_done[idn]:
                    fprob = [fprob_dict[n][idn] for n in range(N)]
                    logweight[p] += fprob_dict[N][idn]
                else:
                    fprob = [logprob[particle[n][p][k]][k] for n in range(N)]
                    mx = max(fprob)
                    for n in range(N):
                        fprob[n] -= mx
                        fprob[n] = math.exp(fprob[n])
                        fprob[n] *= Pi[n][k]
                    for n in range(1, N):
                        fprob[n] += fprob[n - 1]                      # cumsum!
                    inc = math.log(fprob[N - 1]) + mx
                    fprob_dict[N][idn] = inc
                    logweight[p] += inc
                    tot = fprob[N - 1]
                    for n in range(N):
                        fprob[n] = fprob[n] / tot
                    for n in range(N):
                        fprob_dict[n][idn] = fprob[n]
                    fprob_done[idn] = True
                out["lp"][step][k][p] = [logprob[particle[n][p][k]][k] for n in range(N)]
                if p != 0:
                    new_s, u = 1, tapes["alloc"][step][k][p]
                    for _ in range(N - 1):
                        if fprob[new_s - 1] > u:
                            break
                        new_s += 1
                else:
                    new_s = s[i - 1][k]                               # reference trajectory
                sstar_id[p][k] = particle[new_s - 1][p][k]
                sstar[p][i - 1][k] = new_s
                out["alloc"][step][k][p] = new_s
                if new_id[new_s - 1][idn][k] == -1:
                    curr_id += 1
                    new_id[new_s - 1][idn][k] = curr_id
                    particle_id[p][k] = curr_id
                else:
                    particle_id[p][k] = new_id[new_s - 1][idn][k]
            # add the observation to the chosen clusters, copy-on-write (:241-273)
            updated = [False] * T
            max_k = max(particle[n][p][k] for n in range(N) for p in range(P))
            chosen = [sstar_id[p][k] for p in range(P)]
            for pid in chosen:
                if not updated[pid]:
                    updated[pid] = True
                    ncopies = chosen.count(pid)
                    if ncopies == counts[pid][k]:
                        idn = pid
                    else:
                        idn = max_k + 1
                        counts[pid][k] -= ncopies
                        counts[idn][k] = ncopies
                        clusters[k][idn] = clusters[k][pid].copy()
                        max_k += 1
                    clusters[k][idn].add(obs, flags[k])
                    if idn != pid:
                        for part in range(P):
                            s_id = sstar[part][i - 1][k]
                            if particle[s_id - 1][part][k] == pid:
                                particle[s_id - 1][part][k] = idn
        if K > 1:
            phi_upweight(logweight, [sstar[p][i - 1] for p in range(P)], K, phi, P)   # :275-277
        out["lw"][step] = logweight[:]
        if calc_ess(logweight) <= 0.5 * P:                            # :280
            n_res += 1
            ps = draw_partstar(logweight, P, tapes["resamp"][step], tapes["shuffle"][step])
            out["anc"][step] = ps[:]
            logweight = [1.0] * P
            for k in range(K):
                particle_new = [[None] * P for _ in range(N)]
                for n in range(N):
                    for p in range(P):
                        particle_new[n][p] = particle[n][ps[p] - 1][k]
                pid_new = [particle_id[ps[p] - 1][k] for p in range(P)]
                for n in range(N):
                    for p in range(P):
                        particle[n][p][k] = particle_new[n][p]
                for p in range(P):
                    particle_id[p][k] = pid_new[p]
                for t in range(T):
                    counts[t][k] = 0
                ids = sorted({particle[n][p][k] for n in range(N) for p in range(P)})
                for newi, idn in enumerate(ids):                      # renumber (:291-301)
                    if idn != newi:
                        for n in range(N):
                            for p in range(P):
                                if particle[n][p][k] == idn:
                                    particle[n][p][k] = newi
                        clusters[k][newi] = clusters[k][idn].copy()
                    counts[newi][k] = sum(particle[n][p][k] == newi for n in range(N) for p in range(P))
            sstar = [[row[:] for row in sstar[ps[p] - 1]] for p in range(P)]  # :285 (all datasets)
    # select (:305-311)
    mx = max(logweight)
    w = [math.exp(l - mx) for l in logweight]
    t = tapes["select"][0] * sum_seq(w)
    isel, cw = 0, w[0]
    while cw < t and isel < P - 1:
        isel += 1
        cw += w[isel]
    s_out = [sstar[isel][i][:] for i in range(n_obs)]
    # invariants of test/runtests.jl:147,151
    for k in range(K):
        for p in range(P):
            assert sum(clusters[k][particle[n][p][k]].n for n in range(N)) == n_obs
        for t_ in range(T):
            assert counts[t_][k] == sum(particle[n][p][k] == t_ for n in range(N) for p in range(P))
    out.update(s=s_out, p_star=isel + 1, logweight=logweight, n_resamples=n_res, n_ops=n_ops,
               cluster_n=[[[clusters[k][particle[n][p][k]].n for n in range(N)] for p in range(P)] for k in range(K)])
    return out


def sum_seq(w):
    acc = 0.0
    for v in w:
        acc += v
    return acc
