"""Pins the CPU oracle (test infrastructure) against the identities the reference's own tests
assert (test/runtests.jl:13-54) re-evaluated with scipy, and against scipy closed forms for the
parts the reference never tests (NegBinom, calc_logmarginal, calc_ESS, draw_partstar).

The reference is Julia and cannot run here, and it ships no golden vectors: these
identities are what "the reference's known-answer tests for this path" amount to.
"""
import math

import numpy as np
import pytest
from scipy import special, stats

from oracle import oracle as orc

G, C, NB = orc.GAUSSIAN, orc.CATEGORICAL, orc.NEGBINOM


def _ctx(data, types, N=2, P=2):
    return orc.Oracle(data, types, N, P)


# ------------------------------------------------------------------ Gaussian (runtests.jl:13-36)
def test_gaussian_cluster_add_identities():
    rng = np.random.default_rng(0)
    n = 1000
    x = rng.normal(3.0, 2.0, (n + 1, 1))
    o = _ctx([x], [G])
    cl = o.cluster(0)
    for i in range(1, n + 1):
        cl.add(i)
    xs = x[:n, 0]
    assert cl.n == n
    # runtests.jl:22-30
    np.testing.assert_allclose(cl.get("sum")[0], xs.sum(), rtol=1e-12)
    np.testing.assert_allclose(cl.get("mu")[0], xs.sum() / (n + 0.001), rtol=1e-12)
    xbar = xs.mean()
    beta = 0.5 + 0.5 * (((xs - xbar) ** 2).sum() + 0.001 * n * xbar ** 2 / (n + 0.001))
    np.testing.assert_allclose(cl.get("beta")[0], beta, rtol=1.5e-8)
    lam = (0.5 + 0.5 * n) * (n + 0.001) / (beta * (n + 1.001))
    np.testing.assert_allclose(cl.get("lam")[0], lam, rtol=1.5e-8)
    # runtests.jl:33-36: Student-t predictive
    mu, lam_o = cl.get("mu")[0], cl.get("lam")[0]
    want = stats.t.logpdf((x[n, 0] - mu) * math.sqrt(lam_o), df=n + 1) + 0.5 * math.log(lam_o)
    np.testing.assert_allclose(cl.logprob(n + 1), want, rtol=1.5e-8)


def test_gaussian_empty_cluster_predictive_is_student_t():
    x = np.array([[0.7, -1.2, 2.5]])
    o = _ctx([np.repeat(x, 2, axis=0)], [G])
    cl = o.cluster(0)
    want = sum(stats.t.logpdf(v, df=1) for v in x[0])  # n = 0: mu = 0, lambda = 1
    np.testing.assert_allclose(cl.logprob(1), want, rtol=1e-12)


def test_gaussian_logmarginal_closed_form():
    rng = np.random.default_rng(1)
    x = rng.normal(size=(30, 5))
    o = _ctx([x], [G])
    cl = o.cluster(0)
    for i in range(1, 21):
        cl.add(i)
    n = 20
    beta = cl.get("beta")
    a_n = n / 2 + 0.5
    want = (-a_n * np.log(beta) + 0.5 * math.log(0.5) + special.gammaln(a_n) - special.gammaln(0.5)
            + 0.5 * (math.log(0.001) - math.log(n + 0.001)) - 0.5 * n * math.log(2 * math.pi))
    np.testing.assert_allclose(cl.logmarginal(), want, rtol=1e-13)
    # and it IS the NIG marginal likelihood of the 20 rows: chain rule over the predictive
    # densities (gaussian_cluster.jl:37-52 is the exact NIG posterior predictive for n >= 1; the
    # empty cluster is constructed with lambda = 1 (:20), not the prior predictive, so the chain
    # starts at the second row)
    c1 = o.cluster(0)
    c1.add(1)
    chain = c1.logmarginal()
    for i in range(2, 21):
        for q in range(5):
            oq = _ctx([x[:, q:q + 1]], [G])
            c = oq.cluster(0)
            for j in range(1, i):
                c.add(j)
            chain[q] += c.logprob(i)
    np.testing.assert_allclose(cl.logmarginal(), chain, rtol=1e-9)


# --------------------------------------------------------------- Categorical (runtests.jl:38-54)
def test_categorical_counts_and_predictive():
    rng = np.random.default_rng(2)
    L, n = 10, 1000
    x = rng.integers(1, L + 1, (n + 1, 1)).astype(np.int64)
    x[0, 0] = L
    x[n, 0] = 1  # runtests.jl:52 evaluates level 1
    o = _ctx([x], [C])
    cl = o.cluster(0)
    for i in range(1, n + 1):
        cl.add(i)
    counts = cl.get("counts")[:, 0]
    want_counts = np.bincount(x[:n, 0], minlength=L + 1)[1:]
    np.testing.assert_array_equal(counts, want_counts)
    want = math.log((want_counts[0] + 0.5) / (n + 0.5 * L))
    np.testing.assert_allclose(cl.logprob(n + 1), want, rtol=1.5e-8)


def test_categorical_empty_cluster_and_logmarginal():
    rng = np.random.default_rng(3)
    L, n, D = 3, 40, 4
    x = rng.integers(1, L + 1, (n, D)).astype(np.int64)
    x[0, :] = L
    o = _ctx([x], [C])
    cl = o.cluster(0)
    np.testing.assert_allclose(cl.logprob(5), D * math.log(0.5 / (0.5 * L)), rtol=1e-13)
    for i in range(1, 26):
        cl.add(i)
    cnt = np.stack([np.bincount(x[:25, q], minlength=L + 1)[1:] for q in range(D)], axis=1)
    want = (special.gammaln(0.5 * L * 2) - special.gammaln(0.5 * L * 2 + 25)
            + special.gammaln(cnt + 0.5).sum(axis=0))
    np.testing.assert_allclose(cl.logmarginal(), want, rtol=1e-13)


# ------------------------------------------------------------------ NegBinom (unpinned upstream)
def test_negbinom_predictive_is_beta_geometric():
    rng = np.random.default_rng(4)
    n, D = 30, 6
    x = rng.poisson(6.0, (n + 1, D)).astype(np.int64)
    o = _ctx([x], [NB])
    cl = o.cluster(0)
    # n = 0
    want0 = (special.betaln(0 + 2, 1 + 0 + x[3]) - special.betaln(0 + 1, 1 + 0)).sum()
    np.testing.assert_allclose(cl.logprob(4), want0, rtol=1e-12)
    for i in range(1, n + 1):
        cl.add(i)
    S = x[:n].sum(axis=0)
    np.testing.assert_array_equal(cl.get("isum"), S)
    want = (special.betaln(n + 2, 1 + S + x[n]) - special.betaln(n + 1, 1 + S)).sum()
    np.testing.assert_allclose(cl.logprob(n + 1), want, rtol=1e-11)
    lm = special.gammaln(S + 1) - special.gammaln(S + n + 2) + special.gammaln(1 + n)
    np.testing.assert_allclose(cl.logmarginal(), lm, rtol=1e-13)
    # = log B(n+1, S+1): the marginal of n geometric draws under a Beta(1,1) prior
    np.testing.assert_allclose(cl.logmarginal(), special.betaln(n + 1, S + 1), rtol=1e-12)


# ----------------------------------------------------------------- feature flags (3-arg methods)
@pytest.mark.parametrize("t", [G, C, NB])
def test_flags_restrict_to_flagged_features(t):
    rng = np.random.default_rng(5)
    n, D = 25, 7
    if t == G:
        x = rng.normal(size=(n, D))
    elif t == C:
        x = rng.integers(1, 4, (n, D)).astype(np.int64)
        x[0, :] = 3
    else:
        x = rng.poisson(4.0, (n, D)).astype(np.int64)
    flags = np.array([1, 0, 1, 1, 0, 0, 1], dtype=np.uint8)
    o = _ctx([x], [t])
    o.set_flags(0, flags)
    cl = o.cluster(0)
    for i in range(1, 16):
        cl.add(i)
    sub = np.ascontiguousarray(x[:, flags.astype(bool)])
    if t == C:
        assert (sub.max(axis=0) == 3).all()
    o2 = _ctx([sub], [t])
    cl2 = o2.cluster(0)
    for i in range(1, 16):
        cl2.add(i)
    np.testing.assert_allclose(cl.logprob(20), cl2.logprob(20), rtol=1e-13)


# ------------------------------------------------------------------------- misc.jl:15-47
def test_calc_ess():
    rng = np.random.default_rng(6)
    lw = rng.normal(size=257) * 3
    w = np.exp(lw - lw.max())
    np.testing.assert_allclose(orc.calc_ess(lw), w.sum() ** 2 / (w ** 2).sum(), rtol=1e-13)
    assert orc.calc_ess(np.full(64, 1.0)) == 64.0


def _draw_partstar_py(lw, r, shuffle_u):
    """src/misc.jl:27-47 in plain Python (Fisher-Yates as Random.shuffle!: i = n..2, j = 1 + floor(u*i))."""
    P = len(lw)
    u = r / P
    pprob = np.cumsum(np.exp(lw - lw.max()))
    ps, i = [0] * P, 0
    for p in range(P):
        while i < P and pprob[p] / pprob[-1] >= u:
            u += 1 / P
            ps[i] = p + 1
            i += 1
    for pos in range(P, 1, -1):
        j = min(pos, 1 + int(math.floor(shuffle_u[pos - 1] * pos)))
        ps[pos - 1], ps[j - 1] = ps[j - 1], ps[pos - 1]
    ps[0] = 1
    return sorted(ps)


@pytest.mark.parametrize("P", [2, 5, 32, 333])
def test_draw_partstar(P):
    rng = np.random.default_rng(7 + P)
    for _ in range(20):
        lw = rng.normal(size=P) * 4
        r, su = rng.random(), rng.random(P)
        got = orc.draw_partstar(lw, r, su)
        assert list(got) == _draw_partstar_py(lw, r, su)
        assert got[0] == 1 and (np.diff(got) >= 0).all()  # reference particle pinned, sorted
    # equal weights: every particle is drawn once, then the one shuffled to position 1 is replaced
    # by the reference particle (misc.jl:43-45)
    got = orc.draw_partstar(np.zeros(P), 0.5, rng.random(P))
    assert got[0] == 1 and len(set(got.tolist())) >= P - 1


def test_philox_uniform_range_and_addressing():
    us = [orc.uniform(9, 1, kind, step, k, idx) for kind in range(5) for step in range(3)
          for k in range(2) for idx in range(4)]
    assert all(0.0 <= u < 1.0 for u in us)
    assert len(set(us)) == len(us)
    assert orc.uniform(9, 1, 0, 0, 0, 1) == orc.uniform(9, 1, 0, 0, 0, 1)
