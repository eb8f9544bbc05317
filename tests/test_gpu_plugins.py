"""GPU parity of the cluster-type plugin contract (calc_logprob / cluster_add! / calc_logmarginal,
src/datatypes/*.jl) and of feature selection (src/pmdi.jl:120-128, 354-370) through the C-ABI."""
import math

import numpy as np
import pytest
from scipy import special, stats

from helpers import C, G, NB, problem

pytestmark = pytest.mark.gpu
RTOL = 1e-5  # north_star tolerance; observed ~1e-13


def _data(t, n, D, seed):
    rng = np.random.default_rng(seed)
    if t == G:
        return rng.normal(1.0, 2.0, (n, D))
    if t == C:
        x = rng.integers(1, 5, (n, D)).astype(np.int64)
        x[0, :] = 4
        return x
    return rng.poisson(7.0, (n, D)).astype(np.int64)


@pytest.mark.parametrize("t", [G, C, NB])
@pytest.mark.parametrize("D", [1, 64, 300])
def test_cluster_eval_matches_oracle(t, D):
    from oracle import oracle as orc
    import pmdi_b200.capi as capi
    n = 50
    x = _data(t, n, D, 3 + D)
    o = orc.Oracle([x], [t], 3, 4)
    rng = np.random.default_rng(0)
    with capi.Context([x], [t], 3, 4) as ctx:
        for m in (0, 1, 17, 49):
            rows = rng.permutation(n)[:m] + 1
            cl = o.cluster(0)
            for r in rows:
                cl.add(int(r))
            lp, lm = ctx.cluster_eval(0, rows, obs_1based=n, logmarginal=True)
            np.testing.assert_allclose(lp, cl.logprob(n), rtol=RTOL)
            np.testing.assert_allclose(lm, cl.logmarginal(), rtol=RTOL, atol=1e-9)


def test_gaussian_predictive_is_student_t():
    """The identity the reference's own test asserts (test/runtests.jl:33-36), on the GPU."""
    import pmdi_b200.capi as capi
    rng = np.random.default_rng(1)
    n = 400
    x = rng.normal(3.0, 2.0, (n + 1, 1))
    with capi.Context([x], [G], 3, 4) as ctx:
        lp, _ = ctx.cluster_eval(0, np.arange(1, n + 1), obs_1based=n + 1)
    xs = x[:n, 0]
    xbar = xs.mean()
    beta = 0.5 + 0.5 * (((xs - xbar) ** 2).sum() + 0.001 * n * xbar ** 2 / (n + 0.001))
    lam = (0.5 + 0.5 * n) * (n + 0.001) / (beta * (n + 1.001))
    mu = xs.sum() / (n + 0.001)
    want = stats.t.logpdf((x[n, 0] - mu) * math.sqrt(lam), df=n + 1) + 0.5 * math.log(lam)
    np.testing.assert_allclose(lp, want, rtol=1.5e-8)


def test_categorical_and_negbinom_closed_forms():
    import pmdi_b200.capi as capi
    rng = np.random.default_rng(2)
    n, L = 300, 10
    x = rng.integers(1, L + 1, (n + 1, 1)).astype(np.int64)
    x[0, 0], x[n, 0] = L, 1
    with capi.Context([x], [C], 3, 4) as ctx:
        lp, _ = ctx.cluster_eval(0, np.arange(1, n + 1), obs_1based=n + 1)
    c1 = (x[:n, 0] == 1).sum()
    np.testing.assert_allclose(lp, math.log((c1 + 0.5) / (n + 0.5 * L)), rtol=1.5e-8)  # runtests.jl:52
    y = rng.poisson(9.0, (n + 1, 5)).astype(np.int64)
    with capi.Context([y], [NB], 3, 4) as ctx:
        lp, _ = ctx.cluster_eval(0, np.arange(1, n + 1), obs_1based=n + 1)
    S = y[:n].sum(axis=0)
    want = (special.betaln(n + 2, 1 + S + y[n]) - special.betaln(n + 1, 1 + S)).sum()
    np.testing.assert_allclose(lp, want, rtol=1e-9)


def test_feature_selection_matches_oracle():
    from oracle import oracle as orc
    import pmdi_b200.capi as capi
    pr = problem(sets=[(G, 130, 0), (C, 65, 3), (NB, 100, 0)], n=120, N=12, P=8, seed=4)
    o = orc.Oracle(pr["data"], pr["types"], pr["N"], pr["P"])
    rng = np.random.default_rng(5)
    with capi.Context(pr["data"], pr["types"], pr["N"], pr["P"]) as ctx:
        for k in range(pr["K"]):
            fn_ref, fn_got = o.feature_null(k), ctx.feature_null(k)
            np.testing.assert_allclose(fn_got, fn_ref, rtol=RTOL)
            labels = pr["s"][:, k]
            # Philox draws
            p_ref, f_ref = o.feature_select(k, labels, fn_ref, seed=6, it=2)
            p_got, f_got = ctx.feature_select(k, labels, fn_ref, seed=6, it=2)
            np.testing.assert_allclose(p_got, p_ref, rtol=RTOL, atol=1e-7)
            np.testing.assert_array_equal(f_got, f_ref)
            # taped draws
            tape = rng.random(pr["data"][k].shape[1])
            p_ref, f_ref = o.feature_select(k, labels, fn_ref, tape_f=tape)
            p_got, f_got = ctx.feature_select(k, labels, fn_ref, tape_f=tape)
            np.testing.assert_array_equal(f_got, f_ref)
