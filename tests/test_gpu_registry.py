"""User-defined cluster types on the device (reference README.md:48-88): the reference's own
GaussianCluster restated as a USER type must reproduce the built-in type - allocations, ancestors and selected
particle bit for bit, log-probabilities to 1e-9 (the built-in hoists terms and takes logs of products; the user
type sums its per-feature terms) - through the sweep, the single-cluster plugin calls and feature selection; and a
type the reference does not have runs end to end through pmdi()."""
import numpy as np
import pytest

import user_types as ut
from helpers import G, NB, problem, tapes_for

pytestmark = pytest.mark.gpu


def test_user_gaussian_reproduces_the_builtin_in_a_sweep():
    import pmdi_b200.capi as capi
    tag = capi.register_cluster_type("MyGaussian", ut.USER_GAUSSIAN, "MyGaussian", capi.F64)
    pr = problem(sets=[(G, 130, 0), (NB, 70, 0), (G, 40, 0)], n=110, N=9, P=48, seed=6)
    tapes = tapes_for(pr)
    outs = []
    for types in (pr["types"], [tag, pr["types"][1], tag]):
        with capi.Context(pr["data"], types, pr["N"], pr["P"]) as ctx:
            outs.append(ctx.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"], seed=3, it=2, tapes=tapes, debug=True))
    a, b = outs
    assert b["engine"] == "pool"
    np.testing.assert_array_equal(a["alloc"], b["alloc"])
    np.testing.assert_array_equal(a["anc"], b["anc"])
    np.testing.assert_array_equal(a["s"], b["s"])
    np.testing.assert_array_equal(a["cluster_n"], b["cluster_n"])
    assert a["p_star"] == b["p_star"] and a["n_resamples"] == b["n_resamples"] and a["n_resamples"] > 0
    np.testing.assert_allclose(b["lp"], a["lp"], rtol=1e-9, atol=1e-9)
    np.testing.assert_allclose(b["lw"], a["lw"], rtol=1e-9, atol=1e-9)


def test_user_gaussian_plugin_calls_and_feature_selection():
    import pmdi_b200.capi as capi
    tag = capi.register_cluster_type("MyGaussian", ut.USER_GAUSSIAN, "MyGaussian", capi.F64)
    rng = np.random.default_rng(2)
    n, D, N = 60, 37, 5
    x = rng.normal(size=(n, D)) + 3.0 * (np.arange(n) % 3)[:, None]
    labels = 1 + np.arange(n) % 3
    res = []
    for t in (capi.GAUSSIAN, tag):
        with capi.Context([x], [t], N, 8) as ctx:
            lp, lm = ctx.cluster_eval(0, [3, 9, 12, 30, 31], obs_1based=7, logmarginal=True)
            fn = ctx.feature_null(0)
            prob, flags = ctx.feature_select(0, labels, fn, seed=1, it=1)
            res.append((lp, lm, fn, prob, flags))
    for u, v in zip(res[0][:4], res[1][:4]):
        np.testing.assert_allclose(v, u, rtol=1e-10, atol=1e-10)
    np.testing.assert_array_equal(res[0][4], res[1][4])


def test_a_type_the_reference_does_not_have_runs_through_pmdi(tmp_path):
    import pmdi_b200  # noqa: F401
    from pmdi_b200 import pmdi as host
    rng = np.random.default_rng(0)
    n = 90
    z = np.arange(n) % 3
    counts = rng.poisson(np.array([0.5, 8.0, 60.0])[z][:, None] * np.ones((n, 30))).astype(np.int64)
    poisson = host.UserCluster("MyPoisson", ut.USER_POISSON, integer_data=True)
    out = tmp_path / "o.csv"
    host.pmdi([counts], [poisson], 6, 32, 0.25, 30, str(out), seed=1)
    alloc = host.read_allocations(str(out), 1, n, burnin=10)
    # planted clusters recovered: same-cluster pairs co-clustered, others apart
    psm = host.posterior_similarity(alloc)[0]
    same = z[:, None] == z[None, :]
    assert psm[same].mean() > 0.85 and psm[~same].mean() < 0.1
