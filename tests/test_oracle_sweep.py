"""The oracle's sweep: the reference's structural invariants (test/runtests.jl:136-162) in both of
its data layouts, agreement of the de-duplicated (reference data structures) and dense (what the
GPU does) modes, and the committed golden fixtures."""
import numpy as np
import pytest

from helpers import C, G, GOLDEN, NB, assert_matches_golden, load_golden, problem, tapes_for
from oracle import oracle as orc


def _sweep(pr, mode, **kw):
    o = orc.Oracle(pr["data"], pr["types"], pr["N"], pr["P"])
    return o.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"], mode=mode, **kw)


def _invariants(pr, r):
    # runtests.jl:147: every observation is in exactly one cluster of every particle
    assert (r["cluster_n"].sum(axis=2) == pr["n"]).all()
    assert (r["s"] >= 1).all() and (r["s"] <= pr["N"]).all()


@pytest.mark.parametrize("P,iters", [(2, 1), (1024, 3)])
def test_reference_structural_invariants(P, iters):
    """test/runtests.jl:136-162: 3 Gaussian datasets 100 x 16, N = 10 (the reference runs 100
    iterations at P = 1024; 3 keep the CPU suite short)."""
    pr = problem(sets=[(G, 16, 0)] * 3, n=100, N=10, P=P, seed=8)
    s = pr["s"]
    rng = np.random.default_rng(1)
    for it in range(iters):
        order = rng.permutation(pr["n"]) + 1
        o = orc.Oracle(pr["data"], pr["types"], pr["N"], P)
        for mode in (orc.MODE_DEDUP | orc.MODE_LITERAL_NEWID, orc.MODE_DEDUP, orc.MODE_DENSE):
            r = o.sweep(s, order, pr["n1"], pr["Pi"], pr["phi"], mode=mode, seed=5, it=it,
                        logweight_init=0.0 if it == 0 else 1.0, debug=(P == 2))
            _invariants(pr, r)
            if P == 2:
                # conditional SMC: the reference particle follows s (src/pmdi.jl:251,262)
                steps_obs = order[pr["n1"] - 1:] - 1
                for k in range(pr["K"]):
                    np.testing.assert_array_equal(r["alloc"][:, k, 0], s[steps_obs, k])
        s = r["s"]


CASES = {
    "gauss": dict(sets=[(G, 12, 0)], n=80, N=7, P=24),
    "cat": dict(sets=[(C, 15, 3)], n=60, N=6, P=16),
    "negbinom": dict(sets=[(NB, 11, 0)], n=60, N=6, P=16),
    "mixed_k3": dict(sets=[(G, 20, 0), (C, 9, 3), (NB, 14, 0)], n=90, N=9, P=48),
}


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("use_tapes", [False, True])
def test_dedup_corrected_equals_dense(name, use_tapes):
    """the reference's copy-on-write pool with the cache key made correct is the
    dense evaluation, bit for bit."""
    pr = problem(**CASES[name], seed=12)
    tapes = tapes_for(pr) if use_tapes else None
    a = _sweep(pr, orc.MODE_DENSE, seed=3, it=1, tapes=tapes, debug=True, logweight_init=1.0)
    b = _sweep(pr, orc.MODE_DEDUP, seed=3, it=1, tapes=tapes, debug=True, logweight_init=1.0)
    for key in ("alloc", "anc", "s", "cluster_n", "lp", "lw", "logweight"):
        np.testing.assert_array_equal(a[key], b[key], err_msg=key)
    assert a["p_star"] == b["p_star"] and a["n_resamples"] == b["n_resamples"]
    assert b["n_ops"] <= a["n_ops"]  # the pool evaluates unique clusters only (src/__pmdi.jl:187)
    _invariants(pr, a)


def test_literal_mode_keeps_invariants_but_may_differ():
    pr = problem(**CASES["mixed_k3"], seed=13)
    r = _sweep(pr, orc.MODE_DEDUP | orc.MODE_LITERAL_NEWID, seed=3, it=1)
    _invariants(pr, r)


def test_sstar_compat_only_changes_emitted_allocations():
    """pmdi() does not permute sstar on resampling (src/pmdi.jl:321-324)."""
    pr = problem(**CASES["mixed_k3"], seed=14)
    a = _sweep(pr, orc.MODE_DENSE, seed=3, it=1, debug=True)
    b = _sweep(pr, orc.MODE_DENSE | orc.MODE_SSTAR_COMPAT, seed=3, it=1, debug=True)
    assert a["n_resamples"] > 0
    for key in ("alloc", "anc", "lw", "cluster_n"):
        np.testing.assert_array_equal(a[key], b[key])
    assert a["p_star"] == b["p_star"]


def test_ancestors_sorted_with_reference_pinned():
    pr = problem(**CASES["mixed_k3"], seed=15)
    r = _sweep(pr, orc.MODE_DENSE, seed=4, it=0, debug=True)
    ev = r["anc"][r["anc"][:, 0] > 0]
    assert len(ev) == r["n_resamples"] > 0
    assert (ev[:, 0] == 1).all() and (np.diff(ev, axis=1) >= 0).all()


def test_edge_shapes():
    # N = 2 (smallest allowed, src/pmdi.jl:54), P = 2 (smallest allowed, :55), n1 = 1 (empty prefix)
    pr = problem(sets=[(G, 3, 0), (C, 2, 2)], n=12, N=2, P=2, rho=0.1, seed=16)
    assert pr["n1"] == 1
    for mode in (orc.MODE_DENSE, orc.MODE_DEDUP):
        _invariants(pr, _sweep(pr, mode, seed=1, it=0))
    # rho close to 1: one observation step left
    pr = problem(sets=[(NB, 5, 0)], n=10, N=3, P=4, rho=0.999, seed=17)
    r = _sweep(pr, orc.MODE_DENSE, seed=1, it=0, debug=True)
    assert r["alloc"].shape[0] == 10 - pr["n1"] + 1
    _invariants(pr, r)


@pytest.mark.parametrize("name", GOLDEN)
@pytest.mark.parametrize("mode", [orc.MODE_DENSE, orc.MODE_DEDUP])
def test_oracle_reproduces_golden(name, mode):
    pr, tapes, z = load_golden(name)
    r = _sweep(pr, mode, seed=int(z["seed"]), it=int(z["it"]), tapes=tapes, debug=True,
               logweight_init=float(z["lw0"]))
    assert_matches_golden(r, z, rtol=0.0 if mode == orc.MODE_DENSE else 1e-15)
