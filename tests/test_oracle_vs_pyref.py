"""Cross-check of the C++ oracle against a second, independent restatement of the reference sweep in
plain Python (tests/pyref.py, which keeps the reference's own copy-on-write pool and caches,
src/__pmdi.jl:132-318).  Both are fed the same tapes; everything must agree BIT FOR BIT:
per-step log-probs and log-weights, allocations, ancestors, the selected particle, the emitted
allocations, the number of calc_logprob calls.  Three pairings:

  pyref(corrected cache key)  ==  oracle DEDUP            (same data structures)
  pyref(corrected cache key)  ==  oracle DENSE            (what the GPU implements)
  pyref(literal new_id)       ==  oracle DEDUP | LITERAL  (the reference as written)

Neither side is Julia (it cannot run here); this shows two independent readings of the source
agree, including in the literal mode that the CPU baseline times."""
import numpy as np
import pytest

import pyref
from helpers import C, G, NB, problem, tapes_for
from oracle import oracle as orc

CASES = {
    "gauss": dict(sets=[(G, 5, 0)], n=40, N=4, P=8),
    "cat": dict(sets=[(C, 6, 3)], n=36, N=4, P=6),
    "negbinom": dict(sets=[(NB, 5, 0)], n=36, N=4, P=6),
    "mixed_k3": dict(sets=[(G, 7, 0), (C, 5, 3), (NB, 6, 0)], n=48, N=5, P=12),
    "k2_rho_half": dict(sets=[(G, 4, 0), (G, 3, 0)], n=30, N=3, P=10, rho=0.5),
}


def _lists(pr):
    data = [d.tolist() for d in pr["data"]]
    return dict(data=data, types=pr["types"], N=pr["N"], P=pr["P"], s=pr["s"].tolist(),
                order_obs=[int(v) for v in pr["order"]], n1=pr["n1"], Pi=pr["Pi"].tolist(),
                phi=[float(v) for v in pr["phi"]])


def _compare(py, o, pr, with_ops):
    steps = pr["n"] - pr["n1"] + 1
    np.testing.assert_array_equal(np.array(py["alloc"]), o["alloc"])
    np.testing.assert_array_equal(np.array(py["anc"]), o["anc"])
    np.testing.assert_array_equal(np.array(py["s"]), o["s"])
    assert py["p_star"] == o["p_star"] and py["n_resamples"] == o["n_resamples"]
    np.testing.assert_array_equal(np.array(py["cluster_n"]), o["cluster_n"])
    # floating point: bit for bit
    np.testing.assert_array_equal(np.array(py["lw"]), o["lw"])
    np.testing.assert_array_equal(np.array(py["logweight"]), o["logweight"])
    np.testing.assert_array_equal(np.array(py["lp"]), o["lp"])
    assert np.array(py["lp"]).shape == (steps, pr["K"], pr["P"], pr["N"])
    if with_ops:
        assert py["n_ops"] == o["n_ops"]


@pytest.mark.parametrize("name", list(CASES))
@pytest.mark.parametrize("flagged", [False, True])
def test_two_restatements_agree(name, flagged):
    pr = problem(**CASES[name], seed=31)
    tapes = tapes_for(pr, seed=9)
    rng = np.random.default_rng(2)
    flags = [(rng.random(d.shape[1]) < 0.7).astype(np.uint8) if flagged else np.ones(d.shape[1], np.uint8)
             for d in pr["data"]]
    for f in flags:
        f[0] = 1
    o = orc.Oracle(pr["data"], pr["types"], pr["N"], pr["P"])
    for k, f in enumerate(flags):
        o.set_flags(k, f)
    kw = _lists(pr)
    tp = {k: v.tolist() for k, v in tapes.items()}
    fl = [[bool(v) for v in f] for f in flags]

    py = pyref.sweep(flags=fl, tapes=tp, lw_init=1.0, literal_new_id=False, **kw)
    for mode, ops in ((orc.MODE_DEDUP, True), (orc.MODE_DENSE, False)):
        ref = o.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"], mode=mode, tapes=tapes,
                      logweight_init=1.0, debug=True)
        _compare(py, ref, pr, ops)

    py_lit = pyref.sweep(flags=fl, tapes=tp, lw_init=1.0, literal_new_id=True, **kw)
    ref = o.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"],
                  mode=orc.MODE_DEDUP | orc.MODE_LITERAL_NEWID, tapes=tapes, logweight_init=1.0, debug=True)
    _compare(py_lit, ref, pr, True)


def test_the_cross_check_exercises_resampling_and_the_stale_cache():
    """The comparison above is only worth something if the hard paths run: resampling with pool
    renumbering, copy-on-write splits, and (literal mode) cache hits that change the outcome."""
    pr = problem(**CASES["mixed_k3"], seed=31)
    tapes = tapes_for(pr, seed=9)
    kw = _lists(pr)
    tp = {k: v.tolist() for k, v in tapes.items()}
    fl = [[True] * d.shape[1] for d in pr["data"]]
    a = pyref.sweep(flags=fl, tapes=tp, lw_init=1.0, literal_new_id=False, **kw)
    b = pyref.sweep(flags=fl, tapes=tp, lw_init=1.0, literal_new_id=True, **kw)
    assert a["n_resamples"] > 0
    dense_calls = (pr["n"] - pr["n1"] + 1) * pr["K"] * pr["P"] * pr["N"]
    assert a["n_ops"] < dense_calls                      # the pool really de-duplicates
    assert a["lw"] != b["lw"]                            # the literal cache key changes the weights (F4)
