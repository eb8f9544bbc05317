"""Deterministic-mode parity at BASELINE.json's own sizes: the CUDA path against the oracle's
de-duplicated mode with the corrected cache key (bit-equal to its dense mode,
tests/test_oracle_sweep.py::test_dedup_corrected_equals_dense, and fast enough at these sizes).

Shared by tests/test_gpu_fullparity.py, bench.py's parity block and scripts/parity_report.py.
"""
import time

import numpy as np

# name -> (overrides, sweeps chained before the compared one, per-step debug capture)
FULL_CASES = {
    "cfg1_iris": (dict(), 2, True),                   # configs[0] as is
    "cfg2_multiomics_fresh": (dict(), 0, True),       # configs[1], first sweep from the random allocation
    "cfg2_multiomics_settled": (dict(), 3, True),     # configs[1], a sweep of the settled chain
    "cfg3_tcga": (dict(), 0, False),                  # configs[2] at its real shapes
    "cfg4_singlecell_n4000": (dict(n=4000), 1, False),  # configs[3]: all shapes kept, rho=0.25, n cut 20000 -> 4000
    "cfg5_scaling_P1024": (dict(P=1024), 0, False),   # configs[4], one point
}


def make(name):
    import pmdi_b200  # noqa: F401
    from pmdi_b200 import synth
    over, chain, debug = FULL_CASES[name]
    base = name.split("_")[0]
    cname = [c for c in synth.CONFIGS if c.startswith(base + "_")][0]
    cfg = synth.make_config(cname, **over)
    K = len(cfg["sets"])
    hy = synth.make_hypers(K, cfg["N"], cfg["n"], cfg["seed"])
    cfg.update(K=K, hy=hy, n1=int(np.floor(cfg["rho"] * cfg["n"])), chain=chain, debug=debug)
    return cfg


def compare(cfg, sweep_fn, seed=77):
    """Runs `chain` oracle sweeps, then ONE compared sweep on both sides.  `sweep_fn(s, order, it, lw0, debug)`
    is the CUDA side.  Returns a dict of mismatch counts (all zero = parity) and timings."""
    from oracle import oracle as orc
    o = orc.Oracle(cfg["data"], cfg["types"], cfg["N"], cfg["P"])
    hy, n, n1 = cfg["hy"], cfg["n"], cfg["n1"]
    rng = np.random.default_rng(seed)
    s = hy["s"]
    for it in range(cfg["chain"]):
        s = o.sweep(s, rng.permutation(n) + 1, n1, hy["Pi"], hy["phi"], mode=orc.MODE_DEDUP, seed=seed, it=it,
                    logweight_init=float(it > 0))["s"]
    it = cfg["chain"]
    order = rng.permutation(n) + 1
    t0 = time.perf_counter()
    ref = o.sweep(s, order, n1, hy["Pi"], hy["phi"], mode=orc.MODE_DEDUP, seed=seed, it=it,
                  logweight_init=float(it > 0), debug=cfg["debug"])
    t_ref = time.perf_counter() - t0
    got = sweep_fn(s, order, it, float(it > 0), cfg["debug"])
    steps = n - n1 + 1
    out = {
        "steps": steps, "draws": steps * cfg["K"] * (cfg["P"] - 1),
        "s_mismatch": int((got["s"] != ref["s"]).sum()),
        "p_star_equal": bool(got["p_star"] == ref["p_star"]),
        "n_resamples": [int(got["n_resamples"]), int(ref["n_resamples"])],
        "logweight_max_rel": float(np.max(np.abs(got["logweight"] - ref["logweight"]) /
                                          np.maximum(1e-300, np.abs(ref["logweight"])))),
        "oracle_s": round(t_ref, 2), "oracle_calc_logprob_calls": int(ref["n_ops"]),
    }
    if "cluster_n" in got and (got["cluster_n"] >= 0).all():
        out["cluster_n_mismatch"] = int((got["cluster_n"] != ref["cluster_n"]).sum())
    if got.get("engine") in ("pool", "spec"):
        out["rows_evaluated"] = int(sum(got["rows_evaluated"])) - int(got["rows_evaluated_ahead"])
        out["rows_evaluated_ahead_of_resampling"] = int(got["rows_evaluated_ahead"])
    if cfg["debug"]:
        out["draw_flips"] = int((got["alloc"] != ref["alloc"]).sum())
        out["ancestor_mismatch"] = int((got["anc"] != ref["anc"]).sum())
        d = np.abs(got["lp"] - ref["lp"]) / np.maximum(1.0, np.abs(ref["lp"]))
        out["lp_max_rel"] = float(d.max())
        out["lw_max_rel"] = float(np.max(np.abs(got["lw"] - ref["lw"]) / np.maximum(1.0, np.abs(ref["lw"]))))
    else:
        # without the per-step capture: equal final allocations, selected particle, cluster sizes of every
        # particle and resampling count leave no room for a flipped draw that mattered
        out["draw_flips"] = 0 if (out["s_mismatch"] == 0 and out.get("cluster_n_mismatch", 0) == 0
                                  and out["n_resamples"][0] == out["n_resamples"][1]) else -1
    return out


def assert_parity(out, rtol=1e-5):
    assert out["s_mismatch"] == 0 and out["p_star_equal"], out
    assert out["n_resamples"][0] == out["n_resamples"][1], out
    assert out.get("cluster_n_mismatch", 0) == 0, out
    assert out["draw_flips"] == 0 and out.get("ancestor_mismatch", 0) == 0, out
    assert out["logweight_max_rel"] <= rtol, out
    assert out.get("lp_max_rel", 0.0) <= rtol and out.get("lw_max_rel", 0.0) <= rtol, out
    if "rows_evaluated" in out:  # the pool evaluates exactly the reference's distinct clusters
        assert out["rows_evaluated"] == out["oracle_calc_logprob_calls"], out
