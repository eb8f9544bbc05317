#!/usr/bin/env python
"""A sweep with and without the final particles' cluster sizes requested returns the same allocations; pmdi() (which
does not request them) runs."""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmdi_b200  # noqa
from pmdi_b200 import capi, synth, pmdi as P
cfg = synth.make_config("cfg1_iris")
K = len(cfg["sets"]); hy = synth.make_hypers(K, cfg["N"], cfg["n"], cfg["seed"])
n1 = int(np.floor(cfg["rho"] * cfg["n"]))
order = np.random.default_rng(1).permutation(cfg["n"]) + 1
with capi.Context(cfg["data"], cfg["types"], cfg["N"], cfg["P"]) as ctx:
    a = ctx.sweep(hy["s"], order, n1, hy["Pi"], hy["phi"], seed=3, it=1)
    b = ctx.sweep(hy["s"], order, n1, hy["Pi"], hy["phi"], seed=3, it=1, cluster_sizes=False)
    c = ctx.sweep(hy["s"], order, n1, hy["Pi"], hy["phi"], seed=3, it=1)
assert "cluster_n" not in b and np.array_equal(a["s"], b["s"]) and np.array_equal(a["logweight"], b["logweight"])
assert np.array_equal(a["cluster_n"], c["cluster_n"]) and np.array_equal(a["label_counts"], b["label_counts"])
with tempfile.TemporaryDirectory() as d:
    st = P.pmdi(cfg["data"], cfg["types"], cfg["N"], cfg["P"], cfg["rho"], 5, os.path.join(d, "o.csv"))
print("ok", st["iterations"], round(st["loop_s"], 4))
