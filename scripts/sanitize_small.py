"""A small sweep (all three built-in types, resampling, splits) for compute-sanitizer: every engine named in argv."""
import os, sys
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests"))
import numpy as np
import pmdi_b200  # noqa
from pmdi_b200 import capi
from helpers import C, G, NB, problem
for eng in sys.argv[1:] or ["spec"]:
    os.environ["PMDI_ENGINE"] = eng
    pr = problem(sets=[(G, 70, 0), (C, 65, 3), (NB, 40, 0)], n=48, N=6, P=24, seed=4)
    with capi.Context(pr["data"], pr["types"], pr["N"], pr["P"]) as ctx:
        s = pr["s"]
        for it in range(2):
            r = ctx.sweep(s, pr["order"], pr["n1"], pr["Pi"], pr["phi"], seed=11, it=it, logweight_init=float(it > 0))
            s = r["s"]
    print(eng, "engine", r["engine"], "resamples", r["n_resamples"], "rows", sum(r["rows_evaluated"]))
