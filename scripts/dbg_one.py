import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
import numpy as np
import pmdi_b200
from helpers import *
from oracle import oracle as orc
import pmdi_b200.capi as capi
np.set_printoptions(linewidth=220, precision=5)
pr = problem(sets=[(G,4,0),(G,6,0)], n=60, N=6, P=100, seed=3)
o = orc.Oracle(pr["data"], pr["types"], pr["N"], pr["P"])
ref = o.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"], seed=11, it=3, debug=True)
ctx = capi.Context(pr["data"], pr["types"], pr["N"], pr["P"])
got = ctx.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"], seed=11, it=3, debug=True)
d = np.abs(got["lp"] - ref["lp"])
for st in range(3):
    for k in range(2):
        bad = np.argwhere(d[st, k].max(axis=1) > 1e-6).ravel()
        print("step", st, "k", k, "bad particles", bad[:40], "n", len(bad))
        if len(bad):
            p = bad[0]
            print("  ref", ref["lp"][st, k, p]); print("  got", got["lp"][st, k, p])
    print(" alloc mism", np.argwhere(got["alloc"][st] != ref["alloc"][st])[:10].tolist())
