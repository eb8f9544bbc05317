import sys, os
sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..')); sys.path.insert(0, os.path.join(os.path.dirname(__file__), '..', 'tests'))
import numpy as np
import pmdi_b200
from helpers import *
from oracle import oracle as orc
import pmdi_b200.capi as capi
np.set_printoptions(linewidth=200, precision=6)
case = dict(sets=[(G, 4, 0)], n=60, N=6, P=8)
pr = problem(**case, seed=3)
o = orc.Oracle(pr["data"], pr["types"], pr["N"], pr["P"])
ctx = capi.Context(pr["data"], pr["types"], pr["N"], pr["P"])
ref = o.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"], seed=11, it=3, debug=True)
got = ctx.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"], seed=11, it=3, debug=True, time_phases=True)
print("phase_ms", got["phase_ms"], got["phase_ms_max"], "kernel_ms", got["sweep_kernel_ms"])
steps = ref["lp"].shape[0]
for st in range(min(steps, 4)):
    d = np.abs(got["lp"][st] - ref["lp"][st]).max()
    print("step", st, "max|dlp|", d, "alloc ref", ref["alloc"][st].ravel(), "got", got["alloc"][st].ravel())
    print(" ref lp p0", ref["lp"][st, 0, 0]); print(" got lp p0", got["lp"][st, 0, 0])
    print(" ref lp p1", ref["lp"][st, 0, 1]); print(" got lp p1", got["lp"][st, 0, 1])
    print(" ref lw", ref["lw"][st]); print(" got lw", got["lw"][st])
    print(" anc ref", ref["anc"][st], "got", got["anc"][st])
    for p in range(1, 3):
        print("  u", orc.uniform(11, 3, 0, st, 0, p), capi.uniform(11, 3, 0, st, 0, p))
