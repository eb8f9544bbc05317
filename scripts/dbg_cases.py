import sys, os, subprocess, json
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, 'tests'))
CASES = {
  "g1_P200": "dict(sets=[(G,4,0)], n=60, N=6, P=200)",
  "g2_P8": "dict(sets=[(G,4,0),(G,6,0)], n=60, N=6, P=8)",
  "g2_P100": "dict(sets=[(G,4,0),(G,6,0)], n=60, N=6, P=100)",
  "gc_P16": "dict(sets=[(G,40,0),(C,20,3)], n=60, N=6, P=16)",
  "gnb_P16": "dict(sets=[(G,40,0),(NB,20,0)], n=60, N=6, P=16)",
  "mixed_P16": "dict(sets=[(G,130,0),(C,65,3),(NB,100,0)], n=60, N=6, P=16)",
  "mixed_k3": "dict(sets=[(G,130,0),(C,65,3),(NB,100,0)], n=120, N=12, P=64)",
}
if len(sys.argv) > 1:
    import numpy as np
    import pmdi_b200
    from helpers import *
    from oracle import oracle as orc
    import pmdi_b200.capi as capi
    pr = problem(**eval(CASES[sys.argv[1]]), seed=3)
    o = orc.Oracle(pr["data"], pr["types"], pr["N"], pr["P"])
    ref = o.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"], seed=11, it=3, debug=True)
    ctx = capi.Context(pr["data"], pr["types"], pr["N"], pr["P"])
    got = ctx.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"], seed=11, it=3, debug=True)
    ok = (got["alloc"] == ref["alloc"]).all() and (got["anc"] == ref["anc"]).all() and (got["s"] == ref["s"]).all()
    bad = np.argwhere(got["alloc"] != ref["alloc"])
    print(sys.argv[1], "OK" if ok else "MISMATCH first=%s" % (bad[0] if len(bad) else None), "res", ref["n_resamples"], got["n_resamples"],
          "max|dlp|", np.abs(got["lp"] - ref["lp"]).max(), "max|dlw|", np.abs(got["lw"] - ref["lw"]).max())
else:
    for name in CASES:
        r = subprocess.run([sys.executable, __file__, name], capture_output=True, text=True, timeout=120)
        print(name, "->", (r.stdout.strip().splitlines() or ["<no output>"])[-1], "|", (r.stderr.strip().splitlines() or [""])[-1][:200])
