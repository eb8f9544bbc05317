import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
os.environ.setdefault("PMDI_TRACE_STEP", "200"); os.environ.setdefault("PMDI_TRACE_CTA", "5")
os.environ["PMDI_TRACE_FILE"] = os.path.join(ROOT, "gpurun_out", "trace.txt")
import numpy as np
import bench
from pmdi_b200 import capi
cfg = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2_multiomics")
hy = cfg["hy"]
ctx = capi.Context(cfg["data"], cfg["types"], cfg["N"], cfg["P"])
s = hy["s"]
for it in range(3):
    order = cfg["rng"].permutation(cfg["n"]) + 1
    r = ctx.sweep(s, order, cfg["n1"], hy["Pi"], hy["phi"], seed=1, it=it, logweight_init=1.0)
    s = r["s"]
ev = [l.split() for l in open(os.environ["PMDI_TRACE_FILE"])]
ev = [(int(a), int(b), int(c)) for a, b, c in ev]
t0 = min(c for _, _, c in ev)
names = {30: "p_rc", 31: "p_max", 32: "p_cum", 33: "p_log", 34: "p_lab", 9: "res_start", 10: "res_end", 20: "E", 21: "ld0", 22: "prod", 23: "log", 24: "red", 29: "X", 1: "step_top", 2: "queue_start", 3: "item_done", 4: "prop_start", 5: "prop_end", 6: "leave_list", 7: "cta_arrive", 8: "after_ess"}
for w in range(16):
    row = [(tag, c - t0) for ww, tag, c in ev if ww == w]
    out = []
    for tag, c in row:
        if tag & 0x100:
            out.append(f"[k{(tag >> 12) & 7}{chr(70) if tag & 0x800 else chr(112)} r{(tag >> 5) & 7}q{tag & 31}@{c}")
        else:
            out.append(f"{names.get(tag, tag)}@{c}")
    print("w%02d" % w, " ".join(out))
print("kernel_ms", r["sweep_kernel_ms"])
