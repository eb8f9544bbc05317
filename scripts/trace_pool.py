"""Per-warp event trace of one CTA over one observation step of the pool engine (clock64 cycles).
  python scripts/trace_pool.py cfg2_multiomics [step] [cta]"""
import sys, os
ROOT = os.path.join(os.path.dirname(os.path.abspath(__file__)), '..')
sys.path.insert(0, ROOT)
name = sys.argv[1] if len(sys.argv) > 1 else "cfg2_multiomics"
os.environ["PMDI_TRACE_STEP"] = sys.argv[2] if len(sys.argv) > 2 else "200"
os.environ["PMDI_TRACE_CTA"] = sys.argv[3] if len(sys.argv) > 3 else "5"
os.environ["PMDI_TRACE_FILE"] = os.path.join(ROOT, "gpurun_out", f"trace_{name}.txt")
import numpy as np
import bench
from pmdi_b200 import capi
cfg = bench.make_workload(name)
hy = cfg["hy"]
ctx = capi.Context(cfg["data"], cfg["types"], cfg["N"], cfg["P"])
s = hy["s"]
for it in range(4):
    order = cfg["rng"].permutation(cfg["n"]) + 1
    r = ctx.sweep(s, order, cfg["n1"], hy["Pi"], hy["phi"], seed=1, it=it, logweight_init=float(it > 0))
    s = r["s"]
ev = [l.split() for l in open(os.environ["PMDI_TRACE_FILE"])]
ev = [(int(a), int(b), int(c)) for a, b, c in ev]
t0 = min(c for _, _, c in ev)
names = {40: "top", 41: "svc", 42: "P[", 43: "ld", 44: "exp", 45: "cum", 46: "lab", 47: "atom", 48: "]P", 49: "arrB2",
         50: "B2", 51: "res", 52: "ess", 53: "E[", 54: "meta", 55: "]E", 56: "arrB1"}
if os.environ.get("PMDI_ENGINE", "spec") == "spec":
    names = {40: "top", 41: "dec", 42: "P[", 43: "ld", 44: "exp", 45: "cum", 46: "lab", 47: "]P", 49: "C[", 48: "]C",
             51: "fix", 52: "eval", 56: "arr", 60: "f.b", 61: "f.ld", 62: "f.scan", 63: "e.in", 64: "e.blk", 65: "e.sync", 66: "e.fin"}
for w in range(16):
    row = [(tag, c - t0) for ww, tag, c in ev if ww == w]
    print("w%02d" % w, " ".join(f"{names.get(tag, tag)}@{c}" for tag, c in row))
print("kernel_ms", r["sweep_kernel_ms"], "us/step", 1e3 * r["sweep_kernel_ms"] / (cfg["n"] - cfg["n1"] + 1))
