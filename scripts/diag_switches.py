"""Same-binary experiments (runtime switches only): kernel time of settled sweeps with the allocation
uniforms computed in the kernel (Philox) vs loaded from a tape.  (diagnostic, not a test)"""
import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from pmdi_b200 import capi

cfg = bench.make_workload(sys.argv[1] if len(sys.argv) > 1 else "cfg2_multiomics")
hy, n, K, P = cfg["hy"], cfg["n"], cfg["K"], cfg["P"]
steps = n - cfg["n1"] + 1
ctx = capi.Context(cfg["data"], cfg["types"], cfg["N"], P)
rng = np.random.default_rng(1)
s = hy["s"]
for it in range(8):                                   # settle the chain
    s = ctx.sweep(s, rng.permutation(n) + 1, cfg["n1"], hy["Pi"], hy["phi"], seed=1, it=it,
                  logweight_init=float(it > 0))["s"]
orders = [rng.permutation(n) + 1 for _ in range(12)]
tape = {"alloc": np.random.default_rng(2).random((steps, K, P))}
for label, tapes in (("philox", None), ("tape  ", tape), ("philox", None), ("tape  ", tape)):
    ms = [ctx.sweep(s, o, cfg["n1"], hy["Pi"], hy["phi"], seed=1, it=20 + i, logweight_init=1.0, tapes=tapes)["sweep_kernel_ms"]
          for i, o in enumerate(orders)]
    print("QB=%s uniforms=%s kernel_ms median %.3f  min %.3f" % (os.environ.get("PMDI_QB", "2"), label, np.median(ms), min(ms)))
ctx.close()
