"""Why do bench.py's two CPU legs disagree?  Per-sweep time and calc_logprob calls of the oracle's
de-duplicated literal mode along a chain, for both legs' protocols.  (diagnostic, not a test)"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import bench
from oracle import oracle as orc

def chain(label, rng, seed, nsweeps):
    cfg = bench.make_workload("cfg2_multiomics")
    o = orc.Oracle(cfg["data"], cfg["types"], cfg["N"], cfg["P"])
    hy, n = cfg["hy"], cfg["n"]
    mode = orc.MODE_DEDUP | orc.MODE_LITERAL_NEWID
    s = hy["s"]
    rng = rng if rng is not None else cfg["rng"]
    print(label)
    for it in range(nsweeps):
        order = rng.permutation(n) + 1
        t0 = time.perf_counter()
        r = o.sweep(s, order, cfg["n1"], hy["Pi"], hy["phi"], mode=mode, seed=seed if seed is not None else cfg["seed"],
                    it=it, logweight_init=0.0 if it == 0 else 1.0)
        dt = time.perf_counter() - t0
        occ = [len(set(r["s"][:, k].tolist())) for k in range(cfg["K"])]
        print("  sweep %d: %.3f s  calc_logprob calls %d  resamples %d  occupied labels of s_out per dataset %s"
              % (it, dt, r["n_ops"], r["n_resamples"], occ))
        s = r["s"]

print("host cores:", os.cpu_count(), " load:", os.getloadavg())
chain("cpu_baseline_leg protocol (rng(5), seed 3)", np.random.default_rng(5), 3, 6)
chain("run_reference protocol (cfg rng, cfg seed)", None, None, 6)
