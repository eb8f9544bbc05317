#!/usr/bin/env python
"""configs[4]: particle-count sweep 64 -> 8192 at K=3, n=2000, rho in {0.1, 0.25, 0.5} on one GPU.  Per point: chained
sweeps from the random initial allocation; kernel time of the first sweep and mean of the last two.
  python scripts/cfg5_grid.py > gpurun_out/r02_cfg5_grid.jsonl"""
import json, os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmdi_b200  # noqa
from pmdi_b200 import capi, synth
base = synth.make_config("cfg5_scaling")
K = len(base["sets"])
hy = synth.make_hypers(K, base["N"], base["n"], base["seed"])
n = base["n"]
sweeps = int(sys.argv[1]) if len(sys.argv) > 1 else 5
for rho in (0.1, 0.25, 0.5):
    n1 = int(np.floor(rho * n))
    steps = n - n1 + 1
    for P in (64, 128, 256, 512, 1024, 2048, 4096, 8192):
        rng = np.random.default_rng(1)
        with capi.Context(base["data"], base["types"], base["N"], P) as ctx:
            s, kms, rows = hy["s"], [], []
            for it in range(sweeps):
                r = ctx.sweep(s, rng.permutation(n) + 1, n1, hy["Pi"], hy["phi"], seed=9, it=it, logweight_init=float(it > 0))
                s = r["s"]; kms.append(r["sweep_kernel_ms"]); rows.append(sum(r["rows_evaluated"]))
        dense = steps * P * base["N"] * sum(d.shape[1] for d in base["data"])
        ms = float(np.mean(kms[-2:]))
        print(json.dumps(dict(rho=rho, P=P, steps=steps, first_sweep_ms=round(kms[0], 2), settled_ms=round(ms, 2),
                              us_per_step=round(1e3 * ms / steps, 2), dense_evals_per_s=dense / (ms * 1e-3),
                              distinct_clusters_per_sweep=int(np.mean(rows[-2:])), engine=r["engine"])), flush=True)
