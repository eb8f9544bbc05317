#!/usr/bin/env python
"""Per-sweep timing of a BASELINE configuration on one GPU: chained sweeps from the random initial
allocation, kernel time / distinct rows evaluated / rows referenced per sweep.
  python scripts/time_configs.py cfg4_singlecell 6 [P=...] [rho=...] [n=...]"""
import json
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmdi_b200  # noqa: E402,F401
from pmdi_b200 import capi, synth  # noqa: E402

name, sweeps = sys.argv[1], int(sys.argv[2])
over = {}
for a in sys.argv[3:]:
    k, v = a.split("=")
    over[k] = float(v) if k == "rho" else int(v)
cfg = synth.make_config(name, **over)
K = len(cfg["sets"])
hy = synth.make_hypers(K, cfg["N"], cfg["n"], cfg["seed"])
n1 = int(np.floor(cfg["rho"] * cfg["n"]))
rng = np.random.default_rng(1)
rows = []
with capi.Context(cfg["data"], cfg["types"], cfg["N"], cfg["P"]) as ctx:
    s = hy["s"]
    for it in range(sweeps):
        r = ctx.sweep(s, rng.permutation(cfg["n"]) + 1, n1, hy["Pi"], hy["phi"], seed=9, it=it,
                      logweight_init=float(it > 0), time_phases=(it == sweeps - 1) or bool(os.environ.get('PMDI_TIME_ALL')))
        s = r["s"]
        rows.append(dict(sweep=it, kernel_ms=round(r["sweep_kernel_ms"], 3), device_ms=round(r["device_ms"], 3),
                         rows_evaluated=sum(r["rows_evaluated"]), rows_referenced=sum(r["rows_referenced"]),
                         resamples=r["n_resamples"], engine=r["engine"],
                         **({'phase_ms': [round(v, 3) for v in r['phase_ms']]} if os.environ.get('PMDI_TIME_ALL') else {})))
        print(json.dumps(rows[-1]), flush=True)
    steps = cfg["n"] - n1 + 1
    print(json.dumps(dict(config=name, over=over, steps=steps, us_per_step=round(1e3 * rows[-1]["kernel_ms"] / steps, 2),
                          phase_ms=[round(v, 3) for v in r["phase_ms"]],
                          phase_ms_max=[round(v, 3) for v in r["phase_ms_max"]],
                          labels_occupied_in_pstar=[int(len(np.unique(s[:, k]))) for k in range(K)])))
