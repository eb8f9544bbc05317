#!/usr/bin/env python
"""Generates tests/golden/*.npz: seeded inputs + the oracle's outputs for small sweeps.

The reference (Julia) cannot run in this image and ships no golden vectors, so
these fixtures are produced by the CPU oracle (oracle/pmdi_oracle.cpp, dense mode), whose maths
is pinned against the reference's own test identities in tests/test_oracle_closed_forms.py.
They freeze the oracle's behaviour (any later change to it shows up as a diff) and give the GPU
parity tests a target that does not need the oracle at run time.

    python tests/golden/make_golden.py        # rewrites the fixtures
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tests"))

from helpers import C, G, NB, problem, tapes_for  # noqa: E402
from oracle import oracle as orc  # noqa: E402

CASES = {
    # name: (problem kwargs, use tapes, logweight_init)
    "g_iris_shape": (dict(sets=[(G, 4, 0)], n=150, N=10, P=32, seed=21), False, 0.0),
    "mixed_k3_tapes": (dict(sets=[(G, 40, 0), (C, 24, 3), (NB, 30, 0)], n=90, N=8, P=24, seed=22), True, 1.0),
    "g_k2_rho_half": (dict(sets=[(G, 70, 0), (G, 9, 0)], n=64, N=6, P=16, rho=0.5, seed=23), True, 1.0),
    "nb_cat_k2": (dict(sets=[(NB, 33, 0), (C, 17, 4)], n=70, N=7, P=40, rho=0.1, seed=24), False, 1.0),
}


def generate(name):
    kw, use_tapes, lw0 = CASES[name]
    pr = problem(**kw)
    tapes = tapes_for(pr, seed=100 + kw["seed"]) if use_tapes else None
    o = orc.Oracle(pr["data"], pr["types"], pr["N"], pr["P"])
    r = o.sweep(pr["s"], pr["order"], pr["n1"], pr["Pi"], pr["phi"], mode=orc.MODE_DENSE,
                logweight_init=lw0, seed=77, it=2, tapes=tapes, debug=True)
    out = dict(
        N=pr["N"], P=pr["P"], n1=pr["n1"], types=np.array(pr["types"]), order=pr["order"],
        s_in=pr["s"], Pi=pr["Pi"], phi=pr["phi"], lw0=lw0, seed=77, it=2,
        s_out=r["s"], p_star=r["p_star"], logweight=r["logweight"], alloc=r["alloc"].astype(np.uint8),
        anc=r["anc"].astype(np.int16), lw=r["lw"], cluster_n=r["cluster_n"].astype(np.int16),
        n_resamples=r["n_resamples"],
        # log-probs: the per-step maximum over labels and the sum are enough to pin them
        lp_sum=r["lp"].sum(axis=3), lp_max=r["lp"].max(axis=3),
    )
    for k, d in enumerate(pr["data"]):
        out[f"data{k}"] = d
    if tapes:
        for t, v in tapes.items():
            out[f"tape_{t}"] = v
    return out


if __name__ == "__main__":
    for name in CASES:
        path = os.path.join(HERE, name + ".npz")
        np.savez_compressed(path, **generate(name))
        print(name, os.path.getsize(path), "bytes")
