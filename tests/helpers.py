"""Shared builders for the parity tests: seeded problems the oracle finishes in seconds."""
import numpy as np

import pmdi_b200 as pm
from pmdi_b200 import synth

G, C, NB = synth.GAUSSIAN, synth.CATEGORICAL, synth.NEGBINOM


def problem(sets, n, N, P, rho=0.25, seed=1, c_true=3):
    data, types, truth = synth.make_data(sets, n, c_true, seed)
    K = len(sets)
    hy = synth.make_hypers(K, N, n, seed)
    rng = np.random.default_rng(seed + 99)
    order = rng.permutation(n) + 1
    n1 = int(np.floor(rho * n))
    return dict(data=data, types=types, n=n, N=N, P=P, K=K, n1=n1, order=order,
                s=hy["s"], Pi=hy["Pi"], phi=hy["phi"])


def tapes_for(pr, seed=5):
    rng = np.random.default_rng(seed)
    steps = pr["n"] - pr["n1"] + 1
    return dict(alloc=rng.random((steps, pr["K"], pr["P"])), resamp=rng.random(steps),
                shuffle=rng.random((steps, pr["P"])), select=rng.random(1))
