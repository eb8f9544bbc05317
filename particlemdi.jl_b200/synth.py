"""Synthetic inputs of the BASELINE.json shapes.

Not part of the reference: the reference ships no data.  Shapes, seeds and generators follow
the measurement definition so that the GPU arm, the CPU baseline and the parity tests all see
the same matrices.
"""
from __future__ import annotations

import numpy as np

from .dataprep import gaussian_normalise

GAUSSIAN, CATEGORICAL, NEGBINOM = 0, 1, 2

# name -> (n, N, P, rho, [(type, D, levels)], true clusters)
CONFIGS = {
    "cfg1_iris": dict(n=150, N=10, P=32, rho=0.25, sets=[(GAUSSIAN, 4, 0)], c_true=3, idx=1),
    "cfg2_multiomics": dict(n=500, N=20, P=256, rho=0.25,
                            sets=[(GAUSSIAN, 2000, 0), (CATEGORICAL, 500, 3), (NEGBINOM, 1000, 0)],
                            c_true=5, idx=2),
    "cfg3_tcga": dict(n=1000, N=30, P=512, rho=0.25,
                      sets=[(GAUSSIAN, 2000, 0), (GAUSSIAN, 2000, 0), (GAUSSIAN, 200, 0),
                            (NEGBINOM, 500, 0), (CATEGORICAL, 1000, 3), (CATEGORICAL, 500, 2)],
                      c_true=8, idx=3),
    "cfg4_singlecell": dict(n=20000, N=50, P=1024, rho=0.25,
                            sets=[(GAUSSIAN, 2000, 0), (GAUSSIAN, 2000, 0)], c_true=12, idx=4),
    "cfg5_scaling": dict(n=2000, N=20, P=256, rho=0.25,
                         sets=[(GAUSSIAN, 2000, 0), (CATEGORICAL, 500, 3), (NEGBINOM, 1000, 0)],
                         c_true=5, idx=5),
}


def make_data(sets, n, c_true, seed, shared=0.8):
    """Returns (list of n x D matrices, list of type tags, true membership n x K)."""
    rng = np.random.default_rng(seed)
    base = rng.permutation(np.arange(n) % c_true)
    data, types, truth = [], [], []
    for (t, D, L) in sets:
        z = base.copy()
        own = rng.random(n) > shared
        z[own] = rng.integers(0, c_true, own.sum())
        informative = rng.random(D) < 0.25
        if t == GAUSSIAN:
            means = rng.normal(0.0, 2.0, (c_true, D)) * informative[None, :]
            x = means[z] + rng.normal(0.0, 1.0, (n, D))
            x = gaussian_normalise(x)
        elif t == CATEGORICAL:
            probs = np.full((c_true, D, L), 1.0 / L)
            dir_ = rng.dirichlet(np.full(L, 0.5), (c_true, D))
            probs[:, informative, :] = dir_[:, informative, :]
            cdf = np.cumsum(probs[z], axis=2)
            u = rng.random((n, D, 1))
            x = 1 + (u > cdf).sum(axis=2)
            x = np.minimum(x, L).astype(np.int64)
            x[0, :] = L  # every column attains its top level (nlevels = 0.5*max, categorical_cluster.jl:10)
        else:
            lam = np.full((c_true, D), 5.0)
            lam_inf = rng.uniform(1.0, 20.0, (c_true, D))
            lam[:, informative] = lam_inf[:, informative]
            g = rng.gamma(2.0, 0.5, (n, D))
            x = rng.poisson(lam[z] * g).astype(np.int64)
        data.append(x)
        types.append(t)
        truth.append(z)
    return data, types, np.stack(truth, axis=1)


def make_config(name, seed=None, **override):
    cfg = dict(CONFIGS[name])
    cfg.update(override)
    if seed is None:
        seed = 20260101 + cfg["idx"]
    data, types, truth = make_data(cfg["sets"], cfg["n"], cfg["c_true"], seed)
    cfg.update(data=data, types=types, truth=truth, seed=seed, name=name)
    return cfg


def make_hypers(K, N, n, seed):
    """Fixed hyper-parameters for sweep-only runs: gamma ~ Gamma(1/N,1)+eps, Phi ~ Gamma(1,0.2)
    (as src/pmdi.jl:60-61), initial allocations ~ Categorical(gamma) (src/pmdi.jl:63-66)."""
    rng = np.random.default_rng(seed + 7919)
    gamma = rng.gamma(1.0 / N, 1.0, (N, K)) + np.finfo(np.float64).eps
    Pi = gamma / gamma.sum(axis=0, keepdims=True)
    phi = rng.gamma(1.0, 0.2, max(1, K * (K - 1) // 2)) if K > 1 else np.zeros(1)
    s = np.stack([1 + rng.choice(N, size=n, p=Pi[:, k]) for k in range(K)], axis=1).astype(np.int64)
    return dict(gamma=gamma, Pi=Pi, phi=phi, s=s)
