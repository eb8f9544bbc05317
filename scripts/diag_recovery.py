"""Diagnostic: does pmdi() recover planted clusters, and how fast?  (not a test)"""
import os, sys, tempfile
import numpy as np
sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
import pmdi_b200  # noqa
from pmdi_b200 import pmdi as host, synth
from scipy.special import comb

def ari(a, b):
    ct = np.zeros((a.max() + 1, b.max() + 1))
    for i, j in zip(a, b): ct[i, j] += 1
    s = comb(ct, 2).sum(); sa, sb = comb(ct.sum(1), 2).sum(), comb(ct.sum(0), 2).sum()
    e = sa * sb / comb(len(a), 2)
    return (s - e) / (0.5 * (sa + sb) - e)

def run(name, data, truth, types, N, P, iters, seed):
    K = len(data); n = data[0].shape[0]
    with tempfile.TemporaryDirectory() as d:
        out = os.path.join(d, "o.csv")
        host.pmdi(data, types, N, P, 0.25, iters, out, seed=seed)
        alloc = host.read_allocations(out, K, n)
    pts = [0, 1, 2, 5, 10, 20, 40, 80, 120, 160, 200]
    print(name, "seed", seed)
    for k in range(K):
        print("  k=%d ARI by iteration:" % k, " ".join("%d:%.2f" % (i, ari(alloc[i][:, k], truth[:, k])) for i in pts if i < len(alloc)),
              " #labels at end:", len(set(alloc[-1][:, k].tolist())))

n = 90
sets = [(synth.GAUSSIAN, 30, 0), (synth.CATEGORICAL, 20, 3), (synth.NEGBINOM, 25, 0)]
data, types, truth = synth.make_data(sets, n, 3, 11, shared=1.0)
for seed in (5, 6):
    run("synth (25% informative features)", data, truth, types, 6, 32, 200, seed)
# trivially separable Gaussian: every feature informative, means 0 / 6 / 12 in unit noise
rng = np.random.default_rng(0)
z = np.arange(n) % 3
x = rng.normal(0, 1, (n, 10)) + 6.0 * z[:, None]
run("trivially separable Gaussian, K=1", [x], z[:, None], [0], 6, 32, 60, 1)
run("trivially separable Gaussian x2, K=2", [x, x + rng.normal(0, 0.1, x.shape)], np.stack([z, z], 1), [0, 0], 6, 32, 60, 1)

# trivially separable count data: cluster rates 1 / 30 / 900 (geometric-like counts need orders of magnitude)
rate = np.array([1.0, 30.0, 900.0])[z]
xnb = rng.poisson(rate[:, None] * np.ones((n, 40))).astype(np.int64)
run("trivially separable NegBinom, K=1", [xnb], z[:, None], [2], 6, 32, 60, 1)
# trivially separable categorical: level = cluster + 1 with 5 % noise, 3 levels, 30 features
xc = np.where(rng.random((n, 30)) < 0.05, rng.integers(1, 4, (n, 30)), (z + 1)[:, None]).astype(np.int64)
xc[0, :] = 3
run("trivially separable Categorical, K=1", [xc], z[:, None], [1], 6, 32, 60, 1)
run("separable Gaussian + NegBinom + Categorical, K=3", [x, xnb, xc], np.stack([z, z, z], 1), [0, 2, 1], 6, 32, 60, 1)
