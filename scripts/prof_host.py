#!/usr/bin/env python
"""Host side of one pmdi() iteration without the GPU: hyper-parameter updates, label alignment from the
contingency tables, CSV row - timed per component at a configuration's N, K, n_obs.
  python scripts/prof_host.py 20 3 500 [iters]"""
import os, sys, time
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import pmdi_b200  # noqa
from pmdi_b200 import pmdi as P

N, K, n = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3])
iters = int(sys.argv[4]) if len(sys.argv) > 4 else 50
rng = np.random.default_rng(0)
npairs = K * (K - 1) // 2
M = np.full(K, 2.0)
gamma = rng.gamma(1.0 / N, 1.0, (N, K)) + P.EPS
phi = rng.gamma(1.0, 0.2, npairs) if K > 1 else np.zeros(1)
s = (1 + rng.integers(0, 4, size=(n, K))).astype(np.int64)
factorised = N ** K > 200_000
tables = P.FactorisedZ(N, K) if factorised else P.HyperTables(N, K)
if not factorised:
    tables.refresh(gamma)
Z = P.update_Z(phi, tables, gamma)
v = P.update_v(n, Z, rng)
T = {}
def tick(name, t0):
    T[name] = T.get(name, 0.0) + time.perf_counter() - t0
counts = agree = None
for it in range(iters):
    t0 = time.perf_counter(); order = rng.permutation(n) + 1; tick("permutation", t0)
    t0 = time.perf_counter(); P.update_M(M, gamma, K, N, rng); tick("update_M", t0)
    t0 = time.perf_counter()
    if not factorised: tables.refresh(gamma)
    tick("refresh", t0)
    t0 = time.perf_counter(); P.update_gamma(gamma, phi, v, M, s, tables, rng, counts_all=counts); tick("update_gamma", t0)
    t0 = time.perf_counter(); Pi = gamma / gamma.sum(axis=0, keepdims=True)
    if not factorised: tables.refresh(gamma)
    tick("refresh", t0)
    t0 = time.perf_counter()
    if K > 1: P.update_phi(phi, v, s, tables, rng, agree_all=agree, gamma=gamma)
    tick("update_phi", t0)
    t0 = time.perf_counter(); Z = P.update_Z(phi, tables, gamma); v = P.update_v(n, Z, rng); tick("update_Z_v", t0)
    # stand-in for the sweep's outputs
    s = (1 + rng.integers(0, 4, size=(n, K))).astype(np.int64)
    cont = np.zeros((max(npairs, 1), N, N), dtype=np.int64)
    i = 0
    for a in range(K - 1):
        for b in range(a + 1, K):
            np.add.at(cont[i], (s[:, b] - 1, s[:, a] - 1), 1); i += 1
    t0 = time.perf_counter()
    if K > 1: counts, agree = P.align_labels_tables(s, cont, phi, gamma, N, K, rng)
    tick("align_labels_tables", t0)
    t0 = time.perf_counter(); row = P.csv_row(M, phi, 1.0, s); tick("csv_row", t0)
tot = sum(T.values())
for k_, v_ in sorted(T.items(), key=lambda x: -x[1]):
    print(f"{k_:22s} {1e3 * v_ / iters:8.3f} ms/iter")
print(f"{'total':22s} {1e3 * tot / iters:8.3f} ms/iter")
