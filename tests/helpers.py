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


GOLDEN = ["g_iris_shape", "mixed_k3_tapes", "g_k2_rho_half", "nb_cat_k2"]


def load_golden(name):
    """A committed fixture (tests/golden/make_golden.py): inputs + oracle outputs of one sweep."""
    import os
    z = np.load(os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden", name + ".npz"))
    K = len(z["types"])
    pr = dict(data=[z[f"data{k}"] for k in range(K)], types=[int(t) for t in z["types"]],
              n=int(z["s_in"].shape[0]), N=int(z["N"]), P=int(z["P"]), K=K, n1=int(z["n1"]),
              order=z["order"], s=z["s_in"], Pi=z["Pi"], phi=z["phi"])
    tapes = None
    if "tape_alloc" in z.files:
        tapes = {t: z[f"tape_{t}"] for t in ("alloc", "resamp", "shuffle", "select")}
    return pr, tapes, z


def assert_matches_golden(got, z, rtol):
    np.testing.assert_array_equal(got["alloc"], z["alloc"])
    np.testing.assert_array_equal(got["anc"], z["anc"])
    assert got["p_star"] == int(z["p_star"])
    np.testing.assert_array_equal(got["s"], z["s_out"])
    np.testing.assert_array_equal(got["cluster_n"], z["cluster_n"])
    assert got["n_resamples"] == int(z["n_resamples"])
    np.testing.assert_allclose(got["lw"], z["lw"], rtol=rtol, atol=1e-9)
    np.testing.assert_allclose(got["logweight"], z["logweight"], rtol=rtol, atol=1e-9)
    np.testing.assert_allclose(got["lp"].sum(axis=3), z["lp_sum"], rtol=rtol, atol=1e-9)
    np.testing.assert_allclose(got["lp"].max(axis=3), z["lp_max"], rtol=rtol, atol=1e-9)
