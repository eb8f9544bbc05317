"""Host side of pmdi() (particlemdi.jl_b200/pmdi.py), no GPU: the reference's own tests of the
normalising constant (test/runtests.jl:57-108) and of label alignment (:111-134), the CSV layout
(src/pmdi.jl:147-158, consumed by consensus_map.jl:38-53), the asserts (:50-55)."""
import itertools
import math

import numpy as np
import pytest

import pmdi_b200  # noqa: F401
from pmdi_b200 import pmdi as host


@pytest.mark.parametrize("N,K", [(2, 1), (3, 2), (4, 3), (2, 5), (7, 2), (5, 4), (20, 1), (20, 2), (6, 5)])
def test_update_Z_against_brute_force(N, K):
    """test/runtests.jl:57-108 (the reference loops N = 2..20, K = 1..5; a spread of those here)."""
    rng = np.random.default_rng(N * 10 + K)
    gamma = rng.gamma(1.0 / N, 1.0, (N, K)) + 1e-6
    phi = rng.gamma(1.0, 5.0, K * (K - 1) // 2) if K > 1 else np.zeros(1)
    pairs = host.phi_lab(K)
    Z = 0.0
    for combo in itertools.product(range(N), repeat=K):
        tmp = math.prod(gamma[combo[k], k] for k in range(K))
        for i, (a, b) in enumerate(pairs):
            tmp *= 1.0 + phi[i] * (combo[a] == combo[b])
        Z += tmp
    tables = host.HyperTables(N, K)
    tables.refresh(gamma)
    np.testing.assert_allclose(host.update_Z(phi, tables), Z, rtol=1.5e-8)


def test_combination_table_layout():
    """c_combn[:, K-k+1] = div(0:N^K-1, N^(K-k)) % N + 1 (src/pmdi.jl:70-72)."""
    N, K = 3, 3
    t = host.HyperTables(N, K)
    idx = np.arange(N ** K)
    for k in range(1, K + 1):
        want = (idx // N ** (K - k)) % N
        np.testing.assert_array_equal(t.combn[:, K - k], want)
    assert len({tuple(r) for r in t.combn}) == N ** K


def test_tables_refuse_infeasible_sizes():
    with pytest.raises(MemoryError, match="F6"):
        host.HyperTables(30, 6)  # BASELINE config 3: 7.29e8 rows


def test_align_labels_reference_property():
    """test/runtests.jl:111-134: datasets agree up to a label permutation; whenever the labels are
    aligned the gamma rows are aligned too, and alignment is reached."""
    rng = np.random.default_rng(3)
    K, N, n = 4, 6, 3000
    s = np.empty((n, K), dtype=np.int64)
    s[:, 0] = rng.integers(1, N + 1, n)
    gamma = np.empty((N, K))
    gamma[:, 0] = rng.gamma(1.0 / N, 1.0, N) + 1e-9
    for k in range(1, K):
        shuf = rng.permutation(N) + 1          # new label of old label l is shuf[l-1]
        s[:, k] = shuf[s[:, 0] - 1]
        gamma[shuf - 1, k] = gamma[:, 0]
    phi = np.full(K * (K - 1) // 2, 10.0)
    aligned = False
    for _ in range(10):
        host.align_labels(s, phi, gamma, N, K, rng)
        a_s = bool((s[:, 1:] == s[:, :1]).all())
        a_g = bool((gamma[:, 1:] == gamma[:, :1]).all())
        assert a_s == a_g
        aligned = a_s
    assert aligned
    before = s.copy()
    host.align_labels(s, phi, gamma, N, 1, rng)  # K == 1: no-op (src/misc.jl:62)
    np.testing.assert_array_equal(s, before)


def test_csv_layout_round_trip(tmp_path):
    K, n = 3, 5
    names = ["a", "b", "c"]
    hdr = host.csv_header(K, n, names).split(",")
    assert hdr[:3] == ["MassParameter_1", "MassParameter_2", "MassParameter_3"]
    assert hdr[3:6] == ["phi_1_2", "phi_1_3", "phi_2_3"] and hdr[6] == "ll"
    assert hdr[7] == "a_n1" and hdr[7 + n] == "b_n1" and hdr[-1] == "c_n5"   # dataset-major
    assert len(hdr) == host.n_hyper_columns(K) + K * n
    assert host.csv_header(1, 2, ["x"]).split(",") == ["MassParameter_1", "phi_1_1", "ll", "x_n1", "x_n2"]
    assert host.n_hyper_columns(1) == 3                                       # consensus_map.jl:38
    s = np.arange(1, K * n + 1).reshape(n, K, order="F")
    row = host.csv_row([2.0, 2.5, 1e-5], [0.25, 0.5, 3.0], 0, s).split(",")
    assert row[0] == "2.0" and row[2] == "1.0e-5" and row[6] == "0.0"        # Float64-promoted
    assert row[7:] == [f"{i}.0" for i in range(1, K * n + 1)]                 # labels print as 3.0
    f = tmp_path / "o.csv"
    f.write_text(",".join(hdr) + "\n" + ",".join(row) + "\n" + ",".join(row) + "\n")
    alloc = host.read_allocations(str(f), K, n)
    assert alloc.shape == (2, n, K)
    np.testing.assert_array_equal(alloc[0], s)
    psm = host.posterior_similarity(alloc)
    assert psm.shape == (K, n, n) and (np.diagonal(psm, axis1=1, axis2=2) == 1.0).all()


@pytest.mark.parametrize("kw,msg", [
    (dict(dataTypes=[0]), "Number of datatypes"),
    (dict(dataNames=["a"]), "Number of data names"),
    (dict(rho=1.0), "must be between 0 and 1"),
    (dict(N=1), "Number of clusters"),
    (dict(N=99), "Number of clusters"),
    (dict(particles=1), "2 or more particles"),
])
def test_asserts_of_pmdi(kw, msg, tmp_path):
    """src/pmdi.jl:50-55, raised before anything touches a device."""
    args = dict(dataFiles=[np.zeros((20, 3)), np.zeros((20, 2))], dataTypes=[0, 0], N=4, particles=8,
                rho=0.25, iter=1, outputFile=str(tmp_path / "x.csv"))
    args.update(kw)
    with pytest.raises(AssertionError, match=msg):
        host.pmdi(**args)
    with pytest.raises(AssertionError, match="same number of observations"):
        host.pmdi([np.zeros((20, 3)), np.zeros((19, 2))], [0, 0], 4, 8, 0.25, 1, str(tmp_path / "y.csv"))


def test_hyper_updates_keep_state_valid():
    rng = np.random.default_rng(0)
    N, K, n = 4, 3, 60
    gamma = rng.gamma(1.0 / N, 1.0, (N, K)) + host.EPS
    phi = rng.gamma(1.0, 0.2, 3)
    M = np.full(K, 2.0)
    s = rng.integers(1, N + 1, (n, K))
    tables = host.HyperTables(N, K)
    for _ in range(5):
        tables.refresh(gamma)
        host.update_M(M, gamma, K, N, rng)
        host.update_gamma(gamma, phi, 1.0, M, s, tables, rng)
        tables.refresh(gamma)
        host.update_phi(phi, 1.0, s, tables, rng)
        Z = host.update_Z(phi, tables)
        v = host.update_v(n, Z, rng)
        assert (gamma > 0).all() and (phi >= 0).all() and (M > 0).all() and Z > 0 and v > 0
        assert np.isfinite(gamma).all() and np.isfinite(phi).all()


@pytest.mark.parametrize("N,K", [(3, 2), (4, 3), (3, 4), (2, 5)])
def test_factorised_sums_equal_the_table_sums(N, K):
    """FactorisedZ (no N^K tables) against the literal tables of src/pmdi.jl:69-92: Z (update_hypers.jl:29-39),
    the per-(n, k) sums update_gamma! takes (:75-81) and the per-pair sums update_Phi! takes (:101-107)."""
    rng = np.random.default_rng(N * 10 + K)
    gamma = rng.gamma(1.0, 1.0, (N, K)) + 0.1
    phi = rng.gamma(1.0, 0.5, max(1, K * (K - 1) // 2))
    t = host.HyperTables(N, K)
    t.refresh(gamma)
    f = host.FactorisedZ(N, K)
    norm = t.norm_terms(phi)
    np.testing.assert_allclose(f.Z(gamma, phi), norm.sum(), rtol=1e-12)
    for k in range(K):
        want = np.array([norm[t.combn[:, k] == n].sum() / gamma[n, k] for n in range(N)])
        np.testing.assert_allclose(f.A(gamma, phi, k), want, rtol=1e-11)
    for i in range(K * (K - 1) // 2):
        want = norm[t.phi_index[:, i]].sum() / (1.0 + phi[i])
        np.testing.assert_allclose(f.Q(gamma, phi, i), want, rtol=1e-10)


def test_factorised_handles_config3():
    """BASELINE config 3 (N = 30, K = 6: 7.29e8 table rows in the reference) is a DP over 64 subsets."""
    rng = np.random.default_rng(0)
    gamma = rng.gamma(1.0 / 30, 1.0, (30, 6)) + 1e-6
    phi = rng.gamma(1.0, 0.2, 15)
    f = host.FactorisedZ(30, 6)
    Z = f.Z(gamma, phi)
    assert np.isfinite(Z) and Z > 0
    # multilinearity: Z = sum_n gamma[n, k] * A_k[n] for every k
    for k in range(6):
        np.testing.assert_allclose((gamma[:, k] * f.A(gamma, phi, k)).sum(), Z, rtol=1e-10)


def test_factorised_updates_follow_the_table_updates():
    """The same random stream through update_gamma! / update_Phi! with the tables and with the factorised sums."""
    N, K, n = 4, 3, 40
    rng0 = np.random.default_rng(5)
    s = rng0.integers(1, N + 1, (n, K))
    g0 = rng0.gamma(1.0, 1.0, (N, K)) + 0.1
    p0 = rng0.gamma(1.0, 0.2, 3)
    M = np.full(K, 2.0)
    out = []
    for fact in (False, True):
        gamma, phi = g0.copy(), p0.copy()
        tables = host.FactorisedZ(N, K) if fact else host.HyperTables(N, K)
        rng = np.random.default_rng(9)
        for _ in range(3):
            if not fact:
                tables.refresh(gamma)
            host.update_gamma(gamma, phi, 0.7, M, s, tables, rng)
            if not fact:
                tables.refresh(gamma)
            host.update_phi(phi, 0.7, s, tables, rng, gamma=gamma)
        out.append((gamma, phi))
    np.testing.assert_allclose(out[0][0], out[1][0], rtol=1e-9)
    np.testing.assert_allclose(out[0][1], out[1][1], rtol=1e-9)


@pytest.mark.parametrize("K", [2, 3, 4])
def test_align_labels_from_tables_equals_align_labels(K):
    """align_labels! driven by contingency tables (what the sweep reduces on the device) makes the same
    proposals with the same random numbers as the pass over the observations (src/misc.jl:61-108)."""
    N, n = 6, 90
    rng0 = np.random.default_rng(K)
    base = rng0.integers(1, 4, n)
    s0 = np.stack([np.where(rng0.random(n) < 0.8, (base + k) % 3 + 1, rng0.integers(1, N + 1, n)) for k in range(K)], axis=1)
    phi = rng0.gamma(2.0, 2.0, K * (K - 1) // 2)
    g0 = rng0.gamma(1.0, 1.0, (N, K))
    pairs = host.phi_lab(K)
    cont = np.zeros((len(pairs), N, N), dtype=np.int64)
    for i, (a, b) in enumerate(pairs):
        np.add.at(cont[i], (s0[:, b] - 1, s0[:, a] - 1), 1)   # [pair][lb][la], the device's layout
    s1, g1 = s0.copy(), g0.copy()
    host.align_labels(s1, phi, g1, N, K, np.random.default_rng(3))
    s2, g2 = s0.copy(), g0.copy()
    counts, agree = host.align_labels_tables(s2, cont, phi, g2, N, K, np.random.default_rng(3))
    assert (s1 != s0).any()  # labels did move
    np.testing.assert_array_equal(s1, s2)
    np.testing.assert_array_equal(g1, g2)
    for k in range(K):
        np.testing.assert_array_equal(counts[:, k], np.bincount(s2[:, k] - 1, minlength=N))
    for i, (a, b) in enumerate(pairs):
        assert agree[i] == (s2[:, a] == s2[:, b]).sum()


def test_spelled_out_log_densities_are_scipy_bit_for_bit():
    """update_M! / update_Phi! (src/update_hypers.jl:5-26, 95-128) use the gamma and binomial log-densities; the
    host loop spells scipy's formulas out (the generic wrappers cost ~100 us a call).  Same bits required: the
    values decide Metropolis tests and mixture draws."""
    from scipy import stats
    rng = np.random.default_rng(11)
    for _ in range(300):
        x = rng.gamma(0.05, 1.0, 20) + host.EPS
        a = rng.uniform(0.01, 5.0)
        np.testing.assert_array_equal(stats.gamma.logpdf(x, a=a, scale=1.0), host._gamma_logpdf(x, a, 1.0))
        c = rng.uniform(0.01, 30.0)
        assert stats.gamma.logpdf(c, a=2.0, scale=0.25) == host._gamma_logpdf(c, 2.0, 0.25)
        n = int(rng.integers(0, 2500))
        j = np.arange(n + 1)
        np.testing.assert_array_equal(stats.binom.logpmf(j, n, 0.5), host._binom_logpmf(j, n, 0.5))


def test_row_lists_of_the_tables_are_the_masks():
    """update_gamma! / update_Phi! sum the normalising terms over the rows with c_k = n / c_a = c_b
    (src/update_hypers.jl:75-78, 101-104): the precomputed row lists select what the masks select, in order."""
    t = host.HyperTables(5, 3)
    for k in range(3):
        for n in range(5):
            np.testing.assert_array_equal(t.rows_of[k][n], np.flatnonzero(t.combn[:, k] == n))
    for i in range(3):
        np.testing.assert_array_equal(t.rows_pair[i], np.flatnonzero(t.phi_index[:, i]))
