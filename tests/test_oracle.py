"""CPU tests that pin the oracle (no GPU).

Stage 1 is pinned to the unmodified reference through tests/golden/ns_*.npz (pyNNGP/nngp.py:49-62 run
by tests/golden/make_golden.py).  Stages 2-3 are unpinned by the reference (nngp.py:73-96 are stubs)
and are anchored to the dense-GP identity and closed-form hand cases.
"""
import glob
import os

import numpy as np
import pytest

from oracle import nngp_oracle as orc
from pynngp_b200.synthetic import CONFIGS, synthetic

TIE_FREE = ["test_init_shape", "cfg1", "d2_m15", "d3_m30", "d1_m5", "d3_m32"]


@pytest.mark.parametrize("name", TIE_FREE)
def test_knn_c_oracle_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"ns_{name}.npz"))
    tab = orc.c_knn_ordered(g["coords"], int(g["m"]))
    assert tab.dtype == np.int32
    assert np.array_equal(tab, g["Ns"])  # bit-exact, order included


@pytest.mark.parametrize("name", ["test_init_shape", "d1_m5", "d3_m32"])
def test_knn_numpy_oracle_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"ns_{name}.npz"))
    Ns = orc.np_knn_ordered(g["coords"], int(g["m"]))
    assert Ns[0] == []
    assert np.array_equal(orc.ns_to_table(Ns, int(g["m"])), g["Ns"])


@pytest.mark.parametrize("name", ["test_init_shape", "cfg1", "d1_m5", "lattice"])
def test_sklearn_restatement_of_stage1_is_the_reference(golden_dir, name):
    """oracle.sk_reference_stage1 makes the reference's own scikit-learn calls (nngp.py:55-61), so it must
    reproduce the unmodified reference's tables exactly -- on the tied lattice too, where the KD-tree's
    order among equal distances is whatever scikit-learn does.  bench.py times this function as the
    reference's stage 1 on the GPU box's host."""
    g = np.load(os.path.join(golden_dir, f"ns_{name}.npz"))
    Ns = orc.sk_reference_stage1(g["coords"], int(g["m"]))
    assert Ns[0] == []
    assert np.array_equal(orc.ns_to_table(Ns, int(g["m"])), g["Ns"])


def test_knn_lattice_ties_distance_multiset(golden_dir):
    """On tied inputs the reference's KD-tree order is unspecified: the oracle's
    (d2, j) rule must give the same multiset of neighbour distances per row."""
    g = np.load(os.path.join(golden_dir, "ns_lattice.npz"))
    s, ref = g["coords"], g["Ns"]
    tab = orc.c_knn_ordered(s, int(g["m"]))
    for i in range(len(s)):
        a, b = tab[i][tab[i] >= 0], ref[i][ref[i] >= 0]
        assert len(a) == len(b)
        da = np.sort(orc.np_dist2(s[i], s[a])) if len(a) else np.zeros(0)
        db = np.sort(orc.np_dist2(s[i], s[b])) if len(b) else np.zeros(0)
        assert np.array_equal(da, db)
        # ascending (d2, j) inside the oracle's own rows
        if len(a) > 1:
            d = orc.np_dist2(s[i], s[a])
            assert all((d[k], a[k]) < (d[k + 1], a[k + 1]) for k in range(len(a) - 1))


def test_knn_duplicates_ties_by_index():
    s = np.array([[0.5, 0.5]] * 6 + [[0.25, 0.5]] * 3, dtype=np.float64)
    tab = orc.c_knn_ordered(s, 4)
    assert tab[0].tolist() == [-1, -1, -1, -1]
    assert tab[3].tolist() == [0, 1, 2, -1]
    assert tab[5].tolist() == [0, 1, 2, 3]
    assert tab[8].tolist() == [6, 7, 0, 1]
    assert np.array_equal(tab, orc.ns_to_table(orc.np_knn_ordered(s, 4), 4))


def test_knn_threads_and_ranges():
    s, _ = synthetic(700, 3, 5)
    full = orc.c_knn_ordered(s, 7)
    assert np.array_equal(full, orc.c_knn_ordered(s, 7, threads=4))
    part = orc.c_knn_ordered(s, 7, lo=100, hi=300)
    assert np.array_equal(part[100:300], full[100:300])
    assert (part[:100] == -1).all() and (part[300:] == -1).all()


@pytest.mark.parametrize("kernel_id", [0, 1, 2])
@pytest.mark.parametrize("D", [1, 2, 3])
def test_c_matches_numpy_per_location(kernel_id, D):
    s, y = synthetic(120, D, 40 + D)
    m = 9
    nbr = orc.c_knn_ordered(s, m)
    eps2 = np.linspace(0.0, 0.02, len(s))
    sig, phi, tau = 1.3, 5.0, 0.07
    npo = orc.NumpyNNGP(s, y, nbr, kernel_id, sig, phi, tau, eps2)
    CN, cc, cs = orc.c_cov_blocks(s, nbr, kernel_id, sig, phi, tau, eps2)
    B, F = orc.c_factors(s, y, nbr, kernel_id, sig, phi, tau, eps2)
    for i in [0, 1, 2, 5, 8, 9, 10, 57, 119]:
        p = min(i, m)
        np.testing.assert_allclose(CN[i, :p, :p], npo._CNs(i), rtol=1e-14)
        np.testing.assert_allclose(cc[i, :p], npo._Ccross(i), rtol=1e-14)
        assert cs[i] == npo._Cs(i)
        np.testing.assert_allclose(B[i, :p], npo._Bsi(i), rtol=1e-10, atol=1e-13)
        assert (B[i, p:] == 0).all()
        np.testing.assert_allclose(F[i], npo._Fsi(i), rtol=1e-12)
    slog, squad, bad = orc.c_loglik(s, y, nbr, kernel_id, sig, phi, tau, eps2)
    s2, q2 = npo.loglik_terms()
    assert bad == 0
    np.testing.assert_allclose([slog, squad], [s2, q2], rtol=1e-12)
    # threaded split == single pass (summation order only)
    t = orc.c_loglik(s, y, nbr, kernel_id, sig, phi, tau, eps2, threads=3)
    np.testing.assert_allclose(t[:2], [slog, squad], rtol=1e-13)


@pytest.mark.parametrize("kernel_id", [0, 1, 2])
def test_dense_gp_identity(kernel_id):
    """Known answer: with m = n-1 every location conditions on all predecessors and the NNGP density
    is the exact N(0, sigma2 rho + tau2 I) density."""
    n = 60  # ORACLE_MAX_M = 64 bounds m
    s, y = synthetic(n, 2, 77)
    nbr = orc.c_knn_ordered(s, n - 1)
    sig, phi, tau = 1.0, 6.0, 0.1
    slog, squad, bad = orc.c_loglik(s, y, nbr, kernel_id, sig, phi, tau)
    assert bad == 0
    ll = orc.loglik_from_terms(slog, squad, n)
    exact = orc.dense_gp_loglik(s, y, kernel_id, sig, phi, tau)
    assert abs(ll - exact) <= 1e-12 * abs(exact)


def test_dense_gp_identity_numpy_larger():
    n = 150
    s, y = synthetic(n, 3, 78)
    Ns = orc.np_knn_ordered(s, n - 1)
    nbr = orc.ns_to_table(Ns, n - 1)
    npo = orc.NumpyNNGP(s, y, nbr, 1, 0.8, 4.0, 0.2)
    slog, squad = npo.loglik_terms()
    exact = orc.dense_gp_loglik(s, y, 1, 0.8, 4.0, 0.2)
    assert abs(orc.loglik_from_terms(slog, squad, n) - exact) <= 1e-11 * abs(exact)


def test_hand_cases():
    s = np.array([[0.0, 0.0], [0.3, 0.4]])  # distance 0.5
    y = np.array([1.5, -0.5])
    nbr = np.array([[-1], [0]], dtype=np.int32)
    sig, phi, tau = 2.0, 3.0, 0.5
    B, F = orc.c_factors(s, y, nbr, 0, sig, phi, tau)
    # i = 0: no neighbours -> F = C_ii
    assert B[0, 0] == 0.0 and F[0] == sig + tau
    # p = 1: b = sigma2 rho(d) / (sigma2 + tau2)
    rho = np.exp(-phi * 0.5)
    np.testing.assert_allclose(B[1, 0], sig * rho / (sig + tau), rtol=1e-15)
    np.testing.assert_allclose(F[1], sig + tau - (sig * rho) ** 2 / (sig + tau), rtol=1e-15)
    slog, squad, bad = orc.c_loglik(s, y, nbr, 0, sig, phi, tau)
    r1 = y[1] - B[1, 0] * y[0]
    np.testing.assert_allclose(slog, np.log(F[0]) + np.log(F[1]), rtol=1e-15)
    np.testing.assert_allclose(squad, y[0] ** 2 / F[0] + r1 ** 2 / F[1], rtol=1e-15)


def test_non_spd_is_counted_not_nan():
    s = np.array([[0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [1.0, 1.0]])  # duplicates, no nugget
    y = np.ones(4)
    nbr = orc.c_knn_ordered(s, 3)
    slog, squad, bad = orc.c_loglik(s, y, nbr, 0, 1.0, 1.0, 0.0)
    assert bad >= 1 and np.isfinite(slog) and np.isfinite(squad)


def test_golden_files_present(golden_dir):
    assert len(glob.glob(os.path.join(golden_dir, "ns_*.npz"))) >= 7


def test_krige_known_answers():
    """m = n: nearest-neighbour kriging is exact GP conditioning -- mean and variance equal the dense
    formulas; a new site far from every reference site returns the prior (0, sigma2 + tau2)."""
    rng = np.random.default_rng(3)
    s = rng.random((40, 2))
    y = rng.standard_normal(40)
    tn = rng.random((5, 2))
    for kid in (0, 1, 2):
        mean, var, tab = orc.np_krige(s, y, tn, 40, kid, 1.2, 4.0, 0.05)
        d = np.sqrt(((s[:, None, :] - s[None, :, :]) ** 2).sum(-1))
        C = 1.2 * orc.np_corr(kid, 4.0 * d) + 0.05 * np.eye(40)
        c = 1.2 * orc.np_corr(kid, 4.0 * np.sqrt(((s[None, :, :] - tn[:, None, :]) ** 2).sum(-1)))
        np.testing.assert_allclose(mean, c @ np.linalg.solve(C, y), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(var, 1.25 - np.einsum("qi,qi->q", c, np.linalg.solve(C, c.T).T), rtol=1e-9)
        assert sorted(tab[0].tolist()) == list(range(40))
    mean, var, _ = orc.np_krige(s, y, np.array([[500.0, 500.0]]), 6, 0, 1.2, 4.0, 0.05)
    assert abs(mean[0]) < 1e-200 and var[0] == 1.25


# ---- the cell-grid search's algorithm (CPU model of knn_grid.cu): stopping never changes the result ----
def _grid_cases():
    rng = np.random.default_rng(42)
    lattice = np.stack(np.meshgrid(np.arange(24.0), np.arange(20.0)), -1).reshape(-1, 2)
    yield "uniform2d", rng.random((900, 2)), 15
    yield "uniform3d", rng.random((700, 3)), 30
    yield "uniform1d", rng.random((600, 1)), 7
    yield "lattice_ties", lattice[rng.permutation(len(lattice))] / 7.0, 8
    yield "duplicates", np.repeat(rng.random((60, 2)), 8, axis=0)[rng.permutation(480)], 12
    yield "all_same", np.full((300, 3), 0.25), 5
    yield "collapsed_dim", np.stack([rng.random(500), np.full(500, 2.0)], 1), 9
    yield "thin_dim", np.stack([rng.random(600), 1e-7 * rng.random(600), rng.random(600)], 1), 10
    yield "huge_offset", rng.random((500, 2)) * 1e-3 + 1e6, 10
    yield "negative_box", rng.random((500, 3)) * [5.0, 0.01, 300.0] - [2.5, 1e3, 150.0], 6
    yield "clusters", np.concatenate([0.5 + 0.01 * rng.standard_normal((400, 2)), rng.random((150, 2)),
                                      0.2 + 1e-4 * rng.standard_normal((150, 2))])[rng.permutation(700)], 15
    yield "m32", rng.random((400, 2)), 32
    yield "m1", rng.random((400, 3)), 1


@pytest.mark.parametrize("name,s,m", list(_grid_cases()), ids=[c[0] for c in _grid_cases()])
def test_grid_search_model_is_exact(name, s, m):
    from oracle import grid_knn_model as gm

    want = orc.c_knn_ordered(s, m)
    for lam in (0.15, 1.0, 5.0):
        st = {}
        got = gm.grid_knn_ordered(s, m, lam_scale=lam, brute_rows=128, stats=st)
        assert np.array_equal(got, want), (name, lam)
    # and the walk really prunes on well-spread data: far fewer candidates than the n^2/2 of brute force
    if name == "uniform2d":
        st = {}
        gm.grid_knn_ordered(s, m, 1.0, 128, stats=st)
        assert st["candidates"] < 0.5 * len(s) * len(s) / 2


def test_grid_search_model_random_sweep():
    """Seeded sweep over sizes, dimensions, m, cell sizes, anisotropic boxes and offsets."""
    from oracle import grid_knn_model as gm

    rng = np.random.default_rng(7)
    for _ in range(30):
        D = int(rng.integers(1, 4))
        n = int(rng.integers(200, 1200))
        m = int(rng.integers(1, 33))
        s = rng.random((n, D)) * rng.choice([1e-3, 1.0, 40.0], D) + rng.choice([0.0, -7.0, 1e5], D)
        if rng.random() < 0.3:  # coarse coordinates: many exact ties
            s = np.round(s, 2)
        want = orc.c_knn_ordered(s, m)
        got = gm.grid_knn_ordered(s, m, lam_scale=float(rng.choice([0.2, 1.0, 3.0])), brute_rows=128)
        assert np.array_equal(got, want), (D, n, m)


@pytest.mark.parametrize("kernel_id,D,m", [(0, 2, 5), (1, 2, 7), (2, 3, 4), (1, 1, 3)])
def test_loglik_equals_product_of_gaussian_conditionals_scipy(kernel_id, D, m):
    """Independent known answer for m < n - 1 (the dense-GP identity only covers m = n - 1): by definition the
    NNGP density is prod_i p(y_i | y_N(i)) under the parent GP, and each factor is a ratio of two multivariate
    normal densities, log N([y_N, y_i]; 0, C_joint) - log N(y_N; 0, C_N).  SciPy's multivariate_normal (an
    eigendecomposition, not a Cholesky/LDL^T elimination) evaluates both; the oracle's sum of
    log F_i + r_i^2 / F_i must give the same number.  Covers per-observation eps as well."""
    from scipy.stats import multivariate_normal

    n = 70
    s, y = synthetic(n, D, 50 + kernel_id)
    rng = np.random.default_rng(kernel_id)
    eps2 = rng.random(n) * 0.05
    sigma2, phi, tau2 = 1.3, 4.0, 0.09
    tab = orc.c_knn_ordered(s, m)
    d = np.sqrt(((s[:, None, :] - s[None, :, :]) ** 2).sum(-1))
    C = sigma2 * orc.np_corr(kernel_id, phi * d)
    C[np.diag_indices(n)] = sigma2 + tau2 + eps2
    want = 0.0
    for i in range(n):
        N = tab[i][tab[i] >= 0]
        idx = np.concatenate([N, [i]])
        want += multivariate_normal(np.zeros(len(idx)), C[np.ix_(idx, idx)], allow_singular=False).logpdf(y[idx])
        if len(N):
            want -= multivariate_normal(np.zeros(len(N)), C[np.ix_(N, N)], allow_singular=False).logpdf(y[N])
    slog, squad, bad = orc.c_loglik(s, y, tab, kernel_id, sigma2, phi, tau2, eps2=eps2)
    assert bad == 0
    np.testing.assert_allclose(orc.loglik_from_terms(slog, squad, n), want, rtol=1e-10)
    npo = orc.NumpyNNGP(s, y, tab, kernel_id, sigma2, phi, tau2, eps2=eps2)
    np.testing.assert_allclose(orc.loglik_from_terms(*npo.loglik_terms(), n), want, rtol=1e-10)


def test_oracle_matches_extended_precision_known_answers(golden_dir):
    """tests/golden/kat_cfg1.npz (made by make_kat.py): cfg1 on the unmodified reference's own neighbour sets,
    every number evaluated in 80-bit extended precision with textbook formulas that share no code with the
    oracle.  Both restatements (C and numpy) must reproduce the per-location C_N, c, C_ii, b_i, F_i, r_i of six
    rows and the two halves of Q.  A simultaneous drift of oracle and kernels would show up here."""
    k = np.load(os.path.join(golden_dir, "kat_cfg1.npz"))
    g = np.load(os.path.join(golden_dir, "ns_cfg1.npz"))
    c = CONFIGS["cfg1"]
    s, y = synthetic(c["n"], c["D"], c["seed"])
    tab, eps2, m = g["Ns"], k["eps2"], int(k["m"])
    assert np.array_equal(tab, orc.c_knn_ordered(s, m))  # the reference's sets == the oracle's
    prm = tuple(k["params"])
    for kid in (0, 1):
        slog, squad, bad = orc.c_loglik(s, y, tab, kid, *prm, eps2=eps2)
        assert bad == 0
        np.testing.assert_allclose([slog, squad], [k[f"k{kid}_sum_log_F"], k[f"k{kid}_sum_r2_over_F"]], rtol=1e-12)
        npo = orc.NumpyNNGP(s, y, tab, kid, *prm, eps2=eps2)
        np.testing.assert_allclose(npo.loglik_terms(), [k[f"k{kid}_sum_log_F"], k[f"k{kid}_sum_r2_over_F"]], rtol=1e-12)
        CN, cc, cs = orc.c_cov_blocks(s, tab, kid, *prm, eps2=eps2)
        B, F = orc.c_factors(s, y, tab, kid, *prm, eps2=eps2)
        for i in k["rows"]:
            np.testing.assert_allclose(CN[i], k[f"k{kid}_CN_{i}"], rtol=1e-14, atol=1e-300)
            np.testing.assert_allclose(cc[i], k[f"k{kid}_c_{i}"], rtol=1e-14, atol=1e-300)
            np.testing.assert_allclose(cs[i], k[f"k{kid}_Cii_{i}"], rtol=1e-15)
            np.testing.assert_allclose(B[i], k[f"k{kid}_b_{i}"], rtol=0, atol=1e-11)
            np.testing.assert_allclose(F[i], k[f"k{kid}_F_{i}"], rtol=1e-12)
            p = int((tab[i] >= 0).sum())
            np.testing.assert_allclose(npo._Bsi(i), k[f"k{kid}_b_{i}"][:p], rtol=0, atol=1e-11)
            np.testing.assert_allclose(npo._Fsi(i), k[f"k{kid}_F_{i}"], rtol=1e-12)
            r_i = y[i] - (B[i][:p] @ y[tab[i][:p]] if p else 0.0)
            np.testing.assert_allclose(r_i, k[f"k{kid}_r_{i}"], rtol=1e-10, atol=1e-13)
