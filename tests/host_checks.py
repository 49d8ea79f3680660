"""Assertions on the drop-in class shared by the CPU suite (oracle-backed FakeEngine monkeypatched in,
tests/test_host_logic.py) and the GPU suite (the real engine, tests/test_gpu_parity.py): the same checks
run against both, so the host logic is covered without a GPU and the kernels are covered with one.

Reference sets other than T are pinned to the reference's own calls: ``Nt[i]`` must equal
``KDTree(s).query(t_i.reshape(1, -1), m)`` (nngp.py:68-71) and ``ws`` must equal
``KNeighborsRegressor(5).fit(t, y).predict(s)`` (nngp.py:45-47), both evaluated with scikit-learn here;
``Ns`` on s follows nngp.py:49-62 through the oracle (itself pinned to the reference's golden tables).
"""
import numpy as np
import pytest

from oracle import nngp_oracle as orc

RTOL = 1e-10


def _expected_layout(obj, t, y, eps2=None):
    """Engine rows for ('subset', nRef): [s ; T - S], T - S in the order of t."""
    choice = obj._choice
    rest = np.setdiff1d(np.arange(len(t)), choice)
    rows = np.concatenate([choice, rest])
    return rows, rest


def check_subset(make, t, y, n_ref, m, kernel_id, prm, seed=7, eps=0.0):
    """make(t, y, eps, refType, m, seed=...) -> NNGP.  Returns the object for further checks."""
    from sklearn.neighbors import KDTree, KNeighborsRegressor

    obj = make(t, y, eps, ("subset", n_ref), m, seed=seed)
    n, D = t.shape
    # s: a seeded draw without replacement, reproducible
    want_choice = np.random.default_rng(seed).choice(n, size=n_ref, replace=False)
    assert np.array_equal(obj._choice, want_choice) and len(set(want_choice.tolist())) == n_ref
    assert np.array_equal(obj.s, t[want_choice])
    assert np.array_equal(obj.wt, y) and obj.wt is not y
    # Ns: ordered neighbours within s (nngp.py:49-62)
    s_tab = orc.c_knn_ordered(obj.s, m)
    assert len(obj.Ns) == n_ref and obj.Ns[0] == []
    for i in (1, 2, m, n_ref // 2, n_ref - 1):
        row = s_tab[i]
        assert np.array_equal(obj.Ns[i], row[row >= 0].astype(np.int64))
        assert obj.Ns[i].dtype == np.int64 and i not in obj.Ns[i]
    # Nt: what the reference's call returns (nngp.py:68-71)
    tree = KDTree(obj.s)
    assert len(obj.Nt) == n and obj.Nt is not obj.Ns
    for i in (0, 1, int(want_choice[0]), n // 2, n - 1):
        dist, ind = tree.query(t[i].reshape(1, -1), m)
        got_d, got_i = obj.Nt[i]
        assert got_i.shape == (1, m) and got_i.dtype == np.int64
        assert np.array_equal(got_i, ind)
        np.testing.assert_allclose(got_d, dist, rtol=1e-14, atol=0)
    all_ind = tree.query(t, m)[1]
    assert np.array_equal(obj._nt_table, all_ind)
    assert obj.Nt[int(want_choice[3])][0][0, 0] == 0.0  # a member of S is its own nearest reference site
    # ws: the reference's regressor evaluated at s (nngp.py:45-47)
    want_ws = KNeighborsRegressor(n_neighbors=5, weights="uniform").fit(t, y).predict(obj.s)
    np.testing.assert_allclose(obj.ws, want_ws, rtol=1e-13, atol=1e-15)
    # engine layout and table
    rows, rest = _expected_layout(obj, t, y)
    assert np.array_equal(obj._rows, rows)
    want_tab = np.concatenate([s_tab, all_ind[rest].astype(np.int32)])
    assert np.array_equal(obj._table, want_tab)
    assert (want_tab[n_ref:] < n_ref).all()  # T - S conditions on reference sites only
    # likelihood = the NNGP density of the permuted observations under that table
    eps2 = None
    if np.ndim(eps) or eps:
        eps2 = (np.broadcast_to(np.asarray(eps, dtype=np.float64), y.shape) ** 2)[rows]
    slog, squad, bad = orc.c_loglik(t[rows], y[rows], want_tab, kernel_id, *prm, eps2=eps2)
    assert bad == 0
    got = obj.loglik_terms(*prm)
    np.testing.assert_allclose(got, (slog, squad), rtol=RTOL)
    np.testing.assert_allclose(obj.loglik(*prm), orc.loglik_from_terms(slog, squad, n), rtol=RTOL)
    # accessors address engine rows: rows < nRef are s, later rows are T - S
    npo = orc.NumpyNNGP(t[rows], y[rows], want_tab, kernel_id, *prm, eps2=eps2)
    kw = dict(sigma2=prm[0], phi=prm[1], tau2=prm[2])
    for i in (0, 1, n_ref - 1, n_ref, n - 1):
        np.testing.assert_allclose(obj._CNs(i, **kw), npo._CNs(i), rtol=1e-12)
        np.testing.assert_allclose(obj._Ccross(i, **kw), npo._Ccross(i), rtol=1e-12)
        np.testing.assert_allclose(obj._Bsi(i, **kw), npo._Bsi(i), rtol=1e-8, atol=1e-11)
        np.testing.assert_allclose(obj._Fsi(i, **kw), npo._Fsi(i), rtol=RTOL)
    # kriging conditions on the reference sites and their responses only
    tn = np.random.default_rng(seed + 1).random((9, D))
    mean, var = obj.predict(tn, *prm)
    m0, v0, _ = orc.np_krige(obj.s, y[want_choice], tn, m, kernel_id, *prm,
                             eps2=None if eps2 is None else eps2[:n_ref])
    np.testing.assert_allclose(mean, m0, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(var, v0, rtol=1e-9)
    return obj


def check_subset_equals_dense_gp(make, t, y, kernel_id, prm, seed=3):
    """Known answer: with S a permutation of T and m = n - 1 the NNGP density is the exact zero-mean GP
    density, which does not depend on the ordering."""
    n = len(t)
    obj = make(t, y, 0.0, ("subset", n), n - 1, seed=seed)
    assert sorted(obj._choice.tolist()) == list(range(n)) and not np.array_equal(obj._choice, np.arange(n))
    want = orc.dense_gp_loglik(t, y, kernel_id, *prm)
    np.testing.assert_allclose(obj.loglik(*prm), want, rtol=1e-9)


def check_random(make, t, y, n_ref, m, kernel_id, prm, seed=11):
    from sklearn.neighbors import KDTree, KNeighborsRegressor

    D = t.shape[1]
    bounds = tuple((-0.25 * (k + 1), 1.0 + 0.5 * k) for k in range(D))
    obj = make(t, y, 0.0, ("random", n_ref, bounds), m, seed=seed)
    rng = np.random.default_rng(seed)
    want_s = np.vstack([rng.uniform(lo, hi, n_ref) for lo, hi in bounds]).T  # nngp.py:38-40 as written
    assert np.array_equal(obj.s, want_s) and obj.s.shape == (n_ref, D)
    s_tab = orc.c_knn_ordered(want_s, m)
    assert np.array_equal(obj._table, s_tab) and len(obj.Ns) == n_ref
    ind = KDTree(want_s).query(t, m)[1]
    assert np.array_equal(obj._nt_table, ind) and len(obj.Nt) == len(t)
    d, i = obj.Nt[5]
    np.testing.assert_allclose(d, KDTree(want_s).query(t[5].reshape(1, -1), m)[0], rtol=1e-14)
    want_ws = KNeighborsRegressor(n_neighbors=5, weights="uniform").fit(t, y).predict(want_s)
    np.testing.assert_allclose(obj.ws, want_ws, rtol=1e-13, atol=1e-15)
    npo = orc.NumpyNNGP(want_s, np.zeros(n_ref), s_tab, kernel_id, *prm)
    kw = dict(sigma2=prm[0], phi=prm[1], tau2=prm[2])
    for i in (0, 1, n_ref - 1):
        np.testing.assert_allclose(obj._CNs(i, **kw), npo._CNs(i), rtol=1e-12)
        np.testing.assert_allclose(obj._Fsi(i, **kw), npo._Fsi(i), rtol=RTOL)
        assert obj._Cs(i, **kw) == npo._Cs(i)
    with pytest.raises(NotImplementedError):
        obj.loglik(*prm)
    with pytest.raises(NotImplementedError):
        obj.predict(t[:3], *prm)
    return obj


def check_ref_type_errors(make, t, y):
    with pytest.raises(ValueError):
        make(t, y, 0.0, "T=S", 3)
    with pytest.raises(ValueError):
        make(t, y, 0.0, ("grid", 10), 3)
    with pytest.raises(ValueError):  # KDTree(s).query(t_i, m) raises upstream when m > nRef
        make(t, y, 0.0, ("subset", 4), 5, seed=0)
    with pytest.raises(ValueError):
        make(t, y, 0.0, ("subset", len(t) + 1), 3, seed=0)
    with pytest.raises(ValueError):
        make(t, y, 0.0, ("random", 10, ((0, 1),) * (t.shape[1] + 1)), 3, seed=0)


def _latent_expected(s, w, t, y, m, kernel_id, prm, eps2_t=None):
    """log p(w_S) + log p(y | w_S) computed without the engine's row layout: the oracle's NNGP density of w on s
    (no nugget) plus, per observation, the kriging density from its m nearest reference sites (oracle.np_krige)."""
    s_tab = orc.c_knn_ordered(s, m)
    slog, squad, bad = orc.c_loglik(s, w, s_tab, kernel_id, prm[0], prm[1], 0.0)
    assert bad == 0
    ll = -0.5 * (slog + squad) - 0.5 * len(s) * np.log(2 * np.pi)
    mean, var, _ = orc.np_krige(s, w, t, m, kernel_id, prm[0], prm[1], 0.0)  # var = sigma2 - c^T C_N^-1 c
    F = var + prm[2] + (0.0 if eps2_t is None else eps2_t)
    return ll + float(np.sum(-0.5 * np.log(2 * np.pi * F) - 0.5 * (y - mean) ** 2 / F))


LATENT_RTOL = 1e-8  # no nugget on the latent field: F_i of close reference sites cancels down to ~1e-6 sigma2


def check_latent_density(make, t, y, m, kernel_id, prm, seed=5):
    """loglik_latent for the three reference-set types, against the layout-free restatement above."""
    n, D = t.shape
    rng = np.random.default_rng(seed)
    kw = dict(sigma2=prm[0], phi=prm[1], tau2=prm[2])
    # S = T
    obj = make(t, y, 0.0, "S=T", m)
    w = y + 0.1 * rng.standard_normal(n)
    np.testing.assert_allclose(obj.loglik_latent(w, **kw), _latent_expected(t, w, t, y, m, kernel_id, prm), rtol=LATENT_RTOL)
    # subset, with per-observation eps, evaluated at two nuggets (the second call reuses the resident rows)
    eps = np.linspace(0.05, 0.3, n)
    sub = make(t, y, eps, ("subset", max(m + 2, n // 3)), m, seed=seed)
    ws = np.asarray(sub.ws, dtype=np.float64)
    for tau2 in (prm[2], 2.5 * prm[2]):
        want = _latent_expected(sub.s, ws, t, y, m, kernel_id, (prm[0], prm[1], tau2), eps2_t=eps ** 2)
        np.testing.assert_allclose(sub.loglik_latent(ws, prm[0], prm[1], tau2), want, rtol=LATENT_RTOL)
    # random: no response at the reference sites, the latent density is the one that exists
    bounds = tuple((0.0, 1.0) for _ in range(D))
    rnd = make(t, y, 0.0, ("random", max(m + 2, n // 4), bounds), m, seed=seed)
    wr = np.asarray(rnd.ws, dtype=np.float64)
    np.testing.assert_allclose(rnd.loglik_latent(wr, **kw), _latent_expected(rnd.s, wr, t, y, m, kernel_id, prm), rtol=LATENT_RTOL)
    with pytest.raises(ValueError):
        rnd.loglik_latent(wr[:-1], **kw)


def check_latent_dense_identity(make, t, y, kernel_id, prm):
    """Closed form for S = T with m = n - 1: the dense GP density of w (no nugget) plus independent
    N(y_i | w_i, tau2) terms -- every observation's nearest reference site is itself."""
    n = len(t)
    obj = make(t, y, 0.0, "S=T", n - 1)
    w = 0.7 * y + 0.05
    want = orc.dense_gp_loglik(t, w, kernel_id, prm[0], prm[1], 0.0) + float(
        np.sum(-0.5 * np.log(2 * np.pi * prm[2]) - 0.5 * (y - w) ** 2 / prm[2]))
    np.testing.assert_allclose(obj.loglik_latent(w, *prm), want, rtol=1e-8)
