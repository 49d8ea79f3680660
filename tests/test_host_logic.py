"""CPU tests of the drop-in class's host logic (pynngp_b200/nngp.py) with the oracle-backed FakeEngine
(tests/fake_engine.py) monkeypatched in place of the CUDA engine: row layouts, response columns, eps,
neighbour-table injection, prediction bookkeeping, the Metropolis driver, and the reference sets other
than T (nngp.py:32-40, 68-71).  The same assertions run against the real engine in
tests/test_gpu_parity.py (shared through tests/host_checks.py)."""
import numpy as np
import pytest

import host_checks as hc
from fake_engine import FakeEngine
from oracle import nngp_oracle as orc
from pynngp_b200 import Exponential, Matern, _lib
from pynngp_b200.synthetic import synthetic

PRM = (1.3, 5.0, 0.07)


@pytest.fixture()
def make(monkeypatch):
    monkeypatch.setattr(_lib, "Engine", FakeEngine)
    import pyNNGP

    def _make(t, y, eps, refType, m, cov=None, **kw):
        return pyNNGP.NNGP(t, y, eps, refType, m, cov, **kw)

    return _make


def test_st_class_matches_oracle(make):
    s, y = synthetic(300, 2, 21)
    obj = make(s, y, 0.0, "S=T", 6, Matern(1.5, *PRM))
    tab = orc.c_knn_ordered(s, 6)
    assert obj.s is s and obj.Nt is obj.Ns and np.array_equal(obj._table, tab)
    assert obj.Ns[0] == [] and obj.Ns[3].tolist() == tab[3, :3].tolist() and len(obj.Ns[299]) == 6
    for i in range(300):
        assert i not in obj.Ns[i]  # the reference's own assertion, tests/test_init.py:22-23
    slog, squad, _ = orc.c_loglik(s, y, tab, 1, *PRM)
    np.testing.assert_allclose(obj.loglik_terms(), (slog, squad), rtol=1e-12)
    np.testing.assert_allclose(obj.loglik(phi=4.0), orc.loglik_from_terms(*orc.c_loglik(s, y, tab, 1, 1.3, 4.0, 0.07)[:2], 300),
                               rtol=1e-12)
    from sklearn.neighbors import KNeighborsRegressor

    np.testing.assert_allclose(obj.ws, KNeighborsRegressor(5).fit(s, y).predict(s), rtol=1e-13)
    B, F = obj.factors()
    B0, F0 = orc.c_factors(s, y, tab, 1, *PRM)
    assert np.array_equal(B, B0) and np.array_equal(F, F0)
    with pytest.raises(NotImplementedError):
        obj.oneSample()


def test_two_columns_and_eps_sum_over_columns(make):
    s, y = synthetic(250, 2, 9)
    y2 = np.stack([y, -2 * y], axis=1)
    eps = np.stack([np.full(250, 0.1), np.linspace(0.0, 0.3, 250)], axis=1)
    obj = make(s, y2, eps, "S=T", 5, Exponential(*PRM))
    tab = orc.c_knn_ordered(s, 5)
    want = np.zeros(3)
    for c in range(2):
        want += orc.c_loglik(s, y2[:, c], tab, 0, *PRM, eps2=eps[:, c] ** 2)
    np.testing.assert_allclose(obj.loglik_batch([PRM])[0], want, rtol=1e-12)
    np.testing.assert_allclose(obj.loglik(), -0.5 * (want[0] + want[1]) - 0.5 * 250 * 2 * np.log(2 * np.pi), rtol=1e-12)
    assert obj.ws.shape == (250, 2)
    tn = np.random.default_rng(0).random((7, 2))
    mean, var = obj.predict(tn)
    assert mean.shape == (7, 2)
    for c in range(2):
        m0, v0, _ = orc.np_krige(s, y2[:, c], tn, 5, 0, *PRM, eps2=eps[:, c] ** 2)
        np.testing.assert_allclose(mean[:, c], m0, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(var[:, c], v0, rtol=1e-9)
    # scalar eps enters as one variance for every observation
    obj1 = make(s, y, 0.2, "S=T", 5, Exponential(*PRM))
    np.testing.assert_allclose(obj1.loglik_terms(), orc.c_loglik(s, y, tab, 0, *PRM, eps2=np.full(250, 0.04))[:2], rtol=1e-12)


def test_one_dimensional_sites_and_predict_shapes(make):
    rng = np.random.default_rng(4)
    t = rng.random(120)
    y = np.sin(6 * t)
    obj = make(t, y, 0.0, "S=T", 4, Matern(2.5, *PRM))
    tab = orc.c_knn_ordered(t[:, None], 4)
    assert np.array_equal(obj._table, tab)
    mean, var = obj.predict(np.array([0.5, 0.25]))
    m0, v0, _ = orc.np_krige(t, y, np.array([0.5, 0.25]), 4, 2, *PRM)
    np.testing.assert_allclose(mean, m0, rtol=1e-9)
    assert obj.predict(np.zeros((0, 1)))[0].shape == (0,)
    with pytest.raises(ValueError):
        obj.predict(np.zeros((3, 2)))


def test_table_injection_save_load_and_metropolis(make, tmp_path):
    s, y = synthetic(200, 2, 14)
    spec = Matern(1.5, 1.0, 6.0, 0.1)
    a = make(s, y, 0.0, "S=T", 5, spec)
    path = str(tmp_path / "nbr.npy")
    a.save_neighbors(path)
    before = FakeEngine.launches
    b = make(s, y, 0.0, "S=T", 5, spec, neighbors=path)
    assert FakeEngine.launches == before  # stage 1 skipped
    assert np.array_equal(a._table, b._table) and a.loglik() == b.loglik()
    chain, trace, rate = b.metropolis(25, step=0.1, seed=3)
    assert chain.shape == (25, 3) and (chain > 0).all() and 0.0 <= rate <= 1.0
    assert trace[-1] == b.loglik(*chain[-1])
    chain2, trace2, _ = b.metropolis(25, step=0.1, seed=3)
    assert np.array_equal(chain, chain2) and np.array_equal(trace, trace2)
    # a non-SPD evaluation is reported, never returned as a number
    with pytest.raises(FloatingPointError):
        b.loglik_terms(1.0, 6.0, -1.5)


def test_constructor_errors(make):
    s, y = synthetic(50, 2, 1)
    with pytest.raises(TypeError):
        make(s, y, 0.0, "S=T", 3, lambda a, b: 1.0)
    with pytest.raises(ValueError):
        make(s, y, 0.0, "S=T", 3, None, knn="kdtree")
    bad = s.copy()
    bad[7, 1] = np.nan
    with pytest.raises(ValueError):
        make(bad, y, 0.0, "S=T", 3, None)
    hc.check_ref_type_errors(make, s, y)


@pytest.mark.parametrize("n,D,n_ref,m,kernel_id", [(260, 2, 90, 6, 1), (180, 3, 60, 9, 0), (150, 1, 40, 3, 2)])
def test_subset_reference_set(make, n, D, n_ref, m, kernel_id):
    t, y = synthetic(n, D, 30 + D)
    spec = [Exponential(), Matern(1.5), Matern(2.5)][kernel_id]
    obj = hc.check_subset(lambda *a, **k: make(*a, spec, **k), t, y, n_ref, m, kernel_id, PRM)
    # the global numpy RNG is used when no seed is given, as upstream (nngp.py:36)
    np.random.seed(5)
    o2 = make(t, y, 0.0, ("subset", n_ref), m, spec)
    np.random.seed(5)
    assert np.array_equal(o2._choice, np.random.choice(n, size=n_ref, replace=False))
    assert obj._n_ref == n_ref


def test_subset_with_eps_and_two_columns(make):
    t, y = synthetic(140, 2, 77)
    eps = np.linspace(0.05, 0.4, 140)
    hc.check_subset(lambda *a, **k: make(*a, Exponential(), **k), t, y, 50, 5, 0, PRM, eps=eps)
    y2 = np.stack([y, 0.5 * y + 1.0], axis=1)
    obj = make(t, y2, 0.0, ("subset", 50), 5, Exponential(*PRM), seed=7)
    rows, tab = obj._rows, obj._table
    want = sum(np.array(orc.c_loglik(t[rows], y2[rows, c], tab, 0, *PRM)) for c in range(2))
    np.testing.assert_allclose(obj.loglik_batch([PRM])[0], want, rtol=1e-12)
    assert obj.ws.shape == (50, 2)


def test_subset_of_everything_is_the_dense_gp(make):
    t, y = synthetic(24, 2, 8)
    for kid, spec in ((0, Exponential()), (1, Matern(1.5))):
        hc.check_subset_equals_dense_gp(lambda *a, **k: make(*a, spec, **k), t, y, kid, PRM)


def test_random_reference_set(make):
    t, y = synthetic(200, 2, 12)
    hc.check_random(lambda *a, **k: make(*a, Matern(1.5), **k), t, y, 70, 6, 1, PRM)
    t3, y3 = synthetic(120, 3, 13)
    hc.check_random(lambda *a, **k: make(*a, Exponential(), **k), t3, y3, 40, 4, 0, PRM)


def test_latent_density(make):
    t, y = synthetic(150, 2, 31)
    hc.check_latent_density(lambda *a, **k: make(*a, Matern(1.5), **k), t, y, 6, 1, PRM)
    t1, y1 = synthetic(90, 1, 32)
    hc.check_latent_density(lambda *a, **k: make(*a, Exponential(), **k), t1, y1, 4, 0, PRM)


def test_latent_density_dense_identity(make):
    t, y = synthetic(20, 2, 33)
    hc.check_latent_dense_identity(lambda *a, **k: make(*a, Exponential(), **k), t, y, 0, PRM)


def test_close_releases_both_engines(make):
    t, y = synthetic(60, 2, 3)
    with make(t, y, 0.0, "S=T", 4, Exponential(*PRM)) as obj:
        obj.loglik_latent(y)
        main, latent = obj._engine, obj._latent_eng
        assert not main._closed and not latent._closed
    assert main._closed and latent._closed and obj._latent_eng is None


@pytest.mark.parametrize("seed", range(24))
def test_random_configurations_hold_the_invariants(make, seed):
    """Seeded sweep over (n, D, m, refType, response columns, eps form): whatever the shapes -- m >= n, n below the
    5 neighbours `ws` asks for, a subset of everything, 1-D sites given as a flat array -- the likelihood equals the
    oracle's on the engine's row layout and the reference's attributes keep their shapes."""
    rng = np.random.default_rng(1000 + seed)
    n = int(rng.integers(2, 70))
    D = int(rng.integers(1, 4))
    m = int(rng.integers(1, 13))
    ncol = int(rng.integers(1, 3))
    kid = int(rng.integers(0, 3))
    spec = [Exponential(*PRM), Matern(1.5, *PRM), Matern(2.5, *PRM)][kid]
    t = rng.random((n, D))
    t_arg = t[:, 0] if D == 1 and seed % 2 else t           # flat 1-D sites, as the reference would accept them
    y = rng.standard_normal((n, ncol)) if ncol > 1 else rng.standard_normal(n)
    eps = [0.0, 0.15, np.abs(rng.standard_normal(np.shape(y))) * 0.2][seed % 3]
    kind = ["S=T", "subset", "random"][int(rng.integers(0, 3))]
    if kind == "S=T":
        ref = "S=T"
    else:
        n_ref = int(rng.integers(1, n + 1))
        if m > n_ref:
            with pytest.raises(ValueError):
                make(t_arg, y, eps, (kind, n_ref) + (((0, 1),) * D,) * (kind == "random"), m, spec, seed=seed)
            return
        ref = ("subset", n_ref) if kind == "subset" else ("random", n_ref, ((0.0, 1.0),) * D)
    obj = make(t_arg, y, eps, ref, m, spec, seed=seed)
    n_ref = obj._n_ref
    assert len(obj.Ns) == n_ref and obj.Ns[0] == [] and len(obj.Nt) == (n_ref if kind == "S=T" else n)
    assert all(len(obj.Ns[i]) == min(m, i) for i in range(n_ref))
    assert np.shape(obj.ws) == (n_ref,) + np.shape(y)[1:] and np.array_equal(obj.wt, y)
    if kind != "S=T":
        d, idx = obj.Nt[n - 1]
        assert d.shape == idx.shape == (1, m) and (np.diff(d[0]) >= 0).all() and (idx < n_ref).all()
    y2d = np.asarray(y, dtype=np.float64).reshape(n, -1)
    e2 = np.broadcast_to(np.asarray(eps, dtype=np.float64).reshape((n, -1) if np.ndim(eps) else (1, 1)), y2d.shape) ** 2
    if kind == "random":
        with pytest.raises(NotImplementedError):
            obj.loglik()
    else:
        rows = np.arange(n) if obj._rows is None else obj._rows
        want = np.zeros(3)
        for c in range(ncol):
            want += orc.c_loglik(t[rows], y2d[rows, c], obj._table, kid, *PRM, eps2=e2[rows, c] if np.any(e2) else None)
        got = obj.loglik_batch([PRM])[0]
        np.testing.assert_allclose(got, want, rtol=1e-11)
        mean, var = obj.predict(rng.random((3, D)), m=min(m, n_ref))
        assert mean.shape == var.shape == (3,) + np.shape(y)[1:] and (var > 0).all()
    w = rng.standard_normal((n_ref, ncol))
    ll = obj.loglik_latent(w if ncol > 1 else w[:, 0])
    want_ll = 0.0
    for c in range(ncol):
        want_ll += hc._latent_expected(np.asarray(obj.s, dtype=np.float64).reshape(n_ref, -1), w[:, c], t, y2d[:, c],
                                       m, kid, PRM, eps2_t=e2[:, c] if np.any(e2) else None)
    np.testing.assert_allclose(ll, want_ll, rtol=1e-7)
