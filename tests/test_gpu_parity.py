"""GPU parity tests (run on the B200 box: pytest -m gpu).  Every call goes through the C ABI
(pynngp_b200._lib.Engine -> libnngp_b200.so) or the drop-in class on top of it; the oracle
(oracle/) and the golden fixtures (tests/golden/) are the checkers.

Tolerances (BASELINE.json north_star): neighbour index sets bit-exact; b_i, F_i and the
log-likelihood statistics within 1e-10 relative in fp64 and 1e-4 in fp32.
"""
import os

import numpy as np
import pytest

from oracle import nngp_oracle as orc
from pynngp_b200 import _lib
from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic

pytestmark = pytest.mark.gpu

RTOL64 = 1e-10
RTOL32 = 1e-4
P0 = np.array([PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0])
KIDS = {"exponential": 0, "matern32": 1, "matern52": 2}


def engine(s, y, eps2=None, dtype="float64"):
    e = _lib.Engine(0, dtype)
    e.set_data(s, y, eps2)
    return e


# ---- stage 1: ordered k-NN ------------------------------------------------------------------------
@pytest.mark.parametrize("name", ["test_init_shape", "cfg1", "d2_m15", "d3_m30", "d1_m5", "d3_m32"])
def test_knn_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"ns_{name}.npz"))
    s = g["coords"]
    e = engine(s, np.zeros(len(s)))
    e.build_neighbors(int(g["m"]))
    assert np.array_equal(e.get_neighbors(), g["Ns"])  # bit-exact incl. order and -1 padding


def test_knn_lattice_ties(golden_dir):
    g = np.load(os.path.join(golden_dir, "ns_lattice.npz"))
    s, m = g["coords"], int(g["m"])
    e = engine(s, np.zeros(len(s)))
    e.build_neighbors(m)
    got = e.get_neighbors()
    assert np.array_equal(got, orc.c_knn_ordered(s, m))  # the engine's (d2, j) rule, exactly
    ref = g["Ns"]
    for i in range(len(s)):  # vs the reference: same distance multiset per row
        a, b = got[i][got[i] >= 0], ref[i][ref[i] >= 0]
        assert np.array_equal(np.sort(orc.np_dist2(s[i], s[a])), np.sort(orc.np_dist2(s[i], s[b])))


def test_knn_duplicates_and_tiny_n():
    s = np.array([[0.5, 0.5]] * 6 + [[0.25, 0.5]] * 3)
    e = engine(s, np.zeros(len(s)))
    e.build_neighbors(4)
    assert np.array_equal(e.get_neighbors(), orc.c_knn_ordered(s, 4))
    e1 = engine(np.array([[0.1, 0.2]]), np.zeros(1))
    e1.build_neighbors(3)
    assert e1.get_neighbors().tolist() == [[-1, -1, -1]]


@pytest.mark.parametrize("n,D,m", [(20000, 2, 15), (9000, 3, 30), (5000, 1, 7), (1537, 2, 32)])
def test_knn_matches_oracle_mid_size(n, D, m):
    s, y = synthetic(n, D, 100 + D)
    e = engine(s, y)
    e.build_neighbors(m)
    got = e.get_neighbors()
    want = orc.c_knn_ordered(s, m, threads=os.cpu_count() or 1)
    assert np.array_equal(got, want)


def test_knn_tile_split_assembles():
    s, y = synthetic(3000, 2, 9)
    e = engine(s, y)
    e.build_neighbors(10)
    full = e.get_neighbors()
    parts = []
    for off in range(3):
        e.build_neighbors(10, off, 3)
        parts.append(e.get_neighbors())
    for p in parts:
        assert ((p == full) | (p == _lib.ROW_UNSET)).all()
    assert np.array_equal(np.maximum.reduce(parts), full)
    # each row is owned by exactly one rank
    owned = sum((p[:, 0] != _lib.ROW_UNSET).astype(int) for p in parts)
    assert (owned == 1).all()


# ---- stage 1 through the cell grid: bit-identical to the brute-force kernel --------------------------
def grid_table(s, m, lam=1.0, brute_rows=128, lo=0, hi=None, algo="grid"):
    e = engine(s, np.zeros(len(s)))
    e.set_knn_tuning(lam, brute_rows)
    e.build_neighbors_grid(m, lo, hi, algo)
    return e, e.get_neighbors()


@pytest.mark.parametrize("name", ["test_init_shape", "cfg1", "d2_m15", "d3_m30", "d1_m5", "d3_m32"])
def test_knn_grid_matches_reference_golden(golden_dir, name):
    g = np.load(os.path.join(golden_dir, f"ns_{name}.npz"))
    for lam in (0.3, 1.0, 3.0):
        e, got = grid_table(g["coords"], int(g["m"]), lam)
        assert e.knn_used_grid()
        assert np.array_equal(got, g["Ns"])


def test_knn_grid_ties_duplicates_degenerate(golden_dir):
    g = np.load(os.path.join(golden_dir, "ns_lattice.npz"))
    s, m = g["coords"], int(g["m"])
    for lam in (0.2, 1.0):
        assert np.array_equal(grid_table(s, m, lam, 128)[1], orc.c_knn_ordered(s, m))
    rng = np.random.default_rng(5)
    dup = np.repeat(rng.random((40, 2)), 10, axis=0)[rng.permutation(400)]   # every site 10 times
    assert np.array_equal(grid_table(dup, 12)[1], orc.c_knn_ordered(dup, 12))
    same = np.full((300, 3), 0.25)                                           # zero-extent box
    assert np.array_equal(grid_table(same, 5)[1], orc.c_knn_ordered(same, 5))
    line = np.stack([rng.random(600), np.full(600, 2.0)], axis=1)            # one collapsed dimension
    assert np.array_equal(grid_table(line, 9)[1], orc.c_knn_ordered(line, 9))
    thin = np.stack([rng.random(900), 1e-7 * rng.random(900), rng.random(900)], axis=1)  # extent below a cell
    assert np.array_equal(grid_table(thin, 16)[1], orc.c_knn_ordered(thin, 16))
    far = rng.random((700, 2)) * 1e-3 + 1e6                                  # large offset: slack dominates
    assert np.array_equal(grid_table(far, 10)[1], orc.c_knn_ordered(far, 10))


@pytest.mark.parametrize("n,D,m", [(20000, 2, 15), (9000, 3, 30), (5000, 1, 7), (1537, 2, 32), (3000, 3, 1)])
def test_knn_grid_matches_oracle_mid_size(n, D, m):
    s, _ = synthetic(n, D, 100 + D)
    want = orc.c_knn_ordered(s, m, threads=os.cpu_count() or 1)
    for lam, brute_rows in ((1.0, 128), (0.25, 512), (4.0, 128)):
        assert np.array_equal(grid_table(s, m, lam, brute_rows)[1], want)


def test_knn_grid_clustered_and_auto():
    rng = np.random.default_rng(11)
    # one dense blob over a uniform background: the default cells are rejected by the work estimate, the
    # refined ones (a quarter of the occupancy per attempt) are accepted
    mod = np.concatenate([0.5 + 0.01 * rng.standard_normal((18000, 2)), rng.random((12000, 2))])
    mod = mod[rng.permutation(len(mod))]
    want = orc.c_knn_ordered(mod, 15, threads=os.cpu_count() or 1)
    e, got = grid_table(mod, 15, 1.0, 8192, algo="auto")
    assert e.knn_used_grid() and np.array_equal(got, want)
    # blobs whose densities differ by three orders of magnitude defeat any single cell size: auto may fall
    # back to brute force, a forced grid is slow there -- the table is the same either way
    blobs = np.concatenate([0.5 + 0.01 * rng.standard_normal((15000, 2)), rng.random((5000, 2)),
                            0.2 + 0.0005 * rng.standard_normal((5000, 2))])
    s = blobs[rng.permutation(len(blobs))]
    want = orc.c_knn_ordered(s, 15, threads=os.cpu_count() or 1)
    e, got = grid_table(s, 15, 1.0, 8192, algo="auto")
    assert np.array_equal(got, want)
    assert np.array_equal(grid_table(s, 15, 1.0, 256, algo="grid")[1], want)
    # a single tight clump: the histogram predicts no gain, auto falls back to brute force
    clump = np.concatenate([np.full((30000, 2), 0.5) + 1e-12 * rng.standard_normal((30000, 2)), [[0.0, 0.0], [1.0, 1.0]]])
    e, got = grid_table(clump, 8, 1.0, 8192, algo="auto")
    assert not e.knn_used_grid()
    e2 = engine(clump, np.zeros(len(clump)))
    e2.build_neighbors(8)
    assert np.array_equal(got, e2.get_neighbors())
    # non-finite coordinates never reach the grid
    bad = rng.random((500, 2)); bad[17, 0] = np.inf
    e3 = engine(bad, np.zeros(500))
    e3.build_neighbors_grid(4, 0, None, "auto")
    assert not e3.knn_used_grid()
    with pytest.raises(_lib.NNGPError):
        e3.build_neighbors_grid(4, 0, None, "grid")


def test_knn_grid_row_ranges_assemble():
    s, _ = synthetic(30000, 2, 9)
    e, full = grid_table(s, 10, 1.0, 1024)
    parts = []
    for lo, hi in ((0, 700), (700, 11000), (11000, 30000)):
        p = grid_table(s, 10, 1.0, 1024, lo, hi)[1]
        assert np.array_equal(p[lo:hi], full[lo:hi])
        assert ((p == full) | (p == _lib.ROW_UNSET)).all()
        parts.append(p)
    assert np.array_equal(np.maximum.reduce(parts), full)
    b = grid_table(s, 10, 1.0, 1024, 5000, 6000, algo="brute")[1]
    assert np.array_equal(b[5000:6000], full[5000:6000]) and (b[6000:] == _lib.ROW_UNSET).all()


@pytest.mark.parametrize("n,D,m", [(200000, 2, 15), (120000, 3, 30)])
def test_knn_grid_equals_brute_kernel_large(n, D, m):
    s, _ = synthetic(n, D, 21)
    e = engine(s, np.zeros(n))
    e.build_neighbors(m)
    want = e.get_neighbors()
    e.build_neighbors_grid(m)          # defaults: auto, 8192 brute rows
    assert e.knn_used_grid()
    assert np.array_equal(e.get_neighbors(), want)


def test_knn_plain_through_grid():
    for D, n, k in ((2, 40000, 5), (3, 20000, 8), (1, 9000, 3)):
        s, _ = synthetic(n, D, 70 + D)
        e = engine(s, np.zeros(n))
        idx = e.knn_plain(k)
        assert (idx[:, 0] == np.arange(n)).all()
        for i in list(range(0, n, n // 50)) + [n - 1]:
            d2 = orc.np_dist2(s[i], s)
            assert np.array_equal(idx[i], np.lexsort((np.arange(n), d2))[:k])


# ---- stages 2-3 -----------------------------------------------------------------------------------
CASES = [  # n, D, m, kernel
    (1000, 2, 10, "exponential"),   # cfg1
    (800, 2, 15, "matern32"),       # cfg3's shape
    (700, 3, 30, "matern32"),       # cfg4's shape
    (500, 1, 3, "matern52"),
    (600, 3, 7, "exponential"),
    (400, 2, 20, "matern52"),
    (300, 3, 32, "exponential"),
    (300, 2, 31, "matern32"),
    (200, 2, 1, "exponential"),
]


@pytest.mark.parametrize("n,D,m,kernel", CASES)
def test_cov_blocks_factors_loglik_fp64(n, D, m, kernel):
    s, y = synthetic(n, D, 31 + m)
    kid = KIDS[kernel]
    eps2 = np.linspace(0.0, 0.02, n)
    nbr = orc.c_knn_ordered(s, m)
    e = engine(s, y, eps2)
    e.set_neighbors(nbr)
    prm = np.array([1.3, 5.0, 0.07, 0.0])
    CN, cc, cs = e.cov_blocks(kid, prm)
    CN0, cc0, cs0 = orc.c_cov_blocks(s, nbr, kid, *prm[:3], eps2=eps2)
    np.testing.assert_allclose(CN, CN0, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(cc, cc0, rtol=1e-12, atol=1e-300)
    np.testing.assert_allclose(cs, cs0, rtol=1e-15)
    B, F = e.factors(kid, prm)
    B0, F0 = orc.c_factors(s, y, nbr, kid, *prm[:3], eps2=eps2)
    scale = np.maximum(np.abs(B0).max(axis=1, keepdims=True), 1.0)
    assert (np.abs(B - B0) / scale).max() <= RTOL64
    np.testing.assert_allclose(F, F0, rtol=RTOL64)
    st = e.loglik(kid, prm)[0]
    st0 = orc.c_loglik(s, y, nbr, kid, *prm[:3], eps2=eps2)
    assert st[2] == 0 and st0[2] == 0
    np.testing.assert_allclose(st[:2], st0[:2], rtol=RTOL64)


@pytest.mark.parametrize("n,D,m,kernel", CASES[:4])
def test_loglik_fp32(n, D, m, kernel):
    s, y = synthetic(n, D, 31 + m)
    kid = KIDS[kernel]
    nbr = orc.c_knn_ordered(s, m)
    e = engine(s, y, None, "float32")
    e.set_neighbors(nbr)
    st = e.loglik(kid, P0)[0]
    st0 = orc.c_loglik(s, y, nbr, kid, *P0[:3])
    np.testing.assert_allclose(st[:2], st0[:2], rtol=RTOL32)
    B, F = e.factors(kid, P0)
    B0, F0 = orc.c_factors(s, y, nbr, kid, *P0[:3])
    scale = np.maximum(np.abs(B0).max(axis=1, keepdims=True), 1.0)
    assert (np.abs(B - B0) / scale).max() <= RTOL32
    np.testing.assert_allclose(F, F0, rtol=RTOL32)


def test_batched_params_and_determinism():
    """K parameter vectors in one call: the sweep variant (pair distances computed once per location,
    sigma2 factored out) against one-at-a-time evaluation and the oracle; odd K exercises a ragged chunk."""
    s, y = synthetic(5000, 2, 8)
    nbr = orc.c_knn_ordered(s, 15, threads=4)
    eps2 = np.linspace(0.0, 0.05, 5000)
    nbr30 = orc.c_knn_ordered(s, 30, threads=4)
    for kid, m_use in ((1, 15), (0, 9), (2, 12), (1, 5), (1, 30), (0, 17)):
        e = engine(s, y, eps2)
        src = nbr if m_use <= 15 else nbr30
        e.set_neighbors(np.ascontiguousarray(np.where(np.arange(src.shape[1])[None, :] < m_use, src, -1)[:, :m_use]))
        tab = e.get_neighbors()
        rng = np.random.default_rng(kid)
        K = 11
        prm = np.stack([rng.uniform(0.5, 2, K), rng.uniform(3, 30, K), rng.uniform(0.01, 0.5, K), np.zeros(K)], 1)
        st = e.loglik(kid, prm)
        assert np.array_equal(st, e.loglik(kid, prm))  # same launch shape -> bitwise identical
        for k in range(K):
            one = e.loglik(kid, prm[k])[0]
            np.testing.assert_allclose(st[k, :2], one[:2], rtol=1e-11)
            assert st[k, 2] == one[2] == 0
        for k in (0, 5, 10):
            st0 = orc.c_loglik(s, y, tab, kid, *prm[k, :3], eps2=eps2)
            np.testing.assert_allclose(st[k, :2], st0[:2], rtol=RTOL64)
    # shards and an empty shard through the batched path
    e.set_shard(1000, 3000)
    a = e.loglik(2, prm[:3])
    for k in range(3):
        st0 = orc.c_loglik(s, y, tab, 2, *prm[k, :3], eps2=eps2, lo=1000, hi=3000)
        np.testing.assert_allclose(a[k, :2], st0[:2], rtol=RTOL64)
    e.set_shard(7, 7)
    assert np.array_equal(e.loglik(2, prm[:3]), np.zeros((3, 3)))


def test_batched_non_spd_counted():
    s = np.array([[0.0, 0.0]] * 3 + [[1.0, 1.0], [0.5, 0.5]] + [[0.1 * k, 0.3] for k in range(1, 8)])
    y = np.ones(len(s))
    e = engine(s, y)
    e.build_neighbors(8)
    prm = np.array([[1.0, 1.0, 0.0, 0.0], [2.0, 3.0, 0.1, 0.0]])
    st = e.loglik(0, prm)
    # vector 0: coincident sites without a nugget are exactly singular -- the locations are counted and
    # left out, never a NaN (which of the knife-edge pivots round to <= 0 is arithmetic-dependent, so the
    # count itself is not compared); vector 1 has a nugget: regular, must match the oracle
    assert st[0, 2] >= 1 and np.isfinite(st).all()
    st1 = orc.c_loglik(s, y, e.get_neighbors(), 0, *prm[1, :3])
    assert st[1, 2] == 0 and st1[2] == 0
    np.testing.assert_allclose(st[1, :2], st1[:2], rtol=RTOL64)


def test_shards_sum_to_whole_and_empty_shard():
    s, y = synthetic(10007, 2, 12)
    nbr = orc.c_knn_ordered(s, 15, threads=4)
    e = engine(s, y)
    e.set_neighbors(nbr)
    whole = e.loglik(1, P0)[0]
    from pynngp_b200.dist import shard_bounds

    for world in (2, 4, 8):
        acc = np.zeros(3)
        for r in range(world):
            e.set_shard(*shard_bounds(len(s), r, world))
            acc += e.loglik(1, P0)[0]
        np.testing.assert_allclose(acc[:2], whole[:2], rtol=1e-12)
    e.set_shard(5, 5)
    assert np.array_equal(e.loglik(1, P0)[0], np.zeros(3))
    # shard parity against the oracle on the same rows
    e.set_shard(1234, 4321)
    st0 = orc.c_loglik(s, y, nbr, 1, *P0[:3], lo=1234, hi=4321)
    np.testing.assert_allclose(e.loglik(1, P0)[0][:2], st0[:2], rtol=RTOL64)


def test_dense_gp_identity_through_engine():
    """m = n-1 = 31: the NNGP density equals the exact GP density (known answer)."""
    n = 32
    s, y = synthetic(n, 2, 77)
    e = engine(s, y)
    e.build_neighbors(n - 1)
    for kid in (0, 1, 2):
        st = e.loglik(kid, P0)[0]
        ll = orc.loglik_from_terms(st[0], st[1], n)
        exact = orc.dense_gp_loglik(s, y, kid, *P0[:3])
        assert abs(ll - exact) <= 1e-11 * abs(exact)


def test_non_spd_counted_not_nan():
    s = np.array([[0.0, 0.0], [0.0, 0.0], [0.0, 0.0], [1.0, 1.0], [0.5, 0.5]])
    y = np.ones(5)
    e = engine(s, y)
    e.build_neighbors(3)
    st = e.loglik(0, np.array([1.0, 1.0, 0.0, 0.0]))[0]
    st0 = orc.c_loglik(s, y, e.get_neighbors(), 0, 1.0, 1.0, 0.0)
    assert st[2] >= 1 and np.isfinite(st[:2]).all()
    assert st[2] == st0[2]


def test_nonpositive_sigma2_counted_not_nan():
    """sigma2 is factored out of the matrix in the fp64 reduction kernels: a non-positive or non-finite
    sigma2 must come back as 'every location bad', not as NaN -- one vector, a sweep, and the rolled shape."""
    for n, D, m in ((3000, 2, 15), (1200, 3, 30), (500, 2, 5)):
        s, y = synthetic(n, D, 3)
        e = engine(s, y)
        e.build_neighbors(m)
        good = np.array([1.0, 6.0, 0.1, 0.0])
        want = e.loglik(1, good)[0]
        for bad_s2 in (0.0, -1.0, np.inf, np.nan):
            st = e.loglik(1, np.array([bad_s2, 6.0, 0.1, 0.0]))[0]
            assert np.array_equal(st, [0.0, 0.0, float(n)]), (m, bad_s2, st)
            both = e.loglik(1, np.array([good, [bad_s2, 6.0, 0.1, 0.0], good]))
            assert np.array_equal(both[1], [0.0, 0.0, float(n)])
            np.testing.assert_allclose(both[0], want, rtol=1e-12)
            np.testing.assert_allclose(both[2], want, rtol=1e-12)


def test_set_y_replaces_response():
    s, y = synthetic(3000, 2, 3)
    nbr = orc.c_knn_ordered(s, 8)
    e = engine(s, y)
    e.set_neighbors(nbr)
    y2 = np.cos(7 * y)
    e.set_y(y2)
    st0 = orc.c_loglik(s, y2, nbr, 0, *P0[:3])
    np.testing.assert_allclose(e.loglik(0, P0)[0][:2], st0[:2], rtol=RTOL64)


def test_error_behaviour():
    e = _lib.Engine(0)
    with pytest.raises(_lib.NNGPError):
        e.loglik(0, P0)  # no data yet
    s, y = synthetic(100, 2, 1)
    e.set_data(s, y)
    with pytest.raises(_lib.NNGPError):
        e.loglik(0, P0)  # no neighbours yet
    with pytest.raises(_lib.NNGPError):
        e.build_neighbors(33)
    with pytest.raises(_lib.NNGPError):
        e.set_data(np.zeros((10, 4)), np.zeros(10))  # D > 3


# ---- the drop-in class ---------------------------------------------------------------------------
def test_reference_own_test_headless():
    """tests/test_init.py of the reference, body reproduced against the new class (plot dropped):
    n=200 uniform 2-D sites, y and eps of shape (n, 2), m=3, cov=None."""
    import pyNNGP

    n = 200
    rng = np.random.default_rng(1234)
    t = np.vstack([rng.uniform(size=n), rng.uniform(size=n)]).T
    y = np.zeros_like(t)
    eps = np.ones_like(t) * 0.001
    nngp = pyNNGP.NNGP(t, y, eps, "S=T", 3, None)
    for i in range(n):
        assert i not in nngp.Ns[i]
    assert nngp.s is t and nngp.Nt is nngp.Ns and nngp.Ns[0] == []
    assert nngp.Ns[5].dtype == np.int64 and len(nngp.Ns[2]) == 2
    assert nngp.s[nngp.Ns[7]].shape == (3, 2)
    assert np.array_equal(nngp.wt, y)


def test_class_cfg1_end_to_end():
    import pyNNGP
    from pynngp_b200 import Exponential

    c = CONFIGS["cfg1"]
    s, y = synthetic(c["n"], c["D"], c["seed"])
    obj = pyNNGP.NNGP(s, y, 0.0, "S=T", c["m"], Exponential(**PARAMS))
    want_tab = orc.c_knn_ordered(s, c["m"])
    assert np.array_equal(obj._table, want_tab)
    npo = orc.NumpyNNGP(s, y, want_tab, 0, *P0[:3])
    for i in (0, 1, 5, 10, 500, 999):
        np.testing.assert_allclose(obj._CNs(i), npo._CNs(i).reshape(obj._CNs(i).shape), rtol=1e-12)
        np.testing.assert_allclose(obj._Ccross(i), npo._Ccross(i), rtol=1e-12)
        assert obj._Cs(i) == npo._Cs(i)
        np.testing.assert_allclose(obj._Bsi(i), npo._Bsi(i), rtol=1e-9, atol=1e-12)
        np.testing.assert_allclose(obj._Fsi(i), npo._Fsi(i), rtol=RTOL64)
    slog, squad, _ = orc.c_loglik(s, y, want_tab, 0, *P0[:3])
    got = obj.loglik_terms()
    np.testing.assert_allclose(got, (slog, squad), rtol=RTOL64)
    np.testing.assert_allclose(obj.loglik(), orc.loglik_from_terms(slog, squad, c["n"]), rtol=RTOL64)
    with pytest.raises(TypeError):
        pyNNGP.NNGP(s, y, 0.0, "S=T", 3, lambda a, b: 1.0)
    with pytest.raises(NotImplementedError):
        obj.oneSample()


# ---- either side of the path: prediction, table I/O, sweep driver -----------------
numpy_kriging = orc.np_krige  # the oracle's exact restatement


@pytest.mark.parametrize("n,D,m,kernel", [(4000, 2, 15, "matern32"), (3000, 3, 30, "exponential"), (2500, 1, 6, "matern52"),
                                          (20000, 2, 10, "exponential")])
def test_predict_matches_numpy_kriging(n, D, m, kernel):
    import pyNNGP
    from pynngp_b200 import Exponential, Matern

    s, y = synthetic(n, D, 40 + m)
    rng = np.random.default_rng(n)
    tn = rng.random((257, D)) * 1.2 - 0.1  # some sites outside the reference bounding box
    def mk(tau2):
        return {"exponential": Exponential(1.3, 5.0, tau2), "matern32": Matern(1.5, 1.3, 5.0, tau2),
                "matern52": Matern(2.5, 1.3, 5.0, tau2)}[kernel]

    spec = mk(0.07)
    obj = pyNNGP.NNGP(s, y, 0.0, "S=T", m, spec)
    mean, var = obj.predict(tn)
    mean0, var0, tab0 = numpy_kriging(s, y, tn, m, KIDS[kernel], 1.3, 5.0, 0.07)
    np.testing.assert_allclose(mean, mean0, rtol=1e-9, atol=1e-10)
    np.testing.assert_allclose(var, var0, rtol=1e-9)
    # the neighbour sets themselves, bit-exact, through the C ABI (grid and brute force)
    for algo in ("grid", "brute"):
        e = engine(np.concatenate([s, tn]), np.zeros(n + len(tn)))
        e.set_knn_tuning(1.0, 256)
        e.build_neighbors_capped(m, n, n + len(tn), n, algo)
        assert np.array_equal(e.get_neighbor_rows(n, n + len(tn)), tab0)
    if kernel == "exponential":
        # well-conditioned family: a site that coincides with a reference site reproduces it as the
        # nugget vanishes
        mu, v = pyNNGP.NNGP(s, y, 0.0, "S=T", m, mk(1e-7)).predict(s[100:103])
        np.testing.assert_allclose(mu, y[100:103], rtol=0, atol=1e-3)
        assert (v > 0).all() and (v < 1e-5).all()


def test_predict_two_columns_and_eps():
    import pyNNGP

    s, y = synthetic(1500, 2, 9)
    y2 = np.stack([y, -2 * y], axis=1)
    eps = np.stack([np.full(1500, 0.1), np.linspace(0.0, 0.3, 1500)], axis=1)
    from pynngp_b200 import Exponential

    obj = pyNNGP.NNGP(s, y2, eps, "S=T", 8, Exponential(1.0, 6.0, 0.1))
    tn = np.random.default_rng(0).random((40, 2))
    mean, var = obj.predict(tn)
    assert mean.shape == (40, 2) and var.shape == (40, 2)
    p = obj._params()
    for c in range(2):
        m0, v0, _ = numpy_kriging(s, y2[:, c], tn, 8, 0, p[0], p[1], p[2], eps2=eps[:, c] ** 2)
        np.testing.assert_allclose(mean[:, c], m0, rtol=1e-9, atol=1e-10)
        np.testing.assert_allclose(var[:, c], v0, rtol=1e-9)


def test_save_neighbors_roundtrip_and_metropolis(tmp_path):
    import pyNNGP
    from pynngp_b200 import Matern

    s, y = synthetic(6000, 2, 14)
    spec = Matern(1.5, 1.0, 6.0, 0.1)
    a = pyNNGP.NNGP(s, y, 0.0, "S=T", 10, spec)
    path = str(tmp_path / "nbr.npy")
    a.save_neighbors(path)
    b = pyNNGP.NNGP(s, y, 0.0, "S=T", 10, spec, neighbors=path)
    assert np.array_equal(a._table, b._table) and a.loglik() == b.loglik()
    chain, trace, rate = b.metropolis(60, step=0.08, seed=3)
    assert chain.shape == (60, 3) and np.isfinite(chain).all() and (chain > 0).all()
    assert 0.0 < rate < 1.0
    assert trace[-1] == b.loglik(*chain[-1])                      # the trace is the target at the chain's states
    assert trace.max() >= a.loglik() - 1e-9 or rate > 0             # moved towards higher density or stayed
    chain2, trace2, _ = b.metropolis(60, step=0.08, seed=3)
    assert np.array_equal(chain, chain2) and np.array_equal(trace, trace2)  # deterministic given the seed


# ---- full-size properties (no oracle at this size) -----------------------------------------------
def test_cfg3_full_size_properties():
    c = CONFIGS["cfg3"]
    s, y = synthetic(c["n"], c["D"], c["seed"])
    e = engine(s, y)
    e.build_neighbors(c["m"])                                         # the quadratic brute-force kernel
    tab = e.get_neighbors()
    e.build_neighbors_grid(c["m"])                                    # the production default (class, bench.py)
    assert e.knn_used_grid()
    assert np.array_equal(e.get_neighbors(), tab)                     # grid == brute force at production size
    n, m = tab.shape
    rows = np.arange(n)[:, None]
    assert (tab[m:] >= 0).all() and (tab < rows).all()                # predecessors only
    assert all((tab[i, :i] >= 0).all() and (tab[i, i:] == -1).all() for i in range(m))  # ragged head
    d2 = ((s[tab[m:]] - s[m:, None, :]) ** 2).sum(-1)
    assert (np.diff(d2, axis=1) >= 0).all()                           # ascending distance
    # spot rows against the oracle's exact scan
    for i in (15, 16, 4097, 123456, 999999):
        assert np.array_equal(tab[i], orc.c_knn_ordered(s, m, lo=i, hi=i + 1)[i])
    whole = e.loglik(1, P0)[0]
    assert whole[2] == 0
    acc = np.zeros(3)
    for r in range(8):
        lo, hi = (n * r) // 8, (n * (r + 1)) // 8
        e.set_shard(lo, hi)
        acc += e.loglik(1, P0)[0]
    np.testing.assert_allclose(acc[:2], whole[:2], rtol=1e-12)        # shards sum to the whole
    lo, hi = 500000, 540000                                           # a slab against the oracle
    e.set_shard(lo, hi)
    st0 = orc.c_loglik(s, y, tab, 1, *P0[:3], lo=lo, hi=hi, threads=os.cpu_count() or 1)
    np.testing.assert_allclose(e.loglik(1, P0)[0][:2], st0[:2], rtol=RTOL64)


def test_cfg2_whole_search_and_likelihood_vs_oracle():
    """BASELINE.json configs[1]: n = 1e5, m = 15, 2-D exponential, search + likelihood on one GPU, the whole of
    both against the threaded C oracle (no sampling)."""
    import pyNNGP
    from pynngp_b200 import Exponential

    c = CONFIGS["cfg2"]
    s, y = synthetic(c["n"], c["D"], c["seed"])
    obj = pyNNGP.NNGP(s, y, 0.0, "S=T", c["m"], Exponential(**PARAMS))
    assert obj._engine.knn_used_grid()
    th = os.cpu_count() or 1
    want_tab = orc.c_knn_ordered(s, c["m"], threads=th)
    assert np.array_equal(obj._table, want_tab)
    st0 = orc.c_loglik(s, y, want_tab, 0, *P0[:3], threads=th)
    np.testing.assert_allclose(obj.loglik_terms(), st0[:2], rtol=RTOL64)
    o32 = pyNNGP.NNGP(s, y, 0.0, "S=T", c["m"], Exponential(**PARAMS), dtype="float32", neighbors=want_tab)
    np.testing.assert_allclose(o32.loglik_terms(), st0[:2], rtol=RTOL32)


def test_cfg3_fp32_slab_and_whole():
    """fp32 mode at cfg3's size (north_star: within 1e-4 relative of the fp64 oracle): a 40 000-row slab against the
    oracle, and the whole evaluation against the engine's own fp64 result."""
    c = CONFIGS["cfg3"]
    s, y = synthetic(c["n"], c["D"], c["seed"])
    e64 = engine(s, y)
    e64.build_neighbors_grid(c["m"])
    tab = e64.get_neighbors()
    e32 = engine(s, y, None, "float32")
    e32.set_neighbors(tab)
    whole64, whole32 = e64.loglik(1, P0)[0], e32.loglik(1, P0)[0]
    assert whole32[2] == 0
    np.testing.assert_allclose(whole32[:2], whole64[:2], rtol=RTOL32)
    lo, hi = 700000, 740000
    e32.set_shard(lo, hi)
    st0 = orc.c_loglik(s, y, tab, 1, *P0[:3], lo=lo, hi=hi, threads=os.cpu_count() or 1)
    np.testing.assert_allclose(e32.loglik(1, P0)[0][:2], st0[:2], rtol=RTOL32)


def test_cfg4_full_size():
    """BASELINE.json configs[3]: n = 1e7, m = 30, 3-D Matern 3/2 through the drop-in class at full size.  No oracle
    finishes the whole of it, so: structural invariants of the table, spot rows against the oracle's exact scan,
    shards summing to the whole, and a 20 000-row slab of the likelihood against the oracle at 1e-10."""
    import pyNNGP
    from pynngp_b200 import Matern

    c = CONFIGS["cfg4"]
    s, y = synthetic(c["n"], c["D"], c["seed"])
    n, m = c["n"], c["m"]
    obj = pyNNGP.NNGP(s, y, 0.0, "S=T", m, Matern(1.5, **PARAMS))
    eng = obj._engine
    assert eng.knn_used_grid()
    tab = obj._table
    assert tab.shape == (n, m)
    for a in range(0, n, 1 << 20):  # predecessors only, no padding past the ragged head (chunked: 3e8 entries)
        blk = tab[a:a + (1 << 20)]
        rows = np.arange(a, a + len(blk))[:, None]
        assert (blk < rows).all() and (blk[max(m - a, 0):] >= 0).all()
    assert all((tab[i, :i] >= 0).all() and (tab[i, i:] == -1).all() for i in range(m))
    rng = np.random.default_rng(0)
    pick = np.sort(rng.choice(np.arange(m, n), size=200000, replace=False))
    d2 = ((s[tab[pick]] - s[pick, None, :]) ** 2).sum(-1)
    assert (np.diff(d2, axis=1) >= 0).all()                           # ascending distance
    assert all(len(set(r.tolist())) == m for r in tab[pick[:2000]])   # no repeated neighbour
    for i in (30, 31, 4097, 5000001, n - 1):                          # spot rows: the oracle's exact scan
        assert np.array_equal(tab[i], orc.c_knn_ordered(s, m, lo=i, hi=i + 1)[i]), i
    whole = np.array(obj.loglik_batch([P0[:3]])[0])
    assert whole[2] == 0
    one = _lib.Engine(0)                                              # one device, explicit shards
    one.set_data(s, y)
    one.set_neighbors(tab)
    acc = np.zeros(3)
    for r in range(8):
        one.set_shard((n * r) // 8, (n * (r + 1)) // 8)
        acc += one.loglik(1, P0)[0]
    np.testing.assert_allclose(acc[:2], whole[:2], rtol=1e-12)        # shards sum to the whole
    lo, hi = 6000000, 6020000                                         # a slab against the oracle
    one.set_shard(lo, hi)
    st0 = orc.c_loglik(s, y, tab, 1, *P0[:3], lo=lo, hi=hi, threads=os.cpu_count() or 1)
    got = one.loglik(1, P0)[0]
    assert got[2] == 0 and st0[2] == 0
    np.testing.assert_allclose(got[:2], st0[:2], rtol=RTOL64)
    B, F = one.factors(1, P0, lo, lo + 64)                            # per-location factors of the same rows
    B0, F0 = orc.c_factors(s, y, tab, 1, *P0[:3], lo=lo, hi=lo + 64)
    np.testing.assert_allclose(F, F0, rtol=RTOL64)
    assert (np.abs(B - B0) / np.maximum(np.abs(B0).max(axis=1, keepdims=True), 1.0)).max() <= RTOL64


def test_known_answers_extended_precision(golden_dir):
    """tests/golden/kat_cfg1.npz: cfg1 on the unmodified reference's own neighbour sets, evaluated in 80-bit
    extended precision by textbook formulas (make_kat.py) -- independent of the oracle.  The engine's accessors
    and statistics must reproduce it."""
    k = np.load(os.path.join(golden_dir, "kat_cfg1.npz"))
    g = np.load(os.path.join(golden_dir, "ns_cfg1.npz"))
    c = CONFIGS["cfg1"]
    s, y = synthetic(c["n"], c["D"], c["seed"])
    e = engine(s, y, k["eps2"])
    e.build_neighbors_grid(int(k["m"]))
    tab = e.get_neighbors()
    assert np.array_equal(tab, g["Ns"])                               # the reference's own sets
    prm = np.array(list(k["params"]) + [0.0])
    for kid in (0, 1):
        st = e.loglik(kid, prm)[0]
        assert st[2] == 0
        np.testing.assert_allclose(st[:2], [k[f"k{kid}_sum_log_F"], k[f"k{kid}_sum_r2_over_F"]], rtol=RTOL64)
        np.testing.assert_allclose(e.loglik_terms(kid, *prm[:3])[:2], st[:2], rtol=0)
        for i in k["rows"]:
            i = int(i)
            CN, cc, cs = e.cov_blocks(kid, prm, i, i + 1)
            B, F = e.factors(kid, prm, i, i + 1)
            np.testing.assert_allclose(CN[0], k[f"k{kid}_CN_{i}"], rtol=1e-12, atol=1e-300)
            np.testing.assert_allclose(cc[0], k[f"k{kid}_c_{i}"], rtol=1e-12, atol=1e-300)
            np.testing.assert_allclose(cs[0], k[f"k{kid}_Cii_{i}"], rtol=1e-15)
            np.testing.assert_allclose(B[0], k[f"k{kid}_b_{i}"], rtol=0, atol=RTOL64)
            np.testing.assert_allclose(F[0], k[f"k{kid}_F_{i}"], rtol=RTOL64)


def test_injected_table_is_validated():
    """nngp_set_neighbors checks the table on the device: an entry outside [-1, i) or padding before a valid entry
    is refused (NNGP_EINVAL) instead of being gathered from."""
    s, y = synthetic(500, 2, 2)
    e = engine(s, y)
    good = orc.c_knn_ordered(s, 6)
    e.set_neighbors(good)
    for mutate in (lambda t: t.__setitem__((400, 2), 500),      # past the records
                   lambda t: t.__setitem__((400, 2), 10**9),
                   lambda t: t.__setitem__((7, 0), 7),          # itself
                   lambda t: t.__setitem__((7, 0), 300),        # a successor: not causal
                   lambda t: t.__setitem__((9, 1), -1),         # padding before a valid entry
                   lambda t: t.__setitem__((9, 1), -5)):
        bad = good.copy()
        mutate(bad)
        with pytest.raises(_lib.NNGPError):
            e.set_neighbors(bad)
    e.set_neighbors(good)
    st0 = orc.c_loglik(s, y, good, 0, *P0[:3])
    np.testing.assert_allclose(e.loglik(0, P0)[0][:2], st0[:2], rtol=RTOL64)
    import pyNNGP

    with pytest.raises(ValueError):
        pyNNGP.NNGP(s, y, 0.0, "S=T", 5, None, neighbors=good)   # m of the table != m of the object


def test_shard_window_table():
    """nngp_build_neighbors_shard: the handle searches and holds the rows of its shard only (what a rank of a
    multi-GPU run does); evaluations and per-location outputs work inside the window, nothing outside exists."""
    s, y = synthetic(30000, 2, 5)
    full = engine(s, y)
    full.build_neighbors_grid(15)
    tab = full.get_neighbors()
    for lo, hi, algo in ((0, 3000, "auto"), (3000, 9000, "auto"), (9000, 30000, "auto"), (100, 5000, "brute"), (12345, 12345, "auto")):
        e = engine(s, y)
        e.set_knn_tuning(1.0, 1024)
        e.set_shard(lo, hi)
        e.build_neighbors_shard(15, algo)
        assert e.neighbor_window() == (lo, hi - lo)
        assert np.array_equal(e.get_neighbor_rows(lo, hi), tab[lo:hi])
        full.set_shard(lo, hi)
        assert np.array_equal(e.loglik(1, P0), full.loglik(1, P0))   # same rows, same launch shape: bitwise
        if hi > lo:
            with pytest.raises(_lib.NNGPError):
                e.get_neighbors()                                      # the whole table is not here
            B, F = e.factors(1, P0, lo, min(lo + 50, hi))
            B0, F0 = full.factors(1, P0, lo, min(lo + 50, hi))
            assert np.array_equal(B, B0) and np.array_equal(F, F0)
            e.set_shard(0, 30000)
            if (lo, hi) != (0, 30000):
                with pytest.raises(_lib.NNGPError):
                    e.loglik(1, P0)                                    # a shard the table does not cover


def test_host_result_path_stress():
    """1000 back-to-back host-pointer evaluations with alternating K: the parameters travel in the kernel arguments
    (K <= 8) or through the staging copy (K > 8), the statistics come back through stamped lines in mapped host
    memory.  Every repetition must return bitwise the same numbers."""
    s, y = synthetic(4000, 2, 6)
    e = engine(s, y)
    e.build_neighbors_grid(15)
    rng = np.random.default_rng(1)
    K = 11
    prm = np.stack([rng.uniform(0.5, 2, K), rng.uniform(3, 30, K), rng.uniform(0.01, 0.5, K), np.zeros(K)], 1)
    want = {k: e.loglik(1, prm[:k]) for k in (1, 2, 5, 8, 9, 11)}
    one = e.loglik_terms(1, *prm[0, :3])
    assert one == tuple(want[1][0])
    st0 = orc.c_loglik(s, y, e.get_neighbors(), 1, *prm[0, :3])
    np.testing.assert_allclose(one[:2], st0[:2], rtol=RTOL64)
    for it in range(1000):
        k = (1, 5, 2, 11, 8, 9)[it % 6]
        assert np.array_equal(e.loglik(1, prm[:k]), want[k]), (it, k)
        if it % 7 == 0:
            assert e.loglik_terms(1, *prm[0, :3]) == one


def test_handles_reuse_pooled_memory():
    """Handles come and go with different sizes and shapes: device blocks, pinned result lines and the exp table are
    recycled between them (stream-ordered pool, per-process caches), so nothing may depend on fresh, zeroed memory."""
    rng = np.random.default_rng(5)
    for it in range(14):
        n = int(rng.integers(300, 6000))
        D = int(rng.integers(1, 4))
        m = int(rng.choice([3, 7, 10, 15, 30]))
        kid = int(rng.integers(0, 3))
        s, y = synthetic(n, D, 100 + it)
        e = engine(s, y, dtype="float64" if it % 3 else "float32")
        e.build_neighbors_grid(m)
        tab = e.get_neighbors()
        assert np.array_equal(tab, orc.c_knn_ordered(s, m)), (it, n, D, m)
        prm = np.array([[1.1, 5.0 + it, 0.05, 0.0], [0.9, 8.0, 0.2, 0.0]])
        got = e.loglik(kid, prm[: 1 + it % 2])
        for k in range(1 + it % 2):
            want = orc.c_loglik(s, y, tab, kid, *prm[k, :3])
            np.testing.assert_allclose(got[k][:2], want[:2], rtol=RTOL64 if it % 3 else 1e-3)  # (this test hunts stale memory, not fp32 precision)
            assert got[k][2] == want[2]
        if it % 4 == 0:
            e.set_data(s[: n // 2], y[: n // 2])  # the same handle again, smaller
            e.build_neighbors_grid(m)
            assert np.array_equal(e.get_neighbors(), tab[: n // 2])
        e.close()


def _need_gpus(k):
    if _lib.device_count() < k:
        pytest.skip(f"needs {k} GPUs")


def test_multi_device_handle_matches_single():
    """nngp_create_multi: ONE process, one handle over several GPUs -- skipped on a 1-GPU box.  Table,
    statistics, per-location outputs and the host-path stress must agree with the single-device engine."""
    _need_gpus(2)
    ndev = min(_lib.device_count(), 8)
    s, y = synthetic(200003, 2, 17)
    eps2 = np.linspace(0.0, 0.02, len(s))
    one = engine(s, y, eps2)
    one.build_neighbors_grid(15)
    tab = one.get_neighbors()
    multi = _lib.Engine(list(range(ndev)))
    multi.set_data(s, y, eps2)
    multi.build_neighbors_grid(15)
    assert multi.knn_used_grid() and np.array_equal(multi.get_neighbors(), tab)
    assert np.array_equal(multi.get_neighbor_rows(99990, 100010), tab[99990:100010])
    rng = np.random.default_rng(2)
    K = 11
    prm = np.stack([rng.uniform(0.5, 2, K), rng.uniform(3, 30, K), rng.uniform(0.01, 0.5, K), np.zeros(K)], 1)
    for k in (1, 3, 11):
        a, b = multi.loglik(1, prm[:k]), one.loglik(1, prm[:k])
        np.testing.assert_allclose(a[:, :2], b[:, :2], rtol=1e-12)
        assert np.array_equal(a[:, 2], b[:, 2])
    np.testing.assert_allclose(multi.loglik_terms(1, *prm[0, :3])[:2], one.loglik(1, prm[0])[0][:2], rtol=1e-12)
    B, F = multi.factors(1, prm[0], 99990, 100010)                    # rows that straddle two devices
    B0, F0 = one.factors(1, prm[0], 99990, 100010)
    assert np.array_equal(B, B0) and np.array_equal(F, F0)
    y2 = np.cos(5 * y)
    multi.set_y(y2); one.set_y(y2)
    np.testing.assert_allclose(multi.loglik(1, prm[0])[0][:2], one.loglik(1, prm[0])[0][:2], rtol=1e-12)
    multi.set_shard(1000, 150001); one.set_shard(1000, 150001)
    multi.build_neighbors_grid(15)
    np.testing.assert_allclose(multi.loglik(1, prm[0])[0][:2], one.loglik(1, prm[0])[0][:2], rtol=1e-12)
    multi.set_shard(0, len(s)); multi.set_neighbors(tab)
    want = {k: multi.loglik(1, prm[:k]) for k in (1, 5, 11)}
    for it in range(1000):                                            # the fused exchange, generation after generation
        k = (1, 5, 11)[it % 3]
        assert np.array_equal(multi.loglik(1, prm[:k]), want[k]), (it, k)
    multi.close()
    # tiny inputs: more devices than rows leaves some devices with an empty shard
    s3, y3 = synthetic(5, 2, 1)
    tiny = _lib.Engine(list(range(ndev)))
    tiny.set_data(s3, y3)
    tiny.build_neighbors_grid(3)
    st0 = orc.c_loglik(s3, y3, orc.c_knn_ordered(s3, 3), 0, *P0[:3])
    np.testing.assert_allclose(tiny.loglik(0, P0)[0][:2], st0[:2], rtol=RTOL64)


def test_class_on_all_devices_of_the_process():
    """NNGP(t, y, eps, refType, m, cov) in a plain process uses every visible GPU (devices=None) -- skipped on a
    1-GPU box; devices=<int> pins one.  Same table, same statistics, same predictions."""
    _need_gpus(2)
    import pyNNGP
    from pynngp_b200 import Matern

    s, y = synthetic(120000, 3, 23)
    spec = Matern(1.5, **PARAMS)
    every = pyNNGP.NNGP(s, y, 0.0, "S=T", 30, spec)
    single = pyNNGP.NNGP(s, y, 0.0, "S=T", 30, spec, devices=0)
    assert len(every.devices) == min(_lib.device_count(), 8) and single.devices == [0]
    assert np.array_equal(every._table, single._table)
    np.testing.assert_allclose(every.loglik_terms(), single.loglik_terms(), rtol=1e-12)
    np.testing.assert_allclose(every.loglik_batch(np.array([[1.0, 6.0, 0.1], [1.5, 9.0, 0.2]])),
                               single.loglik_batch(np.array([[1.0, 6.0, 0.1], [1.5, 9.0, 0.2]])), rtol=1e-12)
    assert np.array_equal(every._Bsi(70000), single._Bsi(70000))
    tn = np.random.default_rng(3).random((50, 3))
    for a, b in zip(every.predict(tn), single.predict(tn)):
        assert np.array_equal(a, b)
    sub = pyNNGP.NNGP(s[:9000], y[:9000], 0.0, ("subset", 3000), 10, spec, seed=1)
    sub1 = pyNNGP.NNGP(s[:9000], y[:9000], 0.0, ("subset", 3000), 10, spec, seed=1, devices=0)
    np.testing.assert_allclose(sub.loglik_terms(), sub1.loglik_terms(), rtol=1e-12)


def test_two_gpu_sharded_equals_single():
    """NCCL path: skipped on a 1-GPU box; `gpurun --gpus 2` exercises it."""
    import subprocess
    import sys

    import torch

    if torch.cuda.device_count() < 2:
        pytest.skip("needs 2 GPUs")
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    r = subprocess.run([sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
                        "--master-addr", "127.0.0.1", "--master-port", "29631",
                        os.path.join(root, "tools", "check_multi_gpu.py"), "cfg2"],
                       capture_output=True, text=True, timeout=600)
    assert r.returncode == 0, r.stdout[-2000:] + r.stderr[-4000:]
    assert "multi-gpu ok" in r.stdout


def test_ws_warm_start_matches_sklearn():
    """`ws` (nngp.py:45-47) = KNeighborsRegressor(5, 'uniform').fit(t, y).predict(s) with S = T; the
    plain k-NN table itself is checked against an exact numpy scan with the (d2, j) order."""
    import pyNNGP
    from sklearn.neighbors import KNeighborsRegressor

    for D, n in ((2, 3000), (3, 1500), (1, 700)):
        s, y = synthetic(n, D, 60 + D)
        obj = pyNNGP.NNGP(s, y, 0.0, "S=T", 4, None)
        want = KNeighborsRegressor(n_neighbors=5, weights="uniform").fit(s, y).predict(s)
        np.testing.assert_allclose(obj.ws, want, rtol=1e-13, atol=1e-15)
        idx = obj._engine.knn_plain(5)
        for i in (0, 1, n // 2, n - 1):
            d2 = orc.np_dist2(s[i], s)
            assert np.array_equal(idx[i], np.lexsort((np.arange(n), d2))[:5])
            assert idx[i, 0] == i  # the site itself comes first (distance 0)
    # 2-column response as in the reference's own test (tests/test_init.py:11)
    s, y = synthetic(500, 2, 9)
    y2 = np.stack([y, -2 * y], axis=1)
    obj = pyNNGP.NNGP(s, y2, 0.0, "S=T", 3, None)
    want = KNeighborsRegressor(n_neighbors=5).fit(s, y2).predict(s)
    np.testing.assert_allclose(obj.ws, want, rtol=1e-13, atol=1e-15)
