"""GPU tests of the reference sets other than T (nngp.py:32-40, 68-71) through the drop-in
class on the real engine.  The assertions are the ones tests/test_host_logic.py runs on the CPU against the
oracle-backed engine (tests/host_checks.py): Nt equals the reference's own KDTree(s).query call, ws its
KNeighborsRegressor call, Ns the ordered search on s, the likelihood the oracle's on the engine's row
layout, all within the tolerances of tests/test_gpu_parity.py."""
import numpy as np
import pytest

import host_checks as hc
from oracle import nngp_oracle as orc
from pynngp_b200 import Exponential, Matern
from pynngp_b200.synthetic import synthetic

pytestmark = pytest.mark.gpu

PRM = (1.3, 5.0, 0.07)
SPECS = [Exponential(), Matern(1.5), Matern(2.5)]


def make_with(spec):
    import pyNNGP

    def _make(t, y, eps, refType, m, **kw):
        return pyNNGP.NNGP(t, y, eps, refType, m, spec, **kw)

    return _make


# n_ref above 4096 sends the rows of s through the grid search, below it through the brute-force kernel
@pytest.mark.parametrize("n,D,n_ref,m,kernel_id", [(12000, 2, 6000, 10, 1), (5000, 3, 1500, 15, 0), (900, 1, 300, 4, 2),
                                                   (9000, 2, 8999, 30, 0)])
def test_subset_reference_set(n, D, n_ref, m, kernel_id):
    t, y = synthetic(n, D, 30 + D)
    obj = hc.check_subset(make_with(SPECS[kernel_id]), t, y, n_ref, m, kernel_id, PRM)
    assert obj._engine.launch_count() >= 2


def test_subset_with_eps_and_fp32():
    t, y = synthetic(3000, 2, 77)
    eps = np.linspace(0.05, 0.4, 3000)
    obj = hc.check_subset(make_with(Exponential()), t, y, 1000, 8, 0, PRM, eps=eps)
    import pyNNGP

    o32 = pyNNGP.NNGP(t, y, eps, ("subset", 1000), 8, Exponential(*PRM), seed=7, dtype="float32")
    assert np.array_equal(o32._table, obj._table)  # stage 1 is always fp64
    np.testing.assert_allclose(o32.loglik_terms(), obj.loglik_terms(*PRM), rtol=1e-4)


def test_subset_of_everything_is_the_dense_gp():
    t, y = synthetic(24, 2, 8)
    for kid in (0, 1, 2):
        hc.check_subset_equals_dense_gp(make_with(SPECS[kid]), t, y, kid, PRM)


@pytest.mark.parametrize("n,D,n_ref,m,kernel_id", [(7000, 2, 5000, 12, 1), (2000, 3, 600, 6, 0)])
def test_random_reference_set(n, D, n_ref, m, kernel_id):
    t, y = synthetic(n, D, 12 + D)
    hc.check_random(make_with(SPECS[kernel_id]), t, y, n_ref, m, kernel_id, PRM)


def test_ref_type_errors_and_injected_table():
    t, y = synthetic(400, 2, 5)
    hc.check_ref_type_errors(make_with(None), t, y)
    import pyNNGP

    a = pyNNGP.NNGP(t, y, 0.0, ("subset", 150), 5, Matern(1.5, *PRM), seed=2)
    b = pyNNGP.NNGP(t, y, 0.0, ("subset", 150), 5, Matern(1.5, *PRM), seed=2, neighbors=a._table)
    assert a.loglik() == b.loglik()
    slog, squad, _ = orc.c_loglik(t[a._rows], y[a._rows], a._table, 1, *PRM)
    np.testing.assert_allclose(a.loglik_terms(), (slog, squad), rtol=1e-10)


@pytest.mark.parametrize("n,D,m,kernel_id", [(1500, 2, 10, 1), (700, 3, 6, 0), (400, 1, 4, 0), (600, 2, 5, 2)])
def test_latent_density(n, D, m, kernel_id):
    """loglik_latent: reference rows without a nugget, observation rows with tau2 + eps^2 through nngp_set_eps2
    (the records' z slot for D < 3, its own array for D = 3)."""
    t, y = synthetic(n, D, 31 + D)
    hc.check_latent_density(make_with(SPECS[kernel_id]), t, y, m, kernel_id, PRM)


def test_latent_density_dense_identity():
    t, y = synthetic(20, 2, 33)
    for kid in (0, 1):
        hc.check_latent_dense_identity(make_with(SPECS[kid]), t, y, kid, PRM)
