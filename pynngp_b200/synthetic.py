"""Seeded synthetic spatial workloads; tests and bench.py draw inputs from here so
the CUDA path, the oracle and the golden fixtures all see identical arrays."""
import numpy as np

# name -> (n, D, m, kernel, seed); BASELINE.json `configs` in order
CONFIGS = {
    "cfg1": dict(n=1_000, D=2, m=10, kernel="exponential", seed=1),
    "cfg2": dict(n=100_000, D=2, m=15, kernel="exponential", seed=2),
    "cfg3": dict(n=1_000_000, D=2, m=15, kernel="matern32", seed=3),
    "cfg4": dict(n=10_000_000, D=3, m=30, kernel="matern32", seed=4),
}
PARAMS = dict(sigma2=1.0, phi=6.0, tau2=0.1)


def synthetic(n, D, seed):
    """coords uniform on [0,1)^D in generated order (the reference imposes no ordering,
    nngp.py:31,51); y = smooth field + 0.3 N(0,1)."""
    rng = np.random.default_rng(seed)
    s = rng.random((n, D))
    f = np.sin(2 * np.pi * s[:, 0])
    if D > 1:
        f = f * np.cos(2 * np.pi * s[:, 1])
    if D > 2:
        f = f * np.cos(2 * np.pi * s[:, 2])
    y = f + 0.3 * rng.standard_normal(n)
    return s, y


def sweep_params(K, seed=5):
    """cfg5: K parameter vectors (sigma2, phi, tau2, nu=0) for the MCMC-like sweep."""
    rng = np.random.default_rng(seed)
    out = np.zeros((K, 4))
    out[:, 1] = rng.uniform(3.0, 30.0, K)
    out[:, 0] = rng.uniform(0.5, 2.0, K)
    out[:, 2] = rng.uniform(0.01, 0.5, K)
    return out
