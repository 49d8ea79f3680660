"""Known-answer snapshot for stages 2-3 at cfg1 ("seeded cfg1 snapshot: Ns, Q, a few b_i / F_i").

The reference only names these quantities (pyNNGP/nngp.py:73-96 are stubs), so there is nothing upstream to
record.  What is stored instead is an answer computed INDEPENDENTLY of the oracle and of the CUDA kernels:

  * the neighbour sets are the unmodified reference's own (tests/golden/ns_cfg1.npz, made by make_golden.py);
  * every number is evaluated in 80-bit extended precision (numpy longdouble, 64-bit mantissa) with textbook
    formulas -- distances, the exponential / Matern 3/2 correlation, an unblocked Cholesky, two triangular
    solves -- sharing no code, no elimination order and no rounding with oracle/ or csrc/.

Stored (tests/golden/kat_cfg1.npz), for the exponential kernel (cfg1's) and Matern 3/2, at BASELINE's parameters
(sigma2 = 1, phi = 6, tau2 = 0.1) with a per-observation eps: sum log F, sum r^2 / F over all 1000 rows, and for
six rows (0, 1, 5, 10, 500, 999) C_N, c, C_ii, b_i, F_i, r_i, all rounded to float64.
    python tests/golden/make_kat.py
"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic  # noqa: E402

LD = np.longdouble
ROWS = (0, 1, 5, 10, 500, 999)


def corr(kernel, u):
    e = np.exp(-u)
    return e if kernel == "exponential" else (LD(1) + u) * e


def chol_solve(A, b):
    """x = A^-1 b by an unblocked Cholesky in extended precision."""
    p = len(b)
    L = np.zeros((p, p), dtype=LD)
    for j in range(p):
        L[j, j] = np.sqrt(A[j, j] - np.dot(L[j, :j], L[j, :j]))
        for i in range(j + 1, p):
            L[i, j] = (A[i, j] - np.dot(L[i, :j], L[j, :j])) / L[j, j]
    z = np.zeros(p, dtype=LD)
    for i in range(p):
        z[i] = (b[i] - np.dot(L[i, :i], z[:i])) / L[i, i]
    x = np.zeros(p, dtype=LD)
    for i in range(p - 1, -1, -1):
        x[i] = (z[i] - np.dot(L[i + 1:, i], x[i + 1:])) / L[i, i]
    return x


def location(s, y, eps2, nbrs, i, kernel, sigma2, phi, tau2):
    p = len(nbrs)
    cii = sigma2 + tau2 + eps2[i]
    if p == 0:
        return np.zeros((0, 0), LD), np.zeros(0, LD), cii, np.zeros(0, LD), cii, y[i]
    pts = s[nbrs]
    d = np.sqrt(((pts[:, None, :] - pts[None, :, :]) ** 2).sum(-1))
    CN = sigma2 * corr(kernel, phi * d)
    CN[np.diag_indices(p)] = sigma2 + tau2 + eps2[nbrs]
    c = sigma2 * corr(kernel, phi * np.sqrt(((pts - s[i]) ** 2).sum(-1)))
    b = chol_solve(CN, c)
    F = cii - np.dot(c, b)
    r = y[i] - np.dot(b, y[nbrs])
    return CN, c, cii, b, F, r


def main():
    c = CONFIGS["cfg1"]
    g = np.load(os.path.join(HERE, "ns_cfg1.npz"))
    s64, y64 = synthetic(c["n"], c["D"], c["seed"])
    assert np.array_equal(s64, g["coords"])
    tab = g["Ns"]  # the reference's own neighbour sets
    n, m = tab.shape
    eps2_64 = np.linspace(0.0, 0.02, n)
    s, y, eps2 = s64.astype(LD), y64.astype(LD), eps2_64.astype(LD)
    sigma2, phi, tau2 = LD(PARAMS["sigma2"]), LD(PARAMS["phi"]), LD(PARAMS["tau2"])
    out = {"rows": np.array(ROWS), "eps2": eps2_64, "m": m, "params": np.array([PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"]])}
    for kid, kernel in ((0, "exponential"), (1, "matern32")):
        slog, squad = LD(0), LD(0)
        for i in range(n):
            nbrs = tab[i][tab[i] >= 0]
            CN, cc, cii, b, F, r = location(s, y, eps2, nbrs, i, kernel, sigma2, phi, tau2)
            slog += np.log(F)
            squad += r * r / F
            if i in ROWS:
                p = len(nbrs)
                CNp = np.zeros((m, m)); CNp[:p, :p] = CN.astype(np.float64)
                ccp = np.zeros(m); ccp[:p] = cc.astype(np.float64)
                bp = np.zeros(m); bp[:p] = b.astype(np.float64)
                out[f"k{kid}_CN_{i}"], out[f"k{kid}_c_{i}"], out[f"k{kid}_b_{i}"] = CNp, ccp, bp
                out[f"k{kid}_Cii_{i}"], out[f"k{kid}_F_{i}"], out[f"k{kid}_r_{i}"] = float(cii), float(F), float(r)
        out[f"k{kid}_sum_log_F"], out[f"k{kid}_sum_r2_over_F"] = float(slog), float(squad)
        print(kernel, float(slog), float(squad))
    np.savez_compressed(os.path.join(HERE, "kat_cfg1.npz"), **out)
    print("mantissa bits:", np.finfo(LD).nmant)


if __name__ == "__main__":
    main()
