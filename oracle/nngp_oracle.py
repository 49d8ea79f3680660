"""CPU ORACLE for the NNGP likelihood hot path -- TEST INFRASTRUCTURE ONLY.

Only ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s ``cpu_baseline`` / ``--impl reference``
legs may import this module; the product package ``pynngp_b200`` never does.

Two restatements of the same algorithm live here:

* ``np_*``  -- plain numpy, one function per reference symbol, following the *structure* of the
  reference's per-location accessors (``pyNNGP/nngp.py:73-96``) and of its ordered neighbour search
  (``pyNNGP/nngp.py:49-62``).  Slow; for small cases.
* ``c_*``   -- ctypes wrappers over ``oracle/nngp_oracle.c`` (same algorithm in C) for sizes up to
  1e5..1e6 locations.  ``tests/test_oracle.py`` checks ``c_*`` == ``np_*``.

Pinning: stage 1 (neighbour sets) is pinned bit-exactly to the unmodified reference through
``tests/golden/ns_*.npz`` (made by ``tests/golden/make_golden.py``, which imports /root/reference).
Stages 2-3 (C_N, c, b, F, log-likelihood) are PARITY UNPINNED by the reference -- ``nngp.py:73-96``
are empty stubs with no test vectors -- and are anchored to the dense-GP identity, to the product of the
parent GP's conditionals evaluated by SciPy's multivariate normal (m < n - 1), and to closed forms.
"""
from __future__ import annotations

import ctypes
import os
import subprocess
from concurrent.futures import ThreadPoolExecutor

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "_build", "liboracle.so")

KERNEL_IDS = {"exponential": 0, "matern32": 1, "matern52": 2}


# --------------------------------------------------------------------------------------------------
# numpy restatement
# --------------------------------------------------------------------------------------------------
def np_dist2(a, b):
    """sklearn euclidean_rdist (sklearn/metrics/_dist_metrics.pxd.tp:39-49): sequential fp64
    accumulation over dimensions, no FMA.  ``a``: (D,), ``b``: (k, D)."""
    d = np.zeros(b.shape[0], dtype=np.float64)
    for k in range(b.shape[1]):
        t = a[k] - b[:, k]
        d = d + t * t
    return d


def np_knn_ordered(s, m):
    """``_make_s_neighbor_sets`` (nngp.py:49-62): for each i the min(m, i) nearest j < i, ascending
    distance; ties on d2 broken by smaller j (the engine's documented rule; the reference's KD-tree
    leaves ties unspecified -- SURVEY 0.7).  Returns the reference's shape: a list, ``Ns[0] == []``,
    ``Ns[i]`` an int64 array."""
    s = np.ascontiguousarray(s, dtype=np.float64)
    Ns = []
    for i in range(len(s)):
        if i == 0:
            Ns.append([])
            continue
        d2 = np_dist2(s[i], s[0:i])
        order = np.lexsort((np.arange(i), d2))[: min(m, i)]
        Ns.append(order.astype(np.int64))
    return Ns


def sk_reference_stage1(s, m):
    """The reference's stage 1 AS WRITTEN (nngp.py:49-62): a scikit-learn ``KDTree(s[0:i])`` rebuilt for
    every i and queried for k = min(m, i) -- the same third-party calls, restated so that bench.py can
    time the reference's own algorithm on the GPU box's host (where /root/reference does not exist).
    tests/test_oracle.py checks it against the golden tables made by the unmodified reference."""
    from sklearn.neighbors import KDTree

    s = np.asarray(s)
    Ns = []
    for i, si in enumerate(s):
        if i == 0:
            Ns.append([])
            continue
        Ns.append(KDTree(s[0:i]).query(si.reshape(1, -1), k=min(m, i), return_distance=False)[0])
    return Ns


def ns_to_table(Ns, m):
    """list-of-arrays (reference layout) -> dense (n, m) int32 table padded with -1."""
    tab = np.full((len(Ns), m), -1, dtype=np.int32)
    for i, row in enumerate(Ns):
        tab[i, : len(row)] = row
    return tab


def np_corr(kernel_id, u):
    if kernel_id == 0:
        return np.exp(-u)
    if kernel_id == 1:
        return (1.0 + u) * np.exp(-u)
    return (1.0 + u + u * u / 3.0) * np.exp(-u)


class NumpyNNGP:
    """Per-location accessors with the reference's names (nngp.py:73-96), on numpy."""

    def __init__(self, s, y, nbr, kernel_id, sigma2, phi, tau2, eps2=None):
        self.s = np.ascontiguousarray(s, dtype=np.float64)
        self.y = np.asarray(y, dtype=np.float64)
        self.nbr = np.asarray(nbr)
        self.kernel_id, self.sigma2, self.phi, self.tau2 = kernel_id, sigma2, phi, tau2
        self.eps2 = None if eps2 is None else np.asarray(eps2, dtype=np.float64)

    def _N(self, i):
        row = self.nbr[i]
        return row[row >= 0]

    def _cov(self, a, b):
        d = np.sqrt(np_dist2(self.s[a], self.s[b][None, :])[0])
        return self.sigma2 * np_corr(self.kernel_id, self.phi * d)

    def _Cs(self, i):  # nngp.py:92-96
        return self.sigma2 + self.tau2 + (0.0 if self.eps2 is None else self.eps2[i])

    def _CNs(self, i):  # nngp.py:78-82
        N = self._N(i)
        C = np.empty((len(N), len(N)))
        for a in range(len(N)):
            for b in range(len(N)):
                C[a, b] = self._Cs(N[a]) if a == b else self._cov(N[a], N[b])
        return C

    def _Ccross(self, i):  # nngp.py:84-86
        return np.array([self._cov(i, j) for j in self._N(i)], dtype=np.float64)

    def _Bsi(self, i):  # nngp.py:73-76
        N = self._N(i)
        if len(N) == 0:
            return np.zeros(0)
        L = np.linalg.cholesky(self._CNs(i))
        z = np.linalg.solve(L, self._Ccross(i))
        return np.linalg.solve(L.T, z)

    def _Fsi(self, i):  # nngp.py:88-90
        return self._Cs(i) - float(self._Ccross(i) @ self._Bsi(i))

    def loglik_terms(self, lo=0, hi=None):
        hi = len(self.s) if hi is None else hi
        slog = squad = 0.0
        for i in range(lo, hi):
            N = self._N(i)
            b, F = self._Bsi(i), self._Fsi(i)
            r = self.y[i] - float(b @ self.y[N])
            slog += np.log(F)
            squad += r * r / F
        return slog, squad


def np_krige(s, y, tn, m, kernel_id, sigma2, phi, tau2, eps2=None):
    """Kriging at new sites `tn` from the m nearest reference sites, (d2, j) order -- the step after the
    path (the reference names it only through oneSample's undefined update_y_unobserved,
    nngp.py:98-101).  mean = c^T C_N^-1 y_N, var = sigma2 + tau2 - c^T C_N^-1 c, C_N with the same
    diagonal as _CNs (nngp.py:78-82 restated above).  Returns (mean, var, neighbour table)."""
    s = np.asarray(s, dtype=np.float64)
    if s.ndim == 1:
        s = s[:, None]
    tn = np.asarray(tn, dtype=np.float64).reshape(-1, s.shape[1])
    mean, var, tabs = [], [], []
    for t in tn:
        nb = np.lexsort((np.arange(len(s)), np_dist2(t, s)))[:m]
        dn = np.sqrt(((s[nb][:, None, :] - s[nb][None, :, :]) ** 2).sum(-1))
        CN = sigma2 * np_corr(kernel_id, phi * dn)
        CN[np.diag_indices(len(nb))] = sigma2 + tau2 + (0.0 if eps2 is None else np.asarray(eps2)[nb])
        c = sigma2 * np_corr(kernel_id, phi * np.sqrt(((s[nb] - t) ** 2).sum(-1)))
        b = np.linalg.solve(CN, c)
        mean.append(b @ np.asarray(y)[nb])
        var.append(sigma2 + tau2 - c @ b)
        tabs.append(nb)
    return np.array(mean), np.array(var), np.array(tabs)


def loglik_from_terms(slog, squad, n):
    """log N(y; 0, C_nngp) from the north_star reduction."""
    return -0.5 * (slog + squad) - 0.5 * n * np.log(2.0 * np.pi)


def dense_gp_loglik(s, y, kernel_id, sigma2, phi, tau2, eps2=None):
    """Exact zero-mean GP log density (the known answer the NNGP equals when m >= n-1)."""
    s = np.asarray(s, dtype=np.float64)
    n = len(s)
    d = np.sqrt(((s[:, None, :] - s[None, :, :]) ** 2).sum(-1))
    C = sigma2 * np_corr(kernel_id, phi * d)
    C[np.diag_indices(n)] = sigma2 + tau2 + (0.0 if eps2 is None else np.asarray(eps2))
    L = np.linalg.cholesky(C)
    z = np.linalg.solve(L, np.asarray(y, dtype=np.float64))
    return -np.log(np.diag(L)).sum() - 0.5 * float(z @ z) - 0.5 * n * np.log(2.0 * np.pi)


# --------------------------------------------------------------------------------------------------
# C restatement through ctypes
# --------------------------------------------------------------------------------------------------
_lib = None


def build():
    """compile oracle/nngp_oracle.c -> oracle/_build/liboracle.so (gcc via oracle/Makefile)."""
    subprocess.run(["make", "-s", "-C", _HERE], check=True)


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        lib = ctypes.CDLL(_LIB_PATH)
        dp, ip = ctypes.POINTER(ctypes.c_double), ctypes.POINTER(ctypes.c_int32)
        i64, i32 = ctypes.c_int64, ctypes.c_int
        lib.oracle_knn_ordered.argtypes = [dp, i64, i32, i32, i64, i64, ip]
        lib.oracle_knn_ordered.restype = None
        lib.oracle_loglik.argtypes = [dp, dp, dp, ip, i64, i32, i32, i32, dp, i64, i64, dp]
        lib.oracle_loglik.restype = None
        lib.oracle_factors.argtypes = [dp, dp, dp, ip, i64, i32, i32, i32, dp, i64, i64, dp, dp]
        lib.oracle_factors.restype = None
        lib.oracle_cov_blocks.argtypes = [dp, dp, ip, i64, i32, i32, i32, dp, i64, i64, dp, dp, dp]
        lib.oracle_cov_blocks.restype = None
        lib.oracle_max_m.restype = i32
        _lib = lib
    return _lib


def _dp(a):
    return None if a is None else a.ctypes.data_as(ctypes.POINTER(ctypes.c_double))


def _ip(a):
    return a.ctypes.data_as(ctypes.POINTER(ctypes.c_int32))


def _prep(s, y=None, eps2=None, nbr=None):
    s = np.ascontiguousarray(s, dtype=np.float64)
    y = None if y is None else np.ascontiguousarray(y, dtype=np.float64)
    eps2 = None if eps2 is None else np.ascontiguousarray(eps2, dtype=np.float64)
    nbr = None if nbr is None else np.ascontiguousarray(nbr, dtype=np.int32)
    return s, y, eps2, nbr


def _chunks(lo, hi, threads):
    edges = np.linspace(lo, hi, threads + 1).astype(np.int64)
    return [(int(a), int(b)) for a, b in zip(edges[:-1], edges[1:]) if b > a]


def c_knn_ordered(s, m, lo=0, hi=None, threads=1):
    """dense (n, m) int32 neighbour table (rows outside [lo, hi) are -1)."""
    lib = _load()
    s, _, _, _ = _prep(s)
    n, D = s.shape
    hi = n if hi is None else hi
    out = np.full((n, m), -1, dtype=np.int32)
    if threads <= 1:
        lib.oracle_knn_ordered(_dp(s), n, D, m, lo, hi, _ip(out))
    else:
        # work per row grows with i: many small interleaved chunks balance the threads
        parts = _chunks(lo, hi, threads * 16)
        with ThreadPoolExecutor(threads) as ex:
            list(ex.map(lambda ab: lib.oracle_knn_ordered(_dp(s), n, D, m, ab[0], ab[1], _ip(out)),
                        parts))
    return out


def c_knn_rows(s, m, rows, threads=1):
    """(len(rows), m) int32: the neighbour sets of the listed rows only (each an exact O(i) scan); nothing of
    size n x m is allocated -- the C routine is handed the address row 0 of a full table would have."""
    lib = _load()
    s, _, _, _ = _prep(s)
    n, D = s.shape
    rows = np.asarray(rows, dtype=np.int64)
    out = np.full((len(rows), m), -1, dtype=np.int32)

    def one(k):
        i = int(rows[k])
        base = ctypes.cast(out.ctypes.data + (k - i) * m * 4, ctypes.POINTER(ctypes.c_int32))
        lib.oracle_knn_ordered(_dp(s), n, D, m, i, i + 1, base)

    with ThreadPoolExecutor(max(threads, 1)) as ex:
        list(ex.map(one, range(len(rows))))
    return out


def c_loglik(s, y, nbr, kernel_id, sigma2, phi, tau2, eps2=None, lo=0, hi=None, threads=1):
    """(sum log F, sum r^2/F, n_bad) over rows [lo, hi)."""
    lib = _load()
    s, y, eps2, nbr = _prep(s, y, eps2, nbr)
    n, D = s.shape
    m = nbr.shape[1]
    hi = n if hi is None else hi
    params = np.array([sigma2, phi, tau2, 0.0], dtype=np.float64)

    def one(ab):
        out = np.zeros(3, dtype=np.float64)
        lib.oracle_loglik(_dp(s), _dp(y), _dp(eps2), _ip(nbr), n, D, m, kernel_id, _dp(params),
                          ab[0], ab[1], _dp(out))
        return out

    if threads <= 1:
        return tuple(one((lo, hi)))
    with ThreadPoolExecutor(threads) as ex:
        parts = list(ex.map(one, _chunks(lo, hi, threads)))
    return tuple(np.sum(parts, axis=0))


def c_loglik_rows(s, y, nbr_rows, row0, kernel_id, sigma2, phi, tau2, eps2=None, threads=1):
    """c_loglik over rows [row0, row0 + len(nbr_rows)) when only those rows of the table are at hand (a slab of
    a table too large to hold whole, e.g. n = 1e7): the C routine indexes the table by absolute row, so it is
    handed the address row 0 would have."""
    lib = _load()
    s, y, eps2, nbr_rows = _prep(s, y, eps2, nbr_rows)
    n, D = s.shape
    cnt, m = nbr_rows.shape
    params = np.array([sigma2, phi, tau2, 0.0], dtype=np.float64)
    base = ctypes.cast(nbr_rows.ctypes.data - int(row0) * m * 4, ctypes.POINTER(ctypes.c_int32))

    def one(ab):
        out = np.zeros(3, dtype=np.float64)
        lib.oracle_loglik(_dp(s), _dp(y), _dp(eps2), base, n, D, m, kernel_id, _dp(params), ab[0], ab[1], _dp(out))
        return out

    with ThreadPoolExecutor(max(threads, 1)) as ex:
        parts = list(ex.map(one, _chunks(row0, row0 + cnt, max(threads, 1))))
    return tuple(np.sum(parts, axis=0))


def c_factors(s, y, nbr, kernel_id, sigma2, phi, tau2, eps2=None, lo=0, hi=None):
    lib = _load()
    s, y, eps2, nbr = _prep(s, y, eps2, nbr)
    n, D = s.shape
    m = nbr.shape[1]
    hi = n if hi is None else hi
    params = np.array([sigma2, phi, tau2, 0.0], dtype=np.float64)
    B = np.zeros((hi - lo, m), dtype=np.float64)
    F = np.zeros(hi - lo, dtype=np.float64)
    lib.oracle_factors(_dp(s), _dp(y), _dp(eps2), _ip(nbr), n, D, m, kernel_id, _dp(params), lo, hi,
                       _dp(B), _dp(F))
    return B, F


def c_cov_blocks(s, nbr, kernel_id, sigma2, phi, tau2, eps2=None, lo=0, hi=None):
    lib = _load()
    s, _, eps2, nbr = _prep(s, None, eps2, nbr)
    n, D = s.shape
    m = nbr.shape[1]
    hi = n if hi is None else hi
    params = np.array([sigma2, phi, tau2, 0.0], dtype=np.float64)
    CN = np.zeros((hi - lo, m, m), dtype=np.float64)
    cc = np.zeros((hi - lo, m), dtype=np.float64)
    cs = np.zeros(hi - lo, dtype=np.float64)
    lib.oracle_cov_blocks(_dp(s), _dp(eps2), _ip(nbr), n, D, m, kernel_id, _dp(params), lo, hi,
                          _dp(CN), _dp(cc), _dp(cs))
    return CN, cc, cs
