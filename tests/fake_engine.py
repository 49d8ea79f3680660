"""A stand-in for ``pynngp_b200._lib.Engine`` backed by the CPU oracle -- TEST INFRASTRUCTURE ONLY.

The CPU suite (``-m "not gpu"``) cannot call the CUDA library, so the host logic of
``pynngp_b200/nngp.py`` -- row layouts for S = T and S != T, response columns, eps, table injection,
prediction bookkeeping, the Metropolis driver -- is exercised here against an engine with the same
methods and the same contracts (include/nngp_b200.h) whose arithmetic is the oracle's.  The product
never sees this class: the tests monkeypatch it in.
"""
from __future__ import annotations

import numpy as np

from oracle import nngp_oracle as orc

ROW_UNSET = -2


def knn_capped(s, m, row_lo, row_hi, cand_cap):
    """Rows [row_lo, row_hi): the min(m, jlim) nearest j < jlim = min(i, cand_cap) in ascending (d2, j);
    -1 padded (the contract of nngp_build_neighbors_capped)."""
    out = np.full((row_hi - row_lo, m), -1, dtype=np.int32)
    for i in range(row_lo, row_hi):
        jlim = min(i, cand_cap)
        if jlim == 0:
            continue
        d2 = orc.np_dist2(s[i], s[:jlim])
        best = np.lexsort((np.arange(jlim), d2))[: min(m, jlim)]
        out[i - row_lo, : len(best)] = best
    return out


class FakeEngine:
    launches = 0

    def __init__(self, device=0, dtype="float64"):
        if dtype not in ("float64", "float32"):
            raise ValueError("dtype must be 'float64' or 'float32'")
        # a list of devices is a multi-device handle (nngp_create_multi): the same contract, totals returned
        self.devices = [int(d) for d in device] if isinstance(device, (list, tuple)) else [int(device)]
        self.device, self.dtype = self.devices[0], str(dtype)
        self.n = self.D = self.m = 0
        self._tab = None
        self._win = (0, 0)  # rows of the table this handle holds
        self._closed = False

    def close(self):
        self._closed = True

    # -- data ----------------------------------------------------------------------------------------
    def set_data(self, coords, y, eps2=None):
        coords = np.ascontiguousarray(coords, dtype=np.float64)
        assert coords.ndim == 2
        self._s, self._y = coords, np.ascontiguousarray(y, dtype=np.float64)
        assert self._y.shape == (len(coords),)
        self._eps2 = None if eps2 is None else np.ascontiguousarray(eps2, dtype=np.float64)
        self.n, self.D, self.m = coords.shape[0], coords.shape[1], 0
        self._tab = None
        self._shard = (0, self.n)  # nngp_set_data resets the shard and drops the table

    def set_y(self, y):
        y = np.ascontiguousarray(y, dtype=np.float64)
        assert y.shape == (self.n,)
        self._y = y

    def set_eps2(self, eps2):
        eps2 = np.ascontiguousarray(eps2, dtype=np.float64)
        assert eps2.shape == (self.n,)
        self._eps2 = eps2

    def set_shard(self, lo, hi):
        assert 0 <= lo <= hi <= self.n
        self._shard = (int(lo), int(hi))

    # -- stage 1 -------------------------------------------------------------------------------------
    def build_neighbors(self, m, tile_offset=0, tile_stride=1):
        assert tile_stride == 1
        self.build_neighbors_grid(m)

    def build_neighbors_grid(self, m, row_lo=0, row_hi=None, algo="auto"):
        self.build_neighbors_capped(m, row_lo, self.n if row_hi is None else row_hi, self.n, algo)

    def build_neighbors_capped(self, m, row_lo, row_hi, cand_cap, algo="auto"):
        assert algo in ("auto", "grid", "brute") and 1 <= m <= 32 and cand_cap >= 1
        tab = np.full((self.n, m), ROW_UNSET, dtype=np.int32)  # rows outside the range are unset
        tab[row_lo:row_hi] = knn_capped(self._s, m, row_lo, row_hi, cand_cap)
        self._tab, self.m, self._win = tab, int(m), (0, self.n)
        FakeEngine.launches += 1

    def build_neighbors_shard(self, m, algo="auto"):
        """nngp_build_neighbors_shard: the rows of the shard only; nothing else is held."""
        lo, hi = self._shard
        self.build_neighbors_capped(m, lo, hi, self.n, algo)
        self._win = (lo, hi - lo)

    def neighbor_window(self):
        return self._win if self._tab is not None else (self._win[0], 0)

    def _held(self, i0, i1):
        r0, rows = self._win
        if i1 > i0 and not (r0 <= i0 and i1 <= r0 + rows):
            raise RuntimeError("the table of this handle holds the rows of its shard only")

    def set_neighbors(self, table):
        table = np.ascontiguousarray(table, dtype=np.int32)
        if table.ndim != 2 or table.shape[0] != self.n:
            raise ValueError("neighbour table must be (n, m) int32")
        # the device-side check of nngp_set_neighbors: entries of row i in [-1, i), padding at the tail
        rows = np.arange(self.n)[:, None]
        pad_then_valid = ((table[:, :-1] == -1) & (table[:, 1:] != -1)).any() if table.shape[1] > 1 else False
        if (table < -1).any() or (table >= rows).any() or pad_then_valid:
            raise RuntimeError("nngp_set_neighbors failed (1): invalid neighbour table")
        self._tab, self.m, self._win = table.copy(), table.shape[1], (0, self.n)

    def get_neighbors(self):
        self._held(0, self.n)
        return self._tab.copy()

    def get_neighbor_rows(self, i0, i1):
        self._held(i0, i1)
        return self._tab[i0:i1].copy()

    def knn_plain(self, k):
        out = np.empty((self.n, k), dtype=np.int32)
        for i in range(self.n):
            out[i] = np.lexsort((np.arange(self.n), orc.np_dist2(self._s[i], self._s)))[:k]
        return out

    # -- stages 2-3 ------------------------------------------------------------------------------------
    def _clean(self):
        assert self._tab is not None, "neighbours not set"
        return np.where(self._tab < 0, -1, self._tab).astype(np.int32)

    def loglik(self, kernel_id, params):
        params = np.atleast_2d(np.asarray(params, dtype=np.float64))
        assert params.shape[1] == 4
        lo, hi = self._shard
        self._held(lo, hi)
        FakeEngine.launches += 1
        return np.array([orc.c_loglik(self._s, self._y, self._clean(), kernel_id, p[0], p[1], p[2], eps2=self._eps2,
                                      lo=lo, hi=hi) for p in params])

    def loglik_terms(self, kernel_id, sigma2, phi, tau2):
        st = self.loglik(kernel_id, [sigma2, phi, tau2, 0.0])[0]
        return float(st[0]), float(st[1]), float(st[2])

    def factors(self, kernel_id, params, i0=0, i1=None, want_B=True, want_F=True):
        i1 = self.n if i1 is None else i1
        p = np.asarray(params, dtype=np.float64)
        B, F = orc.c_factors(self._s, self._y, self._clean(), kernel_id, p[0], p[1], p[2], eps2=self._eps2, lo=i0, hi=i1)
        return (B if want_B else None), (F if want_F else None)

    def cov_blocks(self, kernel_id, params, i0=0, i1=None):
        i1 = self.n if i1 is None else i1
        p = np.asarray(params, dtype=np.float64)
        return orc.c_cov_blocks(self._s, self._clean(), kernel_id, p[0], p[1], p[2], eps2=self._eps2, lo=i0, hi=i1)

    def launch_count(self):
        return FakeEngine.launches
