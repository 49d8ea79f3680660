"""ctypes binding of libnngp_b200.so (C ABI: include/nngp_b200.h).

There is no CPU fallback: if the shared library is missing or no B200 is present, loading /
engine creation raises.  Build with ``python -m pynngp_b200.build``.
"""
from __future__ import annotations

import ctypes
import os

import numpy as np

_PKG = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_PKG, "_C", "libnngp_b200.so")

F64, F32 = 0, 1
NPARAM, NSTAT = 4, 3
MAX_M, MAX_D = 32, 3
KNN_TILE = 128
ROW_UNSET = -2
IPC_HANDLE_BYTES = 64

_c_double_p = ctypes.POINTER(ctypes.c_double)
_c_int32_p = ctypes.POINTER(ctypes.c_int32)
_handle_p = ctypes.c_void_p

# name -> (restype, argtypes); every symbol include/nngp_b200.h declares
SIGNATURES = {
    "nngp_version": (ctypes.c_char_p, []),
    "nngp_last_error": (ctypes.c_char_p, [_handle_p]),
    "nngp_create": (ctypes.c_int, [ctypes.POINTER(_handle_p), ctypes.c_int, ctypes.c_int]),
    "nngp_create_multi": (ctypes.c_int, [ctypes.POINTER(_handle_p), ctypes.POINTER(ctypes.c_int), ctypes.c_int, ctypes.c_int]),
    "nngp_device_count": (ctypes.c_int, [_handle_p]),
    "nngp_visible_devices": (ctypes.c_int, []),
    "nngp_destroy": (None, [_handle_p]),
    "nngp_set_data": (ctypes.c_int, [_handle_p, _c_double_p, ctypes.c_int64, ctypes.c_int, _c_double_p, _c_double_p]),
    "nngp_set_y": (ctypes.c_int, [_handle_p, _c_double_p]),
    "nngp_set_eps2": (ctypes.c_int, [_handle_p, _c_double_p]),
    "nngp_set_shard": (ctypes.c_int, [_handle_p, ctypes.c_int64, ctypes.c_int64]),
    "nngp_build_neighbors": (ctypes.c_int, [_handle_p, ctypes.c_int, ctypes.c_int, ctypes.c_int]),
    "nngp_build_neighbors_grid": (ctypes.c_int, [_handle_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]),
    "nngp_build_neighbors_shard": (ctypes.c_int, [_handle_p, ctypes.c_int, ctypes.c_int]),
    "nngp_neighbor_window": (ctypes.c_int, [_handle_p, ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64)]),
    "nngp_build_neighbors_capped": (ctypes.c_int, [_handle_p, ctypes.c_int, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_int]),
    "nngp_set_knn_tuning": (ctypes.c_int, [_handle_p, ctypes.c_double, ctypes.c_int64]),
    "nngp_knn_used_grid": (ctypes.c_int, [_handle_p]),
    "nngp_set_neighbors": (ctypes.c_int, [_handle_p, _c_int32_p, ctypes.c_int]),
    "nngp_get_neighbors": (ctypes.c_int, [_handle_p, _c_int32_p]),
    "nngp_get_neighbor_rows": (ctypes.c_int, [_handle_p, ctypes.c_int64, ctypes.c_int64, _c_int32_p]),
    "nngp_knn_plain": (ctypes.c_int, [_handle_p, ctypes.c_int, _c_int32_p]),
    "nngp_neighbors_device_ptr": (ctypes.c_void_p, [_handle_p]),
    "nngp_neighbor_window_device_ptr": (ctypes.c_void_p, [_handle_p]),
    "nngp_loglik_terms": (ctypes.c_int, [_handle_p, ctypes.c_int, ctypes.c_double, ctypes.c_double, ctypes.c_double, _c_double_p]),
    "nngp_loglik": (ctypes.c_int, [_handle_p, ctypes.c_int, _c_double_p, ctypes.c_int, _c_double_p]),
    "nngp_loglik_device": (ctypes.c_int, [_handle_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "nngp_peer_export": (ctypes.c_int, [_handle_p, ctypes.c_int, ctypes.c_char_p]),
    "nngp_peer_connect": (ctypes.c_int, [_handle_p, ctypes.c_int, ctypes.c_int, ctypes.c_char_p]),
    "nngp_loglik_device_allreduce": (ctypes.c_int, [_handle_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_int, ctypes.c_void_p, ctypes.c_void_p]),
    "nngp_loglik_allreduce": (ctypes.c_int, [_handle_p, ctypes.c_int, _c_double_p, ctypes.c_int, _c_double_p]),
    "nngp_factors": (ctypes.c_int, [_handle_p, ctypes.c_int, _c_double_p, ctypes.c_int64, ctypes.c_int64, _c_double_p, _c_double_p]),
    "nngp_cov_blocks": (ctypes.c_int, [_handle_p, ctypes.c_int, _c_double_p, ctypes.c_int64, ctypes.c_int64, _c_double_p, _c_double_p, _c_double_p]),
    "nngp_launch_count": (ctypes.c_int64, [_handle_p]),
    "nngp_set_timing": (ctypes.c_int, [_handle_p, ctypes.c_int]),
    "nngp_last_eval_ms": (ctypes.c_int, [_handle_p, _c_double_p]),
    "nngp_measure_fma_peak": (ctypes.c_int, [_handle_p, ctypes.c_int, ctypes.c_int, _c_double_p]),
}

_lib = None


class NNGPError(RuntimeError):
    """A libnngp_b200 call returned a nonzero status."""


def load():
    """dlopen the in-tree library and bind every symbol; raises if it is not built."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise NNGPError(
                f"{LIB_PATH} is missing: build it with `python -m pynngp_b200.build` "
                "(pynngp_b200 has no CPU fallback)")
        lib = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in SIGNATURES.items():
            fn = getattr(lib, name)  # AttributeError if the .so does not export it
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = lib
    return _lib


def device_count():
    """Number of CUDA devices this process can see (0 without a driver); no engine is created."""
    return int(load().nngp_visible_devices())


def _dp(a):
    return None if a is None else a.ctypes.data_as(_c_double_p)


def _f64(a, shape=None):
    a = np.ascontiguousarray(a, dtype=np.float64)
    if shape is not None and a.shape != shape:
        raise ValueError(f"expected shape {shape}, got {a.shape}")
    return a


class Engine:
    """One handle = one GPU = one shard of the ordering (``device`` an int), or one multi-device handle
    (``device`` a list of device indices: nngp_create_multi -- data replicated, shard and table split over
    the devices, evaluations return the total).  Thin, 1:1 over the C ABI."""

    def __init__(self, device=0, dtype="float64"):
        self._lib = load()
        self._h = _handle_p()
        code = {"float64": F64, "float32": F32}.get(str(dtype))
        if code is None:
            raise ValueError("dtype must be 'float64' or 'float32'")
        if isinstance(device, (list, tuple)):
            devs = [int(d) for d in device]
            arr = (ctypes.c_int * len(devs))(*devs)
            rc = self._lib.nngp_create_multi(ctypes.byref(self._h), arr, len(devs), code)
            what = "nngp_create_multi"
            self.devices, self.device = devs, (devs[0] if devs else 0)
        else:
            rc = self._lib.nngp_create(ctypes.byref(self._h), int(device), code)
            what = "nngp_create"
            self.devices, self.device = [int(device)], int(device)
        if rc != 0:
            msg = self._lib.nngp_last_error(None).decode()
            self._h = None
            raise NNGPError(f"{what} failed ({rc}): {msg}")
        self.dtype = str(dtype)
        self.n = self.D = self.m = 0
        # preallocated argument / result storage of the one-vector fast path (loglik_terms)
        self._out3 = (ctypes.c_double * NSTAT)()
        self._terms = self._lib.nngp_loglik_terms

    def _check(self, rc, what):
        if rc != 0:
            raise NNGPError(f"{what} failed ({rc}): {self._lib.nngp_last_error(self._h).decode()}")

    def close(self):
        if getattr(self, "_h", None):
            self._lib.nngp_destroy(self._h)
            self._h = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- data ------------------------------------------------------------------------------------
    def set_data(self, coords, y, eps2=None):
        coords = _f64(coords)
        if coords.ndim != 2:
            raise ValueError("coords must be (n, D)")
        n, D = coords.shape
        y = _f64(y, (n,))
        eps2 = None if eps2 is None else _f64(eps2, (n,))
        self._check(self._lib.nngp_set_data(self._h, _dp(coords), n, D, _dp(y), _dp(eps2)), "nngp_set_data")
        self.n, self.D, self.m = n, D, 0

    def set_y(self, y):
        y = _f64(y, (self.n,))
        self._check(self._lib.nngp_set_y(self._h, _dp(y)), "nngp_set_y")

    def set_eps2(self, eps2):
        eps2 = _f64(eps2, (self.n,))
        self._check(self._lib.nngp_set_eps2(self._h, _dp(eps2)), "nngp_set_eps2")

    def set_shard(self, lo, hi):
        self._check(self._lib.nngp_set_shard(self._h, int(lo), int(hi)), "nngp_set_shard")

    # -- stage 1 ---------------------------------------------------------------------------------
    def build_neighbors(self, m, tile_offset=0, tile_stride=1):
        self._check(self._lib.nngp_build_neighbors(self._h, int(m), int(tile_offset), int(tile_stride)),
                    "nngp_build_neighbors")
        self.m = int(m)

    def build_neighbors_grid(self, m, row_lo=0, row_hi=None, algo="auto"):
        """Stage 1 through the cell-grid search (bit-identical to build_neighbors); rows outside
        [row_lo, row_hi) hold ROW_UNSET or their correct value."""
        row_hi = self.n if row_hi is None else row_hi
        code = {"auto": 0, "grid": 1, "brute": 2}[algo]
        self._check(self._lib.nngp_build_neighbors_grid(self._h, int(m), int(row_lo), int(row_hi), code),
                    "nngp_build_neighbors_grid")
        self.m = int(m)

    def build_neighbors_shard(self, m, algo="auto"):
        """Stage 1 for the rows of the handle's shard only; the table keeps only those rows."""
        code = {"auto": 0, "grid": 1, "brute": 2}[algo]
        self._check(self._lib.nngp_build_neighbors_shard(self._h, int(m), code), "nngp_build_neighbors_shard")
        self.m = int(m)

    def neighbor_window(self):
        """(row0, rows): the rows of the n x m table this handle holds."""
        a, b = ctypes.c_int64(0), ctypes.c_int64(0)
        self._check(self._lib.nngp_neighbor_window(self._h, ctypes.byref(a), ctypes.byref(b)), "nngp_neighbor_window")
        return a.value, b.value

    def build_neighbors_capped(self, m, row_lo, row_hi, cand_cap, algo="auto"):
        """Rows [row_lo, row_hi): the m nearest j < min(i, cand_cap) (prediction sites appended after the
        cand_cap reference sites get reference-only neighbours)."""
        code = {"auto": 0, "grid": 1, "brute": 2}[algo]
        self._check(self._lib.nngp_build_neighbors_capped(self._h, int(m), int(row_lo), int(row_hi), int(cand_cap), code),
                    "nngp_build_neighbors_capped")
        self.m = int(m)

    def set_knn_tuning(self, lambda_scale=0.5, brute_rows=128):
        self._check(self._lib.nngp_set_knn_tuning(self._h, float(lambda_scale), int(brute_rows)), "nngp_set_knn_tuning")

    def knn_used_grid(self):
        return bool(self._lib.nngp_knn_used_grid(self._h))

    def set_neighbors(self, table):
        table = np.ascontiguousarray(table, dtype=np.int32)
        if table.ndim != 2 or table.shape[0] != self.n:
            raise ValueError("neighbour table must be (n, m) int32")
        self._check(self._lib.nngp_set_neighbors(self._h, table.ctypes.data_as(_c_int32_p), table.shape[1]),
                    "nngp_set_neighbors")
        self.m = table.shape[1]

    def get_neighbors(self):
        out = np.empty((self.n, self.m), dtype=np.int32)
        self._check(self._lib.nngp_get_neighbors(self._h, out.ctypes.data_as(_c_int32_p)), "nngp_get_neighbors")
        return out

    def get_neighbor_rows(self, i0, i1):
        out = np.empty((int(i1) - int(i0), self.m), dtype=np.int32)
        self._check(self._lib.nngp_get_neighbor_rows(self._h, int(i0), int(i1), out.ctypes.data_as(_c_int32_p)),
                    "nngp_get_neighbor_rows")
        return out

    def knn_plain(self, k):
        """(n, k) int32: the k nearest rows of every row, itself included, ascending (d2, j)."""
        out = np.empty((self.n, int(k)), dtype=np.int32)
        self._check(self._lib.nngp_knn_plain(self._h, int(k), out.ctypes.data_as(_c_int32_p)), "nngp_knn_plain")
        return out

    def neighbors_device_ptr(self):
        return self._lib.nngp_neighbors_device_ptr(self._h)

    def neighbor_window_device_ptr(self):
        return self._lib.nngp_neighbor_window_device_ptr(self._h)

    # -- stages 2-3 --------------------------------------------------------------------------------
    def loglik(self, kernel_id, params):
        """params (K, 4) -> stats (K, 3) = [sum log F, sum r^2/F, n_bad] over the shard."""
        params = _f64(params)
        if params.ndim == 1:
            params = params[None, :]
        if params.shape[1] != NPARAM:
            raise ValueError("params must be (K, 4): sigma2, phi, tau2, nu")
        out = np.empty((params.shape[0], NSTAT), dtype=np.float64)
        self._check(self._lib.nngp_loglik(self._h, int(kernel_id), _dp(params), params.shape[0], _dp(out)),
                    "nngp_loglik")
        return out

    def loglik_terms(self, kernel_id, sigma2, phi, tau2):
        """One parameter vector by value -> (sum log F, sum r^2/F, n_bad); the total over all GPUs for a
        multi-device handle or a connected peer exchange.  No array is allocated on this path."""
        rc = self._terms(self._h, kernel_id, sigma2, phi, tau2, self._out3)
        if rc != 0:
            self._check(rc, "nngp_loglik_terms")
        o = self._out3
        return o[0], o[1], o[2]

    def loglik_allreduce(self, kernel_id, params):
        """Like loglik, but the (K, 3) result is the total over all ranks (fused peer-memory exchange)."""
        params = _f64(params)
        if params.ndim == 1:
            params = params[None, :]
        out = np.empty((params.shape[0], NSTAT), dtype=np.float64)
        self._check(self._lib.nngp_loglik_allreduce(self._h, int(kernel_id), _dp(params), params.shape[0], _dp(out)),
                    "nngp_loglik_allreduce")
        return out

    def loglik_device(self, kernel_id, d_params, K, d_out, stream=None):
        self._check(self._lib.nngp_loglik_device(self._h, int(kernel_id), ctypes.c_void_p(d_params), int(K),
                                                 ctypes.c_void_p(d_out), ctypes.c_void_p(stream or 0)),
                    "nngp_loglik_device")

    def peer_export(self, K_cap=256):
        """Allocates this rank's exchange buffer; returns its CUDA IPC handle (64 bytes)."""
        buf = ctypes.create_string_buffer(IPC_HANDLE_BYTES)
        self._check(self._lib.nngp_peer_export(self._h, int(K_cap), buf), "nngp_peer_export")
        return buf.raw

    def peer_connect(self, rank, world, handles):
        """handles: the ranks' IPC handles concatenated in rank order (world * 64 bytes)."""
        if len(handles) != world * IPC_HANDLE_BYTES:
            raise ValueError("need world * 64 bytes of IPC handles")
        self._check(self._lib.nngp_peer_connect(self._h, int(rank), int(world), bytes(handles)), "nngp_peer_connect")

    def loglik_device_allreduce(self, kernel_id, d_params, K, d_out, stream=None):
        self._check(self._lib.nngp_loglik_device_allreduce(self._h, int(kernel_id), ctypes.c_void_p(d_params), int(K),
                                                           ctypes.c_void_p(d_out), ctypes.c_void_p(stream or 0)),
                    "nngp_loglik_device_allreduce")

    def factors(self, kernel_id, params, i0=0, i1=None, want_B=True, want_F=True):
        i1 = self.n if i1 is None else i1
        params = _f64(params, (NPARAM,))
        B = np.empty((i1 - i0, self.m), dtype=np.float64) if want_B else None
        F = np.empty(i1 - i0, dtype=np.float64) if want_F else None
        self._check(self._lib.nngp_factors(self._h, int(kernel_id), _dp(params), i0, i1, _dp(B), _dp(F)),
                    "nngp_factors")
        return B, F

    def cov_blocks(self, kernel_id, params, i0=0, i1=None):
        i1 = self.n if i1 is None else i1
        params = _f64(params, (NPARAM,))
        CN = np.empty((i1 - i0, self.m, self.m), dtype=np.float64)
        cc = np.empty((i1 - i0, self.m), dtype=np.float64)
        cs = np.empty(i1 - i0, dtype=np.float64)
        self._check(self._lib.nngp_cov_blocks(self._h, int(kernel_id), _dp(params), i0, i1, _dp(CN), _dp(cc), _dp(cs)),
                    "nngp_cov_blocks")
        return CN, cc, cs

    # -- introspection -----------------------------------------------------------------------------
    def launch_count(self):
        return int(self._lib.nngp_launch_count(self._h))

    def set_timing(self, on=True):
        self._check(self._lib.nngp_set_timing(self._h, 1 if on else 0), "nngp_set_timing")

    def last_eval_ms(self):
        v = ctypes.c_double(0.0)
        self._check(self._lib.nngp_last_eval_ms(self._h, ctypes.byref(v)), "nngp_last_eval_ms")
        return v.value

    def measure_fma_peak(self, dtype="float64", iters=4096):
        v = ctypes.c_double(0.0)
        self._check(self._lib.nngp_measure_fma_peak(self._h, F64 if dtype == "float64" else F32, int(iters),
                                                    ctypes.byref(v)), "nngp_measure_fma_peak")
        return v.value
