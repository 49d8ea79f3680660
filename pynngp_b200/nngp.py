"""``NNGP`` -- the reference's public class (pyNNGP/nngp.py:5-101), same constructor and attributes,
with the hot path running on a B200 through libnngp_b200.so.

What is kept from the reference (drop-in contract):
  * ``NNGP(t, y, eps, refType, m, cov)`` positional, all work done eagerly (nngp.py:6-18);
  * attributes ``t, y, eps, refType, m, cov, s, wt, ws, Ns, Nt``; ``s is t`` and ``Nt is Ns`` for
    ``refType == 'S=T'`` (nngp.py:31, 65-67); ``Ns[0] == []`` and ``Ns[i]`` an int64 array of the
    min(m, i) nearest predecessors in ascending distance (nngp.py:49-62);
  * the per-location accessors ``_CNs, _Ccross, _Cs, _Bsi, _Fsi`` (nngp.py:73-96) -- stubs upstream,
    real values here; ``oneSample`` (nngp.py:98-101) calls undefined methods upstream and raises
    ``NotImplementedError`` here.
What is added: ``loglik``, ``loglik_terms``, ``loglik_batch``, ``factors`` and keyword-only engine
options.  There is no CPU fallback: constructing an ``NNGP`` without a B200 raises.

GPUs.  ``devices=None`` (default) uses every visible B200 of this process when it is a plain process (the
reference's caller writes ``NNGP(t, y, eps, refType, m, cov)`` and nothing else, nngp.py:6), and the local
rank's GPU under ``torchrun``; ``devices=int`` pins one GPU, ``devices=[...]`` lists them.  Several GPUs
in one process are one multi-device handle of the C library (``nngp_create_multi``): the ordering is split in
contiguous blocks, each device searches and keeps the neighbour rows of its block, and an evaluation is one
kernel per device -- launched in parallel by the library's own threads -- whose tails add the statistics
up over NVLink peer memory.

Reference sets other than T (nngp.py:32-40, 68-71).  ``refType = ('subset', nRef)`` and
``('random', nRef, bounds)`` crash upstream (``self.typ`` is never set, nngp.py:34); here they build what
the reference's code intends: ``s``, ``ws`` (5-NN regression of (t, y) evaluated at s), ``Ns`` (ordered
neighbours within s) and ``Nt`` (for every t_i the pair ``KDTree(s).query(t_i, m)`` returns: distances and
indices of its m nearest reference sites).  The engine then holds the rows [s ; T - S]: the first nRef
rows condition on their predecessors in s, every remaining observation on its m nearest reference sites --
the NNGP density of Datta et al. (2016) for S a subset of T -- so ``loglik`` stays one launch of the same
fused kernel.  For 'random' no response is observed at s: the neighbour structure and the per-location
accessors exist, ``loglik`` raises.
"""
from __future__ import annotations

import numpy as np

from . import _lib, dist as _dist
from .kernels import Kernel, parse as _parse_cov

LOG_2PI = float(np.log(2.0 * np.pi))


class NeighborSets:
    """Read-only sequence with the reference's ``Ns`` layout (nngp.py:50-62) over the dense
    (n, m) int32 table: ``Ns[0] == []``; ``Ns[i]`` is an int64 array, ascending distance."""

    def __init__(self, table):
        self.table = table

    def __len__(self):
        return self.table.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        if i < 0:
            i += len(self)
        if i == 0:
            return []  # the reference appends a Python list for the first site (nngp.py:52-54)
        row = self.table[i]
        return row[row >= 0].astype(np.int64)

    def __iter__(self):
        return (self[i] for i in range(len(self)))


def _dist_rows(a, b):
    """Euclidean distances from the point a (D,) to the rows of b (k, D), accumulated dimension by
    dimension as scikit-learn's KDTree does (sklearn/metrics/_dist_metrics.pxd.tp:39-49)."""
    d2 = np.zeros(len(b))
    for k in range(b.shape[1]):
        t = a[k] - b[:, k]
        d2 = d2 + t * t
    return np.sqrt(d2)


class QueryNeighborSets:
    """``Nt`` for S != T (nngp.py:68-71): ``Nt[i]`` is what ``KDTree(s).query(t_i.reshape(1, -1), m)``
    returns -- the pair (distances (1, m) float64, indices into s (1, m) int64), nearest first."""

    def __init__(self, table, t, s):
        self.table, self._t, self._s = table, t, s

    def __len__(self):
        return self.table.shape[0]

    def __getitem__(self, i):
        if isinstance(i, slice):
            return [self[k] for k in range(*i.indices(len(self)))]
        idx = self.table[i].astype(np.int64)
        return _dist_rows(self._t[i], self._s[idx])[None, :], idx[None, :]

    def __iter__(self):
        return (self[i] for i in range(len(self)))


class NNGP(object):
    def __init__(self, t, y, eps, refType, m, cov, *, dtype="float64", devices=None, device=None, neighbors=None,
                 group=None, knn="auto", seed=None):
        self.t = t  # ordinates
        self.y = y  # abscissae
        self.eps = eps  # measurement uncertainties in y
        self.refType = refType  # type of reference set to construct
        self.m = m  # number of reference neighbors per point
        self.cov = cov  # covariance of the parent GP: kernel spec (see pynngp_b200.kernels)

        self._kernel: Kernel = _parse_cov(cov)
        if knn not in ("auto", "grid", "brute"):
            raise ValueError("knn must be 'auto', 'grid' or 'brute'")
        self._knn = knn  # stage-1 algorithm: cell-grid search (bit-identical) unless the data defeats it
        self._group = group
        self._rank, self._world = _dist.get_world(group)
        if devices is None:
            devices = device  # `device=` is the older spelling of `devices=int`
        if devices is None:
            devices = self._default_devices()
        elif self._world > 1 and not isinstance(devices, int):
            raise ValueError("under torch.distributed every rank drives ONE GPU: pass devices=<int> (or nothing)")
        if isinstance(devices, (list, tuple)) and len(devices) == 1:
            devices = int(devices[0])
        self._engine = _lib.Engine(device=devices, dtype=dtype)
        self._ycol = None
        self._timings = {}
        self._seed = seed  # S != T only: None = numpy's global RNG as in the reference (nngp.py:36, 40)

        self._init_s()
        self._init_wt()
        self._upload()
        self._make_s_neighbor_sets(neighbors)
        self._make_t_neighbor_sets()
        self._ws = None
        self._peer_ok = self._setup_peer_exchange()
        k = self._kernel
        self._fast_terms = (self._y2d.shape[1] == 1 and self._ref_kind != "random" and (self._world == 1 or self._peer_ok)
                            and None not in (k.sigma2, k.phi, k.tau2))

    # ---- construction steps, named as in the reference -----------------------------------------
    def _default_devices(self):
        """torchrun: the local rank's GPU.  A plain process: every visible GPU (one multi-device handle), so an
        unchanged caller of the reference's constructor gets the whole box; NNGP_DEVICES=0 or =0,1,.. overrides."""
        import os

        if self._world > 1:
            return int(os.environ.get("LOCAL_RANK", "0"))
        env = os.environ.get("NNGP_DEVICES")
        if env:
            devs = [int(d) for d in env.split(",") if d.strip() != ""]
            return devs[0] if len(devs) == 1 else devs
        n = _lib.device_count()
        return 0 if n <= 1 else list(range(min(n, 8)))

    @property
    def devices(self):
        """The CUDA devices this object computes on."""
        return list(self._engine.devices)

    def _setup_peer_exchange(self, K_cap=128):
        """Multi-GPU on one node: map every rank's exchange buffer (CUDA IPC) so the fused kernel sums the
        statistics over NVLink peer memory itself (nngp_loglik_device_allreduce).  Any failure on any rank
        -- ranks on different nodes, no P2P, NNGP_PEER_EXCHANGE=0 -- leaves all ranks on the NCCL allreduce."""
        import os

        if self._world < 2:
            return False
        import torch
        import torch.distributed as dist

        if self._world > 8 or dist.get_backend(self._group) != "nccl":
            return False
        dev = torch.device("cuda", self._engine.device)
        ok = os.environ.get("NNGP_PEER_EXCHANGE", "1") != "0"
        mine = torch.zeros(_lib.IPC_HANDLE_BYTES, dtype=torch.uint8, device=dev)
        if ok:
            try:
                mine = torch.frombuffer(bytearray(self._engine.peer_export(K_cap)), dtype=torch.uint8).to(dev)
            except _lib.NNGPError:
                ok = False
        allh = [torch.empty_like(mine) for _ in range(self._world)]
        dist.all_gather(allh, mine, group=self._group)
        flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
        dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self._group)
        if int(flag.item()) == 1:
            try:
                self._engine.peer_connect(self._rank, self._world, b"".join(bytes(h.cpu().numpy().tobytes()) for h in allh))
            except _lib.NNGPError:
                ok = False
            flag = torch.tensor([1 if ok else 0], dtype=torch.int32, device=dev)
            dist.all_reduce(flag, op=dist.ReduceOp.MIN, group=self._group)
        self._peer_K_cap = K_cap
        return int(flag.item()) == 1

    def _init_s(self):
        # nngp.py:21-40.  Upstream only 'S=T' runs: the tuple branches read `self.typ`, which is never
        # set (nngp.py:34).  They are built here as written otherwise, with two repairs: the subset is
        # drawn without replacement (a repeated reference site makes C_N singular) and the draw can be
        # seeded (`seed=`; required with several ranks, which must all hold the same s).
        rt = self.refType
        self._choice = None
        if isinstance(rt, str):
            if rt != "S=T":
                raise ValueError("refType must be 'S=T', ('subset', nRef) or ('random', nRef, bounds)")
            self._ref_kind = "S=T"
            self.s = self.t
            return
        if not (isinstance(rt, tuple) and rt and rt[0] in ("subset", "random")):
            raise ValueError("refType must be 'S=T', ('subset', nRef) or ('random', nRef, bounds)")
        if self._seed is None and self._world > 1:
            raise ValueError("refType %r needs seed= when several ranks must draw the same reference set" % (rt[0],))
        rng = np.random if self._seed is None else np.random.default_rng(self._seed)
        typ, nRef = rt[0], int(rt[1])
        t = np.asarray(self.t)
        if nRef < 1:
            raise ValueError("nRef must be >= 1")
        if typ == "subset":
            if nRef > len(t):
                raise ValueError("a subset of T cannot hold more than len(t) sites")
            self._choice = np.asarray(rng.choice(len(t), size=nRef, replace=False), dtype=np.int64)
            self.s = t[self._choice]
        else:
            bounds = rt[2]
            if len(bounds) != (1 if t.ndim == 1 else t.shape[1]):
                raise ValueError("bounds needs one (lo, hi) pair per dimension of t")
            self.s = np.vstack([rng.uniform(lo, hi, nRef) for lo, hi in bounds]).T
        self._ref_kind = typ
        if self.m > nRef:
            # what KDTree(s).query(t_i, m) raises upstream (nngp.py:70-71)
            raise ValueError("m must be less than or equal to the number of reference sites")

    def _init_wt(self):
        self.wt = np.copy(self.y)  # nngp.py:42-43

    def _upload(self):
        """Engine row layout.  S = T: the rows of t.  'subset': [s ; T - S] (T - S in the order of t) with
        y and eps permuted alike.  'random': the rows of s only (no response is observed there)."""
        def as2d(a):
            a = np.ascontiguousarray(a, dtype=np.float64)
            return a[:, None] if a.ndim == 1 else a

        s = as2d(self.s)
        if not np.isfinite(s).all():
            raise ValueError("coordinates must be finite")
        self._n_ref = len(s)
        y = np.asarray(self.y, dtype=np.float64)
        nt = len(np.asarray(self.t))
        y2d = y.reshape(nt, -1)
        eps = np.broadcast_to(np.asarray(self.eps, dtype=np.float64).reshape((nt, -1) if np.ndim(self.eps) else (1, 1)),
                              y2d.shape)
        eps2 = None if not np.any(eps) else np.ascontiguousarray(eps * eps)
        self._yt2d = y2d  # the response in the order of t (what `ws` averages)
        self._eps2_t = eps2  # eps^2 in the order of t, or None
        self._rows = None  # engine row -> row of t
        if self._ref_kind == "subset":
            t = as2d(self.t)
            if not np.isfinite(t).all():
                raise ValueError("coordinates must be finite")
            rest = np.setdiff1d(np.arange(nt), self._choice)
            self._rows = np.concatenate([self._choice, rest])
            s = np.ascontiguousarray(t[self._rows])
            y2d = np.ascontiguousarray(y2d[self._rows])
            eps2 = None if eps2 is None else np.ascontiguousarray(eps2[self._rows])
        elif self._ref_kind == "random":
            y2d = np.zeros((len(s), y2d.shape[1]))
            eps2 = None
        n = len(s)
        self._y2d = y2d
        self._eps2 = eps2
        self._coords = s
        import time

        t0 = time.perf_counter()
        self._set_column(0)
        self._timings["upload_s"] = time.perf_counter() - t0
        lo, hi = _dist.shard_bounds(n, self._rank, self._world)
        self._engine.set_shard(lo, hi)
        self._shard = (lo, hi)

    def _set_column(self, c):
        if self._ycol == c:
            return
        eps2 = None if self._eps2 is None else np.ascontiguousarray(self._eps2[:, c])
        if self._ycol is None:
            self._engine.set_data(self._coords, np.ascontiguousarray(self._y2d[:, c]), eps2)
        else:
            # a column switch replaces the response (and its eps) in place: coordinates, neighbour table and
            # shard stay resident (nngp_set_y / nngp_set_eps2)
            self._engine.set_y(np.ascontiguousarray(self._y2d[:, c]))
            if eps2 is not None:
                self._engine.set_eps2(eps2)
        self._ycol = c

    def _make_s_neighbor_sets(self, neighbors=None):
        # nngp.py:49-62 -> stage 1 on the GPU (or an injected (n, m) table)
        import time

        eng = self._engine
        t0 = time.perf_counter()
        if isinstance(neighbors, (str, bytes)) or hasattr(neighbors, "__fspath__"):
            neighbors = np.load(neighbors)  # a table written by save_neighbors()
        if neighbors is not None:
            neighbors = np.asarray(neighbors)
            if neighbors.ndim != 2 or neighbors.shape != (eng.n, self.m) or neighbors.dtype.kind not in "iu":
                raise ValueError(f"neighbors must be an integer table of shape ({eng.n}, {self.m}): got "
                                 f"{neighbors.dtype} {neighbors.shape}")
            eng.set_neighbors(neighbors)  # checked on the device: entries of row i in [-1, i), padding last
            eng.set_shard(*self._shard)
        elif self._ref_kind != "S=T":
            # nngp.py:49-62 on s (rows [0, nRef): ordered search among predecessors) and nngp.py:68-71
            # for the rows of T - S (their m nearest reference sites, any index); every rank builds the
            # whole table (on one device: a multi-device handle only searches whole shards)
            one = eng if len(eng.devices) == 1 else _lib.Engine(device=eng.device, dtype=eng.dtype)
            try:
                if one is not eng:
                    one.set_data(self._coords, np.zeros(eng.n), None)
                one.build_neighbors_grid(self.m, 0, self._n_ref, self._knn)
                table = one.get_neighbor_rows(0, self._n_ref)
            finally:
                if one is not eng:
                    one.close()
            if eng.n > self._n_ref:
                table = np.concatenate([table, self._nt_table[self._rows[self._n_ref:]]])
            eng.set_neighbors(table)
            eng.set_shard(*self._shard)
            neighbors = table
        else:
            # every GPU (a rank under torchrun, a device of a multi-device handle) searches and keeps the rows of
            # its own shard only: the likelihood needs nothing else, so stage 1 involves no exchange at all.  The
            # grid search costs about the same per row wherever the row sits in the ordering.
            eng.build_neighbors_shard(self.m, self._knn)
        self._timings["knn_s"] = time.perf_counter() - t0
        self._table_host = None if neighbors is None else np.ascontiguousarray(neighbors, dtype=np.int32)
        self._Ns = None

    @property
    def _table(self):
        """(n, m) int32 neighbour table on the host; downloaded from the device the first time it is
        read (the likelihood itself never needs it on the host)."""
        if self._table_host is None:
            row0, rows = self._engine.neighbor_window()
            if row0 != 0 or rows != self._engine.n:
                # a rank of a multi-process run holds the rows of its shard only.  Reading the whole table
                # (`Ns`, save_neighbors, the accessors) must not be a collective -- one rank alone may ask -- so
                # the rank searches the missing rows itself: coordinates are replicated, and the whole search
                # costs milliseconds (gather_table() is the collective alternative)
                self._engine.build_neighbors_grid(self.m, 0, self._engine.n, self._knn)
            self._table_host = self._engine.get_neighbors()
        return self._table_host

    def gather_table(self):
        """Collective (every rank must call it): assembles the (n, m) table from the row blocks the ranks hold
        -- one all_gather of ceil(n / world) rows per rank instead of a search of the missing rows."""
        if self._table_host is not None or self._world == 1:
            return self._table
        import torch
        import torch.distributed as dist

        eng = self._engine
        row0, rows = eng.neighbor_window()
        per = -(-eng.n // self._world)
        if (row0, rows) != tuple((self._shard[0], self._shard[1] - self._shard[0])):
            return self._table
        cuda = dist.get_backend(self._group) == "nccl"
        mine = torch.full((per, eng.m), _lib.ROW_UNSET, dtype=torch.int32)
        mine[:rows] = torch.from_numpy(eng.get_neighbor_rows(row0, row0 + rows))
        if cuda:
            mine = mine.to(f"cuda:{eng.device}")
        parts = [torch.empty_like(mine) for _ in range(self._world)]
        dist.all_gather(parts, mine, group=self._group)
        table = np.empty((eng.n, eng.m), dtype=np.int32)
        for r, part in enumerate(parts):
            lo, hi = _dist.shard_bounds(eng.n, r, self._world)
            table[lo:hi] = part[: hi - lo].cpu().numpy()
        self._table_host = table
        return table

    @property
    def Ns(self):
        """The reference's list of neighbour index arrays (nngp.py:49-62)."""
        if self._Ns is None:
            self._Ns = NeighborSets(self._table[: self._n_ref])
        return self._Ns

    @property
    def Nt(self):
        if self._ref_kind == "S=T":
            return self.Ns  # nngp.py:65-67: the same object
        if self._Nt is None:
            t = np.asarray(self.t, dtype=np.float64)
            self._Nt = QueryNeighborSets(self._nt_table, t[:, None] if t.ndim == 1 else t, self._coords_s())
        return self._Nt

    def _make_t_neighbor_sets(self):
        # nngp.py:65-71.  S = T: Nt is Ns.  Otherwise the (len(t), m) table of nearest reference sites is
        # searched on the GPU the first time it is needed (see _nt_table) and wrapped by the Nt property.
        self._Nt = None

    def _coords_s(self):
        return self._coords[: self._n_ref]

    def _nearest_rows(self, ref, queries, k):
        """(len(queries), k) int32: the k nearest rows of `ref` for every query row, ascending (d2, j) --
        the capped search of the hot path (queries appended after the reference rows, candidates j <
        len(ref)), on a scratch engine."""
        eng = _lib.Engine(device=self._engine.device, dtype=self._engine.dtype)
        try:
            n, q = len(ref), len(queries)
            eng.set_data(np.concatenate([ref, queries]), np.zeros(n + q), None)
            eng.build_neighbors_capped(k, n, n + q, n, self._knn)
            return eng.get_neighbor_rows(n, n + q)
        finally:
            eng.close()

    @property
    def _nt_table(self):
        """S != T: row i = the m nearest reference sites of t_i (nngp.py:68-71), indices into s."""
        if getattr(self, "_nt_table_host", None) is None:
            t = np.ascontiguousarray(self.t, dtype=np.float64)
            self._nt_table_host = self._nearest_rows(self._coords_s(), t[:, None] if t.ndim == 1 else t, self.m)
        return self._nt_table_host

    @property
    def ws(self):
        """Warm start of the latent field at the reference sites (nngp.py:45-47): the uniform mean of y
        over the 5 nearest sites of each site, itself included -- what scikit-learn's
        ``KNeighborsRegressor(5, 'uniform').fit(t, y).predict(s)`` returns for S = T.  Computed lazily
        on the GPU (plain k-NN kernel) the first time it is read; not used by the likelihood."""
        if self._ws is None and self._ref_kind == "S=T":
            k = min(5, len(self._coords))
            idx = self._engine.knn_plain(k)
            self._ws = self._y2d[idx].mean(axis=1).reshape(np.shape(self.y))
        elif self._ws is None:
            # S != T: the regressor is fitted on (t, y) and evaluated at s
            t = np.ascontiguousarray(self.t, dtype=np.float64)
            idx = self._nearest_rows(t[:, None] if t.ndim == 1 else t, self._coords_s(), min(5, len(t)))
            self._ws = self._yt2d[idx].mean(axis=1).reshape((self._n_ref,) + np.shape(self.y)[1:])
        return self._ws

    # ---- parameters -------------------------------------------------------------------------------
    def _params(self, sigma2=None, phi=None, tau2=None):
        return np.array(self._kernel.params(sigma2, phi, tau2), dtype=np.float64)

    # ---- the reference's per-location accessors (nngp.py:73-96) ------------------------------------
    def _p(self, i):
        return int((self._table[i] >= 0).sum())

    def _CNs(self, i, **kw):
        """C_{N(s_i)} (nngp.py:78-82): (p, p)."""
        CN, _, _ = self._engine.cov_blocks(self._kernel.kernel_id, self._params(**kw), i, i + 1)
        p = self._p(i)
        return CN[0, :p, :p]

    def _Ccross(self, i, **kw):
        """C_{s_i, N(s_i)} (nngp.py:84-86): (p,)."""
        _, cc, _ = self._engine.cov_blocks(self._kernel.kernel_id, self._params(**kw), i, i + 1)
        return cc[0, : self._p(i)]

    def _Cs(self, i, **kw):
        """C_{s_i, s_i} (nngp.py:92-96)."""
        _, _, cs = self._engine.cov_blocks(self._kernel.kernel_id, self._params(**kw), i, i + 1)
        return float(cs[0])

    def _Bsi(self, i, **kw):
        """B_{s_i} = C_N(i)^{-1} c_i (nngp.py:73-76): (p,)."""
        B, _ = self._engine.factors(self._kernel.kernel_id, self._params(**kw), i, i + 1, want_F=False)
        return B[0, : self._p(i)]

    def _Fsi(self, i, **kw):
        """F_{s_i} = C(i,i) - c_i^T b_i (nngp.py:88-90)."""
        _, F = self._engine.factors(self._kernel.kernel_id, self._params(**kw), i, i + 1, want_B=False)
        return float(F[0])

    def factors(self, sigma2=None, phi=None, tau2=None, i0=0, i1=None):
        """(B (i1-i0, m) zero padded, F (i1-i0,)) for rows [i0, i1)."""
        return self._engine.factors(self._kernel.kernel_id, self._params(sigma2, phi, tau2), i0, i1)

    # ---- the likelihood ----------------------------------------------------------------------------
    def loglik_batch(self, params):
        """params (K, 3|4) rows of (sigma2, phi, tau2[, nu]) -> (K, 3) global statistics
        [sum log F, sum r^2/F, n_bad], summed over response columns and over all ranks."""
        if self._ref_kind == "random":
            raise NotImplementedError("refType ('random', ...) observes no response at the reference sites: the "
                                      "response likelihood needs S = T or a subset of T")
        params = np.atleast_2d(np.asarray(params, dtype=np.float64))
        if params.shape[1] == 3:
            params = np.concatenate([params, np.zeros((params.shape[0], 1))], axis=1)
        if self._world > 1:
            return self._loglik_batch_sharded(params)
        total = np.zeros((params.shape[0], _lib.NSTAT))
        for c in range(self._y2d.shape[1]):
            self._set_column(c)
            total += self._engine.loglik(self._kernel.kernel_id, params)
        return total

    def _loglik_batch_sharded(self, params):
        """Multi-GPU evaluation.  With the peer exchange connected the fused kernel itself returns the
        statistics summed over the ranks (NVLink peer memory).  Otherwise: parameters go up from a pinned
        buffer, the kernel writes this rank's partials, an NCCL allreduce follows on the same stream with no
        host round trip in between, and one pinned D2H copy returns the totals."""
        K = params.shape[0]
        if self._peer_ok:
            # the kernel's last block sums over the ranks through NVLink peer memory: no torch, no NCCL
            total = np.zeros((K, _lib.NSTAT))
            cap = self._peer_K_cap
            for c in range(self._y2d.shape[1]):
                self._set_column(c)
                for k0 in range(0, K, cap):
                    total[k0:k0 + cap] += self._engine.loglik_allreduce(self._kernel.kernel_id, params[k0:k0 + cap])
            if np.isnan(total).any():
                raise RuntimeError("the cross-GPU exchange timed out: a rank never launched its evaluation")
            return total
        import torch
        import torch.distributed as tdist

        if tdist.get_backend(self._group) != "nccl":
            # a host-side group (gloo): this rank's partial statistics, summed by the group's allreduce
            total = np.zeros((K, _lib.NSTAT))
            for c in range(self._y2d.shape[1]):
                self._set_column(c)
                total += self._engine.loglik(self._kernel.kernel_id, params)
            return np.asarray(_dist.allreduce_stats(total, self._group))
        dev = torch.device("cuda", self._engine.device)
        buf = getattr(self, "_shard_bufs", None)
        if buf is None or buf["K"] < K:
            with torch.cuda.device(dev):
                cap = max(K, 16)
                buf = self._shard_bufs = {
                    "K": cap, "stream": torch.cuda.Stream(),
                    "h_prm": torch.empty((cap, _lib.NPARAM), dtype=torch.float64).pin_memory(),
                    "h_out": torch.empty((cap, _lib.NSTAT), dtype=torch.float64).pin_memory(),
                    "d_prm": torch.empty((cap, _lib.NPARAM), dtype=torch.float64, device=dev),
                    "d_out": torch.empty((cap, _lib.NSTAT), dtype=torch.float64, device=dev),
                    "d_tot": torch.empty((cap, _lib.NSTAT), dtype=torch.float64, device=dev),
                }
        stream = buf["stream"]
        ncol = self._y2d.shape[1]
        buf["h_prm"][:K].numpy()[...] = params
        with torch.cuda.device(dev), torch.cuda.stream(stream):
            d_prm, d_out, d_tot = buf["d_prm"][:K], buf["d_out"][:K], buf["d_tot"][:K]
            d_prm.copy_(buf["h_prm"][:K], non_blocking=True)
            for c in range(ncol):
                if ncol > 1:
                    stream.synchronize()  # the previous column's kernel still reads the y it is about to replace
                self._set_column(c)
                dst = d_out if ncol == 1 else d_tot if c == 0 else d_out
                self._engine.loglik_device(self._kernel.kernel_id, d_prm.data_ptr(), K, dst.data_ptr(),
                                           stream.cuda_stream)
                if c > 0:
                    d_tot += d_out
            res = d_out if ncol == 1 else d_tot
            _dist.allreduce_stats(res, self._group)  # NCCL on the same stream
            buf["h_out"][:K].copy_(res, non_blocking=True)
            stream.synchronize()
        return buf["h_out"][:K].numpy().copy()

    def loglik_terms(self, sigma2=None, phi=None, tau2=None):
        """(sum_i log F_i, sum_i r_i^2 / F_i) -- the reduction BASELINE.json's north_star names."""
        if self._fast_terms:
            # one response column on one handle (or a connected peer exchange): the parameters travel by value
            # through nngp_loglik_terms and the statistics come back through mapped memory -- no arrays here
            k = self._kernel
            st = self._engine.loglik_terms(
                k.kernel_id, float(k.sigma2 if sigma2 is None else sigma2), float(k.phi if phi is None else phi),
                float(k.tau2 if tau2 is None else tau2))
            if st[0] != st[0]:
                raise RuntimeError("the cross-GPU exchange timed out: a rank never launched its evaluation")
        else:
            st = self.loglik_batch(self._params(sigma2, phi, tau2)[None, :])[0]
        if st[2] > 0:
            raise FloatingPointError(f"{int(st[2])} location(s) had a non-positive-definite neighbour covariance")
        return float(st[0]), float(st[1])

    def loglik(self, sigma2=None, phi=None, tau2=None):
        """log N(y; 0, C_nngp)."""
        slog, squad = self.loglik_terms(sigma2, phi, tau2)
        ncol = self._y2d.shape[1]
        return -0.5 * (slog + squad) - 0.5 * len(self._coords) * ncol * LOG_2PI

    def loglik_latent(self, w_s, sigma2=None, phi=None, tau2=None):
        """Joint log density of the latent NNGP model the reference's sampler is meant to explore (its
        ``oneSample`` would update ``wt`` / ``ws``, nngp.py:98-101): log p(w_S) + log p(y | w_S) with

            p(w_S)     = prod_i N(w_i | b_i^T w_N(s_i), F_i),   C = sigma2 rho, no nugget on the latent field,
            p(y | w_S) = prod_t N(y_t | b_t^T w_N(t), sigma2 + tau2 + eps_t^2 - c_t^T C_N(t)^-1 c_t),

        N(s_i) from ``Ns`` and N(t) from ``Nt``.  Works for every refType ('random' included: this is the
        density that needs no response at the reference sites).  One launch of the fused kernel per response
        column over the rows [s ; t] of a second engine (built on first use): the latent values ride in the
        response lane of the reference rows, and the nugget tau2 + eps_t^2 is the per-observation variance of
        the observation rows only (``nngp_set_eps2``), so the kernel's own tau2 stays 0.  Every rank evaluates
        the whole density (no sharding)."""
        prm = self._params(sigma2, phi, tau2)
        n_ref, nt, ncol = self._n_ref, self._yt2d.shape[0], self._yt2d.shape[1]
        w = np.asarray(w_s, dtype=np.float64).reshape(n_ref, -1)
        if w.shape[1] != ncol:
            raise ValueError("w_s must hold one latent value per reference site and response column")
        eng = getattr(self, "_latent_eng", None)
        if eng is None:
            t = np.ascontiguousarray(self.t, dtype=np.float64)
            t = t[:, None] if t.ndim == 1 else t
            eng = _lib.Engine(device=self._engine.device, dtype=self._engine.dtype)
            eng.set_data(np.concatenate([self._coords_s(), t]), np.zeros(n_ref + nt), None)
            eng.set_neighbors(np.concatenate([self._table[:n_ref], self._nt_table]))
            self._latent_eng, self._latent_nugget = eng, None
        total = 0.0
        for c in range(ncol):
            nugget = (float(prm[2]), c if self._eps2_t is not None else 0)
            if self._latent_nugget != nugget:  # the observation rows' variances change with tau2 (and the column's eps)
                e2 = np.full(nt, prm[2]) if self._eps2_t is None else prm[2] + self._eps2_t[:, c]
                eng.set_eps2(np.concatenate([np.zeros(n_ref), e2]))
                self._latent_nugget = nugget
            eng.set_y(np.concatenate([w[:, c], self._yt2d[:, c]]))
            st = eng.loglik(self._kernel.kernel_id, np.array([prm[0], prm[1], 0.0, 0.0]))[0]
            if st[2] > 0:
                raise FloatingPointError(f"{int(st[2])} location(s) had a non-positive-definite neighbour covariance")
            total += st[0] + st[1]
        return -0.5 * total - 0.5 * (n_ref + nt) * ncol * LOG_2PI

    # ---- either side of the path: table I/O, kriging at new sites, a sweep driver ---------------------
    def save_neighbors(self, path):
        """Writes the (n, m) int32 neighbour table as .npy; pass the path as ``neighbors=`` to a later
        constructor on the same sites to skip stage 1."""
        np.save(path, self._table)

    def predict(self, t_new, sigma2=None, phi=None, tau2=None, m=None):
        """Kriging at new sites from their m nearest REFERENCE sites (NNGP prediction): returns
        (mean, var) with mean = b^T y_N and var = C(0) + tau2 - c^T C_N^-1 c, the variance of a new noisy
        observation (subtract tau2 for the latent field).  Runs the hot path's own kernels: the new sites
        are appended after the n reference sites, stage 1 is restricted to candidates j < n, and the
        emitting variant of the fused kernel returns b and F for the appended rows."""
        tn = np.ascontiguousarray(t_new, dtype=np.float64)
        if tn.ndim == 1:
            tn = tn[:, None] if self._coords.shape[1] == 1 else tn[None, :]
        if tn.shape[1] != self._coords.shape[1]:
            raise ValueError("t_new must have the reference sites' dimension")
        if not np.isfinite(tn).all():
            raise ValueError("coordinates must be finite")
        if self._ref_kind == "random":
            raise NotImplementedError("refType ('random', ...) observes no response at the reference sites")
        n, q = self._n_ref, len(tn)  # kriging conditions on the reference sites (the first rows of the engine)
        m = self.m if m is None else int(m)
        prm = self._params(sigma2, phi, tau2)
        ncol = self._y2d.shape[1]
        mean = np.zeros((q, ncol))
        var = np.zeros((q, ncol))
        if q == 0:
            return mean.reshape((0,) + np.shape(self.y)[1:]), var.reshape((0,) + np.shape(self.y)[1:])
        eng = _lib.Engine(device=self._engine.device, dtype=self._engine.dtype)
        try:
            coords = np.concatenate([self._coords[:n], tn])
            B = F = tab = None
            for c in range(ncol):
                if B is None or self._eps2 is not None:  # weights depend on the column only through eps
                    if tab is None:
                        eps2 = None if self._eps2 is None else np.concatenate([self._eps2[:n, c], np.zeros(q)])
                        eng.set_data(coords, np.concatenate([self._y2d[:n, c], np.zeros(q)]), eps2)
                        eng.build_neighbors_capped(m, n, n + q, n, self._knn)
                        tab = eng.get_neighbor_rows(n, n + q)
                    else:  # the next column's eps: replaced in place, coordinates and table stay
                        eng.set_eps2(np.concatenate([self._eps2[:n, c], np.zeros(q)]))
                    B, F = eng.factors(self._kernel.kernel_id, prm, n, n + q)
                yn = np.where(tab >= 0, self._y2d[np.maximum(tab, 0), c], 0.0)
                mean[:, c] = (B * yn).sum(axis=1)
                var[:, c] = F
        finally:
            eng.close()
        shape = (q,) + np.shape(self.y)[1:]
        return mean.reshape(shape), var.reshape(shape)

    def metropolis(self, n_steps, init=None, step=0.05, seed=0, log_prior=None):
        """Random-walk Metropolis over log(sigma2, phi, tau2) with the NNGP likelihood as target
        (neighbours fixed, one fused-kernel evaluation per step: BASELINE.json configs[4]).  Returns
        (chain (n_steps, 3), loglik (n_steps,), acceptance rate).  `log_prior(theta)` defaults to flat in
        log-parameters.  The same random stream on every rank keeps sharded runs in lock step."""
        rng = np.random.default_rng(seed)
        theta = np.array(self._kernel.params(None, None, None)[:3] if init is None else init, dtype=np.float64)
        step = np.broadcast_to(np.asarray(step, dtype=np.float64), (3,))
        lp = (lambda th: 0.0) if log_prior is None else log_prior

        def target(th):
            st = self.loglik_batch(np.array([[th[0], th[1], th[2], 0.0]]))[0]
            if st[2] > 0 or not np.isfinite(st[:2]).all():
                return -np.inf
            ncol = self._y2d.shape[1]
            return -0.5 * (st[0] + st[1]) - 0.5 * len(self._coords) * ncol * LOG_2PI

        cur = target(theta)
        chain = np.empty((n_steps, 3))
        trace = np.empty(n_steps)
        acc = 0
        for k in range(n_steps):
            prop = theta * np.exp(step * rng.standard_normal(3))
            val = target(prop)
            # symmetric in log-parameters: the Jacobian of the log transform enters through log_prior
            if np.log(rng.random()) < (val + lp(prop)) - (cur + lp(theta)):
                theta, cur = prop, val
                acc += 1
            chain[k], trace[k] = theta, cur
        return chain, trace, acc / max(n_steps, 1)

    def close(self):
        """Releases the device memory of this object's engines (also done when the object is collected)."""
        for name in ("_latent_eng", "_engine"):
            eng = getattr(self, name, None)
            if eng is not None:
                eng.close()
        self._latent_eng = None

    def __enter__(self):
        return self

    def __exit__(self, *exc):
        self.close()
        return False

    def oneSample(self):
        # nngp.py:98-101 calls update_wt / update_ws / update_y_unobserved, none of which exist upstream
        raise NotImplementedError("the reference's Gibbs sweep is unimplemented upstream (nngp.py:98-101)")
