"""CPU model of the cell-grid ordered k-NN search (pynngp_b200/csrc/knn_grid.cu) -- TEST INFRASTRUCTURE.

Restates, in numpy, the algorithm the CUDA search implements for _make_s_neighbor_sets
(pyNNGP/nngp.py:49-62): levels of the ordering, one uniform grid per level, a ring walk per query and the
conservative stopping rule.  Its purpose is to let the CPU test-suite attack the exactness argument --
"stopping never changes the result" -- with inputs chosen to break it (ties, duplicates, collapsed
dimensions, huge offsets, clusters, tiny cells), independently of any GPU.  It follows the kernel's
arithmetic: fp64, cell = min(G-1, int((x - lo) * inv_h)), d2 summed dimension by dimension, candidates
compared in (d2, j).  Only tests/ may import it.
"""
from __future__ import annotations

import math

import numpy as np

KNN_TILE = 128
BALL = {1: 2.0, 2: math.pi, 3: 4.0 * math.pi / 3.0}


def make_grid(bb_lo, bb_hi, D, nref, lam, max_cells):
    """knn_grid.cu: make_grid.  Returns dict(lo, h, inv_h, G, ncell, slack)."""
    lo = np.zeros(3)
    ext = np.zeros(3)
    lo[:D] = bb_lo
    ext[:D] = np.asarray(bb_hi) - np.asarray(bb_lo)
    active = ext > 0.0
    mag = max([0.0] + [max(abs(a), abs(b)) for a, b in zip(bb_lo, bb_hi)])
    side = 0.0
    for _ in range(4):
        deff = int(active.sum())
        if deff == 0:
            break
        vol = float(np.prod(ext[active]))
        side = (lam * vol / max(nref, 1.0)) ** (1.0 / deff)
        drop = active & ~(ext >= side)
        if not drop.any():
            break
        active &= ~drop
    while True:
        G = np.ones(3, dtype=np.int64)
        for d in range(3):
            if active[d] and side > 0.0:
                g = math.floor(ext[d] / side)
                G[d] = 1 if g < 1 else min(int(g), 1048576)
        if int(np.prod(G)) <= max_cells:
            break
        side *= 1.26
    h = np.zeros(3)
    inv_h = np.zeros(3)
    for d in range(3):
        if G[d] > 1:
            h[d] = ext[d] / G[d]
            inv_h[d] = G[d] / ext[d]
    return dict(lo=lo, h=h, inv_h=inv_h, G=G, ncell=int(np.prod(G)), slack=1e-14 * mag)


def cell_coords(gs, pts3):
    """(n, 3) int cell coordinates, as cell_index() assigns them."""
    t = (pts3 - gs["lo"]) * gs["inv_h"]
    return np.minimum(gs["G"] - 1, t.astype(np.int64))


def dist2(q, c, D):
    """scikit-learn's order of operations (sklearn/metrics/_dist_metrics.pxd.tp:39-49)."""
    d = (q[0] - c[:, 0]) * (q[0] - c[:, 0])
    d = d + (q[1] - c[:, 1]) * (q[1] - c[:, 1])  # the kernel always adds the y term: 0 for D = 1, d unchanged
    if D == 3:
        d = d + (q[2] - c[:, 2]) * (q[2] - c[:, 2])
    return d


def grid_knn_ordered(s, m, lam_scale=1.0, brute_rows=128, stats=None):
    """(n, m) int32 table, -1 padded: row i = the min(m, i) nearest j < i in ascending (d2, j)."""
    s = np.ascontiguousarray(s, dtype=np.float64)
    if s.ndim == 1:
        s = s[:, None]
    n, D = s.shape
    pts = np.zeros((n, 3))
    pts[:, :D] = s
    out = np.full((n, m), -1, dtype=np.int32)
    bb_lo, bb_hi = s.min(axis=0), s.max(axis=0)
    T0 = max(brute_rows, 4 * m)
    T0 = min(n, (T0 + KNN_TILE - 1) // KNN_TILE * KNN_TILE)
    for i in range(min(T0, n)):  # rows below T0: brute force
        if i:
            d2 = dist2(pts[i], pts[:i], D)
            out[i, : min(m, i)] = np.lexsort((np.arange(i), d2))[:m]
    deff = max(1, int((bb_hi > bb_lo).sum()))
    lam = max(1.0, lam_scale * (m + 2.0 * math.sqrt(m)) / BALL[deff])
    a = T0
    while a < n:
        b = min(max(2 * a, 4096 if a == T0 else 0), n)  # the lowest level runs up to 4096 (knn_grid.cu)
        gs = make_grid(bb_lo, bb_hi, D, float(a), lam, max(8 * n, 1024))
        G = gs["G"]
        cc = cell_coords(gs, pts[:b])
        cid = (cc[:, 2] * G[1] + cc[:, 1]) * G[0] + cc[:, 0]
        order = np.argsort(cid, kind="stable")
        starts = np.searchsorted(cid[order], np.arange(gs["ncell"] + 1))
        for i in range(a, b):
            q = pts[i]
            cx, cy, cz = cc[i]
            best_d = np.empty(0)
            best_j = np.empty(0, dtype=np.int64)
            r = 0
            while True:
                r += 1
                x0, x1 = max(cx - r, 0), min(cx + r, G[0] - 1)
                y0, y1 = max(cy - r, 0), min(cy + r, G[1] - 1)
                z0, z1 = max(cz - r, 0), min(cz + r, G[2] - 1)
                spans = []
                for zz in range(z0, z1 + 1):
                    for yy in range(y0, y1 + 1):
                        base = (zz * G[1] + yy) * G[0]
                        if r == 1 or abs(zz - cz) == r or abs(yy - cy) == r:
                            spans.append((starts[base + x0], starts[base + x1 + 1]))
                        else:
                            if cx - r >= 0:
                                spans.append((starts[base + cx - r], starts[base + cx - r + 1]))
                            if cx + r <= G[0] - 1:
                                spans.append((starts[base + cx + r], starts[base + cx + r + 1]))
                cand = np.concatenate([order[u:v] for u, v in spans]) if spans else np.empty(0, dtype=np.int64)
                cand = cand[cand < i]
                if stats is not None:
                    stats["candidates"] = stats.get("candidates", 0) + len(cand)
                if len(cand):
                    d2 = dist2(q, pts[cand], D)
                    best_d = np.concatenate([best_d, d2])
                    best_j = np.concatenate([best_j, cand])
                    keep = np.lexsort((best_j, best_d))[:m]
                    best_d, best_j = best_d[keep], best_j[keep]
                if x0 == 0 and x1 == G[0] - 1 and y0 == 0 and y1 == G[1] - 1 and z0 == 0 and z1 == G[2] - 1:
                    break
                if len(best_j) == m:
                    bound = math.inf
                    for d, c in enumerate((cx, cy, cz)):
                        if c - r > 0:
                            bound = min(bound, q[d] - (gs["lo"][d] + float(c - r) * gs["h"][d]))
                        if c + r < G[d] - 1:
                            bound = min(bound, (gs["lo"][d] + float(c + r + 1) * gs["h"][d]) - q[d])
                    bs = bound * (1.0 - 1e-9) - gs["slack"]
                    if bs > 0.0 and best_d[-1] < bs * bs:
                        break
            out[i, : len(best_j)] = best_j
        a = b
    return out
