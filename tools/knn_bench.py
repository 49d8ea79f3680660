"""Times stage 1 (ordered k-NN) through the cell-grid search against the brute-force kernel.

    python tools/knn_bench.py cfg3 [--brute] [--lams 0.5,1,2] [--rows 8192]
"""
import argparse
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynngp_b200 import _lib  # noqa: E402
from pynngp_b200.synthetic import CONFIGS, synthetic  # noqa: E402

ap = argparse.ArgumentParser()
ap.add_argument("cfg")
ap.add_argument("--brute", action="store_true")
ap.add_argument("--lams", default="1.0")
ap.add_argument("--rows", default="128")
ap.add_argument("--n", type=int, default=0)
a = ap.parse_args()
c = CONFIGS[a.cfg]
n = a.n or c["n"]
s, y = synthetic(n, c["D"], c["seed"])
e = _lib.Engine(0)
t0 = time.perf_counter()
e.set_data(s, y)
print(f"{a.cfg}: n={n} D={c['D']} m={c['m']}  set_data {time.perf_counter() - t0:.3f}s", flush=True)
ref = None
if a.brute:
    t0 = time.perf_counter()
    e.build_neighbors(c["m"])
    print(f"brute: {time.perf_counter() - t0:.4f}s", flush=True)
    ref = e.get_neighbors()
for rows in [int(r) for r in a.rows.split(",")]:
    for lam in [float(v) for v in a.lams.split(",")]:
        e.set_knn_tuning(lam, rows)
        best = 1e9
        for rep in range(3):
            t0 = time.perf_counter()
            e.build_neighbors_grid(c["m"], 0, None, "grid")
            best = min(best, time.perf_counter() - t0)
        tab = e.get_neighbors()
        ok = "" if ref is None else ("  == brute" if np.array_equal(tab, ref) else "  MISMATCH vs brute")
        if ref is None:
            ref = tab
        elif not a.brute:
            ok = "  == first" if np.array_equal(tab, ref) else "  MISMATCH vs first"
        print(f"grid lam={lam} brute_rows={rows}: {best:.4f}s{ok}", flush=True)
if os.environ.get("KNN_CLUSTERED"):
    # clustered sites: Gaussian blobs of very different scales over a uniform background
    rng = np.random.default_rng(1)
    k = n // 4
    sc = np.concatenate([0.5 + 0.01 * rng.standard_normal((2 * k, c["D"])), rng.random((k, c["D"])),
                         0.2 + 0.0005 * rng.standard_normal((n - 3 * k, c["D"]))])[rng.permutation(n)]
    e.set_data(sc, y)
    e.set_knn_tuning(1.0, 4096)
    for algo in ("auto", "brute"):
        t0 = time.perf_counter()
        e.build_neighbors_grid(c["m"], 0, None, algo)
        dt = time.perf_counter() - t0
        tab = e.get_neighbors()
        print(f"clustered {algo}: {dt:.4f}s used_grid={e.knn_used_grid()}" + ("" if algo == "auto" else f"  equal={np.array_equal(tab, prev)}"), flush=True)
        prev = tab
