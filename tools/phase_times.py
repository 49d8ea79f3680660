"""Wall-clock of each C-ABI call of the cold path (create, set_data, build, first evaluation, destroy), repeated.
    python tools/phase_times.py cfg3 [reps]"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import numpy as np  # noqa: E402

from pynngp_b200 import _lib  # noqa: E402
from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 6
c = CONFIGS[name]
s, y = synthetic(c["n"], c["D"], c["seed"])
kid = {"exponential": 0, "matern32": 1, "matern52": 2}[c["kernel"]]
prm = np.array([PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0])
for r in range(reps):
    t = [time.perf_counter()]
    e = _lib.Engine(0, "float64"); t.append(time.perf_counter())
    e.set_data(s, y); t.append(time.perf_counter())
    e.build_neighbors_grid(c["m"]); t.append(time.perf_counter())
    e.loglik(kid, prm); t.append(time.perf_counter())
    e.loglik(kid, prm); t.append(time.perf_counter())
    e.close(); t.append(time.perf_counter())
    d = np.diff(t) * 1e3
    print(f"rep {r}: create {d[0]:.2f}  set_data {d[1]:.2f}  build {d[2]:.2f}  eval#1 {d[3]:.2f}  eval#2 {d[4]:.2f}  destroy {d[5]:.2f}  total {sum(d):.2f} ms", flush=True)
