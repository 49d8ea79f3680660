"""Small driver for ncu captures: cfg3-shaped data, stage 1 once, the fused kernel a few times."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynngp_b200 import _lib, build as _build  # noqa: E402

if os.environ.get("NNGP_TUNE_SHAPE"):  # development library: the shape knob selects the kernel variant under the profiler
    _lib.LIB_PATH = _build.TUNE_LIB
from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dtype = sys.argv[2] if len(sys.argv) > 2 else "float64"
reps = int(sys.argv[3]) if len(sys.argv) > 3 else 5
c = dict(CONFIGS[name])
if len(sys.argv) > 4:
    c["n"] = int(sys.argv[4])
s, y = synthetic(c["n"], c["D"], c["seed"])
e = _lib.Engine(0, dtype)
e.set_data(s, y)
e.build_neighbors_grid(c["m"])
kid = {"exponential": 0, "matern32": 1, "matern52": 2}[c["kernel"]]
prm = np.array([PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0])
for _ in range(reps):
    st = e.loglik(kid, prm)
print(name, dtype, st.tolist(), "launches", e.launch_count())
