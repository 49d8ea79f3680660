"""ncu driver for the sweep variant: cfg3-shaped data, one K = 16 launch repeated."""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynngp_b200 import _lib  # noqa: E402
from pynngp_b200.synthetic import CONFIGS, sweep_params, synthetic  # noqa: E402

c = CONFIGS["cfg3"]
s, y = synthetic(c["n"], c["D"], c["seed"])
e = _lib.Engine(0)
e.set_data(s, y)
e.build_neighbors_grid(c["m"])
prm = sweep_params(16)
for _ in range(int(sys.argv[1]) if len(sys.argv) > 1 else 2):
    st = e.loglik(1, prm)
print(st[:2].tolist())
