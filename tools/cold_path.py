"""Phase times of the cold path: host arrays -> NNGP(...) -> loglik_terms(), repeated.

    python tools/cold_path.py cfg3 [reps]
"""
import os
import sys
import time

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynngp_b200 import NNGP, Exponential, Matern  # noqa: E402
from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 5
c = CONFIGS[name]
s, y = synthetic(c["n"], c["D"], c["seed"])
spec = Exponential(**PARAMS) if c["kernel"] == "exponential" else Matern(1.5, **PARAMS)
for r in range(reps):
    t0 = time.perf_counter()
    mdl = NNGP(s, y, 0.0, "S=T", c["m"], spec)
    t1 = time.perf_counter()
    terms = mdl.loglik_terms()
    t2 = time.perf_counter()
    tm = {k: round(v * 1e3, 2) for k, v in mdl._timings.items()}
    del mdl
    t3 = time.perf_counter()
    print(f"rep {r}: ctor {1e3 * (t1 - t0):.1f} ms {tm}  eval {1e3 * (t2 - t1):.2f} ms  destroy {1e3 * (t3 - t2):.1f} ms", flush=True)
