"""Fixed cost of one evaluation: the fused kernel timed on ONE GPU over shards [0, k) of cfg3's ordering
(k = n/8 is what a rank holds in an 8-GPU strong-scaling run), L2 flushed between launches and warm."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynngp_b200 import _lib, build as _build  # noqa: E402

if os.environ.get("NNGP_USE_DEV_LIB"):  # A/B against the development library (python -m pynngp_b200.build --tune)
    _lib.LIB_PATH = _build.TUNE_LIB
from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic  # noqa: E402

c = dict(CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"])
s, y = synthetic(c["n"], c["D"], c["seed"])
kid = {"exponential": 0, "matern32": 1, "matern52": 2}[c["kernel"]]
prm = torch.tensor([[PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0]], dtype=torch.float64, device="cuda")
out = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
e = _lib.Engine(0, "float64")
e.set_data(s, y)
e.build_neighbors_grid(c["m"])
n = c["n"]
for k in (8, 9472, n // 16, n // 8, n // 4, n // 2, n):
    lo = n // 2 - k // 2
    e.set_shard(lo, lo + k)
    for mode in ("flushed", "warm"):
        ts = []
        for it in range(23):
            if mode == "flushed":
                flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(st)
            e.loglik_device(kid, prm.data_ptr(), 1, out.data_ptr(), st.cuda_stream)
            b.record(st)
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(a.elapsed_time(b))
        print(f"rows {k:8d} {mode:8s}: {np.mean(ts)*1e3:8.1f} us (min {np.min(ts)*1e3:8.1f})  per 1e6 rows {np.mean(ts)/k*1e6:.4f} ms", flush=True)
