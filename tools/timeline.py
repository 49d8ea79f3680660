"""Development only: where the fixed cost of one launch goes (needs a build with NNGP_TIMELINE)."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynngp_b200 import _lib, build as _build  # noqa: E402

_lib.LIB_PATH = _build.TUNE_LIB  # the development library built with NNGP_DEV_DEFINES=NNGP_TIMELINE
from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic  # noqa: E402

c = dict(CONFIGS["cfg3"])
s, y = synthetic(c["n"], c["D"], c["seed"])
prm = torch.tensor([[PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0]], dtype=torch.float64, device="cuda")
out = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
e = _lib.Engine(0, "float64")
e.set_data(s, y)
e.build_neighbors_grid(c["m"])
lib = _lib.load()
lib.nngp_debug_timeline.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
n = c["n"]
names = ["entry", "table+sync", "first records", "loop done", "partials written", "ticket", "last: enter", "last: end"]
for k in (8, 9472, n // 8, n):
    lo = n // 2 - k // 2
    e.set_shard(lo, lo + k)
    for mode in ("flushed", "warm"):
        rows = []
        for it in range(12):
            if mode == "flushed":
                flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(st)
            e.loglik_device(1, prm.data_ptr(), 1, out.data_ptr(), st.cuda_stream)
            b.record(st)
            torch.cuda.synchronize()
            tl = np.zeros(64, dtype=np.uint64)
            lib.nngp_debug_timeline(tl.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
            if it >= 2:
                g = tl[:8].astype(np.int64)
                rows.append(np.concatenate([[a.elapsed_time(b) * 1e3], (g - g[0]) / 1e3]))
        r = np.median(np.array(rows), axis=0)
        print(f"rows {k:8d} {mode:8s}: events {r[0]:7.1f} us | " + " | ".join(f"{nm} {v:6.1f}" for nm, v in zip(names[1:], r[2:])), flush=True)
