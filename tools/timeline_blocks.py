"""Development only: per-block entry / loop-done times of one launch (needs NNGP_TIMELINE)."""
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
lib.nngp_debug_timeline_blocks.argtypes = [ctypes.POINTER(ctypes.c_uint64)]
n = c["n"]
for knob in os.environ.get("KNOBS", "default").split(","):
  os.environ["NNGP_TUNE_SHAPE"] = knob
  for k in (n,):
    lo = n // 2 - k // 2
    e.set_shard(lo, lo + k)
    for it in range(5):
        flush.zero_()
        e.loglik_device(1, prm.data_ptr(), 1, out.data_ptr(), st.cuda_stream)
        torch.cuda.synchronize()
        tl = np.zeros(3072, dtype=np.uint64)
        lib.nngp_debug_timeline_blocks(tl.ctypes.data_as(ctypes.POINTER(ctypes.c_uint64)))
        nb = int((tl.reshape(3, 1024)[0] > 0).sum()) if it == 0 else nb
        t = tl.reshape(3, 1024)[:, :nb].astype(np.int64)
        smid = t[2]
        t0 = t[0].min()
        ent = (t[0] - t0) / 1e3
        done = (t[1] - t0) / 1e3
        if it >= 2:
            qs = np.percentile(done, [0, 10, 50, 90, 100])
            order = np.argsort(done)
            per_sm = {}
            for b in range(nb):
                per_sm.setdefault(int(smid[b]), []).append(done[b])
            sm_mean = sorted((np.mean(v), sm) for sm, v in per_sm.items())
            print(f"{knob} nb={nb} slow SMs " + " ".join(f"{sm}:{v:.0f}" for v, sm in sm_mean[-10:]) + " | fast SMs " + " ".join(f"{sm}:{v:.0f}" for v, sm in sm_mean[:10]))
            print(f"{knob} rows {k}: entry max {ent.max():.1f} us; loop done min/p10/p50/p90/max = " + "/".join(f"{v:.1f}" for v in qs) +
                  f"; mean {done.mean():.1f}; slowest blocks {np.argsort(done)[-6:].tolist()} fastest {np.argsort(done)[:6].tolist()}", flush=True)
