"""Development timing harness: fused kernel at cfg3 (or argv[1]) under tuning env knobs."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynngp_b200 import _lib, build as _build  # noqa: E402

_lib.LIB_PATH = _build.TUNE_LIB  # the development library (python -m pynngp_b200.build --tune): shape knobs compiled in
from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
dtypes = sys.argv[2].split(",") if len(sys.argv) > 2 else ["float64"]
c = dict(CONFIGS[name])
if len(sys.argv) > 3:
    c["n"] = int(sys.argv[3])
for key in ("D", "m"):  # overrides: TUNE_D=2 TUNE_M=30
    if os.environ.get("TUNE_" + key.upper()):
        c[key] = int(os.environ["TUNE_" + key.upper()])
s, y = synthetic(c["n"], c["D"], c["seed"])
kid = {"exponential": 0, "matern32": 1, "matern52": 2}[c["kernel"]]
prm = torch.tensor([[PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0]], dtype=torch.float64, device="cuda")
out = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
table = None
for dtype in dtypes:
    e = _lib.Engine(0, dtype)
    e.set_data(s, y)
    if table is None:
        t0 = torch.cuda.Event(enable_timing=True); t1 = torch.cuda.Event(enable_timing=True)
        import time
        tt = time.perf_counter(); e.build_neighbors_grid(c["m"]); print(f"knn build {time.perf_counter()-tt:.3f}s", flush=True)
        table = e.get_neighbors()
    else:
        e.set_neighbors(table)
    for knob in os.environ.get("KNOBS", "default").split(","):
        os.environ["NNGP_TUNE_SHAPE"] = knob
        ts = []
        for it in range(13):
            flush.zero_()
            a = torch.cuda.Event(enable_timing=True); b = torch.cuda.Event(enable_timing=True)
            a.record(st)
            e.loglik_device(kid, prm.data_ptr(), 1, out.data_ptr(), st.cuda_stream)
            b.record(st)
            torch.cuda.synchronize()
            if it >= 3:
                ts.append(a.elapsed_time(b))
        print(f"{name} {dtype} shape={knob}: {np.mean(ts):.4f} ms (min {np.min(ts):.4f})  stats={out.cpu().numpy()[0].tolist()}", flush=True)
