"""How much of the fused kernel's time is gather locality?  Times one evaluation on cfg3 / cfg4-shaped data in the
given (random) ordering and with the SAME points pre-sorted along a Z-order curve (then the predecessors a location
gathers sit close together in the record array).  The sorted ordering is a different model (other neighbour sets); the
probe only answers whether spatially coherent record storage would pay.
    python tools/locality_probe.py cfg3 [n]
"""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynngp_b200 import _lib  # noqa: E402
from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic  # noqa: E402

c = dict(CONFIGS[sys.argv[1] if len(sys.argv) > 1 else "cfg3"])
if len(sys.argv) > 2:
    c["n"] = int(sys.argv[2])
s, y = synthetic(c["n"], c["D"], c["seed"])


def morton(s, bits=10):
    q = np.minimum((s * (1 << bits)).astype(np.uint64), (1 << bits) - 1)
    code = np.zeros(len(s), dtype=np.uint64)
    D = s.shape[1]
    for b in range(bits):
        for d in range(D):
            code |= ((q[:, d] >> np.uint64(b)) & np.uint64(1)) << np.uint64(b * D + d)
    return code


kid = {"exponential": 0, "matern32": 1, "matern52": 2}[c["kernel"]]
prm = torch.tensor([[PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0]], dtype=torch.float64, device="cuda")
out = torch.zeros((1, 3), dtype=torch.float64, device="cuda")
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)
for name, order in (("given ordering", None), ("z-order sorted", np.argsort(morton(s), kind="stable")),
                    ("blocks of 4096 z-sorted, blocks shuffled", "blocks")):
    if order is None:
        ss, yy = s, y
    elif isinstance(order, str):
        o = np.argsort(morton(s), kind="stable")
        nb = len(o) // 4096
        blocks = np.random.default_rng(0).permutation(nb)
        o = np.concatenate([o[b * 4096:(b + 1) * 4096] for b in blocks] + [o[nb * 4096:]])
        ss, yy = s[o], y[o]
    else:
        ss, yy = s[order], y[order]
    e = _lib.Engine(0, "float64")
    e.set_data(np.ascontiguousarray(ss), np.ascontiguousarray(yy))
    e.build_neighbors_grid(c["m"])
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
    print(f"{sys.argv[1] if len(sys.argv) > 1 else 'cfg3'} n={c['n']} {name}: {np.mean(ts):.4f} ms (min {np.min(ts):.4f})  n_bad={out.cpu().numpy()[0][2]}", flush=True)
    e.close()
