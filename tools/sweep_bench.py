"""cfg5: K parameter vectors at cfg3's data, evaluated one launch each vs one batched launch (sweep variant)."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynngp_b200 import _lib  # noqa: E402
from pynngp_b200.synthetic import CONFIGS, sweep_params, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg3"
K = int(sys.argv[2]) if len(sys.argv) > 2 else 64
c = CONFIGS[name]
s, y = synthetic(c["n"], c["D"], c["seed"])
kid = {"exponential": 0, "matern32": 1, "matern52": 2}[c["kernel"]]
e = _lib.Engine(0)
e.set_data(s, y)
e.build_neighbors_grid(c["m"])
prm = sweep_params(K)
d_prm = torch.tensor(prm, dtype=torch.float64, device="cuda")
d_out = torch.zeros((K, 3), dtype=torch.float64, device="cuda")
d_one = torch.zeros((K, 3), dtype=torch.float64, device="cuda")
st = torch.cuda.Stream()
torch.cuda.set_stream(st)


def timed(fn, reps=5):
    fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(st); fn(); b.record(st)
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return float(np.min(ts))


def seq():
    for k in range(K):
        e.loglik_device(kid, d_prm[k].data_ptr(), 1, d_one[k].data_ptr(), st.cuda_stream)


def batch():
    e.loglik_device(kid, d_prm.data_ptr(), K, d_out.data_ptr(), st.cuda_stream)


t_seq, t_bat = timed(seq), timed(batch)
a, b = d_one.cpu().numpy(), d_out.cpu().numpy()
print(f"{name} K={K}: sequential {t_seq:.3f} ms ({t_seq / K:.4f} ms/eval)  batched {t_bat:.3f} ms ({t_bat / K:.4f} ms/eval)  "
      f"speed-up {t_seq / t_bat:.3f}  max rel diff {np.max(np.abs(a[:, :2] - b[:, :2]) / np.abs(a[:, :2])):.2e}")
