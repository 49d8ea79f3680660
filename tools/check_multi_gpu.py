"""torchrun worker: the sharded engine (one process per GPU, NCCL) must reproduce the single-GPU
result.  Usage: python -m torch.distributed.run --nproc-per-node N tools/check_multi_gpu.py [cfg] [n]"""
import os
import sys

import numpy as np
import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from pynngp_b200 import NNGP, Matern, _lib  # noqa: E402
from pynngp_b200.synthetic import CONFIGS, PARAMS, synthetic  # noqa: E402

name = sys.argv[1] if len(sys.argv) > 1 else "cfg2"
c = dict(CONFIGS[name])
if len(sys.argv) > 2:
    c["n"] = int(sys.argv[2])
local = int(os.environ.get("LOCAL_RANK", "0"))
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
rank, world = dist.get_rank(), dist.get_world_size()
s, y = synthetic(c["n"], c["D"], c["seed"])
spec = Matern(1.5, **PARAMS)
model = NNGP(s, y, 0.0, "S=T", c["m"], spec, devices=local)         # sharded: every rank searches and keeps its own rows
lo, hi = model._shard
assert model._engine.neighbor_window() == (lo, hi - lo)
terms = model.loglik_terms()
# the fused peer-memory exchange against the NCCL allreduce: same totals (summation order differs)
peer = model._peer_ok
model._peer_ok = False
terms_nccl = model.loglik_terms()
model._peer_ok = peer
np.testing.assert_allclose(terms, terms_nccl, rtol=1e-13)
batch = model.loglik_batch(np.array([[1.0, 6.0, 0.1], [1.5, 9.0, 0.2], [0.7, 4.0, 0.05]]))
allb = [None] * dist.get_world_size()
dist.all_gather_object(allb, batch.tobytes())
assert all(b == allb[0] for b in allb), "ranks disagree on the exchanged totals"
# more parameter vectors than the exchange buffer holds (chunks), and 300 generations back to back
from pynngp_b200.synthetic import sweep_params
big = model.loglik_batch(sweep_params(200)[:, :3])
for k in (0, 131, 199):
    np.testing.assert_allclose(big[k], model.loglik_batch(sweep_params(200)[k:k + 1, :3])[0], rtol=1e-11)
for it in range(300):
    assert model.loglik_terms() == terms, it
table_all = model.gather_table()                                        # collective: all_gather of the row blocks
if rank == 0:
    e = _lib.Engine(local)                                              # single-GPU reference on rank 0
    e.set_data(s, y)
    e.build_neighbors(c["m"])
    assert np.array_equal(e.get_neighbors(), table_all), "assembled neighbour table differs"
    one = e.loglik(1, np.array([PARAMS["sigma2"], PARAMS["phi"], PARAMS["tau2"], 0.0]))[0]
    np.testing.assert_allclose(terms, one[:2], rtol=1e-12)
    print(f"multi-gpu ok: world={world} n={c['n']} m={c['m']} D={c['D']} peer_exchange={peer} terms={terms} knn_s={model._timings['knn_s']:.3f}")
dist.barrier()
dist.destroy_process_group()
