"""CPU tests of the host side (no GPU): the C-ABI library loads and exports every symbol the header
declares, the shard arithmetic, kernel-spec parsing, the NeighborSets view, and the multi-rank
combination over gloo (world_size 2)."""
import os
import re
import subprocess
import sys

import numpy as np
import pytest

from oracle import nngp_oracle as orc
from pynngp_b200 import _lib, kernels
from pynngp_b200.dist import shard_bounds
from pynngp_b200.nngp import NeighborSets
from pynngp_b200.synthetic import synthetic

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def built_lib():
    if not os.path.exists(_lib.LIB_PATH):
        from pynngp_b200.build import build

        build()
    return _lib.load()


def test_library_exports_every_declared_symbol(built_lib):
    header = open(os.path.join(ROOT, "include", "nngp_b200.h")).read()
    declared = set(re.findall(r"\b(nngp_[a-z0-9_]+)\s*\(", header))
    declared.discard("nngp_handle")
    assert declared == set(_lib.SIGNATURES), declared ^ set(_lib.SIGNATURES)
    for name in declared:
        assert hasattr(built_lib, name)
    assert b"sm_100a" in built_lib.nngp_version()


def test_no_cpu_fallback_without_device(built_lib):
    import torch

    if torch.cuda.is_available():
        pytest.skip("a GPU is present")
    with pytest.raises(_lib.NNGPError, match="no CPU fallback"):
        _lib.Engine(0)


def test_product_never_imports_oracle():
    for dirpath, _, files in os.walk(os.path.join(ROOT, "pynngp_b200")):
        for f in files:
            if f.endswith((".py", ".cu", ".cuh", ".h")):
                src = open(os.path.join(dirpath, f)).read()
                assert "import oracle" not in src and "from oracle" not in src and "liboracle" not in src, f


def test_shard_bounds_cover_exactly():
    for n in (0, 1, 7, 1000, 10**6 + 3):
        for world in (1, 2, 3, 4, 8):
            edges = [shard_bounds(n, r, world) for r in range(world)]
            assert edges[0][0] == 0 and edges[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(edges[:-1], edges[1:]))
            sizes = [hi - lo for lo, hi in edges]
            assert max(sizes) - min(sizes) <= 1
    with pytest.raises(ValueError):
        shard_bounds(10, 2, 2)


def test_kernel_spec_parsing():
    assert kernels.parse(None).kernel_id == 0
    assert kernels.parse("matern32").kernel_id == 1
    assert kernels.parse(("matern", 2.5)).kernel_id == 2
    assert kernels.parse(("exponential",)).kernel_id == 0
    assert kernels.Matern(1.5, 1.0, 6.0, 0.1).params() == [1.0, 6.0, 0.1, 0.0]
    assert kernels.Exponential(1.0, 6.0, 0.1).params(phi=3.0) == [1.0, 3.0, 0.1, 0.0]
    with pytest.raises(TypeError):
        kernels.parse(lambda a, b: 0.0)
    with pytest.raises(ValueError):
        kernels.Matern(0.7)
    with pytest.raises(ValueError):
        kernels.parse(None).params()


def test_neighbor_sets_view_matches_reference_layout(golden_dir):
    g = np.load(os.path.join(golden_dir, "ns_test_init_shape.npz"))
    Ns = NeighborSets(g["Ns"])
    assert len(Ns) == 200 and Ns[0] == []
    assert Ns[1].dtype == np.int64 and Ns[1].tolist() == [0]
    assert len(Ns[2]) == 2 and len(Ns[199]) == 3
    for i in range(200):
        assert i not in Ns[i]  # the reference's own assertion, tests/test_init.py:22-23
    assert g["coords"][Ns[7]].shape == (3, 2)


_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["NNGP_ROOT"])
from oracle import nngp_oracle as orc
from pynngp_b200.dist import shard_bounds, allreduce_stats, assemble_table_max, get_world
from pynngp_b200.synthetic import synthetic

dist.init_process_group("gloo")
rank, world = get_world()
s, y = synthetic(2000, 2, 3)
m = 9
# stage 1 split by tiles dealt from the heavy end, assembled with an elementwise MAX
full = orc.c_knn_ordered(s, m)
tile = 128
ntiles = (len(s) + tile - 1) // tile
mine = np.full_like(full, -2)
for t in range(ntiles):
    if (ntiles - 1 - t) % world == rank:
        mine[t * tile:(t + 1) * tile] = full[t * tile:(t + 1) * tile]
table = assemble_table_max(mine)
assert np.array_equal(table, full)
# stages 2-3: contiguous shard per rank, partial statistics summed by allreduce
lo, hi = shard_bounds(len(s), rank, world)
part = np.array([orc.c_loglik(s, y, table, 1, 1.0, 6.0, 0.1, lo=lo, hi=hi)])
tot = allreduce_stats(part)
want = np.array(orc.c_loglik(s, y, full, 1, 1.0, 6.0, 0.1))
np.testing.assert_allclose(tot[0], want, rtol=1e-12)
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_combination(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(_WORKER)
    env = dict(os.environ, NNGP_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29617", str(script)],
        env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout + r.stderr
    assert r.stdout.count("ok") == 2


_CLASS_WORKER = r'''
import os, sys
import numpy as np
import torch.distributed as dist
sys.path.insert(0, os.environ["NNGP_ROOT"]); sys.path.insert(0, os.path.join(os.environ["NNGP_ROOT"], "tests"))
from fake_engine import FakeEngine
from oracle import nngp_oracle as orc
from pynngp_b200 import _lib, Matern
from pynngp_b200.synthetic import synthetic
_lib.Engine = FakeEngine  # the oracle-backed stand-in: this test is about the class's sharded host logic
import pyNNGP

dist.init_process_group("gloo")
rank, world = dist.get_rank(), dist.get_world_size()
s, y = synthetic(1501, 2, 3)
y2 = np.stack([y, 0.5 * y], axis=1)
eps = np.stack([np.full(1501, 0.1), np.linspace(0.0, 0.2, 1501)], axis=1)
obj = pyNNGP.NNGP(s, y2, eps, "S=T", 7, Matern(1.5, 1.0, 6.0, 0.1))
full = orc.c_knn_ordered(s, 7)
lo, hi = obj._shard
assert (lo, hi) == ((1501 * rank) // world, (1501 * (rank + 1)) // world)
# stage 1: the rank built and holds the rows of its own shard only -- no exchange happened
assert obj._engine.neighbor_window() == (lo, hi - lo)
assert np.array_equal(obj._engine.get_neighbor_rows(lo, hi), full[lo:hi])
# stages 2-3: partial statistics of the shard, summed over the ranks, both response columns
want = np.zeros(3)
for c in range(2):
    want += orc.c_loglik(s, y2[:, c], full, 1, 1.0, 6.0, 0.1, eps2=eps[:, c] ** 2)
got = obj.loglik_batch([[1.0, 6.0, 0.1]])[0]
np.testing.assert_allclose(got, want, rtol=1e-12)
np.testing.assert_allclose(obj.loglik_terms(), want[:2], rtol=1e-12)
# the whole table: collectively (all_gather of the row blocks) ...
assert np.array_equal(obj.gather_table(), full)
# ... or by ONE rank alone (no collective: it searches the missing rows itself)
obj2 = pyNNGP.NNGP(s, y, 0.0, "S=T", 7, Matern(1.5, 1.0, 6.0, 0.1))
if rank == 1:
    assert np.array_equal(obj2._table, full) and obj2.Ns[5].tolist() == full[5, :5].tolist()
np.testing.assert_allclose(obj2.loglik_terms(), orc.c_loglik(s, y, full, 1, 1.0, 6.0, 0.1)[:2], rtol=1e-12)
dist.barrier()
dist.destroy_process_group()
print("rank", rank, "ok")
'''


def test_two_rank_gloo_class_shards_stage1_and_likelihood(tmp_path):
    """world_size 2 over gloo: the drop-in class shards the ordering, every rank searches and keeps its own rows
    (no table exchange), the statistics are summed over the ranks; `Ns` can be read by one rank alone."""
    script = tmp_path / "class_worker.py"
    script.write_text(_CLASS_WORKER)
    env = dict(os.environ, NNGP_ROOT=ROOT, MASTER_ADDR="127.0.0.1")
    r = subprocess.run(
        [sys.executable, "-m", "torch.distributed.run", "--nnodes=1", "--nproc-per-node=2",
         "--master-addr", "127.0.0.1", "--master-port", "29619", str(script)],
        env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stdout[-3000:] + r.stderr[-3000:]
    assert r.stdout.count("ok") == 2


def test_bench_work_model_matches_the_survey():
    """bench.py's algorithmic figures are SURVEY 8(d4)'s: FP64-pipe instructions and compulsory HBM bytes per
    location at the four configurations."""
    import bench

    assert bench.fp64_instr_per_location(10, 2, "exponential") == 2301
    assert bench.fp64_instr_per_location(15, 2, "exponential") == 5056
    assert bench.fp64_instr_per_location(15, 2, "matern32") == 5176
    assert bench.fp64_instr_per_location(30, 3, "matern32") == 22716
    assert [bench.hbm_bytes_per_location(m, D) for m, D in ((10, 2), (15, 2), (30, 3))] == [64, 84, 152]


def test_bench_reference_arm_prints_the_contract_line():
    """`bench.py --impl reference` (the CPU arm the driver runs beside the GPU arm): one JSON line on stdout
    with the contract's keys; shortened here through --cpu-seconds / NNGP_BENCH_STAGE1_SIZES."""
    import json

    env = dict(os.environ, NNGP_BENCH_STAGE1_SIZES="150,300")
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--config", "cfg2", "--steps", "1",
                        "--warmup", "0", "--cpu-sample", "1024", "--cpu-seconds", "0.3"],
                       env=env, capture_output=True, text=True, timeout=300)
    assert r.returncode == 0, r.stderr[-2000:]
    lines = [l for l in r.stdout.splitlines() if l.strip()]
    assert len(lines) == 1
    d = json.loads(lines[0])
    assert d["impl"] == "reference" and d["metric"] == "nngp_loglik_evals_per_sec" and d["unit"] == "evals/s"
    assert d["higher_is_better"] is True and d["value"] > 0 and d["n_gpus"] == 1
    assert d["cpu_baseline"]["kind"] == "port" and d["cpu_baseline"]["cores"] >= 1 and d["cpu_baseline"]["value"] == d["value"]
    assert d["e2e"] == {"value": d["value"], "unit": "evals/s", "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}
    assert d["cpu_baseline"]["stage1_reference"]["n"] == [150, 300]
    # ranks other than 0 print nothing and exit 0
    r = subprocess.run([sys.executable, os.path.join(ROOT, "bench.py"), "--impl", "reference", "--gpus", "2"],
                       env=dict(env, RANK="1", WORLD_SIZE="2"), capture_output=True, text=True, timeout=120)
    assert r.returncode == 0 and r.stdout.strip() == ""


def test_tools_and_entry_points_compile():
    """The development harnesses only run on the GPU box: keep them at least syntactically alive here."""
    import glob
    import py_compile

    files = glob.glob(os.path.join(ROOT, "tools", "*.py")) + [os.path.join(ROOT, f) for f in ("bench.py", "__graft_entry__.py")]
    assert len(files) >= 10
    for f in files:
        py_compile.compile(f, doraise=True)
