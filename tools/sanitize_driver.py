"""Small end-to-end pass for compute-sanitizer (memcheck / racecheck / initcheck / synccheck): stage 1 through the
grid search and the brute-force kernel, the fused kernel in its production shapes (m = 10, 15, 30, 32; 2-D and 3-D;
fp64 and fp32), the sweep variant, the emitting variant and prediction rows -- all on small inputs, checked
against the oracle so that a sanitizer run is also a parity run.

    compute-sanitizer --tool memcheck python tools/sanitize_driver.py
"""
import os
import sys

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from oracle import nngp_oracle as orc  # noqa: E402  (the checker)
from pynngp_b200 import _lib  # noqa: E402
from pynngp_b200.synthetic import synthetic  # noqa: E402

PRM = np.array([[1.0, 6.0, 0.1, 0.0], [0.7, 9.0, 0.05, 0.0], [1.3, 4.0, 0.2, 0.0]])
cases = [(1000, 2, 10, 0), (2000, 2, 15, 1), (2000, 3, 30, 1), (600, 3, 32, 2), (1500, 1, 6, 0)]
if len(sys.argv) > 1:
    cases = cases[: int(sys.argv[1])]
for n, D, m, kid in cases:
    s, y = synthetic(n, D, 11 + n)
    want_tab = orc.c_knn_ordered(s, m)
    for dtype in ("float64", "float32"):
        e = _lib.Engine(0, dtype)
        e.set_data(s, y)
        e.set_knn_tuning(1.0, 128)
        e.build_neighbors_grid(m, 0, None, "grid")
        assert np.array_equal(e.get_neighbors(), want_tab), "grid table"
        if dtype == "float64":
            e.build_neighbors(m)
            assert np.array_equal(e.get_neighbors(), want_tab), "brute table"
        tol = 1e-10 if dtype == "float64" else 2e-4
        got1 = e.loglik(kid, PRM[0])[0]
        want1 = orc.c_loglik(s, y, want_tab, kid, *PRM[0, :3])
        np.testing.assert_allclose(got1[:2], want1[:2], rtol=tol)
        gotK = e.loglik(kid, PRM)  # K = 3: the sweep variant in fp64
        for k in range(3):
            np.testing.assert_allclose(gotK[k][:2], orc.c_loglik(s, y, want_tab, kid, *PRM[k, :3])[:2], rtol=tol)
        B, F = e.factors(kid, PRM[0], 0, min(n, 300))
        assert np.isfinite(F[1:]).all()
        e.set_shard(n // 3, n // 2)
        part = e.loglik(kid, PRM[0])[0]
        assert np.isfinite(part).all()
        e.close()
    print(f"n={n} D={D} m={m} kernel={kid}: ok", flush=True)
print("sanitize driver ok")
