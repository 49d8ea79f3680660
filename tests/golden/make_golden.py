"""Generate golden neighbour sets by running the UNMODIFIED reference (pyNNGP/nngp.py:49-62).

Run in the build container only (needs /root/reference; the GPU box has no copy):
    python tests/golden/make_golden.py
The reference imports `past.builtins.basestring` (nngp.py:2), which is not installed here; an
in-memory stand-in module is injected before import -- the reference's files are not touched.
Outputs tests/golden/ns_<case>.npz: inputs (seed, n, D, m), coordinates, and the reference's Ns as a
dense (n, m) int32 table padded with -1 (Ns[0] == [] in the reference -> row of -1).
"""
import os
import sys
import types

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))

CASES = {
    # name: (n, D, m, seed, kind)
    "test_init_shape": (200, 2, 3, 11, "uniform"),     # the reference's own test shape, tests/test_init.py:6-16
    "cfg1": (1000, 2, 10, 1, "uniform"),               # BASELINE.json configs[0]
    "d2_m15": (2000, 2, 15, 21, "uniform"),
    "d3_m30": (600, 3, 30, 22, "uniform"),
    "d1_m5": (300, 1, 5, 23, "uniform"),
    "d3_m32": (300, 3, 32, 24, "uniform"),
    "lattice": (400, 2, 8, 0, "lattice"),               # ties: parity only on distance multisets
}


def load_reference():
    past = types.ModuleType("past")
    builtins_ = types.ModuleType("past.builtins")
    builtins_.basestring = str
    past.builtins = builtins_
    sys.modules.setdefault("past", past)
    sys.modules.setdefault("past.builtins", builtins_)
    sys.path.insert(0, "/root/reference")
    import pyNNGP  # the reference package

    assert pyNNGP.__file__.startswith("/root/reference"), pyNNGP.__file__
    return pyNNGP


def make_inputs(n, D, seed, kind):
    if kind == "lattice":
        g = int(round(np.sqrt(n)))
        xx, yy = np.meshgrid(np.arange(g) / g, np.arange(g) / g, indexing="ij")
        return np.stack([xx.ravel(), yy.ravel()], axis=1)
    from pynngp_b200.synthetic import synthetic

    return synthetic(n, D, seed)[0]


def main():
    ref = load_reference()
    for name, (n, D, m, seed, kind) in CASES.items():
        s = make_inputs(n, D, seed, kind)
        n = len(s)
        y = np.zeros(n)
        obj = ref.NNGP(s, y, 0.001 * np.ones(n), "S=T", m, None)
        tab = np.full((n, m), -1, dtype=np.int32)
        for i, row in enumerate(obj.Ns):
            tab[i, : len(row)] = row
        np.savez_compressed(os.path.join(HERE, f"ns_{name}.npz"), coords=s, Ns=tab, n=n, D=D, m=m,
                            seed=seed, kind=kind, ws=np.asarray(obj.ws))
        print(name, tab.shape, "ok")


if __name__ == "__main__":
    main()
