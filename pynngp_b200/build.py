"""Builds pynngp_b200/_C/libnngp_b200.so in-tree with nvcc for sm_100a (no other target exists).

    python -m pynngp_b200.build [--force] [--verbose]

The library is plain CUDA C++ behind the C ABI of include/nngp_b200.h; it links the static CUDA
runtime only (no torch, no NCCL -- the cross-GPU allreduce is the Python host's job).
"""
from __future__ import annotations

import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

PKG = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(PKG, "csrc")
OUT_DIR = os.path.join(PKG, "_C")
OBJ_DIR = os.path.join(OUT_DIR, "obj")
LIB = os.path.join(OUT_DIR, "libnngp_b200.so")

SOURCES = [
    "nngp_api.cu", "knn_ordered.cu", "knn_grid.cu", "pack.cu", "fma_peak.cu",
    "fused_f64_exp.cu", "fused_f64_m32.cu", "fused_f64_m52.cu",
    "fused_f32_exp.cu", "fused_f32_m32.cu", "fused_f32_m52.cu",
]
NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-O3", "-std=c++17", "-lineinfo",
    "-Xcompiler", "-fPIC", "-Xptxas", "-v",
]


def _nvcc():
    for cand in (os.environ.get("NVCC"), "/usr/local/cuda/bin/nvcc", "nvcc"):
        if cand and (os.path.isabs(cand) and os.path.exists(cand) or not os.path.isabs(cand)):
            return cand
    raise RuntimeError("nvcc not found")


def _deps():
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".cuh", ".h"))]
    hdrs.append(os.path.join(os.path.dirname(PKG), "include", "nngp_b200.h"))
    return hdrs


def _stale(target, deps):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(d) > t for d in deps)


def _compile(src, verbose, obj_dir=OBJ_DIR, defines=()):
    obj = os.path.join(obj_dir, src.replace(".cu", ".o"))
    path = os.path.join(CSRC, src)
    if not _stale(obj, [path] + _deps()):
        return obj, ""
    cmd = [_nvcc(), *NVCC_FLAGS, *[f"-D{d}=1" for d in defines], "-c", path, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    if r.returncode != 0:
        raise RuntimeError(f"nvcc failed on {src}:\n{r.stdout}\n{r.stderr}")
    with open(obj.replace(".o", ".ptxas.log"), "w") as f:  # registers / spills per kernel; timings dropped (they churn)
        f.write("".join(l for l in r.stderr.splitlines(True) if "Compile time" not in l))
    return obj, r.stderr if verbose else ""


TUNE_LIB = os.path.join(OUT_DIR, "libnngp_b200_tune.so")


def build_tune(force=False):
    """The DEVELOPMENT library tools/tune.py loads explicitly: same sources with NNGP_TUNE (extra kernel shapes selected
    by NNGP_TUNE_SHAPE).  The product library never contains those knobs."""
    # NNGP_DEV_DEFINES=A,B replaces the default define set (e.g. NNGP_DIRECT alone: one experiment, a fast build)
    defines = tuple(d for d in os.environ.get("NNGP_DEV_DEFINES", "NNGP_TUNE").split(",") if d)
    return build(force=force, lib=TUNE_LIB, obj_dir=os.path.join(OUT_DIR, "obj_tune"), defines=defines)


def build(force=False, verbose=False, lib=LIB, obj_dir=OBJ_DIR, defines=()):
    LIB, OBJ_DIR = lib, obj_dir  # noqa: N806 (shadow the module defaults for a development build)
    os.makedirs(OBJ_DIR, exist_ok=True)
    if force:
        for f in os.listdir(OBJ_DIR):
            os.remove(os.path.join(OBJ_DIR, f))
    with ThreadPoolExecutor(max_workers=min(8, os.cpu_count() or 1)) as ex:
        results = list(ex.map(lambda s: _compile(s, verbose, OBJ_DIR, defines), SOURCES))
    objs = [o for o, _ in results]
    if verbose:
        for _, log in results:
            if log:
                print(log)
    if force or _stale(LIB, objs):
        cmd = [_nvcc(), "-shared", "-o", LIB, *objs, "-gencode", "arch=compute_100a,code=sm_100a"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            raise RuntimeError(f"link failed:\n{r.stdout}\n{r.stderr}")
    return LIB


if __name__ == "__main__":
    if "--tune" in sys.argv:
        print(build_tune(force="--force" in sys.argv))
    else:
        print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
