"""Build voxelengine_b200/libvxl.so in-tree with nvcc for sm_100a.

    python -m voxelengine_b200.build [--force] [--verbose]

--fmad=false is part of the numerical contract (vxl_math.cuh): the kernels must produce the IEEE
single-precision result of the reference shaders' operation sequence, two roundings per a*b+c.
"""
from __future__ import annotations

import glob
import os
import shutil
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OUT = os.path.join(HERE, "libvxl.so")

NVCC_FLAGS = [
    "-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "--fmad=false", "-std=c++17",
    "-Xcompiler", "-fPIC", "-shared",
]


def sources():
    return sorted(glob.glob(os.path.join(CSRC, "*.cu")))


def _deps():
    return sources() + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + [
        os.path.join(HERE, "..", "include", "vxl.h"), os.path.abspath(__file__)]


def is_stale() -> bool:
    if not os.path.exists(OUT):
        return True
    t = os.path.getmtime(OUT)
    return any(os.path.getmtime(p) > t for p in _deps())


def nvcc_path() -> str:
    p = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
    if not os.path.exists(p):
        raise RuntimeError("nvcc not found: libvxl.so cannot be built (there is no CPU fallback)")
    return p


def build(force: bool = False, verbose: bool = False) -> str:
    extra = os.environ.get("VXL_NVCC_EXTRA", "").split()      # experiments only, e.g. -DVXL_AMBIENT_BLOCKS=3
    out = os.environ.get("VXL_LIB_OUT") or OUT                # experiments only: a differently compiled copy for A/B runs (VXL_LIB loads it)
    if out == OUT and not force and not is_stale():
        return OUT
    cmd = [nvcc_path()] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-o", out] + sources()
    env = dict(os.environ)
    if os.path.exists("/usr/bin/g++"):
        cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
    r = subprocess.run(cmd, capture_output=True, text=True, env=env)
    if verbose or r.returncode != 0:
        sys.stderr.write(r.stdout + r.stderr)
    if r.returncode != 0:
        raise RuntimeError("nvcc failed building libvxl.so")
    return out


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
