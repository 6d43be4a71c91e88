"""ctypes front-end of tests/emul/libemul.so (host instantiation of the device traversal code).
TEST INFRASTRUCTURE ONLY -- see emul.cu."""
from __future__ import annotations

import ctypes as C
import os
import shutil
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "libemul.so")
_CSRC = os.path.join(_HERE, "..", "..", "voxelengine_b200", "csrc")

HIT_DTYPE = np.dtype([("t", "<f4"), ("steps", "<i4"), ("vx", "<i4"), ("vy", "<i4"), ("vz", "<i4"),
                      ("status", "<i4"), ("px", "<f4"), ("py", "<f4"), ("pz", "<f4"),
                      ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4")])


def build(force: bool = False) -> str:
    deps = [os.path.join(_HERE, "emul.cu")] + [os.path.join(_CSRC, f) for f in os.listdir(_CSRC) if f.endswith((".cuh", ".h"))]
    stale = force or not os.path.exists(_LIB) or any(os.path.getmtime(d) > os.path.getmtime(_LIB) for d in deps)
    if stale:
        nvcc = shutil.which("nvcc") or "/usr/local/cuda/bin/nvcc"
        cmd = [nvcc, "-O2", "-std=c++17", "-Wno-deprecated-gpu-targets", "-Xcompiler", "-fPIC,-ffp-contract=off", "-shared",
               "-o", _LIB, os.path.join(_HERE, "emul.cu")]
        if os.path.exists("/usr/bin/g++"):
            cmd[1:1] = ["-ccbin", "/usr/bin/g++"]
        subprocess.run(cmd, check=True, capture_output=True)
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        L = C.CDLL(_LIB)
        L.emul_create.restype = C.c_void_p
        L.emul_create.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        L.emul_destroy.argtypes = [C.c_void_p]
        L.emul_level.argtypes = [C.c_void_p, C.c_int, C.c_void_p, C.c_void_p]
        L.emul_trace.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_int, C.c_int, C.c_int, C.c_int, C.c_void_p, C.c_void_p]
        L.emul_classify.argtypes = [C.c_void_p, C.c_void_p, C.c_longlong, C.c_int, C.c_void_p, C.c_float, C.c_void_p]
        _lib = L
    return _lib


class Emul:
    def __init__(self, volume: np.ndarray):
        self.volume = np.ascontiguousarray(volume, np.uint8)
        sz, sy, sx = self.volume.shape
        self.h = lib().emul_create(self.volume.ctypes.data, sx, sy, sz)

    def level(self, shift: int) -> np.ndarray:
        """occupancy level (cell = 2^shift voxels) as 0/1 uint8 [cz][cy][cx], built by straightforward host code"""
        dims = np.zeros(3, np.int32)
        lib().emul_level(self.h, shift, None, dims.ctypes.data)
        out = np.zeros((dims[2], dims[1], dims[0]), np.uint8)
        lib().emul_level(self.h, shift, out.ctypes.data, dims.ctypes.data)
        return out

    GEOMS = {"ambient": 0, "nogroup": 1, "reflection": 2, "scan": 3, "scan_far": 4, "scan_pre": 5}

    def trace(self, rays: np.ndarray, variant: int, center, fast: bool = True, geom: str = "ambient", direct: bool = True, lockstep: bool = True):
        """-> (hit records, probes that read the volume, total probe count)"""
        rays = np.ascontiguousarray(rays)
        n = len(rays)
        out = np.zeros(n, HIT_DTYPE)
        c = np.asarray(center, np.int32)
        cnt = np.zeros(2, np.uint64)
        lib().emul_trace(self.h, rays.ctypes.data, n, int(variant), c.ctypes.data, int(fast), self.GEOMS[geom], int(direct), int(lockstep), out.ctypes.data, cnt.ctypes.data)
        return out, int(cnt[0]), int(cnt[1])

    def close(self):
        if self.h:
            lib().emul_destroy(self.h)
            self.h = None
