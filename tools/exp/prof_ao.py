"""Profiling helper: config-3 ambient pass, selectable rays.  usage: prof_ao.py [ao|sun|both] [reps]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import numpy as np  # noqa: E402
import torch  # noqa: E402
from voxelengine_b200.capi import check  # noqa: E402
from voxelengine_b200.scenes import VIEW_DTYPE  # noqa: E402
from voxelengine_b200.workloads import Workload  # noqa: E402

what = sys.argv[1] if len(sys.argv) > 1 else "ao"
reps = int(sys.argv[2]) if len(sys.argv) > 2 else 2
wl = Workload(3)
ctx, lib = wl.ctx, wl.ctx.lib
v = np.ascontiguousarray(wl.view, dtype=VIEW_DTYPE).reshape(())
f = wl.gb.frame()
sh = ctx.empty(wl.gb.shape, torch.float32)
ao = ctx.empty(wl.gb.shape, torch.float32)
for _ in range(reps):
    check(lib.vxl_pass_ambient(ctx.h, wl.vol.h, v.ctypes.data_as(C.c_void_p), C.byref(f), int(os.environ.get("VXL_EXP_NAO1", wl.n_ao)) if what != "sun" else 0,
                               C.c_void_p(sh.data_ptr()) if what != "ao" else None, C.c_void_p(ao.data_ptr()) if what != "sun" else None), "ambient")
    torch.cuda.synchronize()
wl.close()
