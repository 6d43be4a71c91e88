"""Per-pass probe / volume-read statistics on a config (variant 2 = counting twin of the default kernels)."""
import os, sys
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
import torch
from voxelengine_b200 import engine as E
from voxelengine_b200.workloads import Workload
cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
wl = Workload(cfg)
n = wl.gb.n_tiles; o = wl.out
wl.ctx.set_variant(2)
def run(name, fn):
    wl.ctx.stats_reset(); fn(); torch.cuda.synchronize()
    st = wl.ctx.stats(); f = wl.ctx.fetched_probes()
    print(f"{name:10s} rays {st['rays']:>11d} probes {st['steps']:>12d} ({st['steps']/max(st['rays'],1):6.1f}/ray) volume reads {f:>11d} ({f/max(st['rays'],1):5.2f}/ray, {100.0*f/max(st['steps'],1):4.1f}% of probes)")
import ctypes as C, numpy as np
from voxelengine_b200.capi import check
from voxelengine_b200.scenes import VIEW_DTYPE
v = np.ascontiguousarray(wl.view, dtype=VIEW_DTYPE).reshape(()); f = wl.gb.frame()
sh = wl.ctx.empty(wl.gb.shape, torch.float32)
run("sun", lambda: check(wl.ctx.lib.vxl_pass_ambient(wl.ctx.h, wl.vol.h, v.ctypes.data_as(C.c_void_p), C.byref(f), 0, C.c_void_p(sh.data_ptr()), None)))
run("ao", lambda: check(wl.ctx.lib.vxl_pass_ambient(wl.ctx.h, wl.vol.h, v.ctypes.data_as(C.c_void_p), C.byref(f), wl.n_ao, None, C.c_void_p(sh.data_ptr()))))
if wl.n_point:
    tmp = wl.ctx.empty((wl.n_point, n, wl.gb.tile_h, wl.gb.tile_w), torch.float32)
    run("point", lambda: wl._point(tmp))
if wl.spec:
    run("reflection", lambda: E.LightReflectionPipeline.Get().Use(wl.view, wl.gb, wl.vol, out_spec_t=o[2, :n]))
wl.close()
