"""Experiment: where k_ambient's time goes on config 3 -- sun-shadow ray alone, AO rays alone, both.
Run on the GPU box:  python tools/exp/split_ambient.py [config]"""
import ctypes as C
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from voxelengine_b200 import engine as E  # noqa: E402
from voxelengine_b200.capi import check  # noqa: E402
from voxelengine_b200.scenes import VIEW_DTYPE  # noqa: E402
from voxelengine_b200.workloads import Workload  # noqa: E402


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    wl = Workload(cfg)
    ctx, lib = wl.ctx, wl.ctx.lib
    v = np.ascontiguousarray(wl.view, dtype=VIEW_DTYPE).reshape(())
    vp = v.ctypes.data_as(C.c_void_p)
    f = wl.gb.frame()
    sh = ctx.empty(wl.gb.shape, torch.float32)
    ao = ctx.empty(wl.gb.shape, torch.float32)
    flush_buf = torch.empty(256 << 20, dtype=torch.uint8, device=ctx.torch_device)

    def run(n_ao, o_sh, o_ao):
        check(lib.vxl_pass_ambient(ctx.h, wl.vol.h, vp, C.byref(f), n_ao, C.c_void_p(o_sh.data_ptr()) if o_sh is not None else None,
                                   C.c_void_p(o_ao.data_ptr()) if o_ao is not None else None), "ambient")

    for name, args in (("sun+ao", (wl.n_ao, sh, ao)), ("sun only", (0, sh, None)), ("ao only", (wl.n_ao, None, ao)), ("ao x1", (1, None, ao))):
        for variant in (1, 0):
            ctx.set_variant(variant)
            ctx.stats_reset()
            run(*args)
            torch.cuda.synchronize()
            st = ctx.stats()
            ts = []
            for _ in range(6):
                flush_buf.fill_(1)
                a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
                a.record(); run(*args); b.record()
                torch.cuda.synchronize()
                ts.append(a.elapsed_time(b))
            t = float(np.median(ts))
            print(f"{name:9s} variant {variant}: {t:7.3f} ms  rays {st['rays']:>11d} probes {st['steps']:>12d}  "
                  f"{st['steps'] / t / 1e6:8.1f} Gprobes/s", flush=True)
    ctx.set_variant(1)
    wl.close()


if __name__ == "__main__":
    main()
