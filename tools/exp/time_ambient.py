"""Experiment: k_ambient on a bench config -- sun+AO, sun only, AO only -- median of N runs with an L2 flush between, plus a digest of
the output planes so differently compiled libraries (VXL_LIB) can be compared bit for bit.  usage: time_ambient.py [config] [reps]"""
import ctypes as C
import hashlib
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))

import numpy as np  # noqa: E402
import torch  # noqa: E402

from voxelengine_b200.capi import check  # noqa: E402
from voxelengine_b200.scenes import VIEW_DTYPE  # noqa: E402
from voxelengine_b200.workloads import Workload  # noqa: E402


def main():
    cfg = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    reps = int(sys.argv[2]) if len(sys.argv) > 2 else 7
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

    out = []
    runs = [("sun+ao", (wl.n_ao, sh, ao)), ("sun", (0, sh, None)), ("ao", (wl.n_ao, None, ao))]
    for n in [int(x) for x in os.environ.get("VXL_EXP_NAO", "").split(",") if x]:
        runs.append((f"ao{n}", (n, None, ao)))
    for name, args in runs:
        ctx.stats_reset()
        run(*args)
        torch.cuda.synchronize()
        st = ctx.stats()
        ts = []
        for _ in range(reps):
            flush_buf.fill_(1)
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record(); run(*args); b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        out.append(f"{name} {float(np.median(ts)):.3f} ms (min {min(ts):.3f}) rays {st['rays']} probes {st['steps']} px {st['pixels']} fetched {ctx.fetched_probes()}")
    run(wl.n_ao, sh, ao)
    torch.cuda.synchronize()
    dig = hashlib.sha1(sh.cpu().numpy().tobytes()).hexdigest()[:12] + " " + hashlib.sha1(ao.cpu().numpy().tobytes()).hexdigest()[:12]
    print(f"[{os.environ.get('VXL_LIB', 'default').split('/')[-1]}] cfg {cfg} | " + " | ".join(out) + " | digest " + dig, flush=True)
    wl.close()


if __name__ == "__main__":
    main()
