"""Generates tests/golden/ref_resolve.npz: the out_Color of the REFERENCE'S OWN light-pass shaders (compiled for the host
by oracle/refcheck/build_shaders.py from /root/reference) on the house fixture with the inputs of
tests/scene_util.resolve_case.  Run in the container where /root/reference is mounted:
    python tests/golden/make_ref_resolve_golden.py"""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(HERE))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
import scene_util as U  # noqa: E402
from oracle import vxo_py as O  # noqa: E402


def main():
    O.build()
    assert O.shader_lib() is not None, "reference shaders not built"
    W, H = 64, 48
    sc = U.house_scene(O, width=W, height=H)
    gb, albedo, point, spot = U.resolve_case(sc)
    vol, view = sc["volume"], sc["view"]
    amb = O.shader_pass(O.PASS_AMBIENT, vol, view, gb, albedo=albedo)["color"]
    lit = (gb["depth24"] & 0xFFFFFF).astype(np.float32) / np.float32(16777215.0) < np.float32(0.999)
    amb = np.where(lit[..., None], amb, np.float32(0))            # sky pixels: the sky-box look-up is outside the path
    sums = {}
    for name, which, lights in (("point_sum", O.PASS_POINT, point), ("spot_sum", O.PASS_SPOT, spot)):
        total = np.zeros((H, W, 4), np.float32)
        for li in range(len(lights)):
            rec = O.shader_pass(which, vol, view, gb, lights=lights, light_index=li, albedo=albedo)
            total = total + np.where((rec["discarded"] != 0)[..., None], np.float32(0), rec["color"])
        sums[name] = total
    np.savez_compressed(os.path.join(HERE, "ref_resolve.npz"), width=W, height=H, ambient=amb.astype(np.float32), **sums)
    print("ambient mean", float(amb[lit].mean()), {k: float(v.mean()) for k, v in sums.items()})


if __name__ == "__main__":
    main()
