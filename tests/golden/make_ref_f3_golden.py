"""Generates tests/golden/ref_f3.npz: the out_Color of the REFERENCE'S OWN LightTAA.frag and LightReflection.frag (compiled for
the host by oracle/refcheck/build_shaders.py from /root/reference) on the house fixture with the inputs of
tests/scene_util.f3_case.  Run in the container where /root/reference is mounted:
    python tests/golden/make_ref_f3_golden.py"""
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
    sc = U.house_scene(O, width=64, height=48)
    c = U.f3_case(O, sc)
    taa = O.shader_taa(sc["view"], c["gb"], c["albedo"], c["motion"], c["light"], c["last_light"])
    refl = O.shader_pass(O.PASS_REFLECTION, sc["volume"], sc["view"], c["gb"], light=np.nan_to_num(taa, nan=0.0, posinf=0.0, neginf=0.0),
                         sky=c["sky"])["color"]
    np.savez_compressed(os.path.join(HERE, "ref_f3.npz"), taa=taa.astype(np.float32), reflection=refl.astype(np.float32))
    print("taa mean", float(np.nanmean(taa)), "reflection mean", float(refl.mean()))


if __name__ == "__main__":
    main()
