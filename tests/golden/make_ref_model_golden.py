"""Generates tests/golden/ref_model.npz: hit records of the REFERENCE'S OWN clipToAABB + intersectVolume (GeometryVoxel.frag
compiled for the host by oracle/refcheck/build_shaders.py from /root/reference) for the rays of
tests/scene_util.model_rays on the glassy house model.  Run where /root/reference is mounted:
    python tests/golden/make_ref_model_golden.py"""
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
    m = U.glassy_house(40)
    n, frame = 3000, 1
    rays = U.model_rays(m.shape, n, seed=11)
    h = O.shader_model_trace(m, rays, frame=frame, res=(1280.0, 720.0))
    np.savez_compressed(os.path.join(HERE, "ref_model.npz"), n=n, frame=frame, hit=h["hit"], material=h["material"], fetches=h["fetches"],
                        pos=h["pos"], normal=h["normal"])
    print("hits", float(h["hit"].mean()), "fetches max", int(h["fetches"].max()))


if __name__ == "__main__":
    main()
