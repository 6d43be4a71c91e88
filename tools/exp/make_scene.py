"""Experiment helper (not product): build the config-3 scene on the CPU with the oracle and cache it in /tmp."""
import os, sys, time
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)
from oracle import vxo_py as O
from voxelengine_b200 import scenes as S

def main(out="/tmp/cfg3_scene.npz", tex=512, W=3840, H=2160):
    t = time.time()
    vol = O.gen_terrain(tex, tex, tex)
    print("terrain", time.time() - t, vol.shape, vol.dtype); t = time.time()
    model = S.house_model(40, seed=1)
    e = S.prop_entities(vol, n=200, model_size=40, seed=2, model=0)
    res = O.voxelize(vol, [model], e)
    print("voxelize", time.time() - t, type(res)); t = time.time()
    view = S.default_camera((tex, tex, tex), W, H, 0)
    gb = O.gbuffer_primary(vol, view, W, H)
    print("gbuffer", time.time() - t, type(gb)); t = time.time()
    noise = S.blue_noise(4)
    np.savez(out, volume=vol, view=view, depth24=gb[0], normal=gb[1], material=gb[2], noise=noise)
    print("saved", out)

if __name__ == "__main__":
    main()
