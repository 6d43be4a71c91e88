"""Regenerate tests/golden/*.npz from the CPU oracle:  python tests/golden/make_golden.py

The reference ships no golden vectors for this path and cannot run in this image,
so these fixtures are produced by the oracle (oracle/vxo.cpp), whose fidelity is established by
tests/test_oracle_invariants.py and tests/test_oracle_refcheck.py.  They pin the oracle against
regressions and let the GPU tests check the CUDA path without rebuilding the inputs' outputs.
Inputs are rebuilt from seeds by tests/scene_util.py; only outputs are stored."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import scene_util as U  # noqa: E402
from oracle import vxo_py as O  # noqa: E402
from voxelengine_b200 import scenes as S  # noqa: E402


def golden_lights(sc):
    sz, sy, sx = sc["volume"].shape
    pos = [(sx * 0.2 * f, sy * 0.2 * 0.8, sz * 0.2 * g) for f, g in ((0.3, 0.3), (0.6, 0.4), (0.5, 0.5))]
    return S.point_lights(pos, [sx * 0.2 * r for r in (0.35, 0.5, 2.0)])


def main():
    sc = U.house_scene(O)                       # BASELINE.json configs[0] in miniature
    sh, ao, st_a = O.pass_ambient(sc["volume"], sc["view"], sc["gb"], 4)
    lights = golden_lights(sc)
    pt, st_p = O.pass_point(sc["volume"], sc["view"], sc["gb"], lights)
    sp, st_s = O.pass_spot(sc["volume"], sc["view"], sc["gb"], S.spot_lights(lights["Position"], lights["Range"], [(0, -1, 0)] * 3))
    t, st_r = O.pass_reflection(sc["volume"], sc["view"], sc["gb"])
    rays = U.random_rays(np.random.RandomState(42), 4096, (64, 64, 64))
    hits = [O.trace_rays(sc["volume"], rays, v) for v in (0, 1, 2)]
    np.savez_compressed(
        os.path.join(HERE, "house64.npz"),
        volume=sc["volume"], depth24=sc["gb"]["depth24"], normal=sc["gb"]["normal"], material=sc["gb"]["material"],
        shadow=sh.astype(np.uint8), ao=ao, point=pt.astype(np.uint8), spot=sp.astype(np.uint8), spec_t=t,
        stats=np.array([[s["rays"], s["steps"], s["pixels"]] for s in (st_a, st_p, st_s, st_r)], np.uint64),
        hits_sparse=hits[0], hits_supersparse=hits[1], hits_dda=hits[2])
    tr = U.terrain_scene(O)                     # configs[1] in miniature
    sh, ao, st_a = O.pass_ambient(tr["volume"], tr["view"], tr["gb"], 8)
    t, st_r = O.pass_reflection(tr["volume"], tr["view"], tr["gb"])
    np.savez_compressed(
        os.path.join(HERE, "terrain128.npz"),
        volume_crc=np.array([int(np.bitwise_xor.reduce(tr["volume"].view(np.uint32).ravel())), int(tr["volume"].sum())], np.uint64),
        depth24=tr["gb"]["depth24"], shadow=sh.astype(np.uint8), ao=ao, spec_t=t,
        stats=np.array([[s["rays"], s["steps"], s["pixels"]] for s in (st_a, st_r)], np.uint64))
    for f in ("house64.npz", "terrain128.npz"):
        print(f, os.path.getsize(os.path.join(HERE, f)), "bytes")


if __name__ == "__main__":
    main()
