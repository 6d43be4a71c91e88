"""Regenerate tests/golden/ref_shaders.npz FROM THE REFERENCE'S OWN SHADERS:

    python tests/golden/make_ref_golden.py        (needs /root/reference; see oracle/refcheck/build_shaders.py)

The reference's Light.frag and Light{Ambient,Point,Spot,Reflection}.frag, compiled as C++ on the
reference's vendored glm (oracle/_ref/libvxshader.so), are run on the inputs of the house fixture
(tests/golden/house64.npz: volume + G-buffer, config 0 in miniature) and on 4096 seeded random rays.
Only their outputs are stored.  The oracle is NOT involved in producing the numbers: these are golden
vectors of the reference itself, and they travel to machines where /root/reference does not exist
(tests/test_ref_golden.py checks the oracle on the CPU and the CUDA path on the GPU against them)."""
import os
import sys

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, os.path.dirname(os.path.dirname(HERE)))
sys.path.insert(0, os.path.dirname(HERE))

import scene_util as U  # noqa: E402
from oracle import vxo_py as O  # noqa: E402
from voxelengine_b200 import scenes as S  # noqa: E402

HOUSE_VIEW = dict(width=96, height=64, frame=3)


def house_inputs():
    g = np.load(os.path.join(HERE, "house64.npz"))
    h, w = g["depth24"].shape
    ext = 2 * 32 * 0.1
    view = S.make_view((-ext * 0.2, ext * 0.9, -ext * 0.25), 3.927, -0.5, w, h, HOUSE_VIEW["frame"])
    gb = dict(depth24=g["depth24"], normal=g["normal"], material=g["material"], noise=S.blue_noise(4))
    return g["volume"], view, gb


def lights_for(volume):
    sz, sy, sx = volume.shape
    pos = [(sx * 0.2 * f, sy * 0.2 * 0.8, sz * 0.2 * g) for f, g in ((0.3, 0.3), (0.6, 0.4), (0.5, 0.5))]
    pl = S.point_lights(pos, [sx * 0.2 * r for r in (0.35, 0.5, 2.0)])
    sl = S.spot_lights(pl["Position"], pl["Range"], [(0, -1, 0)] * 3)
    return pl, sl


def compact(rec):
    """(n, discarded, result[2], dist[2], fetches[2], last texel[2][3]) of a SHADER_PIXREC array"""
    r = rec["ray"]
    return dict(n=rec["n"].astype(np.uint8), discarded=rec["discarded"].astype(np.uint8), result=r["result"].copy(), dist=r["dist"].copy(),
                fetches=r["fetches"].copy(), texel=np.stack([r["lx"], r["ly"], r["lz"]], axis=-1).astype(np.int32),
                origin=r["o"].copy(), dir=r["d"].copy())


def main():
    if O.shader_lib() is None:
        O.build(force=True)
    assert O.shader_lib() is not None, "oracle/_ref/libvxshader.so could not be built (is /root/reference mounted?)"
    volume, view, gb = house_inputs()
    pl, sl = lights_for(volume)
    out = {}
    for k, v in compact(O.shader_pass(O.PASS_AMBIENT, volume, view, gb)).items():
        out["ambient_" + k] = v
    for li in range(3):
        for k, v in compact(O.shader_pass(O.PASS_POINT, volume, view, gb, lights=pl, light_index=li)).items():
            out[f"point{li}_" + k] = v[..., 0] if v.ndim >= 3 and v.shape[2] == 2 else v
        for k, v in compact(O.shader_pass(O.PASS_SPOT, volume, view, gb, lights=sl, light_index=li)).items():
            out[f"spot{li}_" + k] = v[..., 0] if v.ndim >= 3 and v.shape[2] == 2 else v
    for k, v in compact(O.shader_pass(O.PASS_REFLECTION, volume, view, gb)).items():
        out["reflection_" + k] = v[..., 0] if v.ndim >= 3 and v.shape[2] == 2 else v
    rays = U.random_rays(np.random.RandomState(42), 4096, (64, 64, 64))
    for v, name in ((0, "sparse"), (1, "supersparse"), (2, "dda")):
        out["hits_" + name] = O.shader_trace(volume, rays, v)
    # the reference's ShadowVoxSystem (oracle/_ref/libvxshadowvox.so) on tests/scene_util.voxeliser_case()
    models, ents, destroy = U.voxeliser_case()
    vbytes, vregions, vorder = O.ref_shadowvox(models, ents, destroy)
    nz = np.flatnonzero(vbytes.ravel())
    out["vox_nonzero_index"] = nz.astype(np.uint32)           # sparse form of the 524x188x524 staging buffer
    out["vox_nonzero_value"] = vbytes.ravel()[nz]
    out["vox_regions"] = vregions
    out["vox_order"] = vorder
    # drop the per-ray origin/dir of the local-light passes (large, and implied by result + fetches)
    for k in list(out):
        if (k.startswith("point") or k.startswith("spot")) and (k.endswith("_origin") or k.endswith("_dir") or k.endswith("_texel")):
            del out[k]
    path = os.path.join(HERE, "ref_shaders.npz")
    np.savez_compressed(path, **out)
    print(path, os.path.getsize(path), "bytes", len(out), "arrays")


if __name__ == "__main__":
    main()
