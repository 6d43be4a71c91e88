"""Shared scene builders for the tests: small versions of the BASELINE.json configs, built with
the CPU oracle (volumes, G-buffers) from the host-side inputs of voxelengine_b200.scenes."""
import numpy as np

from voxelengine_b200 import scenes as S


def house_scene(oracle, vol_texels=32, model_size=40, width=96, height=64, frame=3):
    """Config 1 in miniature: one .vox-style model voxelised into a (2*vol_texels)^3-voxel volume
    with the reference voxeliser semantics, camera looking at it, oracle primary G-buffer."""
    vol = np.zeros((vol_texels,) * 3, np.uint8)
    model = S.house_model(model_size, seed=1)
    e = S.entities(1)
    off = (2 * vol_texels - model_size) // 2
    e[0]["cur"] = S.transform_matrix((off * 0.1, 0.2, off * 0.1))
    oracle.voxelize(vol, [model], e)
    ext = 2 * vol_texels * 0.1
    # camera on the -x/-z side (the side facing away from SUN_DIR), looking toward +x+z and down
    view = S.make_view((-ext * 0.2, ext * 0.9, -ext * 0.25), 3.927, -0.5, width, height, frame)
    d, n, m = oracle.gbuffer_primary(vol, view, width, height)
    gb = dict(depth24=d, normal=n, material=m, noise=S.blue_noise(4))
    return dict(volume=vol, view=view, gb=gb, model=model, entities=e)


def terrain_scene(oracle, texels=(64, 48, 64), width=160, height=90, frame=5, props=6):
    """Config 2 in miniature: FastNoise terrain, default camera, oracle primary G-buffer."""
    sx, sy, sz = texels
    vol = oracle.gen_terrain(sx, sy, sz)
    if props:
        model = S.house_model(24, seed=1)
        e = S.prop_entities(vol, n=props, model_size=24, seed=2)
        oracle.voxelize(vol, [model], e)
    view = S.default_camera(texels, width, height, frame)
    d, n, m = oracle.gbuffer_primary(vol, view, width, height)
    gb = dict(depth24=d, normal=n, material=m, noise=S.blue_noise(4))
    lights = S.quarter_point_lights(vol, 4)
    return dict(volume=vol, view=view, gb=gb, lights=lights)


def random_rays(rs, n, extent, dist_lo=20.0, dist_hi=300.0):
    """Random rays around/inside a volume of `extent` voxels per axis (some start outside, some
    axis-parallel, some zero-length) for the ray-level parity interface."""
    r = np.zeros(n, dtype=S.RAY_DTYPE)
    ex = np.asarray(extent, np.float32)
    o = rs.uniform(-0.15, 1.15, size=(n, 3)).astype(np.float32) * ex
    d = rs.normal(size=(n, 3)).astype(np.float32)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6).astype(np.float32)
    scale = rs.choice([0.25, 1.0, 1.0, 1.0, 3.0], size=(n, 1)).astype(np.float32)
    d *= scale
    k = n // 16
    d[:k, 0] = 0.0                      # axis-parallel component (exercises 0*inf in the DDA)
    d[k:2 * k] = np.round(d[k:2 * k])   # exact axis / diagonal directions
    o[2 * k:3 * k] = np.round(o[2 * k:3 * k])   # origins on voxel boundaries (ties)
    r["ox"], r["oy"], r["oz"] = o[:, 0], o[:, 1], o[:, 2]
    r["dx"], r["dy"], r["dz"] = d[:, 0], d[:, 1], d[:, 2]
    r["dist"] = rs.uniform(dist_lo, dist_hi, size=n).astype(np.float32)
    return r
