"""Golden vectors OF THE REFERENCE ITSELF (tests/golden/ref_shaders.npz, produced by
tests/golden/make_ref_golden.py from the reference's own shader sources compiled on its vendored glm).
CPU: the oracle reproduces them.  GPU: the CUDA path reproduces them, with no oracle on the path.
Neither test needs /root/reference."""
import importlib.util
import os

import numpy as np
import pytest

import scene_util as U
from voxelengine_b200 import scenes as S

HERE = os.path.dirname(os.path.abspath(__file__))


def _mk():
    spec = importlib.util.spec_from_file_location("make_ref_golden", os.path.join(HERE, "golden", "make_ref_golden.py"))
    m = importlib.util.module_from_spec(spec)
    import sys
    saved = list(sys.path)
    try:
        spec.loader.exec_module(m)
    finally:
        sys.path[:] = saved
    return m


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _expected_planes(g):
    """shadow / ao / point / spot / spec planes implied by the reference's logged march calls."""
    n = g["ambient_n"]
    res = g["ambient_result"]
    shadow = np.where((n >= 1) & (res[..., 0] != np.float32(128.0)), np.float32(0), np.float32(1)).astype(np.float32)
    d = (res[..., 1] / np.float32(128.0)).astype(np.float32)
    ao = np.where(n >= 2, (d * d).astype(np.float32) * np.float32(0.05), np.float32(0)).astype(np.float32)
    loc = {}
    for kind in ("point", "spot"):
        loc[kind] = np.stack([np.where((g[f"{kind}{li}_n"] >= 1) & (g[f"{kind}{li}_result"] < g[f"{kind}{li}_dist"]), np.float32(0), np.float32(1))
                              for li in range(3)]).astype(np.float32)
    t = np.where(g["reflection_n"] >= 1, g["reflection_result"], np.float32(256.0)).astype(np.float32)
    rays = dict(ambient=int(n.sum()), point=int(sum(g[f"point{li}_n"].sum() for li in range(3))),
                spot=int(sum(g[f"spot{li}_n"].sum() for li in range(3))), reflection=int(g["reflection_n"].sum()))
    steps = dict(ambient=int(g["ambient_fetches"].sum()), point=int(sum(g[f"point{li}_fetches"].sum() for li in range(3))),
                 spot=int(sum(g[f"spot{li}_fetches"].sum() for li in range(3))), reflection=int(g["reflection_fetches"].sum()))
    return shadow, ao, loc["point"], loc["spot"], t, rays, steps


def _trunc_half(v):
    v = np.asarray(v, np.int64)
    return np.where(v >= 0, v // 2, -((-v) // 2))


def _check_hits(got, want, variant):
    """got: our HIT records (voxel coordinates); want: reference records (last fetched texel, fetch count)."""
    if variant < 2:
        assert np.array_equal(_bits(got["t"]), _bits(want["t"])), "returned distance"
        assert np.array_equal(got["steps"], want["steps"]), "probe count"
        hit = got["status"] != 0
        for a in ("vx", "vy", "vz"):
            tex = np.where(got["status"] == 1, got[a] >> 1, _trunc_half(got[a]))
            assert np.array_equal(tex[hit], want[a][hit].astype(np.int64)), a
    else:
        ok = got["status"] == 1
        assert np.array_equal(ok, want["status"] == 1)
        for a in ("px", "py", "pz", "nx", "ny", "nz"):
            # a NaN is a NaN: its sign / payload bits are not defined by GLSL (x86 produces 0xFFC00000, sm_100 0x7FFFFFFF)
            x, y = got[a][ok], want[a][ok]
            both_nan = np.isnan(x) & np.isnan(y)
            assert np.array_equal(_bits(x)[~both_nan], _bits(y)[~both_nan]), a
        assert np.array_equal(got["steps"] + ok.astype(np.int32), want["steps"]), "DDA probe count"
        for a in ("vx", "vy", "vz"):
            assert np.array_equal(_trunc_half(got[a])[ok], want[a][ok].astype(np.int64)), a


def test_oracle_reproduces_reference_shader_golden(oracle):
    g = np.load(os.path.join(HERE, "golden", "ref_shaders.npz"))
    m = _mk()
    volume, view, gb = m.house_inputs()
    pl, sl = m.lights_for(volume)
    shadow, ao, point, spot, t, rays, steps = _expected_planes(g)
    gsh, gao, st = oracle.pass_ambient(volume, view, gb, 1)
    assert np.array_equal(_bits(gsh), _bits(shadow)) and np.array_equal(_bits(gao), _bits(ao))
    assert st["rays"] == rays["ambient"] and st["steps"] == steps["ambient"]
    gpt, st = oracle.pass_point(volume, view, gb, pl)
    assert np.array_equal(_bits(gpt), _bits(point)) and st["rays"] == rays["point"] and st["steps"] == steps["point"]
    gsp, st = oracle.pass_spot(volume, view, gb, sl)
    assert np.array_equal(_bits(gsp), _bits(spot)) and st["rays"] == rays["spot"] and st["steps"] == steps["spot"]
    gt, st = oracle.pass_reflection(volume, view, gb)
    assert np.array_equal(_bits(gt), _bits(t)) and st["rays"] == rays["reflection"] and st["steps"] == steps["reflection"]
    r = U.random_rays(np.random.RandomState(42), 4096, (64, 64, 64))
    for v, name in ((0, "sparse"), (1, "supersparse"), (2, "dda")):
        _check_hits(oracle.trace_rays(volume, r, v), g["hits_" + name], v)
    # the rays the reference's main() cast, traced by the oracle's ray-level entry
    lit = g["ambient_n"] >= 2
    for k, variant in ((0, 0), (1, 1)):
        rr = np.zeros(int(lit.sum()), dtype=oracle.RAY_DTYPE)
        o, d = g["ambient_origin"][lit][:, k], g["ambient_dir"][lit][:, k]
        rr["ox"], rr["oy"], rr["oz"], rr["dx"], rr["dy"], rr["dz"] = o[:, 0], o[:, 1], o[:, 2], d[:, 0], d[:, 1], d[:, 2]
        rr["dist"] = 128.0
        h = oracle.trace_rays(volume, rr, variant)
        assert np.array_equal(_bits(h["t"]), _bits(g["ambient_result"][lit][:, k])) and np.array_equal(h["steps"], g["ambient_fetches"][lit][:, k])


def _vox_case(g):
    models, e, destroy = U.voxeliser_case()
    order = g["vox_order"]
    cmds = np.concatenate([e[order], e[np.flatnonzero(destroy)]])
    cmds["flags"][len(e):] = 1                      # VXO_ENT_DESTROY / VXL_ENT_DESTROY
    want = np.zeros(524 * 188 * 524, np.uint8)
    want[g["vox_nonzero_index"]] = g["vox_nonzero_value"]
    return models, cmds, want.reshape(524, 188, 524), g["vox_regions"][1:]   # [0] = the constructor's whole-volume region


def test_oracle_reproduces_reference_voxeliser_golden(oracle):
    g = np.load(os.path.join(HERE, "golden", "ref_shaders.npz"))
    models, cmds, want, want_regions = _vox_case(g)
    vol = np.zeros((524, 188, 524), np.uint8)
    regions, valid = oracle.voxelize(vol, models, cmds)
    assert np.array_equal(vol, want)
    assert regions[valid != 0].tobytes() == np.ascontiguousarray(want_regions).tobytes()


@pytest.mark.gpu
def test_cuda_reproduces_reference_voxeliser_golden(gpu_ctx):
    """Device voxeliser (vxl_volume_voxelize through the C ABI) against the reference ShadowVoxSystem's own output."""
    from voxelengine_b200 import engine as E
    g = np.load(os.path.join(HERE, "golden", "ref_shaders.npz"))
    models, cmds, want, want_regions = _vox_case(g)
    vol = E.ShadowVoxSystem(gpu_ctx)                # the reference's 524 x 188 x 524 texels
    ids = [vol.add_model(m) for m in models]
    cmds = cmds.copy()
    cmds["model"] = np.asarray(ids, np.int32)[cmds["model"]]
    regions, valid = vol.OnUpdate(cmds)
    assert np.array_equal(vol.download(), want)
    assert np.ascontiguousarray(regions[valid != 0]).tobytes() == np.ascontiguousarray(want_regions).tobytes()
    vol.close()


@pytest.mark.gpu
def test_cuda_reproduces_reference_shader_golden(gpu_ctx):
    """CUDA path (through the C ABI) against the reference's own outputs; no oracle call on this test's path."""
    from voxelengine_b200 import engine as E
    g = np.load(os.path.join(HERE, "golden", "ref_shaders.npz"))
    m = _mk()
    volume, view, gbd = m.house_inputs()
    pl, sl = m.lights_for(volume)
    shadow, ao, point, spot, t, rays, steps = _expected_planes(g)
    vol = E.ShadowVoxSystem(gpu_ctx, (32, 32, 32))
    vol.upload(volume)
    h, w = gbd["depth24"].shape
    gb = E.GeometryBuffer(gpu_ctx, w, h)
    gb.set_noise(gbd["noise"])
    gb.set_planes(gbd["depth24"], gbd["normal"], gbd["material"])
    for variant in (1, 0):                       # tile march and plain march
        gpu_ctx.set_variant(variant)
        gpu_ctx.stats_reset()
        sh, a = E.LightAmbientPipeline.Get().Use(view, gb, vol, n_ao=1)
        st = gpu_ctx.stats()
        assert np.array_equal(_bits(sh.cpu().numpy()[0]), _bits(shadow)) and np.array_equal(_bits(a.cpu().numpy()[0]), _bits(ao))
        assert st["rays"] == rays["ambient"] and st["steps"] == steps["ambient"]
        gpu_ctx.stats_reset()
        pt = E.LightPointPipeline.Get().Use(view, gb, vol, lambda p: [p.DrawLight(l["Position"], l["Range"], l["Color"], l["Attenuation"]) for l in pl])
        st = gpu_ctx.stats()
        assert np.array_equal(_bits(pt.cpu().numpy()[:, 0]), _bits(point)) and st["rays"] == rays["point"] and st["steps"] == steps["point"]
        gpu_ctx.stats_reset()
        sp = E.LightSpotPipeline.Get().Use(view, gb, vol, lambda p: [p.DrawLight(l["Position"], l["Range"], l["Color"], l["Attenuation"], l["Direction"], l["Angle"], l["AngleAttenuation"]) for l in sl])
        st = gpu_ctx.stats()
        assert np.array_equal(_bits(sp.cpu().numpy()[:, 0]), _bits(spot)) and st["rays"] == rays["spot"] and st["steps"] == steps["spot"]
        gpu_ctx.stats_reset()
        tt = E.LightReflectionPipeline.Get().Use(view, gb, vol)
        st = gpu_ctx.stats()
        assert np.array_equal(_bits(tt.cpu().numpy()[0]), _bits(t)) and st["rays"] == rays["reflection"] and st["steps"] == steps["reflection"]
    gpu_ctx.set_variant(1)
    r = U.random_rays(np.random.RandomState(42), 4096, (64, 64, 64))
    for v, name in ((0, "sparse"), (1, "supersparse"), (2, "dda")):
        _check_hits(vol.trace_rays(r, v), g["hits_" + name], v)
    vol.close()
