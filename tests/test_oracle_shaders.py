"""CPU: the oracle against THE REFERENCE'S OWN SHADERS run here.

oracle/refcheck/build_shaders.py compiles Sources/Shaders/lib/Light.frag and the four light-pass
fragment shaders of the reference, from where they lie under /root/reference, as C++ on the
reference's vendored glm (oracle/_ref/libvxshader.so).  These tests pin oracle/vxo.cpp to it bit for
bit: distance, probe count and hit texel of the three tracers on random and degenerate rays, and for
every pixel of two scenes the rays the shaders' main() actually cast (results, probe counts) against
the oracle's shadow / AO / point / spot / specular planes and its ray and probe totals.
Skipped when the library is absent (it is built whenever the reference tree is mounted; the committed
fixture tests/golden/ref_shaders.npz carries its outputs to machines without the reference).
"""
import numpy as np
import pytest

import scene_util as U
from voxelengine_b200 import scenes as S


@pytest.fixture(scope="module")
def sh(oracle):
    if oracle.shader_lib() is None:
        pytest.skip("oracle/_ref/libvxshader.so not built (reference tree not mounted)")
    return oracle


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _trunc_half(v):
    """C integer division by 2 (truncation toward zero), as getVolumeAt's `pos /= 2`."""
    v = np.asarray(v, np.int64)
    return np.where(v >= 0, v // 2, -((-v) // 2))


def check_rays_against_reference(oracle, volume, rays):
    """Compares oracle.trace_rays with the compiled reference functions; returns the number of rays compared."""
    for variant in (oracle.SPARSE, oracle.SUPERSPARSE):
        want = oracle.shader_trace(volume, rays, variant)
        got = oracle.trace_rays(volume, rays, variant)
        assert np.array_equal(_bits(got["t"]), _bits(want["t"])), f"variant {variant}: returned distance"
        assert np.array_equal(got["steps"], want["steps"]), f"variant {variant}: probe count"
        hit = got["status"] != 0
        # phase 1 hit: voxel = texel*2 + bit; phase 2 hit: voxel = ivec3(pos), fetched at voxel/2 (C division)
        for a in ("vx", "vy", "vz"):
            tex = np.where(got["status"] == 1, got[a] >> 1, _trunc_half(got[a]))
            assert np.array_equal(tex[hit], want[a][hit].astype(np.int64)), f"variant {variant}: hit texel {a}"
    want = oracle.shader_trace(volume, rays, oracle.DDA)
    got = oracle.trace_rays(volume, rays, oracle.DDA)
    ok = got["status"] == 1
    assert np.array_equal(ok, want["status"] == 1), "DDA: hit / miss"
    for a in ("px", "py", "pz", "nx", "ny", "nz"):
        assert np.array_equal(_bits(got[a][ok]), _bits(want[a][ok])), f"DDA: {a}"
    # the oracle's `steps` is nt (completed advances); a hit performed one more probe than that; a walk that left the
    # inclusive bounds returned before probing
    assert np.array_equal(got["steps"] + ok.astype(np.int32), want["steps"]), "DDA: probe count"
    for a in ("vx", "vy", "vz"):
        assert np.array_equal(_trunc_half(got[a])[ok], want[a][ok].astype(np.int64)), f"DDA: hit texel {a}"
    return 3 * len(rays)


def test_tracers_match_reference_shader_source(sh):
    oracle = sh
    rs = np.random.RandomState(11)
    n = 0
    sc = U.terrain_scene(oracle)
    sz, sy, sx = sc["volume"].shape
    n += check_rays_against_reference(oracle, sc["volume"], U.random_rays(rs, 60000, (2 * sx, 2 * sy, 2 * sz)))
    hs = U.house_scene(oracle)
    n += check_rays_against_reference(oracle, hs["volume"], U.random_rays(rs, 60000, (64, 64, 64), dist_lo=1.0, dist_hi=400.0))
    # empty and full volumes, NaN / inf / zero directions
    r = U.random_rays(rs, 4000, (64, 64, 64))
    r["dx"][:50] = 0; r["dy"][:50] = 0; r["dz"][:50] = 0
    r["dist"][50:100] = 0.0
    r["dist"][100:150] = 16.0
    r["dist"][150:200] = 164.0
    n += check_rays_against_reference(oracle, np.zeros((32, 32, 32), np.uint8), r)
    n += check_rays_against_reference(oracle, np.full((32, 32, 32), 255, np.uint8), r)
    assert n > 300000


def _planes_from_records(rec, which):
    """shadow / ao / t planes implied by the logged calls of the shader's main()."""
    n = rec["n"]
    r0, r1 = rec["ray"][..., 0], rec["ray"][..., 1]
    if which == "ambient":
        shadow = np.where((n >= 1) & (r0["result"] != np.float32(128.0)), np.float32(0), np.float32(1))   # LightAmbient.frag:167-169
        d = (r1["result"] / np.float32(128.0)).astype(np.float32)                                        # :121
        ao = np.where(n >= 2, (d * d).astype(np.float32) * np.float32(0.05), np.float32(0))              # :125
        return shadow.astype(np.float32), ao.astype(np.float32)
    if which == "local":
        return np.where((n >= 1) & (r0["result"] < r0["dist"]), np.float32(0), np.float32(1)).astype(np.float32)
    return np.where(n >= 1, r0["result"], np.float32(256.0)).astype(np.float32)


@pytest.mark.parametrize("scene", ["house", "terrain"])
def test_light_passes_match_reference_shader_source(sh, scene):
    oracle = sh
    sc = U.house_scene(oracle) if scene == "house" else U.terrain_scene(oracle)
    vol, view, gb = sc["volume"], sc["view"], sc["gb"]
    h, w = gb["depth24"].shape
    lit = (gb["depth24"] & 0xFFFFFF).astype(np.float32) / np.float32(16777215.0) < np.float32(0.999)

    # ---- ambient: sun shadow (Sparse) + one AO ray (SuperSparse) per lit pixel ----
    rec = oracle.shader_pass(oracle.PASS_AMBIENT, vol, view, gb)
    assert np.array_equal(rec["n"], np.where(lit, 2, 0))
    assert np.all(rec["ray"][..., 0]["variant"][lit] == 0) and np.all(rec["ray"][..., 1]["variant"][lit] == 1)
    wsh, wao = _planes_from_records(rec, "ambient")
    gsh, gao, st = oracle.pass_ambient(vol, view, gb, 1)
    assert np.array_equal(_bits(gsh), _bits(wsh)) and np.array_equal(_bits(gao), _bits(wao))
    assert st["rays"] == int(rec["n"].sum()) and st["pixels"] == int(lit.sum())
    assert st["steps"] == int(rec["ray"]["fetches"][lit].sum())
    assert 0 < float(gsh.mean()) < 1 or scene == "house"
    # the logged rays themselves, re-traced by the oracle's ray-level entry
    for k, variant in ((0, oracle.SPARSE), (1, oracle.SUPERSPARSE)):
        rr = rec["ray"][..., k][lit]
        rays = np.zeros(len(rr), dtype=oracle.RAY_DTYPE)
        rays["ox"], rays["oy"], rays["oz"] = rr["o"][:, 0], rr["o"][:, 1], rr["o"][:, 2]
        rays["dx"], rays["dy"], rays["dz"] = rr["d"][:, 0], rr["d"][:, 1], rr["d"][:, 2]
        rays["dist"] = rr["dist"]
        hits = oracle.trace_rays(vol, rays, variant)
        assert np.array_equal(_bits(hits["t"]), _bits(rr["result"])) and np.array_equal(hits["steps"], rr["fetches"])

    # ---- point lights: one fullscreen draw per light, `discard` outside the range ----
    ext = 2 * vol.shape[2] * 0.1
    lights = sc.get("lights")
    if lights is None:
        lights = S.point_lights([(ext * 0.5, ext * 0.8, ext * 0.5), (ext * 0.2, ext * 0.5, ext * 0.3)], [ext * 0.6, ext * 0.35])
    gpt, st = oracle.pass_point(vol, view, gb, lights)
    total_rays = total_steps = 0
    for li in range(len(lights)):
        rec = oracle.shader_pass(oracle.PASS_POINT, vol, view, gb, lights=lights, light_index=li)
        assert np.array_equal(rec["n"], np.where(rec["discarded"] != 0, 0, 1))
        assert np.array_equal(_bits(gpt[li]), _bits(_planes_from_records(rec, "local"))), f"point light {li}"
        total_rays += int(rec["n"].sum()); total_steps += int(rec["ray"][..., 0]["fetches"].sum())
    assert st["rays"] == total_rays and st["steps"] == total_steps and total_rays > 0

    # ---- spot lights ----
    spots = S.spot_lights([tuple(l["Position"]) for l in lights[:2]], [float(l["Range"]) for l in lights[:2]], [(0.0, -1.0, 0.0)] * min(2, len(lights)))
    gsp, st = oracle.pass_spot(vol, view, gb, spots)
    total_rays = total_steps = 0
    for li in range(len(spots)):
        rec = oracle.shader_pass(oracle.PASS_SPOT, vol, view, gb, lights=spots, light_index=li)
        assert np.all(rec["ray"][..., 0]["variant"][rec["n"] > 0] == 1)
        assert np.array_equal(_bits(gsp[li]), _bits(_planes_from_records(rec, "local"))), f"spot light {li}"
        total_rays += int(rec["n"].sum()); total_steps += int(rec["ray"][..., 0]["fetches"].sum())
    assert st["rays"] == total_rays and st["steps"] == total_steps

    # ---- reflection: specular-occlusion ray length ----
    rec = oracle.shader_pass(oracle.PASS_REFLECTION, vol, view, gb)
    assert np.array_equal(rec["n"], np.where(lit, 1, 0))
    gt, st = oracle.pass_reflection(vol, view, gb)
    assert np.array_equal(_bits(gt), _bits(_planes_from_records(rec, "reflection")))
    assert st["rays"] == int(lit.sum()) and st["steps"] == int(rec["ray"][..., 0]["fetches"].sum())


def test_multiple_frames_noise_indexing(sh):
    """getNoise()'s frame-dependent texel offsets (LightAmbient.frag:49-52) over several frame indices."""
    oracle = sh
    for frame in (0, 7, 15, 16, 63, 1599):
        sc = U.house_scene(oracle, width=48, height=32, frame=frame)
        rec = oracle.shader_pass(oracle.PASS_AMBIENT, sc["volume"], sc["view"], sc["gb"])
        wsh, wao = _planes_from_records(rec, "ambient")
        gsh, gao, _ = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], 1)
        assert np.array_equal(_bits(gsh), _bits(wsh)) and np.array_equal(_bits(gao), _bits(wao)), frame
