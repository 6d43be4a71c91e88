"""The steps after the light passes: LightTAA.frag and the colour of LightReflection.frag.

CPU: the oracle's restatements against THE REFERENCE'S OWN SHADERS' out_Color (oracle/_ref/libvxshader.so, built when the
reference tree is mounted) and against the committed reference-generated fixture tests/golden/ref_f3.npz.
GPU: vxl_light_taa / vxl_resolve_reflection through the C ABI against the oracle and the same fixture.
LightTAA has no transcendental left once cos / sin are pinned as correctly rounded values: bit equality (NaN == NaN).
The reflection colour carries pow(): 1e-5 relative + 1e-6 absolute, as for row f2."""
import os

import numpy as np
import pytest

import scene_util as U

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL, ATOL = 1e-5, 1e-6


def _bits_equal(got, want, what):
    got, want = np.ascontiguousarray(got, np.float32), np.ascontiguousarray(want, np.float32)
    assert got.shape == want.shape, what
    same = (got.view(np.uint32) == want.view(np.uint32)) | (np.isnan(got) & np.isnan(want))
    assert same.all(), f"{what}: {(~same).sum()} of {same.size} values differ, worst {np.nanmax(np.abs(got - want)[~same])}"


def _close(got, want, what):
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    bad = ~np.isclose(got, want, rtol=RTOL, atol=ATOL, equal_nan=True)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} values differ, worst {np.abs(got - want)[bad].max()}"


def _case(oracle, scene):
    sc = U.house_scene(oracle, width=64, height=48) if scene == "house" else U.terrain_scene(oracle, width=96, height=54)
    return sc, U.f3_case(oracle, sc)


@pytest.mark.parametrize("scene", ["house", "terrain"])
def test_oracle_taa_matches_reference_shader(oracle, scene):
    if oracle.shader_lib() is None:
        pytest.skip("oracle/_ref/libvxshader.so not built (reference tree not mounted)")
    sc, c = _case(oracle, scene)
    want = oracle.shader_taa(sc["view"], c["gb"], c["albedo"], c["motion"], c["light"], c["last_light"])
    got = oracle.light_taa(sc["view"], c["gb"], c["albedo"], c["motion"], c["light"], c["last_light"])
    _bits_equal(got, want, "LightTAA out_Color")
    h, w = c["gb"]["depth24"].shape
    sky = c["gb"]["depth24"] == 0xFFFFFF
    assert sky.any() and np.array_equal(got[sky], c["light"][sky])                      # :49-52 pass-through
    oob = np.zeros((h, w), bool); oob[: h // 6] = True
    assert np.all(got[oob & ~sky][:, 3] == 1.0)                                          # :83 current-frame-only branch
    acc = ~oob & ~sky
    assert 0.05 < float((got[acc][:, 3] > 0).mean()) and len(np.unique(got[acc][:, 3])) > 50   # variance channel alive
    assert float(np.abs(got[acc][:, :3] - c["light"][acc][:, :3]).mean()) > 1e-3         # and the history is blended in


@pytest.mark.parametrize("scene", ["house", "terrain"])
def test_oracle_reflection_colour_matches_reference_shader(oracle, scene):
    if oracle.shader_lib() is None:
        pytest.skip("oracle/_ref/libvxshader.so not built (reference tree not mounted)")
    sc, c = _case(oracle, scene)
    taa = oracle.light_taa(sc["view"], c["gb"], c["albedo"], c["motion"], c["light"], c["last_light"])
    taa = np.nan_to_num(taa, nan=0.0, posinf=0.0, neginf=0.0)
    rec = oracle.shader_pass(oracle.PASS_REFLECTION, sc["volume"], sc["view"], c["gb"], light=taa, sky=c["sky"])
    got = oracle.resolve_reflection(sc["view"], c["gb"], c["t"], light=taa, sky=c["sky"])
    _close(got, rec["color"], "LightReflection out_Color")
    lit = c["gb"]["depth24"] != 0xFFFFFF
    miss = lit & (c["t"] == 256.0)
    hit_lit = lit & (c["t"] != 256.0) & (got[..., :3] > 0).any(axis=-1)
    assert miss.sum() > 20 and hit_lit.sum() > 20, (int(miss.sum()), int(hit_lit.sum()))   # both the sky and the light-buffer branch
    assert np.all(got[~lit][:, :3] == 0)


def test_oracle_f3_matches_reference_golden(oracle):
    g = np.load(os.path.join(HERE, "golden", "ref_f3.npz"))
    sc, c = _case(oracle, "house")
    taa = oracle.light_taa(sc["view"], c["gb"], c["albedo"], c["motion"], c["light"], c["last_light"])
    _bits_equal(taa, g["taa"], "LightTAA vs the reference's out_Color")
    got = oracle.resolve_reflection(sc["view"], c["gb"], c["t"], light=np.nan_to_num(taa, nan=0.0, posinf=0.0, neginf=0.0), sky=c["sky"])
    _close(got, g["reflection"], "LightReflection vs the reference's out_Color")


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["house", "terrain"])
def test_cuda_f3_matches_oracle_and_reference_golden(gpu_ctx, oracle, scene):
    import torch
    from voxelengine_b200 import engine as E
    sc, c = _case(oracle, scene)
    gb = c["gb"]
    h, w = gb["depth24"].shape
    dev = gpu_ctx.torch_device
    want_taa = oracle.light_taa(sc["view"], gb, c["albedo"], c["motion"], c["light"], c["last_light"])
    taa_in = np.nan_to_num(want_taa, nan=0.0, posinf=0.0, neginf=0.0)
    want_refl = oracle.resolve_reflection(sc["view"], gb, c["t"], light=taa_in, sky=c["sky"])
    full = E.FullFrame(gpu_ctx, gb["depth24"], gb["normal"], gb["material"], c["albedo"], c["motion"])
    d_light, d_last = (torch.from_numpy(a).to(dev) for a in (c["light"], c["last_light"]))
    for tile, rank, world in ((None, 0, 1), ((32, 16), 1, 2)):
        fb = E.GeometryBuffer(gpu_ctx, w, h) if tile is None else E.GeometryBuffer(gpu_ctx, w, h, tile[0], tile[1], rank=rank, world=world)
        fb.set_noise(gb["noise"])
        fb.set_planes(gb["depth24"], gb["normal"], gb["material"])
        got = E.LightTAAPipeline.Get().Use(sc["view"], fb, full, d_light, d_last).cpu().numpy()
        d_t = torch.from_numpy(fb.to_tiles(c["t"].view(np.uint32)).view(np.float32)).to(dev)
        refl = E.LightReflectionPipeline.Get().Colour(sc["view"], fb, d_t, full, torch.from_numpy(taa_in).to(dev), c["sky"]).cpu().numpy()
        for ch in range(4):
            _bits_equal(got[..., ch], fb.to_tiles(np.ascontiguousarray(want_taa[..., ch]).view(np.uint32)).view(np.float32), f"TAA channel {ch} {tile}")
            _close(refl[..., ch], fb.to_tiles(np.ascontiguousarray(want_refl[..., ch]).view(np.uint32)).view(np.float32), f"reflection channel {ch} {tile}")
    if scene == "house":
        g = np.load(os.path.join(HERE, "golden", "ref_f3.npz"))
        fb = E.GeometryBuffer(gpu_ctx, w, h)
        fb.set_noise(gb["noise"])
        fb.set_planes(gb["depth24"], gb["normal"], gb["material"])
        got = E.LightTAAPipeline.Get().Use(sc["view"], fb, full, d_light, d_last).cpu().numpy()[0]
        _bits_equal(got, g["taa"], "TAA vs the reference's out_Color")
