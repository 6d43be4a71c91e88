"""Light-buffer resolve: the colour LightAmbient / LightPoint / LightSpot.frag compute after the march.

CPU: the oracle's restatement against THE REFERENCE'S OWN SHADERS' out_Color (oracle/_ref/libvxshader.so, built when the
reference tree is mounted) and against the committed reference-generated fixture tests/golden/ref_resolve.npz.
GPU: vxl_resolve_* through the C ABI against the oracle and against the same fixture (no /root/reference needed).
pow() is specified by accuracy only, so the bar is a tolerance, written here: 1e-5 relative + 1e-6 absolute."""
import os

import numpy as np
import pytest

import scene_util as U

HERE = os.path.dirname(os.path.abspath(__file__))
RTOL, ATOL = 1e-5, 1e-6


def _close(got, want, what):
    got, want = np.asarray(got, np.float32), np.asarray(want, np.float32)
    assert got.shape == want.shape, what
    bad = ~np.isclose(got, want, rtol=RTOL, atol=ATOL, equal_nan=True)
    assert not bad.any(), f"{what}: {bad.sum()} of {bad.size} values differ, worst {np.abs(got - want)[bad].max()}"


def _oracle_planes(oracle, sc):
    gb, albedo, point, spot = U.resolve_case(sc)
    vol, view = sc["volume"], sc["view"]
    sh, ao, _ = oracle.pass_ambient(vol, view, gb, 1)
    pt, _ = oracle.pass_point(vol, view, gb, point)
    sp, _ = oracle.pass_spot(vol, view, gb, spot)
    return gb, albedo, point, spot, sh, ao, pt, sp


@pytest.mark.parametrize("scene", ["house", "terrain"])
def test_oracle_resolve_matches_reference_shaders(oracle, scene):
    if oracle.shader_lib() is None:
        pytest.skip("oracle/_ref/libvxshader.so not built (reference tree not mounted)")
    sc = U.house_scene(oracle, width=64, height=48) if scene == "house" else U.terrain_scene(oracle, width=96, height=54)
    gb, albedo, point, spot, sh, ao, pt, sp = _oracle_planes(oracle, sc)
    vol, view = sc["volume"], sc["view"]
    lit = (gb["depth24"] & 0xFFFFFF).astype(np.float32) / np.float32(16777215.0) < np.float32(0.999)
    rec = oracle.shader_pass(oracle.PASS_AMBIENT, vol, view, gb, albedo=albedo)
    got = oracle.resolve_ambient(view, gb, albedo, sh, ao)
    _close(got[lit], rec["color"][lit], "ambient colour")
    assert np.all(got[~lit] == 0) and float(np.abs(got[lit][:, :3]).mean()) > 0.05
    for which, lights, planes, spot_flag in ((oracle.PASS_POINT, point, pt, False), (oracle.PASS_SPOT, spot, sp, True)):
        total = np.zeros(gb["depth24"].shape + (4,), np.float32)
        for li in range(len(lights)):
            rec = oracle.shader_pass(which, vol, view, gb, lights=lights, light_index=li, albedo=albedo)
            want = np.where((rec["discarded"] != 0)[..., None], np.float32(0), rec["color"])
            got = oracle.resolve_local(view, gb, albedo, lights[li:li + 1], planes[li:li + 1], spot=spot_flag)
            _close(got, want, f"light {li} (spot={spot_flag})")
            assert int((want[..., :3] > 0).any(axis=-1).sum()) > 100, "the test light must reach the scene"
            total = total + want                                      # additive blend, list order, float32
        _close(oracle.resolve_local(view, gb, albedo, lights, planes, spot=spot_flag), total, "sum over the lights")


def test_oracle_resolve_matches_reference_golden(oracle):
    g = np.load(os.path.join(HERE, "golden", "ref_resolve.npz"))
    sc = U.house_scene(oracle, width=int(g["width"]), height=int(g["height"]))
    gb, albedo, point, spot, sh, ao, pt, sp = _oracle_planes(oracle, sc)
    _close(oracle.resolve_ambient(sc["view"], gb, albedo, sh, ao), g["ambient"], "ambient colour")
    _close(oracle.resolve_local(sc["view"], gb, albedo, point, pt), g["point_sum"], "point lights")
    _close(oracle.resolve_local(sc["view"], gb, albedo, spot, sp, spot=True), g["spot_sum"], "spot lights")


@pytest.mark.gpu
@pytest.mark.parametrize("scene", ["house", "terrain"])
def test_cuda_resolve_matches_oracle_and_reference_golden(gpu_ctx, oracle, scene):
    import torch
    from voxelengine_b200 import engine as E
    g = np.load(os.path.join(HERE, "golden", "ref_resolve.npz"))
    sc = U.house_scene(oracle, width=int(g["width"]), height=int(g["height"])) if scene == "house" else U.terrain_scene(oracle)
    gb, albedo, point, spot, sh, ao, pt, sp = _oracle_planes(oracle, sc)
    h, w = gb["depth24"].shape
    vol = E.ShadowVoxSystem(gpu_ctx, sc["volume"].shape[::-1])
    vol.upload(sc["volume"])
    fb = E.GeometryBuffer(gpu_ctx, w, h)
    fb.set_noise(gb["noise"])
    fb.set_planes(gb["depth24"], gb["normal"], gb["material"])
    dev = gpu_ctx.torch_device
    d_alb = torch.from_numpy(albedo.view(np.int32)[None]).to(dev)
    # the march planes come from the CUDA passes themselves (bit-exact against the oracle's, checked here again)
    d_sh, d_ao = E.LightAmbientPipeline.Get().Use(sc["view"], fb, vol, n_ao=1)
    assert np.array_equal(d_sh.cpu().numpy()[0], sh) and np.array_equal(d_ao.cpu().numpy()[0], ao)
    d_pt = E.LightPointPipeline.Get().Use(sc["view"], fb, vol, lambda p: [p.DrawLight(l["Position"], l["Range"], l["Color"], l["Attenuation"]) for l in point])
    d_sp = E.LightSpotPipeline.Get().Use(sc["view"], fb, vol, lambda p: [p.DrawLight(l["Position"], l["Range"], l["Color"], l["Attenuation"], l["Direction"], l["Angle"], l["AngleAttenuation"]) for l in spot])
    lb = E.LightBuffer(fb, d_alb)
    amb = lb.Ambient(sc["view"], d_sh, d_ao).cpu().numpy()[0].copy()
    want_amb = oracle.resolve_ambient(sc["view"], gb, albedo, sh, ao)
    _close(amb, want_amb, "ambient colour vs oracle")
    lb.rgba.zero_()
    got_pt = lb.Point(sc["view"], point, d_pt).cpu().numpy()[0].copy()
    _close(got_pt, oracle.resolve_local(sc["view"], gb, albedo, point, pt), "point lights vs oracle")
    got_all = lb.Spot(sc["view"], spot, d_sp).cpu().numpy()[0].copy()             # accumulates on top of the point lights
    _close(got_all, oracle.resolve_local(sc["view"], gb, albedo, spot, sp, spot=True, accumulate=got_pt), "spot lights added to the buffer")
    if scene == "house":
        _close(amb, g["ambient"], "ambient colour vs the reference's out_Color")
        _close(got_pt, g["point_sum"], "point lights vs the reference's out_Color")
    # a tile-sharded frame resolves to the same colours, given the whole frame's depth plane
    fb2 = E.GeometryBuffer(gpu_ctx, w, h, 32, 16, rank=1, world=2)
    fb2.set_noise(gb["noise"])
    fb2.set_planes(gb["depth24"], gb["normal"], gb["material"])
    t_alb = torch.from_numpy(fb2.to_tiles(albedo).view(np.int32)).to(dev)
    t_sh = torch.from_numpy(fb2.to_tiles(sh.view(np.uint32)).view(np.float32)).to(dev)
    t_ao = torch.from_numpy(fb2.to_tiles(ao.view(np.uint32)).view(np.float32)).to(dev)
    d_full = torch.from_numpy(gb["depth24"].view(np.int32)).to(dev)
    lb2 = E.LightBuffer(fb2, t_alb, depth_full=d_full)
    tiles = lb2.Ambient(sc["view"], t_sh, t_ao).cpu().numpy()
    for c in range(3):
        full = np.full((h, w), np.float32(np.nan))
        fb2.from_tiles(np.ascontiguousarray(tiles[..., c]).view(np.uint32), full.view(np.uint32))
        mine = ~np.isnan(full)
        assert mine.any() and np.array_equal(full[mine], amb[..., c][mine])
    with pytest.raises(Exception):
        E.LightBuffer(fb2, t_alb).Ambient(sc["view"], t_sh, t_ao)            # sharded frame without depth_full
    vol.close()
