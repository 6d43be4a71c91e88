"""GPU parity: the CUDA path, called through the C ABI (include/vxl.h), against the CPU oracle on the
same seeded inputs.  Bar: bit-exact for everything integer/byte/index (volumes, regions, voxel
coordinates, step counts, ray counts) AND, because both sides execute the same IEEE single-precision
operation sequence without FMA contraction, bit-exact for the float outputs as well; the tolerance
BASELINE.json's north_star states for the lighting values (1e-3 absolute) is asserted separately so a
report distinguishes "inside tolerance" from "bit-exact"."""
import numpy as np
import pytest

import scene_util as U
from voxelengine_b200 import scenes as S

pytestmark = pytest.mark.gpu

TOL = 1e-3   # north_star: shadow/AO/spec-occlusion values within 1e-3 absolute


def _eng():
    from voxelengine_b200 import engine
    return engine


def _upload_scene(ctx, sc, tile=None, rank=0, world=1):
    E = _eng()
    sz, sy, sx = sc["volume"].shape
    vol = E.ShadowVoxSystem(ctx, (sx, sy, sz))
    vol.upload(sc["volume"])
    h, w = sc["gb"]["depth24"].shape
    gb = E.GeometryBuffer(ctx, w, h, *(tile or (None, None)), rank=rank, world=world)
    gb.set_noise(sc["gb"]["noise"])
    gb.set_planes(sc["gb"]["depth24"], sc["gb"]["normal"], sc["gb"]["material"])
    return vol, gb


def _full(gb, t):
    out = np.zeros((gb.height, gb.width), np.float32)
    return gb.from_tiles(t.cpu().numpy(), out)


def _assert_plane(name, got, want):
    diff = np.abs(got.astype(np.float64) - want.astype(np.float64))
    assert diff.max() <= TOL, f"{name}: max abs err {diff.max()} > {TOL}"
    nbad = int((got.view(np.uint32) != want.view(np.uint32)).sum())
    assert nbad == 0, f"{name}: {nbad} of {got.size} values are inside tolerance but not bit-identical"


@pytest.fixture(scope="module")
def house(oracle):
    return U.house_scene(oracle)


@pytest.fixture(scope="module")
def terrain(oracle):
    return U.terrain_scene(oracle)


# ---- level 1: traversal given rays -----------------------------------------------------------------
@pytest.mark.parametrize("variant", [0, 1, 2])
def test_trace_rays_bit_exact(gpu_ctx, oracle, terrain, variant):
    E = _eng()
    sz, sy, sx = terrain["volume"].shape
    vol = E.ShadowVoxSystem(gpu_ctx, (sx, sy, sz))
    vol.upload(terrain["volume"])
    rays = U.random_rays(np.random.RandomState(10 + variant), 200_000, (2 * sx, 2 * sy, 2 * sz))
    want = oracle.trace_rays(terrain["volume"], rays, variant)
    got = vol.trace_rays(rays, variant)
    for f in ("steps", "vx", "vy", "vz", "status"):
        assert np.array_equal(got[f], want[f]), f"{f}: {(got[f] != want[f]).sum()} rays differ"
    for f in ("t", "px", "py", "pz", "nx", "ny", "nz"):
        a, b = got[f].view(np.uint32), want[f].view(np.uint32)
        nan_both = np.isnan(got[f]) & np.isnan(want[f])
        assert np.array_equal(a[~nan_both], b[~nan_both]), f"{f}: {(a != b).sum()} rays differ bitwise"
    assert (want["status"] != 0).mean() > 0.1   # the fixture really exercises hits
    vol.close()


def test_trace_rays_empty_and_full_volume(gpu_ctx, oracle):
    """Reference invariants: empty volume => exactly `dist`, 31 + (min(dist,164)-16) probes;
    all-ones volume => first probe (Sparse 0.5, SuperSparse 2.5, 1 step)."""
    E = _eng()
    vol = E.ShadowVoxSystem(gpu_ctx, (16, 16, 16))
    rays = np.zeros(4, dtype=S.RAY_DTYPE)
    rays["ox"], rays["oy"], rays["oz"] = 5.3, 6.1, 7.7
    rays["dx"], rays["dy"], rays["dz"] = 0.6, 0.48, 0.64
    rays["dist"] = [128.0, 256.0, 40.0, 10.0]
    h = vol.trace_rays(rays, 0)
    assert h["t"].tolist() == [128.0, 256.0, 40.0, 10.0]
    assert h["steps"].tolist() == [31 + 112, 31 + 148, 31 + 24, 31]
    h = vol.trace_rays(rays, 1)
    assert h["steps"].tolist() == [6 + 23, 6 + 30, 6 + 5, 6]
    vol.upload(np.full((16, 16, 16), 255, np.uint8))
    h = vol.trace_rays(rays, 0)
    assert h["t"].tolist() == [0.5] * 4 and h["steps"].tolist() == [1] * 4 and h["status"].tolist() == [1] * 4
    h = vol.trace_rays(rays, 1)
    assert h["t"].tolist() == [2.5] * 4 and h["steps"].tolist() == [1] * 4
    h = vol.trace_rays(rays[:0], 0)
    assert len(h) == 0
    vol.close()


# ---- level 2: the passes -----------------------------------------------------------------------------
@pytest.mark.parametrize("scene,n_ao", [("house", 1), ("house", 4), ("terrain", 8)])
def test_pass_ambient(gpu_ctx, oracle, house, terrain, scene, n_ao):
    E = _eng()
    sc = house if scene == "house" else terrain
    vol, gb = _upload_scene(gpu_ctx, sc)
    gpu_ctx.stats_reset()
    sh, ao = E.LightAmbientPipeline.Get().Use(sc["view"], gb, vol, n_ao=n_ao)
    st = gpu_ctx.stats()
    wsh, wao, wst = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], n_ao)
    _assert_plane("shadow", _full(gb, sh), wsh)
    _assert_plane("ao", _full(gb, ao), wao)
    assert st == wst
    assert 0.02 < (wsh == 0).mean() < 0.9
    vol.close()


def _test_lights(sc, spot=False):
    vol = sc["volume"]
    sz, sy, sx = vol.shape
    # lights spread over the visible part of the scene with ranges that cull some pixels
    pos = [(sx * 0.2 * f, sy * 0.2 * 0.8, sz * 0.2 * g) for f, g in ((0.3, 0.3), (0.6, 0.4), (0.4, 0.7), (0.7, 0.7), (0.5, 0.5))]
    ranges = [sx * 0.2 * r for r in (0.35, 0.5, 0.3, 0.6, 2.0)]
    if spot:
        return S.spot_lights(pos, ranges, [(0, -1, 0)] * len(pos))
    return S.point_lights(pos, ranges)


@pytest.mark.parametrize("scene", ["house", "terrain"])
def test_pass_point(gpu_ctx, oracle, house, terrain, scene):
    E = _eng()
    sc = house if scene == "house" else terrain
    lights = _test_lights(sc)
    vol, gb = _upload_scene(gpu_ctx, sc)
    gpu_ctx.stats_reset()
    P = E.LightPointPipeline.Get()
    out = P.Use(sc["view"], gb, vol, lambda p: [p.DrawLight(l["Position"], l["Range"], l["Color"], l["Attenuation"]) for l in lights])
    st = gpu_ctx.stats()
    want, wst = oracle.pass_point(sc["volume"], sc["view"], sc["gb"], lights)
    for i in range(len(lights)):
        _assert_plane(f"point[{i}]", _full(gb, out[i]), want[i])
    assert st == wst and wst["rays"] > 0
    assert (want == 0).any()
    vol.close()


def test_pass_spot(gpu_ctx, oracle, terrain):
    E = _eng()
    sc = terrain
    lights = _test_lights(sc, spot=True)
    vol, gb = _upload_scene(gpu_ctx, sc)
    gpu_ctx.stats_reset()
    P = E.LightSpotPipeline.Get()
    out = P.Use(sc["view"], gb, vol, lambda p: [p.DrawLight(l["Position"], l["Range"], l["Color"], l["Attenuation"], l["Direction"], l["Angle"], l["AngleAttenuation"]) for l in lights])
    st = gpu_ctx.stats()
    want, wst = oracle.pass_spot(sc["volume"], sc["view"], sc["gb"], lights)
    for i in range(len(lights)):
        _assert_plane(f"spot[{i}]", _full(gb, out[i]), want[i])
    assert st == wst and wst["rays"] > 0
    vol.close()


@pytest.mark.parametrize("scene", ["house", "terrain"])
def test_pass_reflection(gpu_ctx, oracle, house, terrain, scene):
    E = _eng()
    sc = house if scene == "house" else terrain
    vol, gb = _upload_scene(gpu_ctx, sc)
    gpu_ctx.stats_reset()
    t = E.LightReflectionPipeline.Get().Use(sc["view"], gb, vol)
    st = gpu_ctx.stats()
    want, wst = oracle.pass_reflection(sc["volume"], sc["view"], sc["gb"])
    _assert_plane("spec_t", _full(gb, t), want)
    assert st == wst
    assert 0.005 < (want < 256).mean() < 0.999
    vol.close()


def test_more_than_64_lights_is_dropped_like_the_reference(gpu_ctx, house):
    E = _eng()
    vol, gb = _upload_scene(gpu_ctx, house)
    P = E.LightPointPipeline()
    out = P.Use(house["view"], gb, vol, lambda p: [p.DrawLight((1, 1, 1), 5.0, (1, 1, 1), 2.0) for _ in range(70)])
    assert out.shape[0] == 64 and P.warnings == 6
    vol.close()


# ---- tile-sharded frames (multi-GPU partitioning, exercised on one GPU) ---------------------------------
@pytest.mark.parametrize("world", [2, 3])
def test_tile_sharding_equals_whole_frame(gpu_ctx, oracle, terrain, world):
    E = _eng()
    sc = terrain
    wsh, wao, wst = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], 2)
    wt, _ = oracle.pass_reflection(sc["volume"], sc["view"], sc["gb"])
    h, w = wsh.shape
    sh_full, ao_full, t_full = (np.zeros((h, w), np.float32) for _ in range(3))
    rays = 0
    for rank in range(world):
        vol, gb = _upload_scene(gpu_ctx, sc, tile=(64, 32), rank=rank, world=world)   # 160x90 -> ragged edge tiles
        gpu_ctx.stats_reset()
        sh, ao = E.LightAmbientPipeline.Get().Use(sc["view"], gb, vol, n_ao=2)
        rays += gpu_ctx.stats()["rays"]
        t = E.LightReflectionPipeline.Get().Use(sc["view"], gb, vol)
        gb.from_tiles(sh.cpu().numpy(), sh_full)
        gb.from_tiles(ao.cpu().numpy(), ao_full)
        gb.from_tiles(t.cpu().numpy(), t_full)
        vol.close()
    _assert_plane("shadow", sh_full, wsh)
    _assert_plane("ao", ao_full, wao)
    _assert_plane("spec_t", t_full, wt)
    assert rays == wst["rays"]


# ---- volume build ---------------------------------------------------------------------------------------
def test_voxelize_matches_sequential_reference(gpu_ctx, oracle):
    """A7 semantics incl. first-frame identity clear, overlap between entities (last writer wins),
    glass voxels, out-of-volume parts, OnVoxDestroyed without pivot."""
    E = _eng()
    rs = np.random.RandomState(7)
    dims = (40, 24, 36)   # texels
    sx, sy, sz = dims
    base = rs.randint(0, 256, size=(sz, sy, sx)).astype(np.uint8)
    models = [S.house_model(20, seed=3), S.shell_cube_model(8), rs.randint(0, 40, size=(5, 7, 9)).astype(np.uint8)]
    n = 60
    e = S.entities(n)
    for i in range(n):
        e[i]["model"] = rs.randint(0, 3)
        pos = rs.uniform(-1.0, [sx * 0.2 + 0.5, sy * 0.2 + 0.5, sz * 0.2 + 0.5])
        rot = rs.uniform(-3.2, 3.2, 3) * (rs.rand() < 0.7)
        scale = rs.choice([1.0, 1.0, 0.5, 2.0])
        e[i]["cur"] = S.transform_matrix(pos, rot, (scale,) * 3)
        if i % 3 == 1:   # moved entity: previous transform nearby (clear-then-set overlap)
            e[i]["prev"] = S.transform_matrix(pos + rs.uniform(-0.3, 0.3, 3), rot, (scale,) * 3)
        e[i]["pivot"] = rs.uniform(0, 1.0, 3)
        if i % 11 == 5:
            e[i]["flags"] = S.ENT_DESTROY
    want = base.copy()
    wreg, wvalid = oracle.voxelize(want, models, e)
    vol = E.ShadowVoxSystem(gpu_ctx, dims)
    vol.upload(base)
    ids = [vol.add_model(m) for m in models]
    ge = e.copy()                                          # model ids are per context (other tests registered theirs first)
    ge["model"] = np.asarray(ids, np.int32)[e["model"]]
    reg, valid = vol.OnUpdate(ge)
    got = vol.download()
    assert np.array_equal(got, want), f"{(got != want).sum()} bytes differ"
    assert np.array_equal(valid, wvalid)
    assert np.array_equal(reg[valid == 1], wreg[wvalid == 1])
    assert (want != base).sum() > 1000
    # set then clear at the same matrix => back to the base volume minus the stamped voxels
    vol.close()


def test_voxelize_set_then_destroy_is_empty(gpu_ctx, oracle):
    E = _eng()
    vol = E.ShadowVoxSystem(gpu_ctx, (32, 32, 32))
    mid = vol.add_model(S.shell_cube_model(16))
    e = S.entities(1)
    e[0]["model"] = mid
    e[0]["prev"] = e[0]["cur"] = S.transform_matrix((1.3, 1.7, 2.1), (0.2, 0.4, 0.1))
    vol.OnUpdate(e)
    assert vol.download().any()
    d = e.copy()
    d[0]["flags"] = S.ENT_DESTROY
    vol.OnUpdate(d)
    assert not vol.download().any()
    vol.close()


def test_upload_regions_addressing(gpu_ctx, oracle):
    E = _eng()
    rs = np.random.RandomState(3)
    dims = (37, 21, 29)
    sx, sy, sz = dims
    staging = rs.randint(0, 256, size=(sz, sy, sx)).astype(np.uint8)
    regions = np.zeros(5, dtype=S.REGION_DTYPE)
    for i, (x, y, z, w, h, d) in enumerate([(0, 0, 0, 5, 4, 3), (30, 15, 20, 7, 6, 9), (10, 3, 8, 1, 1, 1), (35, 0, 0, 10, 30, 40), (3, 4, 5, 20, 10, 12)]):
        regions[i] = (x, y, z, w, h, d, 0)
    want = np.zeros_like(staging)
    oracle.upload_regions(want, staging, regions)
    vol = E.ShadowVoxSystem(gpu_ctx, dims)
    vol.upload_regions(staging, regions)
    assert np.array_equal(vol.download(), want)
    vol.close()


def test_terrain_generator_matches_oracle(gpu_ctx, oracle):
    E = _eng()
    vol = E.ShadowVoxSystem(gpu_ctx, (48, 40, 56))
    vol.gen_terrain()
    got = vol.download()
    want = oracle.gen_terrain(48, 40, 56)
    assert np.array_equal(got, want), f"{(got != want).sum()} bytes differ"
    vol.close()


def test_gbuffer_generator_matches_oracle(gpu_ctx, oracle, terrain):
    E = _eng()
    sc = terrain
    vol, gb = _upload_scene(gpu_ctx, sc)
    gb.depth24.zero_(); gb.normal.zero_(); gb.material.zero_()
    gb.synthesize(vol, sc["view"])
    gpu_ctx.sync()
    for k in ("depth24", "normal", "material"):
        got = getattr(gb, k).cpu().numpy().view(np.uint32)[0]
        assert np.array_equal(got, sc["gb"][k]), f"{k}: {(got != sc['gb'][k]).sum()} pixels differ"
    vol.close()


# ---- whole-frame host drop-in ---------------------------------------------------------------------------
def test_lighting_host_matches_passes(gpu_ctx, oracle, terrain):
    import torch
    E = _eng()
    sc = terrain
    lights = _test_lights(sc)
    sz, sy, sx = sc["volume"].shape
    vol = E.ShadowVoxSystem(gpu_ctx, (sx, sy, sz))
    vol.upload(sc["volume"])
    h, w = sc["gb"]["depth24"].shape
    planes = {k: torch.from_numpy(sc["gb"][k].view(np.int32)).pin_memory() for k in ("depth24", "normal", "material", "noise")}
    outs = dict(shadow=torch.empty((h, w), dtype=torch.float32).pin_memory(), ao=torch.empty((h, w), dtype=torch.float32).pin_memory(),
                point_shadow=torch.empty((len(lights), h, w), dtype=torch.float32).pin_memory(),
                spec_t=torch.empty((h, w), dtype=torch.float32).pin_memory())
    desc = dict(width=w, height=h, tile_w=w, tile_h=h, tile_first=0, tile_stride=1, n_tiles=1)
    E.lighting_host(gpu_ctx, vol, sc["view"], desc, planes, outs, n_ao=3, point=lights)
    wsh, wao, _ = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], 3)
    wpt, _ = oracle.pass_point(sc["volume"], sc["view"], sc["gb"], lights)
    wt, _ = oracle.pass_reflection(sc["volume"], sc["view"], sc["gb"])
    _assert_plane("shadow", outs["shadow"].numpy(), wsh)
    _assert_plane("ao", outs["ao"].numpy(), wao)
    _assert_plane("point", outs["point_shadow"].numpy(), wpt)
    _assert_plane("spec_t", outs["spec_t"].numpy(), wt)
    vol.close()


@pytest.mark.parametrize("rank,world", [(0, 2), (1, 3)])
def test_lighting_host_on_a_tile_shard(gpu_ctx, oracle, terrain, rank, world):
    """vxl_lighting_host / _packed on one rank's shard of a sharded frame (16x16 tiles, round-robin): many small tiles, so the call
    walks the shard in bands of TILES (uploads, the three pass kernels and read-backs of consecutive bands overlap).  The planes,
    scattered back into the frame, equal the oracle's on this rank's tiles -- float and packed output alike."""
    import torch
    E = _eng()
    from voxelengine_b200.tiles import TileLayout
    sc = terrain
    lights = _test_lights(sc)
    sz, sy, sx = sc["volume"].shape
    vol = E.ShadowVoxSystem(gpu_ctx, (sx, sy, sz))
    vol.upload(sc["volume"])
    h, w = sc["gb"]["depth24"].shape
    L = TileLayout(w, h, 16, 16, rank=rank, world=world)
    n = L.n_tiles
    assert n >= 8
    planes = {k: torch.from_numpy(np.ascontiguousarray(L.to_tiles(sc["gb"][k])).view(np.int32)).pin_memory() for k in ("depth24", "normal", "material")}
    planes["noise"] = torch.from_numpy(sc["gb"]["noise"].view(np.int32)).pin_memory()
    desc = dict(width=w, height=h, tile_w=16, tile_h=16, tile_first=rank, tile_stride=world, n_tiles=n)
    f32 = lambda *shape: torch.full(shape, -7.0, dtype=torch.float32).pin_memory()
    outs = dict(shadow=f32(n, 16, 16), ao=f32(n, 16, 16), point_shadow=f32(len(lights), n, 16, 16), spec_t=f32(n, 16, 16))
    E.lighting_host(gpu_ctx, vol, sc["view"], desc, planes, outs, n_ao=3, point=lights)
    wsh, wao, _ = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], 3)
    wpt, _ = oracle.pass_point(sc["volume"], sc["view"], sc["gb"], lights)
    wt, _ = oracle.pass_reflection(sc["volume"], sc["view"], sc["gb"])
    bits = lambda a: np.ascontiguousarray(a).reshape(-1).view(np.uint32)
    for name, got, want in [("shadow", outs["shadow"].numpy(), wsh), ("ao", outs["ao"].numpy(), wao), ("spec_t", outs["spec_t"].numpy(), wt)] + [
            (f"point{i}", outs["point_shadow"].numpy()[i], wpt[i]) for i in range(len(lights))]:
        ref = L.to_tiles(want)
        # pixels of edge tiles beyond the frame are not written
        mask = L.to_tiles(np.ones((h, w), np.uint8)).astype(bool)
        assert np.array_equal(bits(got[mask]), bits(ref[mask])), name
    mb = E.mask_bytes(len(lights), 0)
    pk = dict(shadow_mask=torch.zeros((n, 16, 16, mb), dtype=torch.uint8).pin_memory(), spec_code=torch.zeros((n, 16, 16), dtype=torch.uint8).pin_memory(), ao=f32(n, 16, 16))
    E.lighting_host(gpu_ctx, vol, sc["view"], desc, planes, {"packed": pk}, n_ao=3, point=lights)
    un = E.unpack_planes(pk["shadow_mask"].numpy(), pk["spec_code"].numpy(), len(lights), 0)
    mask = L.to_tiles(np.ones((h, w), np.uint8)).astype(bool)
    assert np.array_equal(bits(pk["ao"].numpy()[mask]), bits(outs["ao"].numpy()[mask]))
    assert np.array_equal(bits(un["shadow"].reshape(n, 16, 16)[mask]), bits(outs["shadow"].numpy()[mask]))
    assert np.array_equal(bits(un["spec_t"].reshape(n, 16, 16)[mask]), bits(outs["spec_t"].numpy()[mask]))
    for i in range(len(lights)):
        assert np.array_equal(bits(un["point_shadow"][i].reshape(n, 16, 16)[mask]), bits(outs["point_shadow"].numpy()[i][mask]))
    vol.close()


def test_lighting_host_packed_decodes_to_the_float_planes(gpu_ctx, oracle, terrain):
    """vxl_lighting_host_packed: the shadow planes as a bit mask per pixel, spec_t as its one-byte code (LightAmbient.frag:167-169,
    LightPoint.frag:125, LightReflection.frag:113 take 2 / 2 / 180 values), AO float32.  Decoded, every plane equals what
    vxl_lighting_host returns -- and therefore the oracle -- bit for bit; 9 point + 2 spot lights need a two-byte mask."""
    import torch
    E = _eng()
    from voxelengine_b200 import scenes as S
    sc = terrain
    base = _test_lights(sc)
    pts = np.concatenate([base] * 5)[:9].copy()
    for i in range(len(pts)):
        pts[i]["Position"][0] += 0.37 * i
        pts[i]["Range"] = 4.0 + i
    spots = S.spot_lights([tuple(float(v) for v in pts[0]["Position"]), tuple(float(v) for v in pts[3]["Position"])], 9.0, [(0.2, -1.0, 0.1)] * 2, 0.9)
    sz, sy, sx = sc["volume"].shape
    vol = E.ShadowVoxSystem(gpu_ctx, (sx, sy, sz))
    vol.upload(sc["volume"])
    h, w = sc["gb"]["depth24"].shape
    planes = {k: torch.from_numpy(sc["gb"][k].view(np.int32)).pin_memory() for k in ("depth24", "normal", "material", "noise")}
    desc = dict(width=w, height=h, tile_w=w, tile_h=h, tile_first=0, tile_stride=1, n_tiles=1)
    f32 = lambda *shape: torch.empty(shape, dtype=torch.float32).pin_memory()
    outs = dict(shadow=f32(h, w), ao=f32(h, w), point_shadow=f32(len(pts), h, w), spot_shadow=f32(len(spots), h, w), spec_t=f32(h, w))
    E.lighting_host(gpu_ctx, vol, sc["view"], desc, planes, outs, n_ao=3, point=pts, spot=spots)
    mb = E.mask_bytes(len(pts), len(spots))
    assert mb == 2
    pk = dict(shadow_mask=torch.full((h, w, mb), 0xAA, dtype=torch.uint8).pin_memory(), spec_code=torch.full((h, w), 0xAA, dtype=torch.uint8).pin_memory(), ao=f32(h, w))
    E.lighting_host(gpu_ctx, vol, sc["view"], desc, planes, {"packed": pk}, n_ao=3, point=pts, spot=spots)
    un = E.unpack_planes(pk["shadow_mask"].numpy(), pk["spec_code"].numpy(), len(pts), len(spots))
    bits = lambda a: np.ascontiguousarray(a).reshape(-1).view(np.uint32)
    assert np.array_equal(bits(pk["ao"].numpy()), bits(outs["ao"].numpy()))
    assert np.array_equal(bits(un["shadow"]), bits(outs["shadow"].numpy()))
    assert np.array_equal(bits(un["spec_t"]), bits(outs["spec_t"].numpy()))
    assert np.array_equal(bits(un["point_shadow"]), bits(outs["point_shadow"].numpy()))
    assert np.array_equal(bits(un["spot_shadow"]), bits(outs["spot_shadow"].numpy()))
    # the float planes themselves against the oracle (shadows occur in every plane kind, so the masks are exercised both ways)
    wsh, wao, _ = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], 3)
    wt, _ = oracle.pass_reflection(sc["volume"], sc["view"], sc["gb"])
    _assert_plane("shadow", un["shadow"].reshape(h, w), wsh)
    _assert_plane("spec_t", un["spec_t"].reshape(h, w), wt)
    assert 0 < float(un["point_shadow"].mean()) < 1 and len(np.unique(pk["spec_code"].numpy())) > 3
    # unused bits of the last mask byte stay clear; only some passes requested
    assert int((pk["shadow_mask"].numpy()[..., 1] >> 4).max()) == 0
    pk2 = dict(spec_code=torch.zeros((h, w), dtype=torch.uint8).pin_memory())
    E.lighting_host(gpu_ctx, vol, sc["view"], desc, planes, {"packed": pk2}, n_ao=3)
    assert np.array_equal(pk2["spec_code"].numpy(), pk["spec_code"].numpy())
    vol.close()


# ---- derived occupancy levels and the tile march ------------------------------------------------------
@pytest.mark.parametrize("texels", [(64, 48, 64), (37, 21, 50), (16, 16, 16), (130, 9, 3)])
def test_occupancy_levels_match_block_maxima(gpu_ctx, oracle, texels):
    """vxl_volume_build_occupancy against a numpy block reduction of the same bytes; odd sizes included."""
    E = _eng()
    sx, sy, sz = texels
    host = oracle.gen_terrain(sx, sy, sz)
    rs = np.random.RandomState(5)
    host[rs.randint(0, sz, 20), rs.randint(0, sy, 20), rs.randint(0, sx, 20)] = 1    # isolated specks
    vol = E.ShadowVoxSystem(gpu_ctx, texels)
    vol.upload(host)
    from scipy import ndimage
    for shift, tpc in ((2, 2), (3, 4), (4, 8)):
        n = [-(-s // tpc) for s in (sz, sy, sx)]
        pad = np.zeros([k * tpc for k in n], np.uint8)
        pad[:sz, :sy, :sx] = host
        want = (pad.reshape(n[0], tpc, n[1], tpc, n[2], tpc).max(axis=(1, 3, 5)) != 0).astype(np.uint8)
        got = vol.occupancy(shift)
        assert got.shape == want.shape and np.array_equal(got, want), f"shift {shift}: {(got != want).sum()} cells differ"
        if shift >= 3:      # dilated twin: 3x3x3 maximum of the plain level, with a 1-cell border
            wd = ndimage.maximum_filter(np.pad(want, 1), size=3, mode="constant", cval=0)
            gd = vol.occupancy(10 + shift)
            assert gd.shape == wd.shape and np.array_equal(gd, wd), f"dilated {shift}: {(gd != wd).sum()} cells differ"
    vol.clear()                      # an update must be picked up (dirty flag)
    assert vol.occupancy(2).max() == 0
    vol.close()


def test_plain_and_tile_kernels_agree_and_the_tile_engages(gpu_ctx, oracle, terrain):
    """Variant 0 (plain march on the bytes), 1 (occupancy-bit tile, default) and 2 (tile + read counter) give the
    same bits and the same probe counts; the tile march reads the volume for only a fraction of the probes."""
    E = _eng()
    vol, gb = _upload_scene(gpu_ctx, terrain)
    res = {}
    for variant in (0, 1, 2):
        gpu_ctx.set_variant(variant)
        gpu_ctx.stats_reset()
        sh, ao = E.LightAmbientPipeline.Get().Use(terrain["view"], gb, vol, n_ao=8)
        t = E.LightReflectionPipeline.Get().Use(terrain["view"], gb, vol)
        L = terrain["lights"]
        pt = E.LightPointPipeline.Get().Use(terrain["view"], gb, vol,
                                            lambda p: [p.DrawLight(l["Position"], l["Range"], l["Color"], l["Attenuation"]) for l in L])
        st = gpu_ctx.stats()
        res[variant] = (sh.cpu().numpy(), ao.cpu().numpy(), t.cpu().numpy(), pt.cpu().numpy(), st, gpu_ctx.fetched_probes())
    gpu_ctx.set_variant(1)
    for v in (1, 2):
        for a, b in zip(res[0][:4], res[v][:4]):
            assert np.array_equal(a.view(np.uint32), b.view(np.uint32))
        assert res[0][4] == res[v][4]
    assert res[0][5] == 0 and res[1][5] == 0 and 0 < res[2][5] < 0.5 * res[2][4]["steps"], (res[2][5], res[2][4])
    vol.close()
