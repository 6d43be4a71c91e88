"""CPU: hand-derivable invariants of the reference algorithm that pin the oracle.
The reference has no tests of its own for this path, so these are the known-answer checks."""
import numpy as np

from voxelengine_b200 import scenes as S


def _rays(dists, o=(5.3, 6.1, 7.7), d=(0.6, 0.48, 0.64)):
    r = np.zeros(len(dists), dtype=S.RAY_DTYPE)
    r["ox"], r["oy"], r["oz"] = o
    r["dx"], r["dy"], r["dz"] = d
    r["dist"] = dists
    return r


def test_empty_volume_returns_dist_and_counts_steps(oracle):
    vol = np.zeros((16, 16, 16), np.uint8)
    h = oracle.trace_rays(vol, _rays([128.0, 256.0, 40.0, 10.0]), oracle.SPARSE)
    assert h["t"].tolist() == [128.0, 256.0, 40.0, 10.0]
    # Light.frag:138-156: 31 fine probes d=0.5..15.5 (dist is NOT tested); :160-168: d=16..min(dist,164)-1
    assert h["steps"].tolist() == [31 + 112, 31 + 148, 31 + 24, 31]
    assert (h["status"] == 0).all()
    h = oracle.trace_rays(vol, _rays([128.0, 256.0, 40.0, 10.0]), oracle.SUPERSPARSE)
    # Light.frag:182-200: d=2.5..15 (6 probes); then d=17.5 step 5
    assert h["steps"].tolist() == [6 + 23, 6 + 30, 6 + 5, 6]


def test_full_volume_hits_at_first_probe(oracle):
    vol = np.full((16, 16, 16), 255, np.uint8)
    for variant, first in ((oracle.SPARSE, 0.5), (oracle.SUPERSPARSE, 2.5)):
        h = oracle.trace_rays(vol, _rays([128.0, 5.0]), variant)
        assert h["t"].tolist() == [first, first] and h["steps"].tolist() == [1, 1] and (h["status"] == 1).all()


def test_single_byte_hit_in_phase_two(oracle):
    vol = np.zeros((64, 64, 64), np.uint8)
    # ray along +x from (1.25, 20.5, 30.5): phase 2 starts at d=16, pos.x = 1.25 + 31*0.5 = 16.75
    # and advances by 1 per probe; texel x = int(pos.x)/2. Put one bit in texel (20, 10, 15).
    vol[15, 10, 20] = 0x10
    h = oracle.trace_rays(vol, _rays([128.0], o=(1.25, 20.5, 30.5), d=(1.0, 0.0, 0.0)), oracle.SPARSE)
    # first pos.x with int(pos.x)//2 == 20 is 40.75 = 16.75 + 24 -> d = 16 + 24 = 40
    assert h["status"][0] == 2 and h["t"][0] == 40.0 and (h["vx"][0], h["vy"][0], h["vz"][0]) == (40, 20, 30)
    assert h["steps"][0] == 31 + 25


def test_fine_phase_bit_select_quirk(oracle):
    """Light.frag:144-146: bit = sum(mod(pos.c, 0.5) > 0.25) << c, NOT the voxel parity."""
    vol = np.zeros((8, 8, 8), np.uint8)
    vol[1, 1, 1] = 1 << 0b101   # bit 5 of texel (1,1,1)
    # pos = (2.3, 2.1, 2.45): mod 0.5 -> (0.3, 0.1, 0.45) -> bits x,z -> 0b101; texel int(pos/2) = (1,1,1)
    h = oracle.trace_rays(vol, _rays([64.0], o=(2.3, 2.1, 2.45), d=(0.0, 0.0, 0.0)), oracle.SPARSE)
    assert h["status"][0] == 1 and h["t"][0] == 0.5 and h["steps"][0] == 1
    h = oracle.trace_rays(vol, _rays([64.0], o=(2.3, 2.3, 2.45), d=(0.0, 0.0, 0.0)), oracle.SPARSE)
    # bit 0b111 is not set -> fine phase misses (31 probes), coarse phase hits the byte at d=16
    assert h["status"][0] == 2 and h["t"][0] == 16.0 and h["steps"][0] == 32


def test_negative_positions_truncate_toward_zero(oracle):
    """ivec3(pos/2) truncates, so pos in (-2, 0) reads texel 0; pos <= -2 is out of range."""
    vol = np.zeros((4, 4, 4), np.uint8)
    vol[0, 0, 0] = 0xFF
    h = oracle.trace_rays(vol, _rays([32.0], o=(-1.5, -0.5, -1.9), d=(0, 0, 0)), oracle.SPARSE)
    assert h["status"][0] == 1
    h = oracle.trace_rays(vol, _rays([32.0], o=(-2.5, -0.5, -1.9), d=(0, 0, 0)), oracle.SPARSE)
    assert h["status"][0] == 0 and h["t"][0] == 32.0


def test_set_volume_at_roundtrip_and_bounds(oracle):
    rs = np.random.RandomState(0)
    vol = np.zeros((6, 5, 7), np.uint8)   # sx=7, sy=5, sz=6 texels
    truth = np.zeros((12, 10, 14), bool)
    for _ in range(3000):
        x, y, z = rs.randint(-3, 17), rs.randint(-3, 13), rs.randint(-3, 15)
        v = int(rs.randint(0, 2))
        oracle.set_volume_at(vol, x, y, z, v)
        if 0 <= x < 14 and 0 <= y < 10 and 0 <= z < 12:
            truth[z, y, x] = bool(v)
    for z in range(12):
        for y in range(10):
            for x in range(14):
                assert oracle.get_volume_at(vol, x, y, z, 0) == truth[z, y, x]
    # mip 1 == "byte != 0" (Light.frag:22-23)
    assert oracle.get_volume_at(vol, 0, 0, 0, 1) == bool(vol[0, 0, 0])


def test_dda_hit_normal_is_signed_unit_axis_and_counts(oracle):
    vol = np.zeros((32, 32, 32), np.uint8)
    vol[:, 20:22, :] = 0xFF   # slab at voxel y in [40, 44)
    r = _rays([1000.0], o=(20.5, 10.5, 30.5), d=(0.1, 0.9, 0.2))
    h = oracle.trace_rays(vol, r, oracle.DDA)
    assert h["status"][0] == 1 and h["vy"][0] == 40
    # the reference reports the face of the NEXT boundary crossing (Light.frag:51,61), a signed unit axis
    n = np.array([h["nx"][0], h["ny"][0], h["nz"][0]])
    assert sorted(np.abs(n).tolist()) == [0.0, 0.0, 1.0]
    # miss: 4 replays x 256 steps
    empty = np.zeros((400, 4, 4), np.uint8)
    h = oracle.trace_rays(empty, _rays([1e9], o=(1.5, 1.5, 1.5), d=(0.0001, 0.0001, 1.0)), oracle.DDA)
    assert h["status"][0] == 0 and h["steps"][0] == 4 * 256
    # leaving the inclusive bounds [0, 2*dim+1] is a miss with status 3
    h = oracle.trace_rays(np.zeros((4, 4, 4), np.uint8), _rays([1e9], o=(1.5, 1.5, 1.5), d=(0.3, 0.2, 1.0)), oracle.DDA)
    assert h["status"][0] == 3 and h["vz"][0] == 10
    # axis-parallel component: 0*inf = NaN ends every walk after one step
    h = oracle.trace_rays(empty, _rays([1e9], o=(1.5, 1.5, 1.5), d=(0.0, 0.0, 1.0)), oracle.DDA)
    assert h["status"][0] == 0 and h["steps"][0] == 4 and np.isnan(h["t"][0])


def test_voxelize_set_then_clear_is_empty_and_first_frame_quirk(oracle):
    vol = np.zeros((32, 32, 32), np.uint8)
    model = S.shell_cube_model(16)
    e = S.entities(1)
    m = S.transform_matrix((1.3, 1.7, 2.1), (0.2, 0.4, 0.1))
    e[0]["prev"] = e[0]["cur"] = m
    reg, valid = oracle.voxelize(vol, [model], e)
    n_set = int(np.unpackbits(vol).sum())
    assert n_set > 500 and valid[0] == 1
    d = e.copy()
    d[0]["flags"] = S.ENT_DESTROY
    oracle.voxelize(vol, [model], d)
    assert not vol.any()
    # first frame: PreviousWorldMatrix is identity (Components.h:61) => a block at the origin is cleared
    vol[:] = 0xFF
    e2 = S.entities(1)
    e2[0]["cur"] = S.transform_matrix((3.0, 3.0, 3.0))
    oracle.voxelize(vol, [model], e2)
    assert not oracle.get_volume_at(vol, 0, 0, 0, 0) and not oracle.get_volume_at(vol, 15, 0, 7, 0)
    assert oracle.get_volume_at(vol, 5, 5, 5, 0)          # interior of the shell is not touched
    assert oracle.get_volume_at(vol, 30, 30, 30, 0)       # set at the current transform (already 1)


def test_glass_is_not_an_occluder(oracle):
    vol = np.zeros((8, 8, 8), np.uint8)
    model = np.full((4, 4, 4), 15, np.uint8)              # palette index 15 < 16: glass (ShadowVoxSystem.cpp:145)
    e = S.entities(1)
    oracle.voxelize(vol, [model], e)
    assert not vol.any()
    model[:] = 16
    oracle.voxelize(vol, [model], e)
    assert int(np.unpackbits(vol).sum()) == 64


def test_region_rule_matches_reference_quirks(oracle):
    """ShadowVoxSystem.cpp:128-189: AABB seeded with (texDim-1, 0) in mixed units, halved with C
    truncation, clamped; no region when max stays (0,0,0)."""
    vol = np.zeros((16, 16, 16), np.uint8)   # 32^3 voxels
    model = np.full((2, 2, 2), 200, np.uint8)
    e = S.entities(2)
    e[0]["prev"] = e[0]["cur"] = S.transform_matrix((2.0, 1.0, 0.4))      # voxels x 20..21, y 10..11, z 4..5
    e[1]["prev"] = e[1]["cur"] = S.transform_matrix((-5.0, -5.0, -5.0))   # entirely negative: max stays 0
    reg, valid = oracle.voxelize(vol, [model], e)
    assert valid.tolist() == [1, 0]
    r = reg[0]
    # min = min(15, 20)//2 = 7 (the seed wins: sic), max = 21//2 = 10
    assert (r["x"], r["w"]) == (7, 4) and (r["y"], r["h"]) == (5, 1) and (r["z"], r["d"]) == (2, 1)


def test_passes_sky_pixels_generate_no_rays(oracle):
    vol = np.full((8, 8, 8), 255, np.uint8)
    view = S.make_view((0.5, 0.5, 0.5), 0.3, -0.2, 16, 8, 0)
    gb = dict(depth24=np.full((8, 16), 0xFFFFFF, np.uint32), normal=np.zeros((8, 16), np.uint32),
              material=np.zeros((8, 16), np.uint32), noise=S.blue_noise(4))
    sh, ao, st = oracle.pass_ambient(vol, view, gb, 4)
    assert st == dict(rays=0, steps=0, pixels=0) and (sh == 1).all() and (ao == 0).all()
    t, st = oracle.pass_reflection(vol, view, gb)
    assert st["rays"] == 0 and (t == 256).all()
    # depth 0.999 threshold: 0.999*16777215 = 16760437.8 -> 16760437 is lit, 16760438 is sky
    gb["depth24"][:] = 16760438
    assert oracle.pass_ambient(vol, view, gb, 1)[2]["rays"] == 0
    gb["depth24"][:] = 16760437
    assert oracle.pass_ambient(vol, view, gb, 1)[2]["rays"] == 2 * 16 * 8


def test_ao_value_for_a_miss_is_ambient_factor(oracle):
    vol = np.zeros((8, 8, 8), np.uint8)
    view = S.make_view((0.5, 0.5, 0.5), 0.3, -0.2, 16, 8, 2)
    gb = dict(depth24=np.full((8, 16), 5000, np.uint32), normal=np.full((8, 16), 0x007F00, np.uint32),
              material=np.zeros((8, 16), np.uint32), noise=S.blue_noise(4))
    sh, ao, st = oracle.pass_ambient(vol, view, gb, 3)
    assert (sh == 1).all() and np.allclose(ao, 0.05) and st["steps"] == 16 * 8 * (143 + 3 * 29)


def test_noise_lookup_and_hemisphere_luts(oracle):
    c, s = oracle.luts()
    assert c[0] == 1.0 and s[0] == 0.0
    k = np.arange(256)
    theta = (np.float32(6.283) * (k.astype(np.float32) / np.float32(255.0))).astype(np.float32)
    assert np.array_equal(c, np.cos(theta.astype(np.float64)).astype(np.float32))
    assert np.array_equal(s, np.sin(theta.astype(np.float64)).astype(np.float32))


def test_row_subset_equals_full_pass(oracle):
    import scene_util as U
    sc = U.house_scene(oracle, width=48, height=32)
    full, fao, fst = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], 2)
    part, pao, pst = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], 2, rows=(3, 30, 4))
    rows = list(range(3, 30, 4))
    assert np.array_equal(part[rows], full[rows]) and np.array_equal(pao[rows], fao[rows])
    assert pst["rays"] < fst["rays"]
