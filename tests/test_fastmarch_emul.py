"""The tile march (occupancy-bit tile in front of the reference's texel test, vxl_bitmarch.cuh) compiled for the
HOST must reproduce the oracle's plain march bit-for-bit: distance, probe count, hit voxel, hit position.
Runs without a GPU; the same header is what the CUDA kernels instantiate."""
import os
import sys

import numpy as np
import pytest

sys.path.insert(0, os.path.dirname(__file__))
import scene_util as U  # noqa: E402
from emul import emul_py  # noqa: E402
from voxelengine_b200 import scenes as S  # noqa: E402


@pytest.fixture(scope="module")
def terrain(oracle):
    return U.terrain_scene(oracle)


@pytest.fixture(scope="module")
def emul(terrain):
    e = emul_py.Emul(terrain["volume"])
    yield e
    e.close()


def _compare(got, want):
    for f in ("steps", "vx", "vy", "vz", "status"):
        assert np.array_equal(got[f], want[f]), f"{f}: {(got[f] != want[f]).sum()} rays differ"
    for f in ("t", "px", "py", "pz"):
        a, b = got[f].view(np.uint32), want[f].view(np.uint32)
        nan_both = np.isnan(got[f]) & np.isnan(want[f])
        assert np.array_equal(a[~nan_both], b[~nan_both]), f"{f}: {(a != b).sum()} rays differ bitwise"


def _surface_rays(rs, vol, n, center, spread, dist_choices):
    """Rays shaped like the light passes': origins clustered within `spread` voxels of `center`, unit-ish directions."""
    r = np.zeros(n, dtype=S.RAY_DTYPE)
    o = np.asarray(center, np.float32) + rs.uniform(-spread, spread, size=(n, 3)).astype(np.float32)
    d = rs.normal(size=(n, 3)).astype(np.float32)
    d /= np.maximum(np.linalg.norm(d, axis=1, keepdims=True), 1e-6).astype(np.float32)
    d *= rs.choice([1.0, 1.0, 1.0, 0.7, 1.5], size=(n, 1)).astype(np.float32)
    r["ox"], r["oy"], r["oz"] = o[:, 0], o[:, 1], o[:, 2]
    r["dx"], r["dy"], r["dz"] = d[:, 0], d[:, 1], d[:, 2]
    r["dist"] = rs.choice(dist_choices, size=n).astype(np.float32)
    return r


def _surface_points(vol, k, rs):
    """k voxel positions just above solid terrain."""
    sz, sy, sx = vol.shape
    pts = []
    while len(pts) < k:
        x, z = rs.randint(4, 2 * sx - 4), rs.randint(4, 2 * sz - 4)
        col = vol[z // 2, :, x // 2]
        ys = np.nonzero(col)[0]
        if len(ys):
            pts.append((x, 2 * int(ys.max()) + 3, z))
    return pts


@pytest.mark.parametrize("variant", [0, 1])
@pytest.mark.parametrize("geom", ["ambient", "nogroup", "reflection"])
def test_emulated_tile_march_matches_oracle_on_pass_like_rays(oracle, terrain, emul, variant, geom):
    vol = terrain["volume"]
    rs = np.random.RandomState(100 + variant)
    total_fetched = total_steps = 0
    for center in _surface_points(vol, 8, rs):
        rays = _surface_rays(rs, vol, 20_000, center, 6.0, [128.0, 256.0, 40.0, 73.3, 17.0, 10.0, 164.0, 500.0])
        want = oracle.trace_rays(vol, rays, variant)
        got, fetched, steps = emul.trace(rays, variant, center, geom=geom)
        _compare(got, want)
        assert steps == int(want["steps"].sum())
        got2, fetched2, _ = emul.trace(rays, variant, center, geom=geom, direct=False, lockstep=False)   # bounds-checked fetch, per-lane exit
        _compare(got2, want)
        assert fetched2 == fetched
        total_fetched += fetched
        total_steps += steps
    # the tile really engages: most probes are answered by a clear occupancy bit
    assert 0 < total_fetched < 0.5 * total_steps, (total_fetched, total_steps)


def test_emulated_scan_march_matches_oracle_on_ao_like_rays(oracle, terrain, emul):
    """AO rays (SuperSparse, dist 128) through the scan + resolve march: candidate mask, approximate-position texel
    lookup with its exact-replay fallback, hit index arithmetic.  Large coordinates make eps large enough that the
    fallback runs; tiny origins spread exercises the phase-1 replay."""
    vol = terrain["volume"]
    rs = np.random.RandomState(321)
    total_fetched = total_steps = 0
    for center in _surface_points(vol, 10, rs):
        rays = _surface_rays(rs, vol, 30_000, center, 5.0, [128.0])
        rays["oy"] -= rs.uniform(0.0, 3.0, size=len(rays)).astype(np.float32)        # some origins inside the terrain
        want = oracle.trace_rays(vol, rays, 1)
        got, fetched, steps = emul.trace(rays, 1, center, geom="scan")
        _compare(got, want)
        assert steps == int(want["steps"].sum())
        got, fetched_far, _ = emul.trace(rays, 1, center, geom="scan_far")       # without the near (texel-level) tile
        _compare(got, want)
        got, fetched_pre, _ = emul.trace(rays, 1, center, geom="scan_pre")       # eligibility from the bundle precheck
        _compare(got, want)
        assert fetched_pre <= fetched_far
        assert fetched < fetched_far
        total_fetched += fetched
        total_steps += steps
    assert 0 < total_fetched < 0.5 * total_steps, (total_fetched, total_steps)
    # other distances take the per-probe march, ineligible rays the plain one
    rays = _surface_rays(rs, vol, 20_000, center, 200.0, [128.0, 40.0, 164.0])
    got, _, _ = emul.trace(rays, 1, center, geom="scan")
    _compare(got, oracle.trace_rays(vol, rays, 1))


@pytest.mark.parametrize("variant", [0, 1])
def test_emulated_fast_march_matches_oracle_on_arbitrary_rays(oracle, terrain, emul, variant):
    """Rays that start outside the volume, at negative coordinates, axis-parallel, on voxel boundaries, far from
    the tile centre, with NaN / inf / zero directions: eligibility must route them correctly."""
    vol = terrain["volume"]
    sz, sy, sx = vol.shape
    rs = np.random.RandomState(7 + variant)
    rays = U.random_rays(rs, 100_000, (2 * sx, 2 * sy, 2 * sz), dist_lo=5.0)
    k = 64
    rays["dx"][:k] = np.nan
    rays["oy"][k:2 * k] = np.inf
    rays["dx"][2 * k:3 * k] = 0.0; rays["dy"][2 * k:3 * k] = 0.0; rays["dz"][2 * k:3 * k] = 0.0
    rays["dist"][3 * k:4 * k] = np.nan
    rays["dist"][4 * k:5 * k] = np.inf
    rays["dist"][5 * k:6 * k] = -3.0
    want = oracle.trace_rays(vol, rays, variant)
    rays["dist"][6 * k:] = np.where(rs.uniform(size=len(rays) - 6 * k) < 0.5, np.float32(128.0), rays["dist"][6 * k:])
    want = oracle.trace_rays(vol, rays, variant)
    for geom, center in [("ambient", (sx, sy, sz)), ("nogroup", (10, 2 * sy - 5, 2 * sz - 3)), ("reflection", (-40, 50, 300)),
                         ("ambient", (3, 3, 3)), ("scan", (sx, sy, sz)), ("scan", (3, 3, 3)), ("scan_pre", (sx, sy, sz)), ("scan_pre", (2, 5, 3))]:
        got, _, steps = emul.trace(rays, variant, center, geom=geom)
        _compare(got, want)


def test_emulated_plain_march_is_the_oracle(oracle, terrain, emul):
    vol = terrain["volume"]
    sz, sy, sx = vol.shape
    rays = U.random_rays(np.random.RandomState(3), 50_000, (2 * sx, 2 * sy, 2 * sz))
    for variant in (0, 1):
        got, fetched, _ = emul.trace(rays, variant, (sx, sy, sz), fast=False)
        _compare(got, oracle.trace_rays(vol, rays, variant))
        assert fetched == 0


def test_host_occupancy_levels_are_block_maxima(terrain, emul):
    """emul's occupancy levels (the reference for the GPU build test) against a numpy block reduction."""
    vol = terrain["volume"]
    sz, sy, sx = vol.shape
    for shift, tpc in ((1, 1), (2, 2), (3, 4), (4, 8), (5, 16)):
        n = [-(-s // tpc) for s in (sz, sy, sx)]
        pad = np.zeros([k * tpc for k in n], np.uint8)
        pad[:sz, :sy, :sx] = vol
        want = (pad.reshape(n[0], tpc, n[1], tpc, n[2], tpc).max(axis=(1, 3, 5)) != 0).astype(np.uint8)
        assert np.array_equal(emul.level(shift), want)
