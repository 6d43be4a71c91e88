"""The G-buffer producer's traversal of one model volume (core): VoxAsset::Upload's mip rule and
GeometryVoxel.frag's clipToAABB + intersectVolume (the reference's hierarchical-mip DDA).

CPU: the oracle against THE REFERENCE'S OWN FUNCTIONS compiled for the host (oracle/_ref/libvxshader.so) and against the
committed reference-generated fixture tests/golden/ref_model.npz.  GPU: vxl_trace_model_rays through the C ABI against the
oracle (every ray, bit for bit) and against the fixture.  Rays with an exactly zero direction component make the reference
arithmetic produce NaN and then convert it to int, which GLSL leaves undefined (x86: INT_MIN, sm_100: 0): those rays are
compared between oracle and CUDA (both cvt.rzi) but not against the host-run reference."""
import os

import numpy as np
import pytest

import scene_util as U

HERE = os.path.dirname(os.path.abspath(__file__))


def _bits(a):
    return np.ascontiguousarray(a, np.float32).view(np.uint32)


def _same(got, want, sel, steps=True):
    for f in ("hit", "material", "fetches") + (("steps",) if steps else ()):
        assert np.array_equal(got[f][sel], want[f][sel]), f
    h = sel & (want["hit"] == 1)
    for f in ("pos", "normal"):
        assert np.array_equal(_bits(got[f][h]), _bits(want[f][h])), f


def test_mip_rule_is_first_nonzero_child(oracle):
    m = U.glassy_house(40)
    mips = oracle.model_mips(m)
    assert [x.shape for x in mips] == [(40, 40, 40), (20, 20, 20), (10, 10, 10)]
    for lvl in (1, 2):
        p = mips[lvl - 1]
        c = np.stack([p[dz::2, dy::2, dx::2] for dz in (0, 1) for dy in (0, 1) for dx in (0, 1)])      # child order vi = x + 2y + 4z
        first = np.argmax(c != 0, axis=0)
        want = np.take_along_axis(c, first[None], axis=0)[0]
        assert np.array_equal(mips[lvl], want)
    assert np.array_equal(oracle.model_mips(np.zeros((8, 4, 12), np.uint8))[2], np.zeros((2, 1, 3), np.uint8))


@pytest.mark.parametrize("frame", [0, 1])
def test_oracle_traversal_matches_reference_function(oracle, frame):
    if oracle.shader_lib() is None:
        pytest.skip("oracle/_ref/libvxshader.so not built (reference tree not mounted)")
    m = U.glassy_house(40)
    rays = U.model_rays(m.shape, 120_000, seed=3 + frame)
    want = oracle.shader_model_trace(m, rays, frame=frame, res=(1280.0, 720.0))
    got = oracle.trace_model_rays(m, rays, frame=frame, res=(1280.0, 720.0))
    regular = ~(rays["dir"] == 0).any(axis=1)
    _same(got, want, regular, steps=False)
    h = want["hit"][regular]
    assert 0.3 < h.mean() < 0.95 and (want["material"][regular][h == 1] < 16).any()              # hits, misses, glass
    assert want["fetches"][regular].max() > 40


def test_oracle_traversal_matches_reference_golden(oracle):
    g = np.load(os.path.join(HERE, "golden", "ref_model.npz"))
    m = U.glassy_house(40)
    rays = U.model_rays(m.shape, int(g["n"]), seed=11)
    got = oracle.trace_model_rays(m, rays, frame=int(g["frame"]), res=(1280.0, 720.0))
    regular = ~(rays["dir"] == 0).any(axis=1)
    want = np.zeros(len(rays), got.dtype)
    for f in ("hit", "material", "fetches", "pos", "normal"):
        want[f] = g[f]
    _same(got, want, regular, steps=False)


@pytest.mark.gpu
def test_cuda_traversal_matches_oracle_and_reference_golden(gpu_ctx, oracle):
    from voxelengine_b200 import engine as E
    g = np.load(os.path.join(HERE, "golden", "ref_model.npz"))
    m = U.glassy_house(40)
    vol = E.ShadowVoxSystem(gpu_ctx, (16, 16, 16))
    mid = vol.add_model(m)
    every = np.ones(1, bool)
    for frame, seed, n in ((int(g["frame"]), 11, int(g["n"])), (0, 5, 300_000), (1, 6, 300_000)):
        rays = U.model_rays(m.shape, n, seed=seed)
        got = E.trace_model_rays(gpu_ctx, mid, rays, frame=frame, res=(1280.0, 720.0))
        want = oracle.trace_model_rays(m, rays, frame=frame, res=(1280.0, 720.0))
        nan_ok = np.isnan(want["pos"]).any(axis=1) | np.isnan(want["normal"]).any(axis=1)         # NaN payloads are not compared
        _same(got, want, np.broadcast_to(every, (n,)) & ~nan_ok)
        for f in ("hit", "material", "fetches", "steps"):
            assert np.array_equal(got[f], want[f]), f
        if seed == 11:
            ref = np.zeros(n, got.dtype)
            for f in ("hit", "material", "fetches", "pos", "normal"):
                ref[f] = g[f]
            _same(got, ref, ~(rays["dir"] == 0).any(axis=1), steps=False)
    # a flat model: sizes that stop halving
    flat = np.zeros((8, 4, 12), np.uint8)
    flat[2:6, 1:3, 3:9] = 200
    mid2 = vol.add_model(flat)
    rays = U.model_rays(flat.shape, 50_000, seed=9)
    got = E.trace_model_rays(gpu_ctx, mid2, rays, frame=0)
    want = oracle.trace_model_rays(flat, rays, frame=0)
    for f in ("hit", "material", "fetches", "steps"):
        assert np.array_equal(got[f], want[f]), f
    vol.close()
