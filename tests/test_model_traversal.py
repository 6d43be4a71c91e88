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


# ---- one fragment of GeometryVoxel.frag's main(), and the geometry pass over a draw list (row f1) ----------------------------

def _frag_case(oracle, n=60_000):
    from voxelengine_b200 import scenes as S
    m = U.glassy_house(40)
    rs = np.random.RandomState(2)
    view = S.make_view((6.0, 5.0, -3.0), 2.2, -0.4, 1280, 720, 5)
    cmd = np.zeros((), oracle.VOX_CMD_DTYPE)
    cmd["WorldMatrix"] = S.transform_matrix((1.0, 0.5, 2.0), (0.1, 0.7, -0.2))
    cmd["LastWorldMatrix"] = S.transform_matrix((1.05, 0.5, 2.0), (0.1, 0.69, -0.2))
    cmd["VolumeRID"], cmd["PalleteIndex"] = 3, 1
    pal_c = rs.randint(0, 2 ** 32, size=(2, 256), dtype=np.uint64).astype(np.uint32)
    pal_m = rs.randint(0, 2 ** 32, size=(2, 256), dtype=np.uint64).astype(np.uint32)
    rays = U.model_rays(m.shape, n, seed=4)
    fr = np.zeros(n, oracle.FRAG_IN_DTYPE)
    fr["cam"], fr["dir"] = rays["cam"], rays["dir"]
    W = cmd["WorldMatrix"].reshape(4, 4).T.astype(np.float64)
    P, V = (view[k].reshape(4, 4).T.astype(np.float64) for k in ("ProjectionMatrix", "ViewMatrix"))
    fr["mvp"] = S.cm(P @ V @ W)
    return m, view, cmd, pal_c, pal_m, fr, ~(rays["dir"] == 0).any(axis=1)


def test_oracle_fragment_matches_reference_main(oracle):
    """vxo_geometry_fragment against GeometryVoxel.frag's own main() compiled for the host: every output, bit for bit."""
    if oracle.shader_lib() is None:
        pytest.skip("oracle/_ref/libvxshader.so not built (reference tree not mounted)")
    m, view, cmd, pal_c, pal_m, fr, regular = _frag_case(oracle)
    want = oracle.geometry_fragment(m, view, cmd, pal_c, pal_m, fr, reference=True)
    got = oracle.geometry_fragment(m, view, cmd, pal_c, pal_m, fr)
    assert np.array_equal(got["hit"][regular], want["hit"][regular]) and np.array_equal(got["fetches"][regular], want["fetches"][regular])
    h = regular & (want["hit"] == 1)
    assert 0.5 < h.mean() < 0.99
    for f in ("color", "normal", "material", "motion", "depth"):
        assert np.array_equal(_bits(got[f][h]), _bits(want[f][h])), f
    assert len(np.unique(want["color"][h][:, 3])) == 2                   # alpha = step(hitMat, 16): both sides of 16 present


def _oracle_gbuffer(oracle, models, cmds, pc, pm, view, w, h):
    c2 = np.zeros(len(cmds), oracle.VOX_CMD_DTYPE)
    for k in ("WorldMatrix", "LastWorldMatrix", "VolumeRID", "PalleteIndex"):
        c2[k] = cmds[k]
    c2["_pad"][:, 0] = cmds["model"]
    return oracle.gbuffer_models(view, w, h, c2, models, pc, pm)


def test_oracle_geometry_pass_depth_test_and_coverage(oracle):
    """Draw-list semantics on the oracle: nearest model wins (LESS on D24, first draw wins ties), a camera inside a box sees
    nothing of it (back faces are culled), sky stays at depth 0xFFFFFF."""
    from voxelengine_b200 import scenes as S
    models, cmds, pc, pm, view = U.model_scene()
    w, h = 160, 96
    g = _oracle_gbuffer(oracle, models, cmds, pc, pm, view, w, h)
    cov = g["depth24"] != 0xFFFFFF
    assert 0.2 < cov.mean() < 0.8 and np.all(g["normal"][~cov] == 0) and np.all(g["albedo"][~cov] == 0)
    single = [_oracle_gbuffer(oracle, models, cmds[i:i + 1], pc, pm, view, w, h) for i in range(len(cmds))]
    dmin = np.minimum.reduce([s["depth24"] for s in single])
    assert np.array_equal(g["depth24"], dmin)
    first = np.argmax(np.stack([s["depth24"] == dmin for s in single]), axis=0)           # first draw reaching the minimum
    for k in ("normal", "material", "albedo"):
        want = np.choose(first, [s[k] for s in single])
        assert np.array_equal(g[k][cov], want[cov]), k
    assert np.array_equal(_oracle_gbuffer(oracle, models, cmds[::-1].copy(), pc, pm, view, w, h)["depth24"], g["depth24"])
    inside = S.make_view((2.0, 2.0, 2.0), 0.3, -0.1, w, h, 0)                           # inside the first (40^3 voxels = 4 units) model
    assert np.all(_oracle_gbuffer(oracle, models, cmds[:1], pc, pm, inside, w, h)["depth24"] == 0xFFFFFF)


@pytest.mark.gpu
def test_cuda_geometry_pass_matches_oracle_and_feeds_the_light_passes(gpu_ctx, oracle):
    import torch
    from voxelengine_b200 import engine as E
    models, cmds, pc, pm, view = U.model_scene()
    w, h = 160, 96
    want = _oracle_gbuffer(oracle, models, cmds, pc, pm, view, w, h)
    vol = E.ShadowVoxSystem(gpu_ctx, (64, 48, 64))
    ids = [vol.add_model(m) for m in models]
    dc = cmds.copy()
    dc["model"] = np.asarray(ids, np.int32)[cmds["model"]]
    dev = gpu_ctx.torch_device
    d_pc, d_pm = (torch.from_numpy(a.view(np.int32)).to(dev) for a in (pc, pm))
    for tile, rank, world in ((None, 0, 1), ((32, 16), 1, 3)):
        fb = E.GeometryBuffer(gpu_ctx, w, h) if tile is None else E.GeometryBuffer(gpu_ctx, w, h, tile[0], tile[1], rank=rank, world=world)
        motion = torch.zeros(fb.shape + (2,), dtype=torch.float32, device=dev)
        alb = E.GeometryVoxelPipeline.Get().Use(view, fb, dc, d_pc, d_pm, motion=motion)
        got = dict(depth24=fb.depth24, normal=fb.normal, material=fb.material, albedo=alb)
        for k, t in got.items():
            assert np.array_equal(t.cpu().numpy().view(np.uint32), fb.to_tiles(want[k])), (k, tile)
        mt = motion.cpu().numpy()
        for c in range(2):
            assert np.array_equal(_bits(mt[..., c]), _bits(fb.to_tiles(np.ascontiguousarray(want["motion"][..., c]).view(np.uint32)).view(np.float32))), tile
    # the produced G-buffer drives the light passes: voxelise the same instances into the shadow volume and compare with the oracle
    ents = np.zeros(len(cmds), E.ENTITY_DTYPE)
    ents["model"] = dc["model"]
    ents["cur"] = cmds["WorldMatrix"]
    ents["prev"] = cmds["WorldMatrix"]
    vol.OnUpdate(ents, want_regions=False)
    fb = E.GeometryBuffer(gpu_ctx, w, h)
    E.GeometryVoxelPipeline.Get().Use(view, fb, dc, d_pc, d_pm)
    from voxelengine_b200 import scenes as S
    fb.set_noise(S.blue_noise(4))
    sh, ao = E.LightAmbientPipeline.Get().Use(view, fb, vol, n_ao=2)
    gbo = dict(depth24=want["depth24"], normal=want["normal"], material=want["material"], noise=S.blue_noise(4))
    osh, oao, _ = oracle.pass_ambient(vol.download(), view, gbo, 2)
    assert np.array_equal(sh.cpu().numpy()[0], osh) and np.array_equal(ao.cpu().numpy()[0], oao)
    assert 0.0 < float(osh[want["depth24"] != 0xFFFFFF].mean()) < 1.0
    vol.close()
