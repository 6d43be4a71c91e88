"""CPU: the oracle's pinned operation orders and its restated terrain noise against the REFERENCE's
own vendored code (glm 0.9.9.9, FastNoise, Sources/Util/Noise.cpp) compiled from where it lies under
/root/reference into oracle/_ref/libvxref.so (oracle/Makefile).  Skipped when that library has not
been built (it is built by __graft_entry__.build() whenever the reference tree is mounted)."""
import ctypes as C

import numpy as np
import pytest


@pytest.fixture(scope="module")
def ref(oracle):
    r = oracle.ref_lib()
    if r is None:
        pytest.skip("oracle/_ref/libvxref.so not built (reference tree not mounted)")
    return r


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def test_glm_operation_orders_bitwise(oracle, ref):
    l = oracle.lib()
    rs = np.random.RandomState(0)
    o1, o2 = np.zeros(3, np.float32), np.zeros(3, np.float32)
    p1, p2 = np.zeros(4, np.float32), np.zeros(4, np.float32)
    b1, b2 = np.zeros(12, np.float32), np.zeros(12, np.float32)
    for i in range(4000):
        a = (rs.randn(3) * rs.choice([0.01, 1, 100])).astype(np.float32)
        b = rs.randn(3).astype(np.float32)
        l.vxo_dbg_normalize(_p(a), _p(o1)); ref.ref_normalize(_p(a), _p(o2))
        assert o1.tobytes() == o2.tobytes()
        for f in ("cross", "reflect"):
            getattr(l, "vxo_dbg_" + f)(_p(a), _p(b), _p(o1)); getattr(ref, "ref_" + f)(_p(a), _p(b), _p(o2))
            assert o1.tobytes() == o2.tobytes(), f
        t = float(rs.rand())
        l.vxo_dbg_mix(_p(a), _p(b), t, _p(o1)); ref.ref_mix(_p(a), _p(b), t, _p(o2))
        assert o1.tobytes() == o2.tobytes()
        assert np.float32(l.vxo_dbg_dot(_p(a), _p(b))) == np.float32(ref.ref_dot(_p(a), _p(b)))
        m = rs.randn(16).astype(np.float32)
        v = rs.randn(4).astype(np.float32)
        l.vxo_dbg_matvec(_p(m), _p(v), _p(p1)); ref.ref_matvec(_p(m), _p(v), _p(p2))
        assert p1.tobytes() == p2.tobytes()
        l.vxo_dbg_basis(_p(m), _p(a), _p(b1)); ref.ref_basis(_p(m), _p(a), _p(b2))
        assert b1.tobytes() == b2.tobytes()
        x = float(np.float32(rs.randn() * 50))
        assert l.vxo_dbg_mod(x, 0.5) == ref.ref_mod(x, 0.5)
        assert l.vxo_dbg_smoothstep(0.0, 0.2, abs(x) / 100) == ref.ref_smoothstep(0.0, 0.2, abs(x) / 100)


def test_voxel_of_model_cell_matches_glm(oracle, ref):
    """ShadowVoxSystem.cpp:146-147 through glm vs the oracle's voxeliser (one solid voxel per call)."""
    from voxelengine_b200 import scenes as S
    rs = np.random.RandomState(1)
    for i in range(200):
        m = S.transform_matrix(rs.uniform(0, 6, 3), rs.uniform(-3, 3, 3), (rs.choice([0.5, 1.0, 2.0]),) * 3)
        piv = rs.uniform(0, 1, 3).astype(np.float32)
        x, y, z = (int(v) for v in rs.randint(0, 12, 3))
        basis = np.zeros(12, np.float32)
        ref.ref_basis(_p(m), _p(piv), _p(basis))
        want = np.zeros(3, np.int32)
        ref.ref_voxel_of(_p(basis), x, y, z, _p(want))
        vol = np.zeros((64, 64, 64), np.uint8)
        model = np.zeros((12, 12, 12), np.uint8)
        model[z, y, x] = 255
        e = S.entities(1)
        e[0]["prev"] = e[0]["cur"] = m
        e[0]["pivot"] = piv
        oracle.voxelize(vol, [model], e)
        inside = all(0 <= want[k] < 128 for k in range(3))
        assert int(np.unpackbits(vol).sum()) == (1 if inside else 0)
        if inside:
            assert oracle.get_volume_at(vol, int(want[0]), int(want[1]), int(want[2]), 0)


def test_terrain_noise_matches_fastnoise(oracle, ref):
    rs = np.random.RandomState(2)
    for i in range(5000):
        x, y, z = (float(np.float32(v)) for v in rs.uniform(-50, 2100, 3))
        if i % 2 == 0:
            x, y, z = float(int(x)), float(int(y)), float(int(z))
        assert oracle.terrain_noise(x, y, z) == ref.ref_terrain_noise(x, y, z)


def test_host_camera_matrices_match_glm(oracle, ref):
    """scenes.py builds ViewData in float64 and rounds once; glm works in float32.  They are inputs (both
    sides consume the same bytes), so only closeness is required."""
    from voxelengine_b200 import scenes as S
    P = np.zeros(16, np.float32)
    ref.ref_perspective(0.8, 16 / 9, 0.1, 4096.0, _p(P))
    assert np.abs(P - S.cm(S.perspective(0.8, 16 / 9, 0.1, 4096.0))).max() < 1e-6
    Cm = np.zeros(16, np.float32)
    pos = np.array([26, 15, 25], np.float32)
    ref.ref_camera(_p(pos), 0.81, -0.43, _p(Cm))
    assert np.abs(Cm - S.cm(S.camera_matrix((26, 15, 25), 0.81, -0.43))).max() < 1e-6


def test_voxeliser_matches_reference_shadowvoxsystem(oracle):
    """vxo_voxelize against the REFERENCE's own ShadowVoxSystem.cpp (+ vendored entt, glm) compiled from where it
    lies (oracle/_ref/libvxshadowvox.so): staging bytes and dirty regions, moving / rotated / overlapping / glass /
    out-of-volume / destroyed entities, commands given to the oracle in entt's visiting order."""
    import scene_util as U
    models, e, destroy = U.voxeliser_case()
    n = len(e)
    got = oracle.ref_shadowvox(models, e, destroy)
    if got is None:
        pytest.skip("oracle/_ref/libvxshadowvox.so not built (reference tree not mounted)")
    want_bytes, want_regions, order = got
    assert sorted(order.tolist()) == list(range(n))
    cmds = np.concatenate([e[order], e[np.flatnonzero(destroy)]])
    cmds["flags"][n:] = oracle.ENT_DESTROY
    vol = np.zeros((524, 188, 524), np.uint8)
    regions, valid = oracle.voxelize(vol, models, cmds)
    assert int(vol.sum()) > 0
    assert np.array_equal(vol, want_bytes)
    mine = regions[valid != 0]
    # the constructor queues one whole-volume region (ShadowVoxSystem.cpp:78) that the first OnUpdate uploads first
    assert tuple(want_regions[0]) == (0, 0, 0, 524, 188, 524, 0)
    want_regions = want_regions[1:]
    assert len(mine) == len(want_regions)
    for a in ("x", "y", "z", "w", "h", "d", "mip"):
        assert np.array_equal(mine[a], want_regions[a]), a
