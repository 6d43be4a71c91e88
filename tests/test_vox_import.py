"""The MagicaVoxel .vox importer (Sources/Editor/Importer/VoxImporter.cpp).

vxl_vox_import (voxelengine_b200/csrc/vxl_voximport.cu, host code) against the restatement oracle/assets_py.vox_import, and both
against the reference's OWN import results: Assets/Mods/default ships FarmHouse.vox / ModernHouse.vox / Player.vox next to the
.v / .p / .pf files the reference's importer wrote from them (79 models).  Those comparisons need /root/reference; the synthetic
files (every `_r` rotation byte, named and unnamed shapes, nested groups, all material types) and the committed digests of the shipped
imports (tests/golden/vox_import_digests.json, made by tools/make_vox_import_golden.py) run anywhere."""
import hashlib
import json
import os
import struct

import numpy as np
import pytest

REF_DEFAULT = "/root/reference/Assets/Mods/default"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_DEFAULT), reason="reference tree not mounted")
GOLDEN = os.path.join(os.path.dirname(__file__), "golden", "vox_import_digests.json")


@pytest.fixture(scope="module")
def E():
    from voxelengine_b200.build import build
    build()
    from voxelengine_b200 import engine
    return engine


@pytest.fixture(scope="module")
def A():
    from oracle import assets_py
    return assets_py


# ---- a .vox writer (MagicaVoxel's published chunk layout) ----------------------------------------------------------------------
def _s(x: str) -> bytes:
    b = x.encode()
    return struct.pack("<i", len(b)) + b


def _d(d: dict) -> bytes:
    return struct.pack("<i", len(d)) + b"".join(_s(k) + _s(v) for k, v in d.items())


def _chunk(cid: bytes, content: bytes, children: bytes = b"") -> bytes:
    return cid + struct.pack("<ii", len(content), len(children)) + content + children


def make_vox(shapes, nodes, rgba=None, matl=()):
    """shapes: [(size3, uint8[n][4] xyzi)]; nodes: list in id order of ("T", name, child, t or None, r or None) | ("G", [ids]) | ("S", shape)."""
    body = b""
    for size, xyzi in shapes:
        body += _chunk(b"SIZE", struct.pack("<iii", *size))
        body += _chunk(b"XYZI", struct.pack("<i", len(xyzi)) + np.ascontiguousarray(xyzi, np.uint8).tobytes())
    for i, n in enumerate(nodes):
        if n[0] == "T":
            _, name, child, t, r = n
            attr = {"_name": name} if name else {}
            frame = {}
            if t is not None:
                frame["_t"] = "%d %d %d" % tuple(t)
            if r is not None:
                frame["_r"] = str(r)
            body += _chunk(b"nTRN", struct.pack("<i", i) + _d(attr) + struct.pack("<iiii", child, -1, 0, 1) + _d(frame))
        elif n[0] == "G":
            body += _chunk(b"nGRP", struct.pack("<i", i) + _d({}) + struct.pack("<i", len(n[1])) + b"".join(struct.pack("<i", c) for c in n[1]))
        else:
            body += _chunk(b"nSHP", struct.pack("<i", i) + _d({}) + struct.pack("<ii", 1, n[1]) + _d({}))
    body += _chunk(b"LAYR", struct.pack("<i", 0) + _d({"_name": "layer"}) + struct.pack("<i", -1))
    if rgba is not None:
        body += _chunk(b"RGBA", np.ascontiguousarray(rgba, np.uint8).tobytes())
    for mid, props in matl:
        body += _chunk(b"MATL", struct.pack("<i", mid) + _d(props))
    body += _chunk(b"rOBJ", _d({"_type": "_bounce", "_diffuse": "2"}))
    return b"VOX " + struct.pack("<i", 150) + _chunk(b"MAIN", b"", body)


# every `_r` byte that encodes a signed permutation: rows 0 and 1 name different axes, three sign bits
ROTATIONS = [rx | (ry << 2) | (s << 4) for rx in range(3) for ry in range(3) if rx != ry for s in range(8)]


def _synthetic(seed=0):
    rs = np.random.RandomState(seed)
    shapes, nodes = [], [("T", "", 1, None, None), ("G", [])]
    children = []

    def add_shape(parent_children, name, r):
        size = tuple(int(v) for v in rs.randint(1, 14, 3))
        n = int(rs.randint(1, 60))
        xyzi = np.stack([rs.randint(0, size[0], n), rs.randint(0, size[1], n), rs.randint(0, size[2], n), rs.randint(1, 256, n)], 1).astype(np.uint8)
        shapes.append((size, xyzi))
        t = tuple(int(v) for v in rs.randint(-300, 300, 3))
        tid = len(nodes)
        nodes.append(("T", name, tid + 1, t, r))
        nodes.append(("S", len(shapes) - 1))
        parent_children.append(tid)

    for k, r in enumerate(ROTATIONS):
        add_shape(children, "part%d" % k if k % 3 == 0 else "", r)
    # a nested group with two more shapes, no `_r` (identity byte 0b0100) and no `_t` on the group's transform
    gt = len(nodes)
    nodes.append(("T", "grp", gt + 1, (10, -20, 30), None))
    nodes.append(("G", []))
    inner = []
    add_shape(inner, "", None)
    add_shape(inner, "leaf.name", None)
    nodes[gt + 1] = ("G", inner)
    children.append(gt)
    nodes[1] = ("G", children)
    rgba = rs.randint(0, 256, (256, 4)).astype(np.uint8)
    matl = [(1, {"_type": "_diffuse", "_rough": "0.1"}), (2, {"_type": "_metal", "_rough": "0.35", "_metal": "1"}),
            (3, {"_type": "_emit", "_emit": "0.5", "_flux": "2"}), (4, {"_type": "_blend", "_rough": "1", "_metal": "0.25"}),
            (5, {"_type": "_glass", "_rough": "0.2"}), (256, {"_type": "_diffuse"}), (7, {"_type": "_emit", "_emit": "1.5"})]
    return make_vox(shapes, nodes, rgba, matl)


def _compare(E, A, data: bytes):
    want = A.vox_import(data)
    imp = E.VoxImporter(data)
    try:
        assert imp.n_models == len(want["models"]) and imp.n_entities == len(want["entities"])
        for i, (name, vox) in enumerate(want["models"]):
            gname, got = imp.model(i)
            assert gname == name and got.shape == vox.shape and np.array_equal(got, vox), (i, name)
        assert np.array_equal(imp.pallete_records(), want["records"])
        ents = imp.entities()
        for g, w in zip(ents, want["entities"]):
            assert int(g["parent"]) == w["parent"] and int(g["model"]) == w["model"] and g["name"].decode("latin-1") == w["name"]
            assert np.array_equal(g["position"].view(np.uint32), np.array(w["position"], np.float32).view(np.uint32))
    finally:
        imp.close()
    return want


def test_synthetic_all_rotations(E, A):
    """48 rotation bytes, named / unnamed shapes, a nested group, five material types: C importer == restatement, bit for bit."""
    for seed in range(6):
        want = _compare(E, A, _synthetic(seed))
        assert len(want["models"]) == len(ROTATIONS) + 2
        # known answers of the restatement itself: sizes are rounded up to multiples of 4, every XYZI voxel lands on exactly one cell
        for _, vox in want["models"]:
            assert all(s % 4 == 0 for s in vox.shape)
        rec = want["records"]
        assert tuple(rec[1, 4:7]) == (229, 0, 0)                  # _diffuse: roughness 0.9 whatever _rough says
        assert tuple(rec[2, 4:7]) == (int(np.float32(0.35) * np.float32(255)), 255, 0)
        assert tuple(rec[3, 4:7]) == (0, 0, 127)
        assert tuple(rec[4, 4:7]) == (255, 63, 0)
        assert tuple(rec[5, 4:7]) == (0, 0, 0)                    # _glass: not handled by the reference
        assert tuple(rec[7, 4:7]) == (0, 0, (382) & 0xFF)         # 1.5 * 255 = 382.5 -> (uint8) keeps the low byte


def test_rotation_is_a_signed_permutation(A):
    """Each `_r` byte moves the voxels by the signed permutation it encodes: the occupied-voxel count is kept and the placement
    equals rotating the dense z-up volume with numpy (transpose + flips), then the importer's z-up -> y-up change of axes."""
    rs = np.random.RandomState(5)
    size = (5, 7, 3)
    dense = (rs.rand(*size) < 0.4) * rs.randint(1, 256, size)     # [x][y][z], MagicaVoxel axes
    xs, ys, zs = np.nonzero(dense)
    xyzi = np.stack([xs, ys, zs, dense[xs, ys, zs]], 1).astype(np.uint8)
    for r in ROTATIONS:
        data = make_vox([(size, xyzi)], [("T", "", 1, None, None), ("G", [2]), ("T", "m", 3, (0, 0, 0), r), ("S", 0)])
        (_, vox), = A.vox_import(data)["models"]
        M = A._vox_matrix(r)
        rot = np.transpose(dense, (M["rx"], M["ry"], M["rz"]))   # rot[x'][y'][z'] = dense at the mapped axes
        for ax, s in enumerate((M["sx"], M["sy"], M["sz"])):
            if s:
                rot = np.flip(rot, ax)
        tx, ty, tz = rot.shape
        want = np.zeros_like(vox)
        want[:ty, :tz, :tx] = np.transpose(rot[:, ::-1, :], (1, 2, 0))     # out[Z = ty-1-y'][Y = z'][X = x']
        assert np.array_equal(vox, want), r
        assert np.count_nonzero(vox) == len(xyzi)


def test_malformed(E):
    from voxelengine_b200.capi import VxlError
    good = _synthetic(0)
    for bad in (b"", b"VOX", b"NOPE" + good[4:], good[:8],                               # no nodes at all
                good[:8] + b"nTRN" + struct.pack("<iii", 0, 0, 0) + struct.pack("<i", 1 << 30),          # dictionary count beyond the file
                good[:8] + b"XYZI" + struct.pack("<iii", 0, 0, 1 << 29)):
        with pytest.raises(VxlError):
            E.VoxImporter(bad)
    # a truncated file stops the chunk loop like the reference's FileReader does; what was read so far is imported or rejected, never a crash
    for cut in (len(good) // 3, len(good) // 2, len(good) - 7):
        try:
            E.VoxImporter(good[:cut]).close()
        except VxlError:
            pass


def test_mutated_files_never_crash(E):
    """600 random mutations of a good file (byte flips, hostile 32-bit counts, deletions): every one is either imported or rejected with
    an error code -- reads are bounded, counts and sizes are checked, nothing crosses the C boundary as an exception."""
    from voxelengine_b200.capi import VxlError
    good = _synthetic(0)
    rs = np.random.RandomState(1)
    ok = bad = 0
    for _ in range(600):
        b = bytearray(good)
        for _ in range(rs.randint(1, 6)):
            mode, pos = rs.randint(3), rs.randint(8, len(b))
            if mode == 0:
                b[pos] = rs.randint(256)
            elif mode == 1 and pos + 4 <= len(b):
                b[pos:pos + 4] = struct.pack("<i", int(rs.choice([-1, 0, 1, 255, 256, 4096, 1 << 20, 1 << 30, -(1 << 31)])))
            else:
                del b[pos:pos + rs.randint(1, 40)]
        try:
            imp = E.VoxImporter(bytes(b))
            for i in range(imp.n_models):
                imp.model(i)
            imp.entities()
            imp.close()
            ok += 1
        except VxlError:
            bad += 1
    assert ok > 0 and bad > 0 and ok + bad == 600


def _digest(res):
    h = hashlib.sha256()
    for name, vox in res["models"]:
        h.update(name.encode() + np.array(vox.shape, "<i4").tobytes() + vox.tobytes())
    rec = res["records"].copy()
    rec[:, 3] = 0
    h.update(rec.tobytes())
    for e in res["entities"]:
        h.update(struct.pack("<ii3f", e["parent"], e["model"], *[float(v) for v in e["position"]]) + e["name"].encode())
    return h.hexdigest()


@needs_ref
@pytest.mark.parametrize("name", ["ModernHouse", "FarmHouse", "Player"])
def test_shipped_imports(E, A, name, tmp_path):
    """The reference's own golden pairs: importing the shipped .vox reproduces the shipped .v files byte for byte, the shipped .p records
    (except the `a` byte, which the reference never initialises: 0xCD / 0 / stack garbage in the three files) and the imported part of
    the shipped .pf scene (names, parents, positions, GUIDs); vxl_vox_scene_write's files equal them too."""
    data = open(f"{REF_DEFAULT}/{name}.vox", "rb").read()
    want = _compare(E, A, data)
    golden = json.load(open(GOLDEN))
    assert _digest(want) == golden[name]["digest"]
    assert len(want["models"]) == golden[name]["models"]
    imp = E.VoxImporter(f"{REF_DEFAULT}/{name}.vox")
    mods = tmp_path / "Mods"
    imp.write(str(mods), "default", name)
    imp.close()
    for mname, vox in want["models"]:
        shipped = open(f"{REF_DEFAULT}/{name}/{mname}.v", "rb").read()
        assert shipped == open(mods / "default" / name / f"{mname}.v", "rb").read(), mname
        assert A.read_v(f"{REF_DEFAULT}/{name}/{mname}.v").tobytes() == vox.tobytes()
    shipped_p = np.fromfile(f"{REF_DEFAULT}/{name}/{name}.p", np.uint8)[:1792].reshape(256, 7)
    ours_p = np.fromfile(mods / "default" / name / f"{name}.p", np.uint8).reshape(256, 7)
    diff = np.argwhere(shipped_p[:, [0, 1, 2, 4, 5, 6]] != ours_p[:, [0, 1, 2, 4, 5, 6]])
    # one record of one file differs: FarmHouse material 249 (`_blend`, `_rough 1`) holds roughness 229 = 0.9 * 255 in the shipped .p while
    # the importer source at this revision gives 255 -- edited after import or written by an older importer; reported, not hidden
    assert [tuple(d) for d in diff] == ([(249, 3)] if name == "FarmHouse" else [])
    pf = f"{REF_DEFAULT}/{name}.pf"
    if os.path.exists(pf):
        shipped = json.load(open(pf))
        ours = json.load(open(mods / "default" / f"{name}.pf"))
        assert len(ours) == len(want["entities"]) <= len(shipped)        # the shipped scenes gained entities in the editor afterwards
        for s, o in zip(shipped, ours):                                   # (a light; instanced furniture)
            assert {k: s[k] for k in o} == o, (s, o)
        # and the text itself: json11's dump + fmt's floats reproduce the shipped bytes of the imported entities
        raw = open(pf, "rb").read().decode()
        text = open(mods / "default" / f"{name}.pf").read()
        assert raw.startswith(text[:-1])


@needs_ref
def test_written_scene_loads(E, A, tmp_path):
    """.vox -> vxl_vox_scene_write -> vxl_scene_load: the written prefab resolves its GUIDs to the written files."""
    imp = E.VoxImporter(f"{REF_DEFAULT}/Player.vox")
    mods = tmp_path / "Mods"
    imp.write(str(mods), "default", "Player")
    ents = E.load_scene(str(mods), "default/Player.pf")
    assert len(ents) == imp.n_entities
    got = imp.entities()
    imp.close()
    for i, e in enumerate(ents):
        assert np.array_equal(e["position"], got[i]["position"])
        if got[i]["model"] >= 0:
            assert int(e["vox_guid"]) == E.asset_guid("default/Player/%s.v" % got[i]["name"].decode())
            assert int(e["pallete_guid"]) == E.asset_guid("default/Player/Player.p")
