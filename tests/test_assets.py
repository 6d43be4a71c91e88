"""On-disk formats: .v volumes, .p palettes, asset GUIDs, .pf scenes with TransformSystem's matrices.

The C-ABI readers (voxelengine_b200/csrc/vxl_assets.cu, host code) against the numpy / json restatement in oracle/assets_py.py:
on files written by the tests (anywhere), on the reference's own shipped assets and against glm itself (when /root/reference is
mounted), and with known answers taken from the reference's data (the GUIDs stored inside its prefabs, SURVEY 8c's voxel counts).
GPU: a scene written to disk, loaded through the C ABI, rendered by the geometry pass and lit -- equal to the oracle on the same files."""
import ctypes as C
import json
import os

import numpy as np
import pytest

import scene_util as U

REF_ASSETS = "/root/reference/Assets/Mods"
needs_ref = pytest.mark.skipif(not os.path.isdir(REF_ASSETS), reason="reference tree not mounted")


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


def test_guid_known_answers(E, A):
    """Assets::Hash == FNV-1a 64: the GUIDs the reference's shipped ModernHouse.pf stores for its first models and its palette."""
    kat = {"default/ModernHouse/0.v": 0x44B7A418296B6797, "default/ModernHouse/1.v": 0x3CF5791825481EF8,
           "default/ModernHouse/ModernHouse.p": 0x5EAD52E114AB9ABC, "": 14695981039346656037}
    for path, want in kat.items():
        assert E.asset_guid(path) == want == A.guid(path), path


def _write_scene(tmp_path, A, rotations=True):
    """The draw list of scene_util.model_scene as files: three .v models, two .p palettes, one .pf scene with a hierarchy."""
    models, cmds, pal_c, pal_m, view = U.model_scene()
    root = tmp_path / "mods" / "t"
    root.mkdir(parents=True)
    for i, m in enumerate(models):
        A.write_v(str(root / f"{i}.v"), m)
    rs = np.random.RandomState(3)
    for k in range(2):
        rec = rs.randint(0, 256, size=(256, 7)).astype(np.uint8)
        A.write_p(str(root / f"pal{k}.p"), rec)
    ents = [{"Id": 10, "Name": "root", "Transform": {"Position": "0.5 0.25 -0.5", "Rotation": "0.0 0.2 0.0" if rotations else "0.0 0.0 0.0", "Scale": "1.0 1.0 1.0"}}]
    place = [((0.0, 0.0, 0.0), (0.0, 0.3, 0.0), 0), ((3.0, 0.2, 1.0), (0.1, -0.8, 0.05), 1), ((1.0, 0.5, 1.5), (0.0, 1.1, 0.0), 1),
             ((-1.5, 1.0, 2.5), (0.4, 0.2, -0.3), 2), ((2.0, 3.0, 2.0), (0.0, 0.0, 0.0), 0)]
    for i, (pos, rot, mi) in enumerate(place):
        if not rotations:
            rot = (0.0, 0.0, 0.0)
        ents.append({"Id": 20 + i, "Name": f"m{i}", "Parent": 10 if i != 3 else 20,            # one grandchild
                     "Transform": {"Position": "%r %r %r" % pos, "Rotation": "%r %r %r" % rot, "Scale": "1.0 1.0 1.0"},
                     "VoxRenderer": {"Pallete": "%X" % A.guid(f"t/pal{i % 2}.p"), "Pivot": "0.1 0.0 0.2" if i == 1 else "0.0 0.0 0.0",
                                     "Vox": "%X" % A.guid(f"t/{mi}.v")},
                     "IKChain": {"Target": "x", "Pole": "y", "Depth": 2}})
    ents.append({"Id": 99, "Name": 'lamp "A"', "Parent": 10, "Transform": {"Position": "2.0 3.5 1.0", "Rotation": "0.0 0.0 0.0", "Scale": "1.0 1.0 1.0"},
                 "Light": {"Angle": 0.30000001192092896, "AngleAttenuation": 1, "Attenuation": 2, "Color": "1.0 0.5 0.25", "Intensity": 2.5, "LightType": 0, "Range": 8}})
    with open(root / "scene.pf", "w") as f:
        json.dump(ents, f)
    return root, models, view


def _check_prefab(got, want):
    assert len(got) == len(want)
    for g, w in zip(got, want):
        assert (int(g["id"]), int(g["parent"]), int(g["has"])) == (w["id"], w["parent"], w["has"])
        assert g["name"].decode() == w["name"][:63]
        for k in ("position", "rotation", "scale"):
            assert np.array_equal(g[k], np.array(w[k], np.float32)), k
        if w["has"] & 2:
            assert (int(g["vox_guid"]), int(g["pallete_guid"])) == (w["vox_guid"], w["pallete_guid"])
            assert np.array_equal(g["pivot"], np.array(w["pivot"], np.float32))
        if w["has"] & 4:
            l = w["light"]
            assert int(g["light_type"]) == l["light_type"] and np.array_equal(g["color"], np.array(l["color"], np.float32))
            for k in ("intensity", "attenuation", "range", "angle", "angle_attenuation"):
                assert g[k] == l[k], k


def test_files_written_here_round_trip(E, A, tmp_path):
    root, models, _ = _write_scene(tmp_path, A, rotations=False)
    for i, m in enumerate(models):
        assert np.array_equal(E.read_vox_file(str(root / f"{i}.v")), m) and np.array_equal(A.read_v(str(root / f"{i}.v")), m)
    for k in range(2):
        c, m = E.read_pallete_file(str(root / f"pal{k}.p"))
        wc, wm = A.read_p(str(root / f"pal{k}.p"))
        assert np.array_equal(c, wc) and np.array_equal(m, wm) and np.all(c >> 24 == 255) and np.all(m >> 24 == 0)
    got, want = E.read_prefab_file(str(root / "scene.pf")), A.read_pf(str(root / "scene.pf"))
    _check_prefab(got, want)
    for g, w in zip(got, want):                                   # no rotation: every product is exact, so bit equality
        assert np.array_equal(g["matrix"], w["matrix"]) and np.array_equal(g["world"], w["world"])
    assert got[4]["parent"] == 1 and got[-1]["name"] == b'lamp "A"'
    # error behaviour: codes, no exceptions across the boundary
    lib = E.capi.load()
    n = C.c_int()
    assert lib.vxl_prefab_file_read(str(root / "0.v").encode(), None, 0, C.byref(n)) != 0
    assert lib.vxl_prefab_file_read(str(root / "missing.pf").encode(), None, 0, C.byref(n)) != 0
    dims = np.zeros(3, np.int32)
    assert lib.vxl_vox_file_read(str(root / "pal0.p").encode(), dims.ctypes.data_as(C.c_void_p), None, 0) != 0      # dims fail the sanity check
    with open(root / "bad.pf", "w") as f:
        json.dump([{"Id": 1, "Instance": "ABCDEF"}], f)
    with pytest.raises(Exception):
        E.read_prefab_file(str(root / "bad.pf"))                    # instances need a Mods directory
    with pytest.raises(Exception):
        E.load_scene(str(root.parent), "t/bad.pf")                  # ... and a GUID that resolves


def test_nested_prefab_instances(E, A, tmp_path):
    """PrefabAsset.cpp:47-56: an entity with "Instance" spawns that prefab under its parent; the nested root takes the entity's Id,
    Name and components (its Transform replaces the nested root's)."""
    root, _, _ = _write_scene(tmp_path, A, rotations=False)
    mods = root.parent
    outer = [{"Id": 0, "Name": "world", "Transform": {"Position": "1.0 2.0 3.0", "Rotation": "0.0 0.0 0.0", "Scale": "1.0 1.0 1.0"}},
             {"Id": 1, "Name": "first", "Parent": 0, "Instance": "%X" % A.guid("t/scene.pf"),
              "Transform": {"Position": "10.0 0.0 0.0", "Rotation": "0.0 0.0 0.0", "Scale": "2.0 2.0 2.0"}},
             {"Id": 2, "Name": "second", "Parent": 0, "Instance": "%X" % A.guid("t/scene.pf")},            # keeps the nested root's Transform
             {"Id": 3, "Name": "after", "Parent": 1, "Transform": {"Position": "0.0 1.0 0.0", "Rotation": "0.0 0.0 0.0", "Scale": "1.0 1.0 1.0"}}]
    with open(root / "outer.pf", "w") as f:
        json.dump(outer, f)
    got, want = E.load_scene(str(mods), "t/outer.pf"), A.load_scene(str(mods), "t/outer.pf")
    inner = A.read_pf(str(root / "scene.pf"))
    assert len(got) == 2 + 2 * len(inner)
    _check_prefab(got, want)
    for g, w in zip(got, want):
        assert np.array_equal(g["matrix"], w["matrix"]) and np.array_equal(g["world"], w["world"])
        assert int(g["instance_guid"]) == w.get("instance_guid", 0)
    first, second, after = got[1], got[1 + len(inner)], got[-1]
    assert first["name"] == b"first" and first["has"] & 8 and np.array_equal(first["scale"], [2, 2, 2]) and first["parent"] == 0
    assert second["name"] == b"second" and np.array_equal(second["position"], inner[0]["position"])
    assert after["parent"] == 1 and np.array_equal(after["world"][12:15], [11.0, 4.0, 3.0])               # (1,2,3) + (10,0,0) + 2 * (0,1,0)
    assert np.array_equal(got[2]["world"][12:15], np.float32([1, 2, 3]) + np.float32([10, 0, 0]) + np.float32(2) * np.asarray(inner[1]["position"], np.float32))


def test_matrices_match_glm(E, A, tmp_path):
    """TransformSystem's matrix chain with rotations: the C-ABI reader against glm itself (bit for bit: same libm) and against
    the numpy restatement (1e-6: its cos / sin are rounded from double)."""
    root, _, _ = _write_scene(tmp_path, A, rotations=True)
    got, want = E.read_prefab_file(str(root / "scene.pf")), A.read_pf(str(root / "scene.pf"))
    _check_prefab(got, want)
    for g, w in zip(got, want):
        assert np.allclose(g["matrix"], w["matrix"], rtol=0, atol=1e-6) and np.allclose(g["world"], w["world"], rtol=0, atol=2e-6)
    ref = os.path.join(os.path.dirname(os.path.abspath(A.__file__)), "_ref", "libvxref.so")
    if not os.path.exists(ref):
        pytest.skip("oracle/_ref/libvxref.so not built (reference tree not mounted)")
    L = C.CDLL(ref)
    fp = lambda a: a.ctypes.data_as(C.c_void_p)
    L.ref_transform.argtypes = [C.c_void_p] * 6
    ident = np.eye(4, dtype=np.float32).reshape(16)
    for g in got:
        parent = ident if g["parent"] < 0 else np.ascontiguousarray(got[int(g["parent"])]["world"])
        m, w = np.zeros(16, np.float32), np.zeros(16, np.float32)
        pos, rot, scl = (np.ascontiguousarray(g[k]) for k in ("position", "rotation", "scale"))
        L.ref_transform(fp(pos), fp(rot), fp(scl), fp(parent), fp(m), fp(w))
        assert np.array_equal(g["matrix"], m) and np.array_equal(g["world"], w), g["name"]


@needs_ref
def test_reference_assets(E, A):
    """Every shipped .v / .p / .pf through the C-ABI readers == the oracle's; SURVEY 8c's known counts; every GUID a prefab
    stores is the hash of a shipped file."""
    base = os.path.join(REF_ASSETS, "default")
    shipped = {}
    n_v = 0
    for dp, _, files in os.walk(base):
        for f in files:
            full = os.path.join(dp, f)
            rel = os.path.relpath(full, REF_ASSETS).replace(os.sep, "/")
            shipped[A.guid(rel)] = rel
            if f.endswith(".v"):
                n_v += 1
                assert np.array_equal(E.read_vox_file(full), A.read_v(full)), rel
            elif f.endswith(".p"):
                (c, m), (wc, wm) = E.read_pallete_file(full), A.read_p(full)
                assert np.array_equal(c, wc) and np.array_equal(m, wm), rel
    assert n_v == 112
    big = E.read_vox_file(os.path.join(base, "ModernHouse", "1.v"))
    assert big.shape == (204, 88, 160) and int((big != 0).sum()) == 226069 and int((big >= 16).sum()) == 219622
    assert E.read_vox_file(os.path.join(base, "ModernHouse", "4.v")).shape == (40, 40, 40)
    assert np.array_equal(E.read_prefab_file(os.path.join(base, "FarmHouse.pf")), E.load_scene(REF_ASSETS, "default/FarmHouse.pf"))
    n_player = sum("VoxRenderer" in e for e in json.load(open(os.path.join(base, "player_ik.pf"))))
    assert n_player > 0
    for name, n_models, n_lights in (("FarmHouse.pf", 67, 1), ("ModernHouse.pf", 6 + n_player, 0)):       # ModernHouse instances player_ik.pf
        got, want = E.load_scene(REF_ASSETS, "default/" + name), A.load_scene(REF_ASSETS, "default/" + name)
        _check_prefab(got, want)
        for g, w in zip(got, want):
            exact = all(float(r) == 0.0 for r in w["rotation"]) and (w["parent"] < 0 or np.array_equal(got[w["parent"]]["world"], want[w["parent"]]["world"]))
            assert np.allclose(g["matrix"], w["matrix"], rtol=0, atol=1e-6) and np.allclose(g["world"], w["world"], rtol=0, atol=1e-5)
            if exact:                                                # no rotation on the way: every product is exact
                assert np.array_equal(g["matrix"], w["matrix"]) and np.array_equal(g["world"], w["world"])
        vox = got[(got["has"] & 2) != 0]
        assert len(vox) == n_models and int(((got["has"] & 4) != 0).sum()) == n_lights, (len(vox), name)
        for g in vox:
            if int(g["vox_guid"]) != 0:                                                     # Asset::NullGUID: an empty slot
                assert shipped[int(g["vox_guid"])].endswith(".v") and shipped[int(g["pallete_guid"])].endswith(".p")
    lamp = E.read_prefab_file(os.path.join(base, "FarmHouse.pf"))
    lamp = lamp[(lamp["has"] & 4) != 0][0]
    assert np.allclose(lamp["world"][12:15], (35.43537, 4.567088, 28.338522)) and lamp["range"] == 10 and lamp["intensity"] == 2


@pytest.mark.gpu
def test_scene_files_to_lit_frame(gpu_ctx, oracle, A, tmp_path):
    """Files -> C-ABI loaders -> voxeliser + geometry pass + ambient pass on the GPU == the oracle on what the oracle's readers
    return for the same files."""
    import torch
    from voxelengine_b200 import engine as E
    from voxelengine_b200 import scenes as S
    root, _, view = _write_scene(tmp_path, A, rotations=True)
    w, h = 160, 96
    ents = E.read_prefab_file(str(root / "scene.pf"))
    vox = ents[(ents["has"] & S.PF_VOX) != 0]
    by_guid = {A.guid(f"t/{n}"): n for n in ("0.v", "1.v", "2.v", "pal0.p", "pal1.p")}
    vol = E.ShadowVoxSystem(gpu_ctx, (64, 48, 64))
    lib = gpu_ctx.lib
    model_id, o_models = {}, {}
    for g in sorted(set(int(v) for v in vox["vox_guid"])):
        mid = C.c_int()
        E.check(lib.vxl_model_load_v(gpu_ctx.h, str(root / by_guid[g]).encode(), C.byref(mid)), "vxl_model_load_v")
        model_id[g] = int(mid.value)
        o_models[g] = A.read_v(str(root / by_guid[g]))
    pals = sorted(set(int(v) for v in vox["pallete_guid"]))
    pc = np.stack([E.read_pallete_file(str(root / by_guid[g]))[0] for g in pals])
    pm = np.stack([E.read_pallete_file(str(root / by_guid[g]))[1] for g in pals])
    cmds = np.zeros(len(vox), S.VOX_CMD_DTYPE)
    cmds["WorldMatrix"] = vox["world"]; cmds["LastWorldMatrix"] = vox["world"]
    cmds["VolumeRID"] = 3 + np.arange(len(vox)); cmds["PalleteIndex"] = [pals.index(int(g)) for g in vox["pallete_guid"]]
    cmds["model"] = [model_id[int(g)] for g in vox["vox_guid"]]
    dev = gpu_ctx.torch_device
    fb = E.GeometryBuffer(gpu_ctx, w, h)
    alb = E.GeometryVoxelPipeline.Get().Use(view, fb, cmds, torch.from_numpy(pc.view(np.int32)).to(dev), torch.from_numpy(pm.view(np.int32)).to(dev))
    # oracle side, from its own readers
    o_ents = [e for e in A.read_pf(str(root / "scene.pf")) if e["has"] & 2]
    order = sorted(o_models)
    oc = np.zeros(len(o_ents), oracle.VOX_CMD_DTYPE)
    for i, e in enumerate(o_ents):
        oc[i]["WorldMatrix"] = e["world"]; oc[i]["LastWorldMatrix"] = e["world"]
        oc[i]["VolumeRID"] = 3 + i; oc[i]["PalleteIndex"] = pals.index(e["pallete_guid"]); oc[i]["_pad"][0] = order.index(e["vox_guid"])
    wpc = np.stack([A.read_p(str(root / by_guid[g]))[0] for g in pals]); wpm = np.stack([A.read_p(str(root / by_guid[g]))[1] for g in pals])
    # rotated entities: the reader's cos / sin come from libm, the numpy restatement's from double -- drive the oracle with the
    # reader's matrices where they differ in the last bit (checked against glm in test_matrices_match_glm)
    for i in range(len(o_ents)):
        assert np.allclose(oc[i]["WorldMatrix"], vox[i]["world"], rtol=0, atol=2e-6)
        oc[i]["WorldMatrix"] = vox[i]["world"]; oc[i]["LastWorldMatrix"] = vox[i]["world"]
    want = oracle.gbuffer_models(view, w, h, oc, [o_models[g] for g in order], wpc, wpm)
    for k, t in dict(depth24=fb.depth24, normal=fb.normal, material=fb.material, albedo=alb).items():
        assert np.array_equal(t.cpu().numpy().view(np.uint32)[0], want[k]), k
    assert 0.15 < float((want["depth24"] != 0xFFFFFF).mean()) < 0.9
    # the same instances into the shadow volume (pivot from the file), then the ambient pass on the produced G-buffer
    ve = np.zeros(len(vox), E.ENTITY_DTYPE)
    ve["model"] = cmds["model"]; ve["cur"] = vox["world"]; ve["prev"] = vox["world"]; ve["pivot"] = vox["pivot"]
    vol.OnUpdate(ve, want_regions=False)
    fb.set_noise(S.blue_noise(4))
    sh, ao = E.LightAmbientPipeline.Get().Use(view, fb, vol, n_ao=2)
    gbo = dict(depth24=want["depth24"], normal=want["normal"], material=want["material"], noise=S.blue_noise(4))
    osh, oao, _ = oracle.pass_ambient(vol.download(), view, gbo, 2)
    assert np.array_equal(sh.cpu().numpy()[0], osh) and np.array_equal(ao.cpu().numpy()[0], oao)
    vol.close()


def test_mutated_asset_files_never_crash(E, tmp_path):
    """Random mutations of a .pf scene (byte flips, inserted JSON punctuation, deletions) and random / truncated .v and .p files:
    each is either read or rejected with an error code; the readers never crash or read out of bounds."""
    from voxelengine_b200.capi import VxlError
    ents = [{"Id": 0, "Name": "root", "Transform": {"Position": "0.0 0.0 0.0", "Rotation": "0.0 0.1 0.0", "Scale": "1.0 1.0 1.0"}}]
    for i in range(1, 6):
        ents.append({"Id": i, "Name": "m%d" % i, "Parent": i - 1 if i % 2 else 0,
                     "Transform": {"Position": "1.5 2.0 -3.25", "Rotation": "0.0 0.0 0.3", "Scale": "1.0 1.0 1.0"},
                     "VoxRenderer": {"Pallete": "5EAD52E114AB9ABC", "Pivot": "0.0 0.0 0.0", "Vox": "44B7A418296B6797"},
                     "Light": {"LightType": 0, "Intensity": 2.0, "Color": "1.0 1.0 1.0", "Attenuation": 2.0, "Range": 10.0, "Angle": 0.3, "AngleAttenuation": 1.0}})
    good = json.dumps(ents).encode()
    rs = np.random.RandomState(2)
    ok = bad = 0
    p = str(tmp_path / "t.pf")
    for _ in range(500):
        b = bytearray(good)
        for _ in range(rs.randint(1, 5)):
            mode, pos = rs.randint(3), rs.randint(0, len(b))
            if mode == 0:
                b[pos] = rs.randint(256)
            elif mode == 1:
                b[pos:pos] = bytes(rs.choice(list(b'{}[]",:0123456789-e. \\u'), rs.randint(1, 6)).astype(np.uint8))
            else:
                del b[pos:pos + rs.randint(1, 30)]
        open(p, "wb").write(bytes(b))
        try:
            E.read_prefab_file(p)
            ok += 1
        except VxlError:
            bad += 1
    assert ok > 0 and bad > 0
    pv, pp = str(tmp_path / "t.v"), str(tmp_path / "t.p")
    for _ in range(200):
        raw = rs.randint(0, 256, rs.randint(0, 64)).astype(np.uint8).tobytes()
        if rs.rand() < 0.5:
            raw = np.array([rs.randint(-2, 70), rs.randint(-2, 70), rs.randint(-2, 70)], "<i4").tobytes() + raw
        open(pv, "wb").write(raw)
        open(pp, "wb").write(raw * rs.randint(1, 60))
        for fn, path in ((E.read_vox_file, pv), (E.read_pallete_file, pp)):
            try:
                fn(path)
            except VxlError:
                pass
