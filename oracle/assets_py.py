"""oracle/assets_py.py -- TEST INFRASTRUCTURE: numpy / json restatement of the reference's on-disk formats, the
checker for voxelengine_b200/csrc/vxl_assets.cu.  Follows, relative to /root/reference:
    Sources/Asset/VoxAsset.h:42-50        .v  = int32 SizeX, SizeY, SizeZ + raw bytes, x fastest
    Sources/Asset/PalleteAsset.h:7-24,63-69 + Sources/Vox/PalleteCache.cpp:5-25   .p = 256 x 7-byte VoxMaterial -> colour / material texels
    Sources/Asset/Assets.h:207-210        GUID = std::hash<std::string>(path) = FNV-1a 64 (MSVC)
    Sources/Asset/PrefabAsset.cpp:8-141   .pf = JSON array of entities
    Sources/World/Systems/TransformSystem.cpp:124-135   Matrix = T * Rz * Ry * Rx * S, World = Parent * Matrix
Pinned against the reference's own shipped assets (the GUIDs inside ModernHouse.pf / FarmHouse.pf are the hashes of the shipped file
paths; SURVEY 8c's voxel counts) and, for the matrix arithmetic, against glm itself (oracle/_ref/libvxref.so: ref_transform)."""
from __future__ import annotations

import json
import math
import os

import numpy as np

F = np.float32


def guid(path: str) -> int:
    h = 14695981039346656037
    for b in path.encode():
        h ^= b
        h = (h * 1099511628211) & 0xFFFFFFFFFFFFFFFF
    return h


def read_v(path: str) -> np.ndarray:
    raw = np.fromfile(path, dtype=np.uint8)
    sx, sy, sz = (int(v) for v in raw[:12].view("<i4"))
    return raw[12:12 + sx * sy * sz].reshape(sz, sy, sx).copy()


def write_v(path: str, voxels: np.ndarray):
    sz, sy, sx = voxels.shape
    with open(path, "wb") as f:
        f.write(np.array([sx, sy, sz], "<i4").tobytes())
        f.write(np.ascontiguousarray(voxels, np.uint8).tobytes())


def read_p(path: str):
    m = np.fromfile(path, dtype=np.uint8)[:256 * 7].reshape(256, 7).astype(np.uint32)
    color = m[:, 0] | (m[:, 1] << 8) | (m[:, 2] << 16) | np.uint32(255 << 24)
    material = m[:, 4] | (m[:, 5] << 8) | (m[:, 6] << 16)
    return color.astype(np.uint32), material.astype(np.uint32)


def write_p(path: str, records: np.ndarray):
    """records: uint8 (256, 7) = r g b a roughness metallic emit"""
    np.ascontiguousarray(records, np.uint8).reshape(256, 7).tofile(path)


def _vec3(s: str):
    return [F(float(x)) for x in s.split(" ")[:3]]         # float(x) -> nearest double, F() -> nearest float: from_chars for these inputs


def _rotate(m, angle, axis_index):
    """glm::rotate (ext/matrix_transform.inl:14-43) about a unit axis, float32 step by step; m is [col][row]."""
    c, s = F(math.cos(float(angle))), F(math.sin(float(angle)))
    axis = [F(0), F(0), F(0)]
    axis[axis_index] = F(1)
    temp = [F(F(1) - c) * a for a in axis]
    R = [[None] * 3 for _ in range(3)]
    R[0][0] = F(c + F(temp[0] * axis[0])); R[0][1] = F(F(temp[0] * axis[1]) + F(s * axis[2])); R[0][2] = F(F(temp[0] * axis[2]) - F(s * axis[1]))
    R[1][0] = F(F(temp[1] * axis[0]) - F(s * axis[2])); R[1][1] = F(c + F(temp[1] * axis[1])); R[1][2] = F(F(temp[1] * axis[2]) + F(s * axis[0]))
    R[2][0] = F(F(temp[2] * axis[0]) + F(s * axis[1])); R[2][1] = F(F(temp[2] * axis[1]) - F(s * axis[0])); R[2][2] = F(c + F(temp[2] * axis[2]))
    out = [list(col) for col in m]
    for j in range(3):
        for r in range(4):
            out[j][r] = F(F(F(m[0][r] * R[j][0]) + F(m[1][r] * R[j][1])) + F(m[2][r] * R[j][2]))
    return out


def transform(position, rotation, scale, parent_world=None):
    """-> (matrix, world) as 16 float32 column-major each."""
    m = [[F(1 if i == j else 0) for i in range(4)] for j in range(4)]
    v = [F(x) for x in position]
    m[3] = [F(F(F(F(m[0][r] * v[0]) + F(m[1][r] * v[1])) + F(m[2][r] * v[2])) + m[3][r]) for r in range(4)]
    for ax in (2, 1, 0):
        m = _rotate(m, F(rotation[ax]), ax)
    for j in range(3):
        m[j] = [F(m[j][r] * F(scale[j])) for r in range(4)]
    P = [[F(1 if i == j else 0) for i in range(4)] for j in range(4)] if parent_world is None else [[F(parent_world[j * 4 + r]) for r in range(4)] for j in range(4)]
    w = [[F(F(F(F(P[0][r] * m[j][0]) + F(P[1][r] * m[j][1])) + F(P[2][r] * m[j][2])) + F(P[3][r] * m[j][3])) for r in range(4)] for j in range(4)]
    flat = lambda M: np.array([M[j][r] for j in range(4) for r in range(4)], np.float32)
    return flat(m), flat(w)


def _apply(e, o):
    if e.get("Name") is not None:
        o["name"] = e["Name"]
    t = e.get("Transform")
    if t is not None:
        o["has"] |= 1
        o["position"], o["rotation"], o["scale"] = _vec3(t["Position"]), _vec3(t["Rotation"]), _vec3(t["Scale"])
    r = e.get("VoxRenderer")
    if r is not None:
        o["has"] |= 2
        o["vox_guid"], o["pallete_guid"] = int(r["Vox"], 16), int(r["Pallete"], 16)
        o["pivot"] = _vec3(r["Pivot"]) if r.get("Pivot") is not None else [F(0)] * 3
    l = e.get("Light")
    if l is not None:
        o["has"] |= 4
        o["light"] = dict(light_type=int(l.get("LightType", 0)), intensity=F(l.get("Intensity", 0)), color=_vec3(l["Color"]),
                          attenuation=F(l.get("Attenuation", 0)), range=F(l.get("Range", 0)), angle=F(l.get("Angle", 0)),
                          angle_attenuation=F(l.get("AngleAttenuation", 0)))


def _spawn(path, parent_index, paths, mods_dir, out):
    """PrefabAsset::Spawn (PrefabAsset.cpp:30-141): appends to `out`, returns the index of the prefab's root."""
    with open(path, "r", encoding="utf-8") as f:
        items = json.load(f)
    index_of, root = {}, -1
    for e in items:
        parent = index_of[int(e["Parent"])] if e.get("Parent") is not None else -1
        is_root = parent < 0
        if is_root:
            parent = parent_index
        if e.get("Instance") is not None:
            if paths is None:
                raise ValueError("nested prefab instance without a Mods directory")
            g = int(e["Instance"], 16)
            idx = _spawn(os.path.join(mods_dir, paths[g]), parent, paths, mods_dir, out)
            out[idx]["has"] |= 8
            out[idx]["instance_guid"] = g
            out[idx]["parent"] = parent
        else:
            out.append(dict(parent=parent, name="", has=0, position=[F(0)] * 3, rotation=[F(0)] * 3, scale=[F(1)] * 3))
            idx = len(out) - 1
        out[idx]["id"] = int(e.get("Id", 0))
        index_of[out[idx]["id"]] = idx
        if is_root:
            root = idx
        _apply(e, out[idx])
    return root


def _finish(out):
    for o in out:
        o["matrix"], o["world"] = transform(o["position"], o["rotation"], o["scale"], None if o["parent"] < 0 else out[o["parent"]]["world"])
    return out


def read_pf(path: str):
    """-> list of dicts in file order: id, parent (index or -1), name, position/rotation/scale, matrix, world, vox / light fields."""
    out = []
    _spawn(path, -1, None, "", out)
    return _finish(out)


def load_scene(mods_dir: str, prefab_path: str):
    """read_pf with nested instances expanded; GUID -> path like ModLoader (hash of every relative file path under mods_dir)."""
    paths = {}
    for dp, _, files in os.walk(mods_dir):
        for f in files:
            rel = os.path.relpath(os.path.join(dp, f), mods_dir).replace(os.sep, "/")
            paths[guid(rel)] = rel
    out = []
    _spawn(os.path.join(mods_dir, prefab_path), -1, paths, mods_dir, out)
    return _finish(out)


# ---- MagicaVoxel .vox importer (Sources/Editor/Importer/VoxImporter.cpp) --------------------------------------------------------
# Pinned against the reference's own import results: Assets/Mods/default ships FarmHouse.vox / ModernHouse.vox / Player.vox NEXT TO
# the .v / .p / .pf files the reference's importer wrote from them (tests/test_vox_import.py compares them byte for byte).

def _vox_matrix(r):
    """VoxTransformMatrix (VoxImporter.cpp:37-84): axis map + sign bits of a row-major signed permutation packed in a byte."""
    r &= 0xFF
    rx, ry = r & 3, (r >> 2) & 3
    return dict(rx=rx, ry=ry, rz=3 - (rx | ry), sx=(r >> 4) & 1, sy=(r >> 5) & 1, sz=(r >> 6) & 1)


def _from_chars_int(s, default=0):
    """std::from_chars<int> on a prefix: optional '-', decimal digits; no leading white space; no match leaves the value."""
    i, n = 0, len(s)
    if i < n and s[i] == "-":
        i += 1
    j = i
    while j < n and s[j].isdigit():
        j += 1
    return int(s[:j]) if j > i else default


def _from_chars_float(s, default=0.0):
    """std::from_chars<float> on a prefix (general format: no leading '+' / white space), correctly rounded."""
    import re
    m = re.match(r"-?(\d+\.?\d*([eE][-+]?\d+)?|\.\d+([eE][-+]?\d+)?|inf(inity)?|nan)", s, re.I)
    return F(float(m.group(0))) if m else F(default)


def _u8_of_float(x):
    """float -> uint8 conversion of `emit * 255.0f` (VoxImporter.cpp:200-202): truncation, values in [0, 256) only."""
    return int(F(x)) & 0xFF


def vox_import(data: bytes):
    """VoxImportContext::Import (:284-394) + CreateEntity (:397-476) + VoxImporter::Import (:478-520).
    -> dict(models=[(name, uint8[sz][sy][sx])], records=uint8[256][7] (.p file contents), entities=[dict(name, parent, model, position)])
    Entities are in creation order (depth first, a group before its children) -- the order PrefabAsset::FromWorld numbers them."""
    import struct
    pos = 0

    def take(n):
        nonlocal pos
        b = data[pos:pos + n]
        pos += n
        return b + b"\0" * (n - len(b))          # a FileReader past the end leaves the destination untouched: zeros / blanks stop the loop

    def i32():
        return struct.unpack("<i", take(4))[0]

    def string():
        n = i32()
        return take(max(n, 0)).decode("latin-1")

    def dictionary():
        d = {}
        for _ in range(i32()):
            k = string()
            d[k] = string()
        return d

    if take(4) != b"VOX ":
        raise ValueError("not a .vox file")
    i32()                                         # version
    pallete = np.zeros((257, 4), np.uint8)
    surfaces = np.zeros((257, 3), np.uint8)       # e, r, m  (uninitialised in the reference; zero here and in vxl_vox_import)
    size = (0, 0, 0)
    shapes, nodes = [], []
    while True:
        header = take(4) if pos < len(data) else b"    "
        if pos >= len(data) and header == b"    ":
            break
        i32(); i32()                              # chunk content size, children size (not used to skip: every known chunk is read field by field)
        h0, h1, h2 = chr(header[0]), chr(header[1]), chr(header[2])
        if h0 == "M":
            if h2 == "T":                         # MATL
                mid = i32()
                p = dictionary()
                rough = emit = metal = F(0)
                typ = p.get("_type", "")
                if typ in ("_metal", "_blend"):
                    if "_rough" in p: rough = _from_chars_float(p["_rough"])
                    if "_metal" in p: metal = _from_chars_float(p["_metal"])
                elif typ == "_emit":
                    if "_emit" in p: emit = _from_chars_float(p["_emit"])
                elif typ == "_diffuse":
                    rough = F(0.9)
                if 0 <= mid <= 256:
                    surfaces[mid] = (_u8_of_float(F(emit) * F(255)), _u8_of_float(F(rough) * F(255)), _u8_of_float(F(metal) * F(255)))
            # 'I' = MAIN: nothing
        elif h0 == "P":
            i32()
        elif h0 == "S":
            size = (i32(), i32(), i32())
        elif h0 == "X":
            n = i32()
            vox = np.frombuffer(take(4 * max(n, 0)), np.uint8).reshape(-1, 4).copy()
            shapes.append((size, vox))
        elif h0 == "R":
            pallete[1:257] = np.frombuffer(take(1024), np.uint8).reshape(256, 4)
        elif h0 == "L":
            i32(); dictionary(); i32()
        elif h0 == "I":
            take(256)
        elif h0 == "r":
            dictionary()
        elif h0 == "n":
            i32()                                 # node id
            nd = dictionary()
            name = nd.get("_name", "")
            if h1 == "T":
                child = i32(); i32(); i32(); i32()
                m = dictionary()
                x = y = z = 0
                t = m.get("_t", "")
                if t:
                    e1 = t.find(" ", 1) + 1
                    x = _from_chars_int(t[0:e1])
                    e2 = t.find(" ", e1 + 1) + 1
                    y = _from_chars_int(t[e1:e2])
                    z = _from_chars_int(t[e2:])
                r = 0b0100
                if m.get("_r", ""):
                    r = _from_chars_int(m["_r"], r)
                nodes.append(dict(kind="T", name=name, child=child, t=(x, y, z), m=_vox_matrix(r)))
            elif h1 == "G":
                nodes.append(dict(kind="G", children=[i32() for _ in range(i32())]))
            elif h1 == "S":
                i32()
                nodes.append(dict(kind="S", shape=i32()))
                dictionary()
        else:
            break

    records = np.zeros((256, 7), np.uint8)        # VoxImporter::Import :489-497; `a` keeps VoxMaterial's default
    records[:, 0:3] = pallete[:256, 0:3]
    records[:, 3] = _VOXMATERIAL_DEFAULT_A
    records[:, 4] = surfaces[:256, 1]
    records[:, 5] = surfaces[:256, 2]
    records[:, 6] = surfaces[:256, 0]

    models, entities = [], []
    counter = [0]

    def create(root, parent):
        node = nodes[root["child"]]
        tx, ty, tz = root["t"]
        p = [F(F(tx) * F(0.1)), F(F(tz) * F(0.1)), F(F(-ty) * F(0.1))]
        e = dict(name="", parent=parent, model=-1, position=p)
        idx = len(entities)
        if node["kind"] == "G":
            entities.append(e)
            for c in node["children"]:
                create(nodes[c], idx)
            return idx
        if node["kind"] == "S":
            (sx, sy, sz), vox = shapes[node["shape"]]
            M = root["m"]
            size = (sx, sy, sz)
            ts = (size[M["rx"]], size[M["ry"]], size[M["rz"]])
            center = (ts[0] - ts[0] // 2 if M["sx"] else ts[0] // 2,
                      ts[2] - ts[2] // 2 if M["sz"] else ts[2] // 2,
                      ts[1] - ts[1] // 2 if not M["sy"] else ts[1] // 2)
            e["position"] = [F(p[i] - F(F(center[i]) * F(0.1))) for i in range(3)]
            pad = lambda n: ((n - 1) & ~3) + 4                           # VoxAsset(int32, int32, int32) rounds every size up to a multiple of 4 (VoxAsset.h:26-30)
            out = np.zeros((pad(ts[1]), pad(ts[2]), pad(ts[0])), np.uint8)  # VoxAsset(tSize.x, tSize.z, tSize.y): [z][y][x]; voxels placed by the unpadded tSize
            v = vox.astype(np.int64)
            comp = (v[:, 0], v[:, 1], v[:, 2])
            cx, cy, cz = comp[M["rx"]], comp[M["ry"]], comp[M["rz"]]
            if M["sx"]: cx = ts[0] - cx - 1
            if M["sy"]: cy = ts[1] - cy - 1
            if M["sz"]: cz = ts[2] - cz - 1
            for i in range(len(v)):                                       # in file order: a later record overwrites an earlier one
                out[ts[1] - 1 - cy[i], cz[i], cx[i]] = v[i, 3]
            if root["name"]:
                e["name"] = root["name"]
            else:
                e["name"] = str(counter[0])
                counter[0] += 1
            e["model"] = len(models)
            models.append((e["name"], out))
            entities.append(e)
            return idx
        return -1

    create(nodes[0], -1)
    return dict(models=models, records=records, entities=entities)


_VOXMATERIAL_DEFAULT_A = 0
