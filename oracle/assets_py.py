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
