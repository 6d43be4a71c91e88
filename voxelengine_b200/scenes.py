"""Host-side input synthesis for the voxel-lighting pass: cameras, ViewData,
light lists, .vox-style models, entity transforms and blue noise.

Everything here produces *inputs* (plain numpy arrays with the reference's byte layouts); the same
bytes are handed to the CUDA path and, in tests, to the CPU checker.  Nothing in this module is path
arithmetic.  Layouts cite /root/reference paths.
"""
from __future__ import annotations

import math

import numpy as np

# --- byte layouts ---------------------------------------------------------------------------------
# Sources/Graphics/Renderer/View.h:16-30 (380 B, column-major mat4)
VIEW_DTYPE = np.dtype([("LastViewMatrix", "<f4", (16,)), ("ViewMatrix", "<f4", (16,)),
                       ("InverseViewMatrix", "<f4", (16,)), ("ProjectionMatrix", "<f4", (16,)),
                       ("InverseProjectionMatrix", "<f4", (16,)), ("Res", "<f4", (2,)),
                       ("iRes", "<f4", (2,)), ("CameraPosition", "<f4", (3,)), ("_pad0", "<i4"),
                       ("Jitter", "<f4", (2,)), ("Frame", "<i4"), ("ColorTextureRID", "<i4"),
                       ("DepthTextureRID", "<i4"), ("PalleteColorRID", "<i4"),
                       ("PalleteMaterialRID", "<i4")])
# Sources/Graphics/Pipelines/LightPointPipeline.h:20-25
POINT_LIGHT_DTYPE = np.dtype([("Position", "<f4", (3,)), ("Range", "<f4"), ("Color", "<f4", (3,)),
                              ("Attenuation", "<f4")])
# Sources/Graphics/Pipelines/LightSpotPipeline.h:19-28
SPOT_LIGHT_DTYPE = np.dtype([("Position", "<f4", (3,)), ("Range", "<f4"), ("Color", "<f4", (3,)),
                             ("Attenuation", "<f4"), ("Direction", "<f4", (3,)), ("Angle", "<f4"),
                             ("AngleAttenuation", "<f4"), ("_pad", "<f4", (3,))])
# one voxelisation command (include/vxl.h vxl_entity)
ENTITY_DTYPE = np.dtype([("model", "<i4"), ("flags", "<i4"), ("prev", "<f4", (16,)),
                         ("cur", "<f4", (16,)), ("pivot", "<f4", (3,)), ("_pad", "<i4")])
REGION_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("z", "<i4"), ("w", "<u4"), ("h", "<u4"),
                         ("d", "<u4"), ("mip", "<i4")])
RAY_DTYPE = np.dtype([("ox", "<f4"), ("oy", "<f4"), ("oz", "<f4"), ("dx", "<f4"), ("dy", "<f4"),
                      ("dz", "<f4"), ("dist", "<f4"), ("pad", "<f4")])
HIT_DTYPE = np.dtype([("t", "<f4"), ("steps", "<i4"), ("vx", "<i4"), ("vy", "<i4"), ("vz", "<i4"),
                      ("status", "<i4"), ("px", "<f4"), ("py", "<f4"), ("pz", "<f4"),
                      ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4")])
ENT_DESTROY = 1

NEAR, FAR, FOV = 0.1, 4096.0, 0.8   # Sources/Shaders/lib/Common.frag:12-13; Components.h Camera::Fov

# TAA jitter table is not consumed by the lighting passes; kept zero.


# --- matrices (column-major, glm conventions; computed in float64, rounded once) -------------------
def _translate(v):
    m = np.eye(4)
    m[:3, 3] = v
    return m


def _rot(axis, a):
    c, s = math.cos(a), math.sin(a)
    m = np.eye(4)
    if axis == 0:
        m[1, 1], m[1, 2], m[2, 1], m[2, 2] = c, -s, s, c
    elif axis == 1:
        m[0, 0], m[0, 2], m[2, 0], m[2, 2] = c, s, -s, c
    else:
        m[0, 0], m[0, 1], m[1, 0], m[1, 1] = c, -s, s, c
    return m


def perspective(fov, aspect, near, far):
    """glm::perspective, GL clip convention (Sources/Editor/EditorCamera.cpp:73)."""
    t = math.tan(fov / 2.0)
    m = np.zeros((4, 4))
    m[0, 0] = 1.0 / (aspect * t)
    m[1, 1] = 1.0 / t
    m[2, 2] = -(far + near) / (far - near)
    m[3, 2] = -1.0
    m[2, 3] = -(2.0 * far * near) / (far - near)
    return m


def camera_matrix(pos, yaw, pitch):
    """T(pos) * Ry(yaw) * Rx(pitch)  (EditorCamera.cpp:62-67)."""
    return _translate(pos) @ _rot(1, yaw) @ _rot(0, pitch)


def transform_matrix(position, rotation=(0.0, 0.0, 0.0), scale=(1.0, 1.0, 1.0)):
    """T * Rz * Ry * Rx * S (Sources/World/Systems/TransformSystem.cpp:124-135) as 16 f32, column-major."""
    m = _translate(position) @ _rot(2, rotation[2]) @ _rot(1, rotation[1]) @ _rot(0, rotation[0]) @ np.diag(
        [scale[0], scale[1], scale[2], 1.0])
    return cm(m)


def cm(m):
    """4x4 row/col indexed matrix -> 16 float32 in glm column-major order."""
    return np.ascontiguousarray(np.asarray(m, dtype=np.float64).T.reshape(16).astype(np.float32))


IDENTITY16 = cm(np.eye(4))


def make_view(pos, yaw, pitch, width, height, frame=0, last_view=None):
    """Fill a ViewData block the way WorldRenderer::DrawWorld does (WorldRenderer.cpp:189-206)."""
    cam = camera_matrix(pos, yaw, pitch)
    view = np.linalg.inv(cam)
    proj = perspective(FOV, width / height, NEAR, FAR)
    v = np.zeros((), dtype=VIEW_DTYPE)
    v["ViewMatrix"] = cm(view)
    v["LastViewMatrix"] = cm(view) if last_view is None else last_view
    v["InverseViewMatrix"] = cm(np.linalg.inv(view))
    v["ProjectionMatrix"] = cm(proj)
    v["InverseProjectionMatrix"] = cm(np.linalg.inv(proj))
    v["Res"] = (width, height)
    v["iRes"] = (1.0 / width, 1.0 / height)
    v["CameraPosition"] = pos
    v["Frame"] = frame
    return v


def default_camera(dims_texels, width, height, frame=0):
    """SURVEY 8d camera: volume centre in xz at 0.75*height, yaw 0.81, pitch -0.43 (reference
    defaults, Sources/Editor/Window/ViewportWindow.cpp:31-34).  dims in texels (sx, sy, sz)."""
    sx, sy, sz = dims_texels
    pos = (sx * 2 * 0.5 * 0.1, sy * 2 * 0.75 * 0.1, sz * 2 * 0.5 * 0.1)   # world units = voxels / 10
    return make_view(pos, 0.81, -0.43, width, height, frame)


# --- noise / lights ----------------------------------------------------------------------------------
def blue_noise(seed=4):
    """512x512 RGBA8 stand-in for Assets/.../LDR_RGBA_0.png: mt19937(seed) words."""
    rs = np.random.RandomState(seed)
    return rs.randint(0, 2 ** 32, size=(512, 512), dtype=np.uint64).astype(np.uint32)


MODEL_RAY_DTYPE = np.dtype([("cam", "<f4", (3,)), ("dir", "<f4", (3,)), ("uv", "<f4", (2,))])                     # vxl_model_ray
MODEL_HIT_DTYPE = np.dtype([("hit", "<i4"), ("material", "<u4"), ("fetches", "<i4"), ("steps", "<i4"), ("pos", "<f4", (3,)),
                            ("normal", "<f4", (3,))])                                                                 # vxl_model_hit


VOX_CMD_DTYPE = np.dtype([("WorldMatrix", "<f4", (16,)), ("LastWorldMatrix", "<f4", (16,)), ("VolumeRID", "<i4"), ("PalleteIndex", "<i4"),
                          ("model", "<i4"), ("_pad", "<i4")])                                                           # vxl_vox_cmd


PREFAB_ENTITY_DTYPE = np.dtype([("id", "<i4"), ("parent", "<i4"), ("has", "<u4"), ("light_type", "<i4"), ("position", "<f4", (3,)),
                                ("rotation", "<f4", (3,)), ("scale", "<f4", (3,)), ("pivot", "<f4", (3,)), ("matrix", "<f4", (16,)),
                                ("world", "<f4", (16,)), ("vox_guid", "<u8"), ("pallete_guid", "<u8"), ("instance_guid", "<u8"), ("intensity", "<f4"),
                                ("color", "<f4", (3,)), ("attenuation", "<f4"), ("range", "<f4"), ("angle", "<f4"),
                                ("angle_attenuation", "<f4"), ("name", "S64")])                                          # vxl_prefab_entity
PF_TRANSFORM, PF_VOX, PF_LIGHT, PF_INSTANCE = 1, 2, 4, 8


def point_lights(positions, ranges, color=(2.0, 2.0, 2.0), attenuation=2.0):
    n = len(positions)
    a = np.zeros(n, dtype=POINT_LIGHT_DTYPE)
    a["Position"] = np.asarray(positions, np.float32).reshape(n, 3)
    a["Range"] = np.broadcast_to(np.asarray(ranges, np.float32), (n,))
    a["Color"] = color
    a["Attenuation"] = attenuation
    return a


def spot_lights(positions, ranges, directions, angle=0.3, angle_attenuation=1.0, color=(2.0, 2.0, 2.0),
                attenuation=2.0):
    n = len(positions)
    a = np.zeros(n, dtype=SPOT_LIGHT_DTYPE)
    a["Position"] = np.asarray(positions, np.float32).reshape(n, 3)
    a["Range"] = np.broadcast_to(np.asarray(ranges, np.float32), (n,))
    a["Direction"] = np.asarray(directions, np.float32).reshape(n, 3)
    a["Color"] = color
    a["Attenuation"] = attenuation
    a["Angle"] = angle
    a["AngleAttenuation"] = angle_attenuation
    return a


def surface_height(volume, vx, vz):
    """Highest solid voxel y (or -1) of column (vx, vz) in a packed (sz, sy, sx) uint8 volume."""
    sz, sy, sx = volume.shape
    tx, tz = min(max(vx // 2, 0), sx - 1), min(max(vz // 2, 0), sz - 1)
    col = volume[tz, :, tx]
    bx, bz = vx & 1, vz & 1
    lo = (col >> (bx | (bz << 2))) & 1            # y bit 0
    hi = (col >> (bx | 2 | (bz << 2))) & 1        # y bit 1
    ys = np.empty(2 * sy, np.uint8)
    ys[0::2], ys[1::2] = lo, hi
    nz = np.nonzero(ys)[0]
    return int(nz[-1]) if nz.size else -1


def quarter_point_lights(volume, n=4):
    """SURVEY 8d lights: point lights at the quarter points of the volume 20 voxels above the
    surface, range = 0.25 * extent (world units), attenuation 2."""
    sz, sy, sx = volume.shape
    ext_vox = 2 * sx
    pts = [(0.25, 0.25), (0.75, 0.25), (0.25, 0.75), (0.75, 0.75), (0.5, 0.5), (0.5, 0.25), (0.25, 0.5), (0.75, 0.5)][:n]
    pos = []
    for fx, fz in pts:
        vx, vz = int(fx * 2 * sx), int(fz * 2 * sz)
        h = surface_height(volume, vx, vz)
        pos.append((vx * 0.1, (h + 20) * 0.1, vz * 0.1))
    return point_lights(pos, 0.25 * ext_vox * 0.1)


# --- models (.v layout: Sources/Asset/VoxAsset.h:42-56; 0 empty, 1..15 glass, >=16 solid) ------------
def house_model(size=40, seed=1):
    """Procedural '.vox-style' house: hollow box, walls 2 thick, 8 random solid cuboids,
    a strip of glass (palette index < 16, rendered but not voxelised: ShadowVoxSystem.cpp:145)."""
    rs = np.random.RandomState(seed)
    m = np.zeros((size, size, size), np.uint8)          # (z, y, x)
    m[:, :, :] = 0
    m[:2], m[-2:], m[:, :2], m[:, :, :2], m[:, :, -2:] = 32, 32, 48, 64, 64
    m[:, -2:] = 80
    for _ in range(8):
        lo = rs.randint(2, size - 10, size=3)
        ext = rs.randint(3, 9, size=3)
        m[lo[0]:lo[0] + ext[0], lo[1]:lo[1] + ext[1], lo[2]:lo[2] + ext[2]] = rs.randint(16, 256)
    m[size // 3: size // 2, size // 3: size // 2, :2] = 7   # glass window in the -x wall
    m[:2, size // 3: size // 2, size // 3: size // 2] = 0   # an opening in the -z wall
    return m


def shell_cube_model(size=16):
    m = np.zeros((size, size, size), np.uint8)
    m[:] = 200
    m[1:-1, 1:-1, 1:-1] = 0
    return m


def entities(n):
    e = np.zeros(n, dtype=ENTITY_DTYPE)
    e["prev"] = IDENTITY16      # Transform::PreviousWorldMatrix{1.0f} (Sources/World/Components.h:61)
    e["cur"] = IDENTITY16
    return e


def prop_entities(volume, n=200, model_size=40, seed=2, model=0):
    """SURVEY 8d props: n instances dropped to the terrain surface, yaw in {0, pi/2, pi, 3pi/2}."""
    sz, sy, sx = volume.shape
    rs = np.random.RandomState(seed)
    e = entities(n)
    for i in range(n):
        vx = int(rs.randint(model_size, 2 * sx - model_size))
        vz = int(rs.randint(model_size, 2 * sz - model_size))
        yaw = float(rs.randint(0, 4)) * (math.pi / 2)
        h = max(surface_height(volume, vx, vz), 0)
        e[i]["model"] = model
        e[i]["cur"] = transform_matrix((vx * 0.1, (h + 1) * 0.1, vz * 0.1), (0.0, yaw, 0.0))
        e[i]["pivot"] = (model_size * 0.05, 0.0, model_size * 0.05)
    return e


def dynamic_entities(dims_texels, n=1000, seed=3, model=0):
    """SURVEY 8d dynamic scene: n entities with mt19937(seed) positions; see advance_entities()."""
    sx, sy, sz = dims_texels
    rs = np.random.RandomState(seed)
    e = entities(n)
    pos = np.stack([rs.uniform(2.0, 2 * sx * 0.1 - 4.0, n), rs.uniform(0.6 * 2 * sy * 0.1, 0.9 * 2 * sy * 0.1, n),
                    rs.uniform(2.0, 2 * sz * 0.1 - 4.0, n)], axis=1)
    yaw = rs.uniform(0, 2 * math.pi, n)
    for i in range(n):
        e[i]["model"] = model
        e[i]["cur"] = transform_matrix(pos[i], (0.0, yaw[i], 0.0))
        e[i]["pivot"] = (0.8, 0.8, 0.8)
    return e, pos, yaw


def advance_entities(e, pos, yaw):
    """Per-frame motion: translation (0.1, 0, 0.05) world units, yaw += 0.01; prev <- cur."""
    pos = pos + np.array([0.1, 0.0, 0.05])
    yaw = yaw + 0.01
    e = e.copy()
    e["prev"] = e["cur"]
    for i in range(len(e)):
        e[i]["cur"] = transform_matrix(pos[i], (0.0, yaw[i], 0.0))
    return e, pos, yaw
