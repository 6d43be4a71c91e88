"""ctypes front-end of the CPU oracle (oracle/liboracle.so).  TEST INFRASTRUCTURE ONLY.

Only tests/, bench.py's cpu_baseline / --impl reference legs and __graft_entry__.smoke() import
this module.  The product package (voxelengine_b200) never does.
"""
from __future__ import annotations

import ctypes as C
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB = os.path.join(_HERE, "liboracle.so")
_REF = os.path.join(_HERE, "_ref", "libvxref.so")
_SHADER = os.path.join(_HERE, "_ref", "libvxshader.so")

HIT_DTYPE = np.dtype([("t", "<f4"), ("steps", "<i4"), ("vx", "<i4"), ("vy", "<i4"), ("vz", "<i4"),
                      ("status", "<i4"), ("px", "<f4"), ("py", "<f4"), ("pz", "<f4"),
                      ("nx", "<f4"), ("ny", "<f4"), ("nz", "<f4")])
RAY_DTYPE = np.dtype([("ox", "<f4"), ("oy", "<f4"), ("oz", "<f4"), ("dx", "<f4"), ("dy", "<f4"),
                      ("dz", "<f4"), ("dist", "<f4"), ("pad", "<f4")])
REGION_DTYPE = np.dtype([("x", "<i4"), ("y", "<i4"), ("z", "<i4"), ("w", "<u4"), ("h", "<u4"),
                         ("d", "<u4"), ("mip", "<i4")])
ENTITY_DTYPE = np.dtype([("model", "<i4"), ("flags", "<i4"), ("prev", "<f4", (16,)),
                         ("cur", "<f4", (16,)), ("pivot", "<f4", (3,)), ("_pad", "<i4")])
VIEW_DTYPE = np.dtype([("LastViewMatrix", "<f4", (16,)), ("ViewMatrix", "<f4", (16,)),
                       ("InverseViewMatrix", "<f4", (16,)), ("ProjectionMatrix", "<f4", (16,)),
                       ("InverseProjectionMatrix", "<f4", (16,)), ("Res", "<f4", (2,)),
                       ("iRes", "<f4", (2,)), ("CameraPosition", "<f4", (3,)), ("_pad0", "<i4"),
                       ("Jitter", "<f4", (2,)), ("Frame", "<i4"), ("ColorTextureRID", "<i4"),
                       ("DepthTextureRID", "<i4"), ("PalleteColorRID", "<i4"),
                       ("PalleteMaterialRID", "<i4")])
POINT_LIGHT_DTYPE = np.dtype([("Position", "<f4", (3,)), ("Range", "<f4"), ("Color", "<f4", (3,)),
                              ("Attenuation", "<f4")])
SPOT_LIGHT_DTYPE = np.dtype([("Position", "<f4", (3,)), ("Range", "<f4"), ("Color", "<f4", (3,)),
                             ("Attenuation", "<f4"), ("Direction", "<f4", (3,)), ("Angle", "<f4"),
                             ("AngleAttenuation", "<f4"), ("_pad", "<f4", (3,))])
assert HIT_DTYPE.itemsize == 48 and RAY_DTYPE.itemsize == 32 and VIEW_DTYPE.itemsize == 380
assert POINT_LIGHT_DTYPE.itemsize == 32 and SPOT_LIGHT_DTYPE.itemsize == 64
assert ENTITY_DTYPE.itemsize == 152 and REGION_DTYPE.itemsize == 28

SPARSE, SUPERSPARSE, DDA = 0, 1, 2
ENT_DESTROY = 1


class _Vol(C.Structure):
    _fields_ = [("data", C.c_void_p), ("sx", C.c_int32), ("sy", C.c_int32), ("sz", C.c_int32)]


class _GB(C.Structure):
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("depth24", C.c_void_p),
                ("normal", C.c_void_p), ("material", C.c_void_p), ("noise", C.c_void_p)]


class _Rows(C.Structure):
    _fields_ = [("begin", C.c_int32), ("end", C.c_int32), ("step", C.c_int32)]


class _Stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("steps", C.c_uint64), ("pixels", C.c_uint64)]


class _Model(C.Structure):
    _fields_ = [("voxels", C.c_void_p), ("sx", C.c_int32), ("sy", C.c_int32), ("sz", C.c_int32)]


def build(force: bool = False) -> str:
    """Compile oracle/liboracle.so (and oracle/_ref when /root/reference is mounted)."""
    src = [os.path.join(_HERE, f) for f in ("vxo.cpp", "vxo.h", "Makefile")]
    stale = force or not os.path.exists(_LIB) or any(os.path.getmtime(s) > os.path.getmtime(_LIB) for s in src)
    if stale:
        subprocess.run(["make", "-C", _HERE, "all"], check=True, capture_output=True)
    from .refcheck import build_shaders
    build_shaders.build()          # oracle/_ref/libvxshader.so, only when the reference tree is mounted
    return _LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        build()
        _lib = C.CDLL(_LIB)
        _lib.vxo_terrain_noise.restype = C.c_float
        _lib.vxo_terrain_noise.argtypes = [C.c_float] * 3
        _lib.vxo_dbg_dot.restype = C.c_float
        _lib.vxo_dbg_mod.restype = C.c_float
        _lib.vxo_dbg_mod.argtypes = [C.c_float, C.c_float]
        _lib.vxo_dbg_smoothstep.restype = C.c_float
        _lib.vxo_dbg_smoothstep.argtypes = [C.c_float] * 3
        _lib.vxo_dbg_mix.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
        _lib.vxo_trace_rays.argtypes = [C.c_void_p, C.c_void_p, C.c_int64, C.c_int, C.c_void_p]
    return _lib


def ref_lib():
    """The cross-check library built from the reference's vendored sources, or None."""
    if not os.path.exists(_REF):
        return None
    r = C.CDLL(_REF)
    for name in ("ref_dot", "ref_mod", "ref_smoothstep", "ref_terrain_noise", "ref_perlin3", "ref_perlin2"):
        getattr(r, name).restype = C.c_float
    r.ref_mod.argtypes = [C.c_float, C.c_float]
    r.ref_smoothstep.argtypes = [C.c_float] * 3
    r.ref_terrain_noise.argtypes = [C.c_float] * 3
    r.ref_perlin3.argtypes = [C.c_float] * 3
    r.ref_perlin2.argtypes = [C.c_float] * 2
    r.ref_mix.argtypes = [C.c_void_p, C.c_void_p, C.c_float, C.c_void_p]
    r.ref_perspective.argtypes = [C.c_float] * 4 + [C.c_void_p]
    r.ref_camera.argtypes = [C.c_void_p, C.c_float, C.c_float, C.c_void_p]
    return r


# ---- the reference's own shaders compiled for the host (oracle/refcheck/build_shaders.py) ---------------
SHADER_RAYREC_DTYPE = np.dtype([("o", "<f4", (3,)), ("d", "<f4", (3,)), ("dist", "<f4"), ("result", "<f4"), ("variant", "<i4"),
                                ("fetches", "<i4"), ("lx", "<i4"), ("ly", "<i4"), ("lz", "<i4"), ("pad", "<i4")])
SHADER_PIXREC_DTYPE = np.dtype([("n", "<i4"), ("discarded", "<i4"), ("color", "<f4", (4,)), ("ray", SHADER_RAYREC_DTYPE, (2,))])
assert SHADER_RAYREC_DTYPE.itemsize == 56 and SHADER_PIXREC_DTYPE.itemsize == 136
PASS_AMBIENT, PASS_POINT, PASS_SPOT, PASS_REFLECTION = 0, 1, 2, 3

_shader = None


def shader_lib():
    """libvxshader.so (Light.frag, LightAmbient/Point/Spot/Reflection.frag of the reference compiled as C++), or None."""
    global _shader
    if _shader is None:
        if not os.path.exists(_SHADER):
            return None
        _shader = C.CDLL(_SHADER)
        _shader.vxshader_set_volume.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_int]
        _shader.vxshader_trace.argtypes = [C.c_void_p, C.c_longlong, C.c_int, C.c_void_p]
        _shader.vxshader_pass.argtypes = [C.c_int, C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 5 + [C.c_int] * 5 + [C.c_void_p]
    return _shader


def shader_trace(volume, rays, variant):
    """The reference's raycastShadowVolume{Sparse,SuperSparse,} on `rays`.  Records use HIT_DTYPE with
    steps = texelFetch calls on the volume and (vx, vy, vz) = the last fetched TEXEL."""
    L = shader_lib()
    volume = np.ascontiguousarray(volume, np.uint8)
    sz, sy, sx = volume.shape
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    out = np.zeros(len(rays), dtype=HIT_DTYPE)
    L.vxshader_set_volume(_p(volume), sx, sy, sz)
    L.vxshader_trace(_p(rays), len(rays), int(variant), _p(out))
    return out


def shader_pass(which, volume, view, gb, lights=None, light_index=0, rect=None, albedo=None, light=None, sky=(0.0, 0.0, 0.0)):
    """Run the reference fragment shader `which` over the pixel rectangle (x0, y0, x1, y1); one SHADER_PIXREC per pixel.
    albedo: COLOR_TEXTURE as (H, W) uint32 RGBA8 (default: unbound, reads 0).  light: LIGHT_TEXTURE of the reflection pass as
    (H, W, 4) float32 (default: unbound).  sky: the sky box as a uniform colour."""
    L = shader_lib()
    alb = None if albedo is None else np.ascontiguousarray(albedo, np.uint32)
    L.vxshader_set_albedo.argtypes = [C.c_void_p]
    L.vxshader_set_albedo(None if alb is None else _p(alb))
    lt = None if light is None else np.ascontiguousarray(light, np.float32)
    L.vxshader_set_light.argtypes = [C.c_void_p]
    L.vxshader_set_light(None if lt is None else _p(lt))
    L.vxshader_set_sky.argtypes = [C.c_float] * 3
    L.vxshader_set_sky(float(sky[0]), float(sky[1]), float(sky[2]))
    volume = np.ascontiguousarray(volume, np.uint8)
    sz, sy, sx = volume.shape
    h, w = gb["depth24"].shape
    x0, y0, x1, y1 = rect if rect is not None else (0, 0, w, h)
    out = np.zeros((y1 - y0, x1 - x0), dtype=SHADER_PIXREC_DTYPE)
    L.vxshader_set_volume(_p(volume), sx, sy, sz)
    keep = [np.ascontiguousarray(gb[k], np.uint32) for k in ("depth24", "normal", "material", "noise")]
    lp = None
    if lights is not None:
        dt = POINT_LIGHT_DTYPE if which == PASS_POINT else SPOT_LIGHT_DTYPE
        buf = np.zeros(64, dtype=dt)
        buf[:len(lights)] = np.ascontiguousarray(lights, dtype=dt)
        lp = _p(buf)
    vw = _view(view)
    L.vxshader_pass(int(which), _p(vw), w, h, _p(keep[0]), _p(keep[1]), _p(keep[2]), _p(keep[3]), lp, int(light_index),
                    int(x0), int(y0), int(x1), int(y1), _p(out))
    L.vxshader_set_albedo(None)
    L.vxshader_set_light(None)
    L.vxshader_set_sky(0.0, 0.0, 0.0)
    return out


def shader_taa(view, gb, albedo, motion, light, last_light, rect=None):
    """The reference's LightTAA.frag main() over the pixel rectangle -> out_Color, float32 (y1 - y0, x1 - x0, 4)."""
    L = shader_lib()
    h, w = gb["depth24"].shape
    x0, y0, x1, y1 = rect if rect is not None else (0, 0, w, h)
    keep = [np.ascontiguousarray(gb[k], np.uint32) for k in ("depth24", "normal", "material")] + [np.ascontiguousarray(albedo, np.uint32),
            np.ascontiguousarray(gb["noise"], np.uint32)]
    fl = [np.ascontiguousarray(a, np.float32) for a in (motion, light, last_light)]
    assert fl[0].shape == (h, w, 2) and fl[1].shape == (h, w, 4) and fl[2].shape == (h, w, 4)
    out = np.zeros((y1 - y0, x1 - x0, 4), np.float32)
    vw = _view(view)
    L.vxshader_taa.argtypes = [C.c_void_p, C.c_int, C.c_int] + [C.c_void_p] * 8 + [C.c_int] * 4 + [C.c_void_p]
    L.vxshader_taa(_p(vw), w, h, *[_p(a) for a in keep], *[_p(a) for a in fl], int(x0), int(y0), int(x1), int(y1), _p(out))
    return out


_SHADOWVOX = os.path.join(_HERE, "_ref", "libvxshadowvox.so")


def ref_shadowvox(models, entities, destroy=None):
    """The reference's own ShadowVoxSystem (constructor, OnCreate, OnUpdate, OnVoxDestroyed) on a fresh 524x188x524
    volume -> (staging bytes [sz][sy][sx], regions copied, entt visiting order), or None when the library is absent."""
    if not os.path.exists(_SHADOWVOX):
        return None
    L = C.CDLL(_SHADOWVOX)
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DTYPE)
    n = len(entities)
    keep = [np.ascontiguousarray(m, dtype=np.uint8) for m in models]
    ptrs = (C.c_void_p * len(keep))(*[m.ctypes.data for m in keep])
    dims = np.array([[m.shape[2], m.shape[1], m.shape[0]] for m in keep], np.int32)
    out = np.zeros((524, 188, 524), np.uint8)
    odims = np.zeros(3, np.int32)
    regions = np.zeros(2 * n + 4, dtype=REGION_DTYPE)
    order = np.zeros(n, np.int32)
    d = None if destroy is None else np.ascontiguousarray(destroy, np.int32)
    L.vxref_shadowvox_run.argtypes = [C.c_void_p, C.c_void_p, C.c_int, C.c_void_p, C.c_int, C.c_void_p, C.c_void_p, C.c_void_p, C.c_void_p, C.c_int, C.c_void_p]
    nr = L.vxref_shadowvox_run(ptrs, _p(dims), len(keep), _p(entities), n, None if d is None else _p(d), _p(out), _p(odims), _p(regions),
                               len(regions), _p(order))
    assert odims.tolist() == [524, 188, 524]
    return out, regions[:nr], order


def _p(a):
    return a.ctypes.data_as(C.c_void_p)


def _vol(volume: np.ndarray) -> _Vol:
    """volume: uint8 array of shape (sz, sy, sx), C-contiguous (x fastest)."""
    assert volume.dtype == np.uint8 and volume.ndim == 3 and volume.flags.c_contiguous
    sz, sy, sx = volume.shape
    return _Vol(volume.ctypes.data, sx, sy, sz)


def _gb(gb: dict) -> _GB:
    h, w = gb["depth24"].shape
    for k in ("depth24", "normal", "material"):
        assert gb[k].dtype == np.uint32 and gb[k].shape == (h, w) and gb[k].flags.c_contiguous
    assert gb["noise"].dtype == np.uint32 and gb["noise"].shape == (512, 512)
    return _GB(w, h, gb["depth24"].ctypes.data, gb["normal"].ctypes.data, gb["material"].ctypes.data,
               gb["noise"].ctypes.data)


def _rows(rows, h):
    if rows is None:
        return _Rows(0, h, 1)
    return _Rows(*rows)


def num_threads() -> int:
    return lib().vxo_num_threads()


def set_num_threads(n: int) -> None:
    lib().vxo_set_num_threads(int(n))


def trace_rays(volume, rays, variant):
    rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
    out = np.zeros(rays.shape[0], dtype=HIT_DTYPE)
    v = _vol(volume)
    lib().vxo_trace_rays(C.byref(v), _p(rays), rays.shape[0], variant, _p(out))
    return out


def _view(view):
    view = np.ascontiguousarray(view, dtype=VIEW_DTYPE).reshape(())
    return view


def pass_ambient(volume, view, gb, n_ao=1, rows=None):
    h, w = gb["depth24"].shape
    shadow = np.ones((h, w), np.float32)
    ao = np.zeros((h, w), np.float32)
    st = _Stats()
    v, g, vw = _vol(volume), _gb(gb), _view(view)
    lib().vxo_pass_ambient(C.byref(v), _p(vw), C.byref(g), int(n_ao), _rows(rows, h), _p(shadow), _p(ao), C.byref(st))
    return shadow, ao, dict(rays=st.rays, steps=st.steps, pixels=st.pixels)


MODEL_RAY_DTYPE = np.dtype([("cam", "<f4", (3,)), ("dir", "<f4", (3,)), ("uv", "<f4", (2,))])
MODEL_HIT_DTYPE = np.dtype([("hit", "<i4"), ("material", "<u4"), ("fetches", "<i4"), ("steps", "<i4"), ("pos", "<f4", (3,)), ("normal", "<f4", (3,))])
assert MODEL_RAY_DTYPE.itemsize == 32 and MODEL_HIT_DTYPE.itemsize == 40


def model_mips(voxels):
    """VoxAsset::Upload's mip chain: [level 0, level 1, level 2] as (sz, sy, sx) uint8 arrays."""
    m0 = np.ascontiguousarray(voxels, np.uint8)
    out = [m0]
    for _ in range(2):
        p = out[-1]
        psz, psy, psx = p.shape
        q = np.zeros((max(psz // 2, 1), max(psy // 2, 1), max(psx // 2, 1)), np.uint8)
        lib().vxo_model_mip(_p(p), psx, psy, psz, _p(q))
        out.append(q)
    return out


def trace_model_rays(voxels, rays, frame=0, res=(1920.0, 1080.0), mips=None):
    """GeometryVoxel.frag's clipToAABB + intersectVolume on one model volume (hierarchical-mip DDA)."""
    m = model_mips(voxels) if mips is None else mips
    sz, sy, sx = m[0].shape
    rays = np.ascontiguousarray(rays, dtype=MODEL_RAY_DTYPE)
    out = np.zeros(len(rays), dtype=MODEL_HIT_DTYPE)
    f = lib().vxo_trace_model_rays
    f.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p, C.c_int64, C.c_int, C.c_float, C.c_float, C.c_void_p]
    f(_p(m[0]), _p(m[1]), _p(m[2]), sx, sy, sz, _p(rays), len(rays), int(frame), float(res[0]), float(res[1]), _p(out))
    return out


def shader_model_trace(voxels, rays, frame=0, res=(1920.0, 1080.0), mips=None):
    """The reference's own clipToAABB + intersectVolume (GeometryVoxel.frag compiled for the host); `steps` is not reported."""
    L = shader_lib()
    m = model_mips(voxels) if mips is None else mips
    sz, sy, sx = m[0].shape
    rays = np.ascontiguousarray(rays, dtype=MODEL_RAY_DTYPE)
    out = np.zeros(len(rays), dtype=MODEL_HIT_DTYPE)
    L.vxshader_model_trace.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p, C.c_longlong, C.c_int, C.c_float, C.c_float, C.c_void_p]
    L.vxshader_model_trace(_p(m[0]), _p(m[1]), _p(m[2]), sx, sy, sz, _p(rays), len(rays), int(frame), float(res[0]), float(res[1]), _p(out))
    return out


VOX_CMD_DTYPE = np.dtype([("WorldMatrix", "<f4", (16,)), ("LastWorldMatrix", "<f4", (16,)), ("VolumeRID", "<i4"), ("PalleteIndex", "<i4"), ("_pad", "<i4", (2,))])
FRAG_IN_DTYPE = np.dtype([("cam", "<f4", (3,)), ("dir", "<f4", (3,)), ("mvp", "<f4", (16,))])
FRAG_OUT_DTYPE = np.dtype([("hit", "<i4"), ("material_index", "<u4"), ("fetches", "<i4"), ("_pad", "<i4"), ("color", "<f4", (4,)), ("normal", "<f4", (4,)),
                           ("material", "<f4", (4,)), ("motion", "<f4", (2,)), ("depth", "<f4"), ("_pad2", "<f4")])
assert VOX_CMD_DTYPE.itemsize == 144 and FRAG_IN_DTYPE.itemsize == 88 and FRAG_OUT_DTYPE.itemsize == 80


class _Model(C.Structure):
    _fields_ = [("voxels", C.c_void_p), ("sx", C.c_int32), ("sy", C.c_int32), ("sz", C.c_int32)]


def geometry_fragment(voxels, view, cmd, pal_color, pal_material, frags, mips=None, reference=False):
    """One fragment of GeometryVoxel.frag's main() per entry of `frags` (FRAG_IN_DTYPE) -> FRAG_OUT_DTYPE.
    reference=True runs the reference's own shader compiled for the host instead of the oracle."""
    m = model_mips(voxels) if mips is None else mips
    sz, sy, sx = m[0].shape
    cmd = np.ascontiguousarray(cmd, dtype=VOX_CMD_DTYPE).reshape(())
    pc, pm = np.ascontiguousarray(pal_color, np.uint32), np.ascontiguousarray(pal_material, np.uint32)
    assert pc.ndim == 2 and pc.shape[1] == 256 and pm.shape == pc.shape
    frags = np.ascontiguousarray(frags, dtype=FRAG_IN_DTYPE)
    out = np.zeros(len(frags), dtype=FRAG_OUT_DTYPE)
    vw = _view(view)
    if reference:
        L = shader_lib()
        L.vxshader_geometry_fragment.argtypes = [C.c_void_p] * 3 + [C.c_int] * 3 + [C.c_void_p] * 4 + [C.c_int, C.c_void_p, C.c_longlong, C.c_void_p]
        L.vxshader_geometry_fragment(_p(m[0]), _p(m[1]), _p(m[2]), sx, sy, sz, _p(vw), cmd.ctypes.data_as(C.c_void_p), _p(pc), _p(pm), pc.shape[0],
                                     _p(frags), len(frags), _p(out))
        return out
    mm = (_Model * 3)(*[_Model(a.ctypes.data, a.shape[2], a.shape[1], a.shape[0]) for a in m])
    f = lib().vxo_geometry_fragment
    f.argtypes = [C.c_void_p] * 6 + [C.c_int64, C.c_void_p]
    f(C.cast(mm, C.c_void_p), _p(vw), cmd.ctypes.data_as(C.c_void_p), _p(pc), _p(pm), _p(frags), len(frags), _p(out))
    return out


def gbuffer_models(view, W, H, cmds, models, pal_color, pal_material, want_motion=True):
    """The geometry pass over a draw list: cmds (VOX_CMD_DTYPE, _pad[0] = index into `models`), models = list of voxel arrays.
    -> dict(depth24, normal, material, albedo (uint32 (H, W)), motion (H, W, 2))."""
    cmds = np.ascontiguousarray(cmds, dtype=VOX_CMD_DTYPE)
    keep = [model_mips(m) for m in models]
    flat = (_Model * (3 * len(keep)))(*[_Model(a.ctypes.data, a.shape[2], a.shape[1], a.shape[0]) for mm in keep for a in mm])
    pc, pm = np.ascontiguousarray(pal_color, np.uint32), np.ascontiguousarray(pal_material, np.uint32)
    out = {k: np.zeros((H, W), np.uint32) for k in ("depth24", "normal", "material", "albedo")}
    out["motion"] = np.zeros((H, W, 2), np.float32)
    vw = _view(view)
    f = lib().vxo_gbuffer_models
    f.argtypes = [C.c_void_p, C.c_int, C.c_int, C.c_void_p, C.c_int] + [C.c_void_p] * 8
    f(_p(vw), W, H, _p(cmds), len(cmds), C.cast(flat, C.c_void_p), _p(pc), _p(pm), _p(out["depth24"]), _p(out["normal"]), _p(out["material"]),
      _p(out["albedo"]), _p(out["motion"]) if want_motion else None)
    return out


def resolve_ambient(view, gb, albedo, shadow, ao, rows=None):
    """LightAmbient.frag's out_Color (float32 RGBA, (H, W, 4)) from the march planes + COLOR_TEXTURE (albedo RGBA8)."""
    h, w = gb["depth24"].shape
    alb = np.ascontiguousarray(albedo, np.uint32)
    sh, a = np.ascontiguousarray(shadow, np.float32), np.ascontiguousarray(ao, np.float32)
    out = np.zeros((h, w, 4), np.float32)
    g, vw = _gb(gb), _view(view)
    lib().vxo_resolve_ambient(_p(vw), C.byref(g), _p(alb), _p(sh), _p(a), _rows(rows, h), _p(out))
    return out


def light_taa(view, gb, albedo, motion, light, last_light, rows=None):
    """LightTAA.frag's out_Color (float32 (H, W, 4)) from full-frame planes: motion (H, W, 2), light / last_light (H, W, 4)."""
    h, w = gb["depth24"].shape
    alb = np.ascontiguousarray(albedo, np.uint32)
    mo, li, la = (np.ascontiguousarray(a, np.float32) for a in (motion, light, last_light))
    assert mo.shape == (h, w, 2) and li.shape == (h, w, 4) and la.shape == (h, w, 4)
    out = np.zeros((h, w, 4), np.float32)
    g, vw = _gb(gb), _view(view)
    f = lib().vxo_light_taa
    f.argtypes = [C.c_void_p] * 6 + [_Rows, C.c_void_p]
    f(_p(vw), C.byref(g), _p(alb), _p(mo), _p(li), _p(la), _rows(rows, h), _p(out))
    return out


def resolve_reflection(view, gb, t_plane, light=None, sky=(0.0, 0.0, 0.0), rows=None):
    """LightReflection.frag's out_Color (float32 (H, W, 4)) from the march's t plane, the (TAA) light buffer and a uniform sky."""
    h, w = gb["depth24"].shape
    t = np.ascontiguousarray(t_plane, np.float32)
    li = None if light is None else np.ascontiguousarray(light, np.float32)
    sk = np.asarray(sky, np.float32)
    out = np.zeros((h, w, 4), np.float32)
    g, vw = _gb(gb), _view(view)
    f = lib().vxo_resolve_reflection
    f.argtypes = [C.c_void_p] * 5 + [_Rows, C.c_void_p]
    f(_p(vw), C.byref(g), _p(t), None if li is None else _p(li), _p(sk), _rows(rows, h), _p(out))
    return out


def resolve_local(view, gb, albedo, lights, shadow, spot=False, rows=None, accumulate=None):
    """LightPoint / LightSpot.frag's out_Color summed over `lights` in order, added to `accumulate` (or zeros)."""
    h, w = gb["depth24"].shape
    alb = np.ascontiguousarray(albedo, np.uint32)
    lights = np.ascontiguousarray(lights, dtype=SPOT_LIGHT_DTYPE if spot else POINT_LIGHT_DTYPE)
    sh = np.ascontiguousarray(shadow, np.float32).reshape(len(lights), h, w)
    out = np.zeros((h, w, 4), np.float32) if accumulate is None else np.ascontiguousarray(accumulate, np.float32).copy()
    g, vw = _gb(gb), _view(view)
    lib().vxo_resolve_local(_p(vw), C.byref(g), _p(alb), _p(lights), len(lights), int(bool(spot)), _p(sh), _rows(rows, h), _p(out))
    return out


def pass_point(volume, view, gb, lights, rows=None):
    h, w = gb["depth24"].shape
    lights = np.ascontiguousarray(lights, dtype=POINT_LIGHT_DTYPE)
    out = np.ones((len(lights), h, w), np.float32)
    st = _Stats()
    v, g, vw = _vol(volume), _gb(gb), _view(view)
    lib().vxo_pass_point(C.byref(v), _p(vw), C.byref(g), _p(lights), len(lights), _rows(rows, h), _p(out), C.byref(st))
    return out, dict(rays=st.rays, steps=st.steps, pixels=st.pixels)


def pass_spot(volume, view, gb, lights, rows=None):
    h, w = gb["depth24"].shape
    lights = np.ascontiguousarray(lights, dtype=SPOT_LIGHT_DTYPE)
    out = np.ones((len(lights), h, w), np.float32)
    st = _Stats()
    v, g, vw = _vol(volume), _gb(gb), _view(view)
    lib().vxo_pass_spot(C.byref(v), _p(vw), C.byref(g), _p(lights), len(lights), _rows(rows, h), _p(out), C.byref(st))
    return out, dict(rays=st.rays, steps=st.steps, pixels=st.pixels)


def pass_reflection(volume, view, gb, rows=None):
    h, w = gb["depth24"].shape
    out = np.full((h, w), 256.0, np.float32)
    st = _Stats()
    v, g, vw = _vol(volume), _gb(gb), _view(view)
    lib().vxo_pass_reflection(C.byref(v), _p(vw), C.byref(g), _rows(rows, h), _p(out), C.byref(st))
    return out, dict(rays=st.rays, steps=st.steps, pixels=st.pixels)


def set_volume_at(volume, x, y, z, value):
    sz, sy, sx = volume.shape
    lib().vxo_set_volume_at(_p(volume), sx, sy, sz, int(x), int(y), int(z), int(value))


def get_volume_at(volume, x, y, z, mip=0):
    v = _vol(volume)
    return bool(lib().vxo_get_volume_at(C.byref(v), int(x), int(y), int(z), int(mip)))


def voxelize(volume, models, entities):
    """In-place sequential voxelisation; returns (regions, valid)."""
    sz, sy, sx = volume.shape
    entities = np.ascontiguousarray(entities, dtype=ENTITY_DTYPE)
    keep = [np.ascontiguousarray(m, dtype=np.uint8) for m in models]
    arr = (_Model * len(keep))()
    for i, m in enumerate(keep):
        msz, msy, msx = m.shape
        arr[i] = _Model(m.ctypes.data, msx, msy, msz)
    regions = np.zeros(len(entities), dtype=REGION_DTYPE)
    valid = np.zeros(len(entities), dtype=np.int32)
    lib().vxo_voxelize(_p(volume), sx, sy, sz, arr, _p(entities), len(entities), _p(regions), _p(valid))
    return regions, valid


def upload_regions(image, staging, regions):
    sz, sy, sx = image.shape
    regions = np.ascontiguousarray(regions, dtype=REGION_DTYPE)
    lib().vxo_upload_regions(_p(image), _p(staging), sx, sy, sz, _p(regions), len(regions))


def gen_terrain(sx, sy, sz):
    vol = np.zeros((sz, sy, sx), np.uint8)
    lib().vxo_gen_terrain(_p(vol), sx, sy, sz)
    return vol


def terrain_noise(x, y, z):
    return lib().vxo_terrain_noise(x, y, z)


def perm_table(seed=1337):
    p = np.zeros(512, np.uint8)
    p12 = np.zeros(512, np.uint8)
    lib().vxo_perm_table(int(seed), _p(p), _p(p12))
    return p, p12


def gbuffer_primary(volume, view, width, height):
    d = np.zeros((height, width), np.uint32)
    n = np.zeros((height, width), np.uint32)
    m = np.zeros((height, width), np.uint32)
    v, vw = _vol(volume), _view(view)
    lib().vxo_gbuffer_primary(C.byref(v), _p(vw), int(width), int(height), _p(d), _p(n), _p(m))
    return d, n, m


def luts():
    c = np.zeros(256, np.float32)
    s = np.zeros(256, np.float32)
    lib().vxo_dbg_luts(_p(c), _p(s))
    return c, s
