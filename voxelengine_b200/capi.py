"""ctypes binding of include/vxl.h (libvxl.so).  Loading fails loudly when the CUDA library has not
been built: there is no CPU or PyTorch fallback for any entry point."""
from __future__ import annotations

import ctypes as C
import os

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("VXL_LIB") or os.path.join(_HERE, "libvxl.so")     # VXL_LIB: A/B experiments with differently compiled kernels

# every symbol include/vxl.h declares (tests check the library exports each one)
SYMBOLS = [
    "vxl_abi_version", "vxl_last_error_string", "vxl_ctx_create", "vxl_ctx_destroy", "vxl_ctx_set_stream",
    "vxl_sync", "vxl_stats_reset", "vxl_stats_read", "vxl_launch_count", "vxl_malloc", "vxl_free",
    "vxl_host_alloc", "vxl_host_free", "vxl_memcpy_h2d", "vxl_memcpy_d2h", "vxl_memset",
    "vxl_volume_create", "vxl_volume_destroy", "vxl_volume_dims", "vxl_volume_upload_regions",
    "vxl_volume_upload", "vxl_volume_download", "vxl_volume_clear", "vxl_volume_device_ptr",
    "vxl_volume_mark_dirty", "vxl_volume_build_occupancy", "vxl_model_create", "vxl_volume_voxelize",
    "vxl_pass_ambient", "vxl_pass_point", "vxl_pass_spot", "vxl_pass_reflection", "vxl_trace_rays",
    "vxl_lighting_host", "vxl_lighting_host_packed", "vxl_lighting", "vxl_volume_gen_terrain", "vxl_gbuffer_primary",
    "vxl_debug_set_variant", "vxl_debug_fetched_probes", "vxl_volume_debug_occupancy", "vxl_debug_read_bandwidth",
    "vxl_resolve_ambient", "vxl_resolve_point", "vxl_resolve_spot", "vxl_trace_model_rays", "vxl_gbuffer_models",
    "vxl_light_taa", "vxl_resolve_reflection",
    "vxl_group_create", "vxl_group_handle", "vxl_group_connect", "vxl_group_base", "vxl_group_connect_pointers", "vxl_group_stack",
    "vxl_group_begin_frame", "vxl_group_fence", "vxl_group_end_frame", "vxl_group_status", "vxl_group_destroy",
    "vxl_asset_guid", "vxl_vox_file_read", "vxl_model_load_v", "vxl_pallete_file_read", "vxl_prefab_file_read", "vxl_scene_load",
    "vxl_vox_import", "vxl_vox_import_memory", "vxl_vox_scene_counts", "vxl_vox_scene_entities", "vxl_vox_scene_model",
    "vxl_vox_scene_pallete", "vxl_vox_scene_write", "vxl_vox_scene_free",
    "vxl_ipc_export", "vxl_ipc_open", "vxl_ipc_close", "vxl_ctx_set_output_mirrors", "vxl_ctx_set_light_plane_stride",
]

VXL_MAX_LIGHTS = 64
TRACE_SPARSE, TRACE_SUPERSPARSE, TRACE_DDA = 0, 1, 2


class VxlError(RuntimeError):
    pass


class Frame(C.Structure):
    """vxl_frame"""
    _fields_ = [("width", C.c_int32), ("height", C.c_int32), ("tile_w", C.c_int32), ("tile_h", C.c_int32),
                ("tile_first", C.c_int32), ("tile_stride", C.c_int32), ("n_tiles", C.c_int32), ("_pad", C.c_int32),
                ("depth24", C.c_void_p), ("normal", C.c_void_p), ("material", C.c_void_p), ("noise", C.c_void_p)]


class Stats(C.Structure):
    _fields_ = [("rays", C.c_uint64), ("steps", C.c_uint64), ("pixels", C.c_uint64)]


class Resolve(C.Structure):
    """vxl_resolve"""
    _fields_ = [("albedo", C.c_void_p), ("depth_full", C.c_void_p)]


class GBufferOut(C.Structure):
    """vxl_gbuffer_out"""
    _fields_ = [("depth24", C.c_void_p), ("normal", C.c_void_p), ("material", C.c_void_p), ("albedo", C.c_void_p), ("motion", C.c_void_p)]


class FullPlanes(C.Structure):
    """vxl_full_planes"""
    _fields_ = [(k, C.c_void_p) for k in ("depth24", "normal", "material", "albedo", "motion", "light", "last_light")]


class LightingHostArgs(C.Structure):
    """vxl_lighting_host_args"""
    _fields_ = [("frame", Frame), ("view", C.c_void_p), ("n_ao", C.c_int32), ("n_point", C.c_int32),
                ("n_spot", C.c_int32), ("point", C.c_void_p), ("spot", C.c_void_p), ("out_shadow", C.c_void_p),
                ("out_ao", C.c_void_p), ("out_point_shadow", C.c_void_p), ("out_spot_shadow", C.c_void_p),
                ("out_spec_t", C.c_void_p)]


class PackedPlanes(C.Structure):
    """vxl_packed_planes"""
    _fields_ = [("shadow_mask", C.c_void_p), ("spec_code", C.c_void_p), ("ao", C.c_void_p)]


_lib = None


def load():
    """dlopen libvxl.so and declare the prototypes.  Raises VxlError if it has not been built."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(LIB_PATH):
        raise VxlError(f"{LIB_PATH} is missing: build it with `python -m voxelengine_b200.build` "
                       "(nvcc, sm_100a). There is no CPU fallback.")
    lib = C.CDLL(LIB_PATH)
    vp, i32, i64, sz = C.c_void_p, C.c_int, C.c_int64, C.c_size_t
    P = C.POINTER
    lib.vxl_abi_version.restype = C.c_int
    lib.vxl_last_error_string.restype = C.c_char_p
    protos = {
        "vxl_ctx_create": [i32, P(vp)], "vxl_ctx_destroy": [vp], "vxl_ctx_set_stream": [vp, vp], "vxl_sync": [vp],
        "vxl_stats_reset": [vp], "vxl_stats_read": [vp, P(Stats)], "vxl_launch_count": [vp, P(C.c_uint64)],
        "vxl_malloc": [vp, sz, P(vp)], "vxl_free": [vp, vp], "vxl_host_alloc": [sz, P(vp)], "vxl_host_free": [vp],
        "vxl_memcpy_h2d": [vp, vp, vp, sz], "vxl_memcpy_d2h": [vp, vp, vp, sz], "vxl_memset": [vp, vp, i32, sz],
        "vxl_volume_create": [vp, i32, i32, i32, P(vp)], "vxl_volume_destroy": [vp],
        "vxl_volume_dims": [vp, P(i32), P(i32), P(i32)], "vxl_volume_upload_regions": [vp, vp, vp, i32],
        "vxl_volume_upload": [vp, vp], "vxl_volume_download": [vp, vp], "vxl_volume_clear": [vp],
        "vxl_volume_device_ptr": [vp, P(vp)], "vxl_volume_mark_dirty": [vp], "vxl_volume_build_occupancy": [vp],
        "vxl_model_create": [vp, vp, i32, i32, i32, P(i32)], "vxl_volume_voxelize": [vp, vp, i32, vp, vp],
        "vxl_pass_ambient": [vp, vp, vp, P(Frame), i32, vp, vp],
        "vxl_pass_point": [vp, vp, vp, P(Frame), vp, i32, vp],
        "vxl_pass_spot": [vp, vp, vp, P(Frame), vp, i32, vp],
        "vxl_pass_reflection": [vp, vp, vp, P(Frame), vp],
        "vxl_resolve_ambient": [vp, vp, P(Frame), P(Resolve), vp, vp, vp],
        "vxl_resolve_point": [vp, vp, P(Frame), P(Resolve), vp, i32, vp, vp],
        "vxl_resolve_spot": [vp, vp, P(Frame), P(Resolve), vp, i32, vp, vp],
        "vxl_trace_rays": [vp, vp, vp, i64, i32, vp],
        "vxl_trace_model_rays": [vp, i32, vp, i64, i32, C.c_float, C.c_float, vp],
        "vxl_gbuffer_models": [vp, vp, P(Frame), vp, i32, vp, vp, P(GBufferOut)],
        "vxl_light_taa": [vp, vp, P(Frame), P(FullPlanes), vp],
        "vxl_asset_guid": [C.c_char_p, P(C.c_uint64)], "vxl_vox_file_read": [C.c_char_p, vp, vp, C.c_uint64],
        "vxl_model_load_v": [vp, C.c_char_p, P(C.c_int)], "vxl_pallete_file_read": [C.c_char_p, vp, vp],
        "vxl_prefab_file_read": [C.c_char_p, vp, i32, P(C.c_int)], "vxl_scene_load": [C.c_char_p, C.c_char_p, vp, i32, P(C.c_int)],
        "vxl_ipc_export": [vp, vp, vp], "vxl_ipc_open": [vp, vp, P(vp)], "vxl_ipc_close": [vp, vp],
        "vxl_ctx_set_output_mirrors": [vp, i32, vp], "vxl_ctx_set_light_plane_stride": [vp, C.c_uint64],
        "vxl_group_create": [vp, i32, i32, sz, i32, P(vp)], "vxl_group_handle": [vp, vp], "vxl_group_connect": [vp, vp],
        "vxl_group_base": [vp, P(vp)], "vxl_group_connect_pointers": [vp, vp], "vxl_group_stack": [vp, i32, P(vp)],
        "vxl_group_begin_frame": [vp, C.c_uint64, P(vp)], "vxl_group_fence": [vp], "vxl_group_end_frame": [vp],
        "vxl_group_status": [vp, P(C.c_int)], "vxl_group_destroy": [vp],
        "vxl_vox_import": [C.c_char_p, P(vp)], "vxl_vox_import_memory": [vp, C.c_uint64, P(vp)],
        "vxl_vox_scene_counts": [vp, P(C.c_int), P(C.c_int)], "vxl_vox_scene_entities": [vp, vp, i32],
        "vxl_vox_scene_model": [vp, i32, vp, vp, vp, C.c_uint64], "vxl_vox_scene_pallete": [vp, vp],
        "vxl_vox_scene_write": [vp, C.c_char_p, C.c_char_p, C.c_char_p], "vxl_vox_scene_free": [vp],
        "vxl_resolve_reflection": [vp, vp, P(Frame), vp, vp, vp, vp, vp],
        "vxl_lighting_host": [vp, vp, P(LightingHostArgs)], "vxl_lighting": [vp, vp, P(LightingHostArgs)],
        "vxl_lighting_host_packed": [vp, vp, P(LightingHostArgs), P(PackedPlanes)],
        "vxl_volume_gen_terrain": [vp], "vxl_gbuffer_primary": [vp, vp, vp, P(Frame)],
        "vxl_debug_set_variant": [vp, i32], "vxl_debug_fetched_probes": [vp, P(C.c_uint64)],
        "vxl_volume_debug_occupancy": [vp, i32, vp, vp], "vxl_debug_read_bandwidth": [vp, sz, i32, P(C.c_double)],
    }
    for name, argtypes in protos.items():
        fn = getattr(lib, name)
        fn.argtypes = argtypes
        fn.restype = C.c_int
    _lib = lib
    return lib


def check(rc: int, what: str = "") -> None:
    if rc != 0:
        msg = load().vxl_last_error_string().decode("utf-8", "replace")
        raise VxlError(f"{what or 'vxl call'} failed ({rc}): {msg}")
