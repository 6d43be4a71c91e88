"""Host-side mirror of the reference's operator interface for the voxel-lighting path, over the
C ABI of include/vxl.h.

Names and argument meaning follow the reference (paths relative to /root/reference):
    ShadowVoxSystem            Sources/World/Systems/ShadowVoxSystem.h:9-38
    LightAmbientPipeline.Use   Sources/Graphics/Pipelines/LightAmbientPipeline.h:35-52
    LightPointPipeline.Use     Sources/Graphics/Pipelines/LightPointPipeline.h:58-100   (+ DrawLight)
    LightSpotPipeline.Use      Sources/Graphics/Pipelines/LightSpotPipeline.h:60-104    (+ DrawLight)
    LightReflectionPipeline.Use Sources/Graphics/Pipelines/LightReflectionPipeline.h:34-51
PyTorch supplies device memory, the stream and (in bench.py) torch.distributed; all arithmetic is
in libvxl.so.  There is no fallback: constructing a Context without the library or a GPU raises.
"""
from __future__ import annotations

import ctypes as C

import numpy as np

from . import capi
from .capi import check
from .tiles import TileLayout
from .scenes import (ENTITY_DTYPE, HIT_DTYPE, POINT_LIGHT_DTYPE, RAY_DTYPE, REGION_DTYPE, SPOT_LIGHT_DTYPE,
                     VIEW_DTYPE)


def _torch():
    import torch
    return torch


def _np_ptr(a: np.ndarray):
    return a.ctypes.data_as(C.c_void_p)


def _dev_ptr(t):
    return C.c_void_p(t.data_ptr()) if t is not None else None


class Context:
    """One GPU + one stream (vxl_ctx).  By default it launches on torch's current stream of that
    device so torch.cuda.Event timing and torch tensors are ordered with the kernels."""

    def __init__(self, device: int = 0, use_torch_stream: bool = True):
        self.lib = capi.load()
        torch = _torch()
        if not torch.cuda.is_available():
            raise capi.VxlError("no CUDA device: voxelengine_b200 has no CPU fallback")
        self.device = int(device)
        h = C.c_void_p()
        check(self.lib.vxl_ctx_create(self.device, C.byref(h)), "vxl_ctx_create")
        self.h = h
        self.torch_device = torch.device("cuda", self.device)
        if use_torch_stream:
            s = torch.cuda.current_stream(self.torch_device).cuda_stream
            check(self.lib.vxl_ctx_set_stream(self.h, C.c_void_p(s)), "vxl_ctx_set_stream")

    def sync(self):
        check(self.lib.vxl_sync(self.h), "vxl_sync")

    def stats_reset(self):
        check(self.lib.vxl_stats_reset(self.h), "vxl_stats_reset")

    def stats(self) -> dict:
        s = capi.Stats()
        check(self.lib.vxl_stats_read(self.h, C.byref(s)), "vxl_stats_read")
        return dict(rays=int(s.rays), steps=int(s.steps), pixels=int(s.pixels))

    def set_variant(self, variant: int):
        """Diagnostics: 0 = plain march on the volume bytes, 1 = occupancy-bit tile march (default), 2 = tile march
        that also counts volume reads (fetched_probes). Same results."""
        check(self.lib.vxl_debug_set_variant(self.h, int(variant)), "vxl_debug_set_variant")

    def fetched_probes(self) -> int:
        n = C.c_uint64()
        check(self.lib.vxl_debug_fetched_probes(self.h, C.byref(n)), "vxl_debug_fetched_probes")
        return int(n.value)

    def read_bandwidth(self, nbytes: int, reps: int = 20) -> float:
        """GB/s of 16-byte reads over an nbytes device buffer from every SM (vxl_debug_read_bandwidth): a buffer well below
        the L2 size measures L2 read bandwidth, one of several GB HBM."""
        g = C.c_double()
        check(self.lib.vxl_debug_read_bandwidth(self.h, int(nbytes), int(reps), C.byref(g)), "vxl_debug_read_bandwidth")
        return float(g.value)

    def launch_count(self) -> int:
        n = C.c_uint64()
        check(self.lib.vxl_launch_count(self.h, C.byref(n)), "vxl_launch_count")
        return int(n.value)

    # -- raw device memory and peer mappings (the fused output-tile gather, tiles.PeerStack) --
    def malloc(self, nbytes: int) -> int:
        p = C.c_void_p()
        check(self.lib.vxl_malloc(self.h, int(nbytes), C.byref(p)), "vxl_malloc")
        return int(p.value)

    def free(self, ptr: int):
        check(self.lib.vxl_free(self.h, C.c_void_p(ptr)), "vxl_free")

    def memset(self, ptr: int, value: int, nbytes: int):
        check(self.lib.vxl_memset(self.h, C.c_void_p(ptr), int(value), int(nbytes)), "vxl_memset")

    def ipc_export(self, ptr: int) -> bytes:
        h = (C.c_ubyte * 64)()
        check(self.lib.vxl_ipc_export(self.h, C.c_void_p(ptr), C.byref(h)), "vxl_ipc_export")
        return bytes(h)

    def ipc_open(self, handle: bytes) -> int:
        h = (C.c_ubyte * 64).from_buffer_copy(handle)
        p = C.c_void_p()
        check(self.lib.vxl_ipc_open(self.h, C.byref(h), C.byref(p)), "vxl_ipc_open")
        return int(p.value)

    def ipc_close(self, ptr: int):
        check(self.lib.vxl_ipc_close(self.h, C.c_void_p(ptr)), "vxl_ipc_close")

    def set_output_mirrors(self, byte_deltas):
        """Every output store of the light-pass kernels is repeated at address + delta (vxl_ctx_set_output_mirrors); [] = off."""
        d = np.ascontiguousarray(byte_deltas, dtype=np.int64)
        check(self.lib.vxl_ctx_set_output_mirrors(self.h, len(d), _np_ptr(d) if len(d) else None), "vxl_ctx_set_output_mirrors")

    def group_create(self, rank: int, world: int, stack_bytes: int, n_stacks: int = 2):
        """This rank's member of a group of `world` processes, one per GPU (vxl_group_create): the gathered stack(s) + arrival flags."""
        return Group(self, rank, world, stack_bytes, n_stacks)

    def set_light_plane_stride(self, pixels: int):
        check(self.lib.vxl_ctx_set_light_plane_stride(self.h, int(pixels)), "vxl_ctx_set_light_plane_stride")

    def tensor_view(self, ptr: int, shape, dtype="float32"):
        """A torch tensor over raw device memory of this context's GPU (no copy; the caller keeps the memory alive)."""
        torch = _torch()

        class _Raw:
            pass
        raw = _Raw()
        raw.__cuda_array_interface__ = {"shape": tuple(int(v) for v in shape), "typestr": np.dtype(dtype).str, "data": (int(ptr), False), "version": 2}
        return torch.as_tensor(raw, device=self.torch_device)

    def close(self):
        if getattr(self, "h", None):
            self.lib.vxl_ctx_destroy(self.h)
            self.h = None

    def empty(self, shape, dtype):
        return _torch().empty(shape, dtype=dtype, device=self.torch_device)


class Group:
    """vxl_group: the frame sharded over the GPUs of one box -- output mirrors into every member's stack, a frame fence by peer-written
    arrival flags (no collective).  The handles travel between the processes by whatever the host has (here: torch.distributed)."""

    def __init__(self, ctx, rank, world, stack_bytes, n_stacks=2):
        self.ctx, self.lib, self.rank, self.world, self.n_stacks = ctx, ctx.lib, int(rank), int(world), int(n_stacks)
        h = C.c_void_p()
        check(self.lib.vxl_group_create(ctx.h, self.rank, self.world, int(stack_bytes), self.n_stacks, C.byref(h)), "vxl_group_create")
        self.h = h

    def handle(self) -> bytes:
        b = (C.c_ubyte * 64)()
        check(self.lib.vxl_group_handle(self.h, C.byref(b)), "vxl_group_handle")
        return bytes(b)

    def connect(self, handles):
        buf = (C.c_ubyte * (64 * self.world))()
        for r, hd in enumerate(handles):
            buf[64 * r:64 * r + 64] = (hd or b"").ljust(64, b"\0")[:64]
        check(self.lib.vxl_group_connect(self.h, C.byref(buf)), "vxl_group_connect")

    def base(self) -> int:
        p = C.c_void_p()
        check(self.lib.vxl_group_base(self.h, C.byref(p)), "vxl_group_base")
        return int(p.value)

    def connect_pointers(self, bases):
        arr = (C.c_void_p * self.world)(*[C.c_void_p(int(b)) for b in bases])
        check(self.lib.vxl_group_connect_pointers(self.h, C.byref(arr)), "vxl_group_connect_pointers")

    def stack(self, which: int) -> int:
        p = C.c_void_p()
        check(self.lib.vxl_group_stack(self.h, int(which), C.byref(p)), "vxl_group_stack")
        return int(p.value)

    def begin_frame(self, frame: int) -> int:
        p = C.c_void_p()
        check(self.lib.vxl_group_begin_frame(self.h, int(frame), C.byref(p)), "vxl_group_begin_frame")
        return int(p.value)

    def fence(self):
        check(self.lib.vxl_group_fence(self.h), "vxl_group_fence")

    def end_frame(self):
        check(self.lib.vxl_group_end_frame(self.h), "vxl_group_end_frame")

    def status(self):
        v = C.c_int(0)
        check(self.lib.vxl_group_status(self.h, C.byref(v)), "vxl_group_status")
        return int(v.value)

    def destroy(self):
        if getattr(self, "h", None):
            self.lib.vxl_group_destroy(self.h)
            self.h = None


class ShadowVoxSystem:
    """World occupancy volume.  Default size is the reference's 524 x 188 x 524 texels
    (ShadowVoxSystem.cpp:56-58); here it is a runtime parameter."""

    def __init__(self, ctx: Context, size=(524, 188, 524)):
        self.ctx, self.lib = ctx, ctx.lib
        self.sx, self.sy, self.sz = (int(v) for v in size)
        h = C.c_void_p()
        check(self.lib.vxl_volume_create(ctx.h, self.sx, self.sy, self.sz, C.byref(h)), "vxl_volume_create")
        self.h = h
        self._models = []

    # --- reference-named surface -------------------------------------------------------------
    def GetVolumeImage(self):
        return self

    def OnUpdate(self, entities: np.ndarray, want_regions: bool = True):
        """ShadowVoxSystem::OnUpdate for the given visited entities (array of scenes.ENTITY_DTYPE,
        in view order).  Returns (regions, valid) like the reference's _UpdateRegions pushes."""
        entities = np.ascontiguousarray(entities, dtype=ENTITY_DTYPE)
        n = len(entities)
        if want_regions:
            regions = np.zeros(n, dtype=REGION_DTYPE)
            valid = np.zeros(n, dtype=np.int32)
            check(self.lib.vxl_volume_voxelize(self.h, _np_ptr(entities), n, _np_ptr(regions), _np_ptr(valid)),
                  "vxl_volume_voxelize")
            return regions, valid
        check(self.lib.vxl_volume_voxelize(self.h, _np_ptr(entities), n, None, None), "vxl_volume_voxelize")
        return None, None

    # --- data movement -----------------------------------------------------------------------
    def add_model(self, voxels: np.ndarray) -> int:
        """voxels: uint8 (sz, sy, sx) palette indices (x fastest), VoxAsset layout."""
        voxels = np.ascontiguousarray(voxels, dtype=np.uint8)
        msz, msy, msx = voxels.shape
        mid = C.c_int()
        check(self.lib.vxl_model_create(self.ctx.h, _np_ptr(voxels), msx, msy, msz, C.byref(mid)), "vxl_model_create")
        return int(mid.value)

    def upload(self, host: np.ndarray):
        host = np.ascontiguousarray(host, dtype=np.uint8)
        assert host.shape == (self.sz, self.sy, self.sx)
        check(self.lib.vxl_volume_upload(self.h, _np_ptr(host)), "vxl_volume_upload")
        self.ctx.sync()

    def upload_regions(self, staging: np.ndarray, regions: np.ndarray):
        """CmdBuffer::copy(_Buffer, _Volume, _UpdateRegions) (ShadowVoxSystem.cpp:196)."""
        staging = np.ascontiguousarray(staging, dtype=np.uint8)
        assert staging.shape == (self.sz, self.sy, self.sx)
        regions = np.ascontiguousarray(regions, dtype=REGION_DTYPE)
        check(self.lib.vxl_volume_upload_regions(self.h, _np_ptr(staging), _np_ptr(regions), len(regions)),
              "vxl_volume_upload_regions")
        self.ctx.sync()

    def download(self) -> np.ndarray:
        out = np.empty((self.sz, self.sy, self.sx), np.uint8)
        check(self.lib.vxl_volume_download(self.h, _np_ptr(out)), "vxl_volume_download")
        return out

    def clear(self):
        check(self.lib.vxl_volume_clear(self.h), "vxl_volume_clear")

    def gen_terrain(self):
        check(self.lib.vxl_volume_gen_terrain(self.h), "vxl_volume_gen_terrain")

    def build_occupancy(self):
        check(self.lib.vxl_volume_build_occupancy(self.h), "vxl_volume_build_occupancy")

    def mark_dirty(self) -> None:
        """The next pass rebuilds every occupancy level in full (vxl_volume_mark_dirty)."""
        check(self.lib.vxl_volume_mark_dirty(self.h), "vxl_volume_mark_dirty")

    def occupancy(self, shift: int) -> np.ndarray:
        """Diagnostics: occupancy level as 0/1 uint8 [cz][cy][cx]; 2, 3, 4 = plain (cell = 2^level voxels),
        1 = texel level, 13, 14 = 3x3x3-dilated levels 3, 4 including their 1-cell border, 22 = level 2 decoded from its shifted copy."""
        dims = np.zeros(3, np.int32)
        check(self.lib.vxl_volume_debug_occupancy(self.h, int(shift), None, _np_ptr(dims)), "vxl_volume_debug_occupancy")
        out = np.zeros((dims[2], dims[1], dims[0]), np.uint8)
        check(self.lib.vxl_volume_debug_occupancy(self.h, int(shift), _np_ptr(out), _np_ptr(dims)), "vxl_volume_debug_occupancy")
        return out

    def trace_rays(self, rays: np.ndarray, variant: int) -> np.ndarray:
        """Ray-level entry: host rays in, host hit records out (copies through device buffers)."""
        torch = _torch()
        rays = np.ascontiguousarray(rays, dtype=RAY_DTYPE)
        n = len(rays)
        out = np.zeros(n, dtype=HIT_DTYPE)
        if n == 0:
            return out
        d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(self.ctx.torch_device)
        d_out = torch.empty(n * HIT_DTYPE.itemsize, dtype=torch.uint8, device=self.ctx.torch_device)
        check(self.lib.vxl_trace_rays(self.ctx.h, self.h, _dev_ptr(d_rays), n, int(variant), _dev_ptr(d_out)), "vxl_trace_rays")
        self.ctx.sync()
        return d_out.cpu().numpy().view(HIT_DTYPE).copy()

    def close(self):
        if getattr(self, "h", None):
            self.lib.vxl_volume_destroy(self.h)
            self.h = None


class GeometryBuffer:
    """The G-buffer attachments the light passes read (Sources/Graphics/Graphics.h:51-60) plus the
    blue-noise image, resident in HBM as torch int32 tensors in tile-compact layout.

    A whole frame on one GPU is one tile (tile_w, tile_h = width, height).  For screen-tile
    sharding, rank r of n holds tiles r, r+n, r+2n, ... of the tile grid."""

    def __init__(self, ctx: Context, width: int, height: int, tile_w: int | None = None, tile_h: int | None = None,
                 rank: int = 0, world: int = 1):
        torch = _torch()
        self.ctx = ctx
        self.layout = L = TileLayout(width, height, tile_w, tile_h, rank, world)
        self.width, self.height, self.tile_w, self.tile_h = L.width, L.height, L.tile_w, L.tile_h
        self.tiles_x, self.tiles_y = L.tiles_x, L.tiles_y
        self.tile_first, self.tile_stride, self.n_tiles = L.tile_first, L.tile_stride, L.n_tiles
        shape = (self.n_tiles, self.tile_h, self.tile_w)
        self.depth24 = torch.full(shape, 0xFFFFFF, dtype=torch.int32, device=ctx.torch_device)
        self.normal = torch.zeros(shape, dtype=torch.int32, device=ctx.torch_device)
        self.material = torch.zeros(shape, dtype=torch.int32, device=ctx.torch_device)
        self.noise = torch.zeros((512, 512), dtype=torch.int32, device=ctx.torch_device)

    @property
    def shape(self):
        return (self.n_tiles, self.tile_h, self.tile_w)

    def tile_ids(self):
        return self.layout.ids()

    def frame(self) -> capi.Frame:
        return capi.Frame(self.width, self.height, self.tile_w, self.tile_h, self.tile_first, self.tile_stride,
                          self.n_tiles, 0, self.depth24.data_ptr(), self.normal.data_ptr(), self.material.data_ptr(),
                          self.noise.data_ptr())

    def set_noise(self, noise: np.ndarray):
        torch = _torch()
        self.noise.copy_(torch.from_numpy(np.ascontiguousarray(noise, dtype=np.uint32).view(np.int32)))

    def set_planes(self, depth24: np.ndarray, normal: np.ndarray, material: np.ndarray):
        """Upload full-frame (H, W) uint32 planes, extracting this shard's tiles."""
        torch = _torch()
        for name, a in (("depth24", depth24), ("normal", normal), ("material", material)):
            t = self.to_tiles(np.ascontiguousarray(a, dtype=np.uint32))
            getattr(self, name).copy_(torch.from_numpy(t.view(np.int32)))

    def to_tiles(self, plane: np.ndarray) -> np.ndarray:
        """(H, W) -> this shard's (n_tiles, tile_h, tile_w), zero padded at the frame edge."""
        return self.layout.to_tiles(plane)

    def from_tiles(self, tiles: np.ndarray, out: np.ndarray):
        """Scatter this shard's (n_tiles, tile_h, tile_w) back into a full (H, W) plane."""
        return self.layout.from_tiles(tiles, out)

    def synthesize(self, volume: ShadowVoxSystem, view: np.ndarray):
        """Fill depth/normal/material with the synthetic primary-visibility G-buffer."""
        f = self.frame()
        v = np.ascontiguousarray(view, dtype=VIEW_DTYPE).reshape(())
        check(self.ctx.lib.vxl_gbuffer_primary(self.ctx.h, volume.h, _np_ptr(v), C.byref(f)), "vxl_gbuffer_primary")


def _view_ptr(view):
    v = np.ascontiguousarray(view, dtype=VIEW_DTYPE).reshape(())
    return v, _np_ptr(v)


class LightAmbientPipeline:
    """Sun shadow + ambient occlusion.  Use() keeps the reference argument order minus the Vulkan
    command buffer and the sky box (only sampled for sky pixels' colour)."""
    _inst = None

    @classmethod
    def Get(cls):
        cls._inst = cls._inst or cls()
        return cls._inst

    def Use(self, viewBuffer, geometryFB: GeometryBuffer, shadowVox: ShadowVoxSystem, n_ao: int = 1,
            out_shadow=None, out_ao=None):
        torch = _torch()
        ctx = geometryFB.ctx
        if out_shadow is None:
            out_shadow = ctx.empty(geometryFB.shape, torch.float32)
        if out_ao is None:
            out_ao = ctx.empty(geometryFB.shape, torch.float32)
        v, vp = _view_ptr(viewBuffer)
        f = geometryFB.frame()
        check(ctx.lib.vxl_pass_ambient(ctx.h, shadowVox.h, vp, C.byref(f), int(n_ao), _dev_ptr(out_shadow), _dev_ptr(out_ao)),
              "vxl_pass_ambient")
        return out_shadow, out_ao


class _LocalLightPipeline:
    MAX_LIGHTS = capi.VXL_MAX_LIGHTS
    _dtype = None
    _fn = None

    def __init__(self):
        self._data = np.zeros(self.MAX_LIGHTS, dtype=self._dtype)
        self._current = 0
        self.warnings = 0

    def _reset(self):
        self._current = 0

    def _use(self, viewBuffer, geometryFB, shadowVox, cb, out_shadow):
        torch = _torch()
        ctx = geometryFB.ctx
        self._reset()                       # _CurrentLightIndex = 0
        if cb is not None:
            cb(self)
        n = self._current
        if out_shadow is None:
            out_shadow = ctx.empty((n,) + geometryFB.shape, torch.float32)
        if n == 0:
            return out_shadow
        v, vp = _view_ptr(viewBuffer)
        f = geometryFB.frame()
        fn = getattr(ctx.lib, self._fn)
        check(fn(ctx.h, shadowVox.h, vp, C.byref(f), _np_ptr(self._data), n, _dev_ptr(out_shadow)), self._fn)
        return out_shadow


class LightPointPipeline(_LocalLightPipeline):
    _dtype = POINT_LIGHT_DTYPE
    _fn = "vxl_pass_point"
    _inst = None

    @classmethod
    def Get(cls):
        cls._inst = cls._inst or cls()
        return cls._inst

    def DrawLight(self, position, range_, color, attenuation):
        # LightPointPipeline.h:58-72: beyond 64 lights the reference warns and drops the light
        if self._current >= self.MAX_LIGHTS:
            self.warnings += 1
            return
        l = self._data[self._current]
        l["Position"], l["Range"], l["Color"], l["Attenuation"] = position, range_, color, attenuation
        self._current += 1

    def Use(self, viewBuffer, geometryFB, shadowVox, cb, out_shadow=None):
        return self._use(viewBuffer, geometryFB, shadowVox, cb, out_shadow)


class LightSpotPipeline(_LocalLightPipeline):
    _dtype = SPOT_LIGHT_DTYPE
    _fn = "vxl_pass_spot"
    _inst = None

    @classmethod
    def Get(cls):
        cls._inst = cls._inst or cls()
        return cls._inst

    def DrawLight(self, position, range_, color, attenuation, direction, angle, angleAttenuation):
        if self._current >= self.MAX_LIGHTS:
            self.warnings += 1
            return
        l = self._data[self._current]
        l["Position"], l["Range"], l["Color"], l["Attenuation"] = position, range_, color, attenuation
        l["Direction"], l["Angle"], l["AngleAttenuation"] = direction, angle, angleAttenuation
        self._current += 1

    def Use(self, viewBuffer, geometryFB, shadowVox, cb, out_shadow=None):
        return self._use(viewBuffer, geometryFB, shadowVox, cb, out_shadow)


class LightReflectionPipeline:
    _inst = None

    @classmethod
    def Get(cls):
        cls._inst = cls._inst or cls()
        return cls._inst

    def Use(self, viewBuffer, geometryFB: GeometryBuffer, shadowVox: ShadowVoxSystem, out_spec_t=None):
        torch = _torch()
        ctx = geometryFB.ctx
        if out_spec_t is None:
            out_spec_t = ctx.empty(geometryFB.shape, torch.float32)
        v, vp = _view_ptr(viewBuffer)
        f = geometryFB.frame()
        check(ctx.lib.vxl_pass_reflection(ctx.h, shadowVox.h, vp, C.byref(f), _dev_ptr(out_spec_t)), "vxl_pass_reflection")
        return out_spec_t

    def Colour(self, viewBuffer, geometryFB: GeometryBuffer, spec_t, full, light_full, sky=(0.0, 0.0, 0.0), out=None):
        """LightReflection.frag's out_Color (:60-139) from the march's t plane: the TAA light buffer (`light_full`, torch float32
        (H, W, 4) or None) where the reflected ray ends on the visible surface, `sky` on a miss.  full: FullFrame (its depth)."""
        torch = _torch()
        ctx = geometryFB.ctx
        if out is None:
            out = torch.zeros(geometryFB.shape + (4,), dtype=torch.float32, device=ctx.torch_device)
        v, vp = _view_ptr(viewBuffer)
        f = geometryFB.frame()
        sk = np.ascontiguousarray(sky, dtype=np.float32)
        check(ctx.lib.vxl_resolve_reflection(ctx.h, vp, C.byref(f), _dev_ptr(spec_t), _dev_ptr(full.depth24) if full is not None else None,
                                             _dev_ptr(light_full), _np_ptr(sk), _dev_ptr(out)), "vxl_resolve_reflection")
        return out


class FullFrame:
    """Whole-frame row-major planes in HBM for the passes that sample other pixels (vxl_full_planes): the G-buffer attachments
    (depth, normal, material, colour, motion).  On several GPUs this is the all-gathered frame."""

    def __init__(self, ctx: Context, depth24, normal, material, albedo, motion):
        torch = _torch()
        self.ctx = ctx
        dev = ctx.torch_device

        def u32(a):
            return a if hasattr(a, "data_ptr") else torch.from_numpy(np.ascontiguousarray(a, dtype=np.uint32).view(np.int32)).to(dev)

        def f32(a):
            return a if hasattr(a, "data_ptr") else torch.from_numpy(np.ascontiguousarray(a, dtype=np.float32)).to(dev)
        self.depth24, self.normal, self.material, self.albedo = u32(depth24), u32(normal), u32(material), u32(albedo)
        self.motion = f32(motion)
        self.height, self.width = self.depth24.shape[-2:]
        assert tuple(self.motion.shape[-3:]) == (self.height, self.width, 2)


class LightTAAPipeline:
    """Temporal + spatial accumulation of the light buffer (Sources/Graphics/Pipelines/LightTAAPipeline.h:34-53, LightTAA.frag;
    SURVEY 8f row f3).  Use() keeps the reference's argument order minus the command buffer: view, (blue noise: the frame's),
    last TAA light buffer, current light buffer, G-buffer."""
    _inst = None

    @classmethod
    def Get(cls):
        cls._inst = cls._inst or cls()
        return cls._inst

    def Use(self, viewBuffer, geometryFB: GeometryBuffer, full: FullFrame, light_full, last_light_full, out=None):
        """light_full / last_light_full: torch float32 (H, W, 4) whole-frame planes.  Returns the shard's TAA light (tile-compact)."""
        torch = _torch()
        ctx = geometryFB.ctx
        assert tuple(light_full.shape) == (full.height, full.width, 4) and tuple(last_light_full.shape) == (full.height, full.width, 4)
        assert (full.width, full.height) == (geometryFB.width, geometryFB.height)
        if out is None:
            out = torch.zeros(geometryFB.shape + (4,), dtype=torch.float32, device=ctx.torch_device)
        v, vp = _view_ptr(viewBuffer)
        f = geometryFB.frame()
        fp = capi.FullPlanes(full.depth24.data_ptr(), full.normal.data_ptr(), full.material.data_ptr(), full.albedo.data_ptr(),
                             full.motion.data_ptr(), light_full.data_ptr(), last_light_full.data_ptr())
        check(ctx.lib.vxl_light_taa(ctx.h, vp, C.byref(f), C.byref(fp), _dev_ptr(out)), "vxl_light_taa")
        return out


def trace_model_rays(ctx: Context, model_id: int, rays: np.ndarray, frame: int = 0, res=(1920.0, 1080.0)) -> np.ndarray:
    """GeometryVoxel.frag's clipToAABB + intersectVolume on a registered model (core): host rays
    (scenes.MODEL_RAY_DTYPE) in, host hit records (scenes.MODEL_HIT_DTYPE) out, through device buffers."""
    from .scenes import MODEL_HIT_DTYPE, MODEL_RAY_DTYPE
    torch = _torch()
    rays = np.ascontiguousarray(rays, dtype=MODEL_RAY_DTYPE)
    n = len(rays)
    if n == 0:
        return np.zeros(0, dtype=MODEL_HIT_DTYPE)
    d_rays = torch.from_numpy(rays.view(np.uint8).reshape(-1)).to(ctx.torch_device)
    d_out = torch.empty(n * MODEL_HIT_DTYPE.itemsize, dtype=torch.uint8, device=ctx.torch_device)
    check(ctx.lib.vxl_trace_model_rays(ctx.h, int(model_id), _dev_ptr(d_rays), n, int(frame), float(res[0]), float(res[1]), _dev_ptr(d_out)),
          "vxl_trace_model_rays")
    ctx.sync()
    return d_out.cpu().numpy().view(MODEL_HIT_DTYPE).copy()


class GeometryVoxelPipeline:
    """The geometry pass over a draw list (Sources/Graphics/Pipelines/GeometryVoxelPipeline.h:49-71): fills a GeometryBuffer's
    depth / normal / material planes plus an albedo plane (and optionally motion vectors) from model instances."""
    _inst = None

    @classmethod
    def Get(cls):
        cls._inst = cls._inst or cls()
        return cls._inst

    def Use(self, viewBuffer, geometryFB: GeometryBuffer, cmds: np.ndarray, pal_color, pal_material, albedo=None, motion=None):
        """cmds: scenes.VOX_CMD_DTYPE (model = id from ShadowVoxSystem.add_model).  pal_*: torch int32 [n_palettes][256] RGBA8."""
        from .scenes import VOX_CMD_DTYPE
        torch = _torch()
        ctx = geometryFB.ctx
        cmds = np.ascontiguousarray(cmds, dtype=VOX_CMD_DTYPE)
        if albedo is None:
            albedo = torch.zeros(geometryFB.shape, dtype=torch.int32, device=ctx.torch_device)
        v, vp = _view_ptr(viewBuffer)
        f = geometryFB.frame()
        o = capi.GBufferOut(geometryFB.depth24.data_ptr(), geometryFB.normal.data_ptr(), geometryFB.material.data_ptr(), albedo.data_ptr(),
                            motion.data_ptr() if motion is not None else None)
        check(ctx.lib.vxl_gbuffer_models(ctx.h, vp, C.byref(f), _np_ptr(cmds), len(cmds), _dev_ptr(pal_color), _dev_ptr(pal_material), C.byref(o)),
              "vxl_gbuffer_models")
        return albedo


class LightBuffer:
    """The light buffer the reference's light passes blend into (WorldRenderer.cpp:242-258), as float32 RGBA in
    tile-compact layout: what LightAmbient / LightPoint / LightSpot.frag compute after the march.
    `albedo`: torch int32 tensor shaped like the frame's planes (COLOR_TEXTURE, RGBA8 UNORM).  `depth_full`: the whole
    frame's depth plane (H, W) for tile-sharded frames (the screen-space occlusion samples other pixels)."""

    def __init__(self, geometryFB: GeometryBuffer, albedo, depth_full=None):
        torch = _torch()
        self.gb, self.ctx = geometryFB, geometryFB.ctx
        self.albedo, self.depth_full = albedo, depth_full
        self.rgba = torch.zeros(geometryFB.shape + (4,), dtype=torch.float32, device=self.ctx.torch_device)
        self._r = capi.Resolve(albedo.data_ptr(), depth_full.data_ptr() if depth_full is not None else None)

    def Ambient(self, viewBuffer, shadow, ao):
        v, vp = _view_ptr(viewBuffer)
        f = self.gb.frame()
        check(self.ctx.lib.vxl_resolve_ambient(self.ctx.h, vp, C.byref(f), C.byref(self._r), _dev_ptr(shadow), _dev_ptr(ao), _dev_ptr(self.rgba)),
              "vxl_resolve_ambient")
        return self.rgba

    def _local(self, fn, dtype, viewBuffer, lights, shadow):
        lights = np.ascontiguousarray(lights, dtype=dtype)
        v, vp = _view_ptr(viewBuffer)
        f = self.gb.frame()
        check(getattr(self.ctx.lib, fn)(self.ctx.h, vp, C.byref(f), C.byref(self._r), _np_ptr(lights), len(lights), _dev_ptr(shadow), _dev_ptr(self.rgba)), fn)
        return self.rgba

    def Point(self, viewBuffer, lights, shadow):
        return self._local("vxl_resolve_point", POINT_LIGHT_DTYPE, viewBuffer, lights, shadow)

    def Spot(self, viewBuffer, lights, shadow):
        return self._local("vxl_resolve_spot", SPOT_LIGHT_DTYPE, viewBuffer, lights, shadow)


def lighting_host(ctx: Context, shadowVox: ShadowVoxSystem, view, frame_desc: dict, planes: dict, outs: dict,
                  n_ao: int = 1, point=None, spot=None):
    """vxl_lighting_host: HOST (pinned torch / numpy) G-buffer planes in, HOST output planes out.
    frame_desc: width,height,tile_w,tile_h,tile_first,tile_stride,n_tiles.  planes: depth24, normal,
    material, noise (host arrays with .ctypes or torch CPU tensors).  outs: any of shadow, ao,
    point_shadow, spot_shadow, spec_t."""
    def hp(x):
        if x is None:
            return None
        if hasattr(x, "data_ptr"):
            return x.data_ptr()
        return x.ctypes.data
    v, vp = _view_ptr(view)
    fr = capi.Frame(frame_desc["width"], frame_desc["height"], frame_desc["tile_w"], frame_desc["tile_h"],
                    frame_desc["tile_first"], frame_desc["tile_stride"], frame_desc["n_tiles"], 0,
                    hp(planes["depth24"]), hp(planes["normal"]), hp(planes.get("material")), hp(planes["noise"]))
    pt = np.ascontiguousarray(point, dtype=POINT_LIGHT_DTYPE) if point is not None else None
    sp = np.ascontiguousarray(spot, dtype=SPOT_LIGHT_DTYPE) if spot is not None else None
    a = capi.LightingHostArgs(fr, vp, int(n_ao), 0 if pt is None else len(pt), 0 if sp is None else len(sp),
                              None if pt is None else pt.ctypes.data, None if sp is None else sp.ctypes.data,
                              hp(outs.get("shadow")), hp(outs.get("ao")), hp(outs.get("point_shadow")),
                              hp(outs.get("spot_shadow")), hp(outs.get("spec_t")))
    if "packed" in outs:                      # vxl_lighting_host_packed: outs["packed"] = dict(shadow_mask=, spec_code=, ao=)
        pk = outs["packed"]
        pp = capi.PackedPlanes(hp(pk.get("shadow_mask")), hp(pk.get("spec_code")), hp(pk.get("ao")))
        check(ctx.lib.vxl_lighting_host_packed(ctx.h, shadowVox.h, C.byref(a), C.byref(pp)), "vxl_lighting_host_packed")
        return
    check(ctx.lib.vxl_lighting_host(ctx.h, shadowVox.h, C.byref(a)), "vxl_lighting_host")


def mask_bytes(n_point: int = 0, n_spot: int = 0) -> int:
    """bytes per pixel of vxl_packed_planes.shadow_mask"""
    return (1 + int(n_point) + int(n_spot) + 7) // 8


def unpack_planes(shadow_mask: np.ndarray, spec_code: np.ndarray, n_point: int = 0, n_spot: int = 0) -> dict:
    """Invert vxl_lighting_host_packed's encoding (include/vxl.h) into the float planes of vxl_lighting_host:
    shadow [px], point_shadow [n_point][px], spot_shadow [n_spot][px], spec_t [px]."""
    out = {}
    if shadow_mask is not None:
        m = np.asarray(shadow_mask, dtype=np.uint8).reshape(-1, mask_bytes(n_point, n_spot))
        plane = lambda p: ((m[:, p >> 3] >> (p & 7)) & 1).astype(np.float32)
        out["shadow"] = plane(0)
        out["point_shadow"] = np.stack([plane(1 + i) for i in range(n_point)]) if n_point else np.zeros((0, m.shape[0]), np.float32)
        out["spot_shadow"] = np.stack([plane(1 + n_point + i) for i in range(n_spot)]) if n_spot else np.zeros((0, m.shape[0]), np.float32)
    if spec_code is not None:
        c = np.asarray(spec_code, dtype=np.uint8).reshape(-1).astype(np.int32)
        if np.any((c > 178) & (c != 255)):
            raise ValueError("spec_code holds a value outside the code set")
        out["spec_t"] = np.where(c < 31, 0.5 * (c + 1), np.where(c == 255, 256.0, 16.0 + (c - 31))).astype(np.float32)
    return out


def lighting(ctx: Context, shadowVox: ShadowVoxSystem, view, geometryFB: "GeometryBuffer", outs: dict, n_ao: int = 1, point=None, spot=None):
    """vxl_lighting: all requested passes of one frame on DEVICE planes, the three kernels on concurrent streams (joined on the
    context's stream).  outs: any of shadow, ao, point_shadow, spot_shadow, spec_t (device tensors in the frame's tile-compact layout)."""
    dp = lambda x: None if x is None else x.data_ptr()
    v, vp = _view_ptr(view)
    fr = geometryFB.frame()
    pt = np.ascontiguousarray(point, dtype=POINT_LIGHT_DTYPE) if point is not None else None
    sp = np.ascontiguousarray(spot, dtype=SPOT_LIGHT_DTYPE) if spot is not None else None
    a = capi.LightingHostArgs(fr, vp, int(n_ao), 0 if pt is None else len(pt), 0 if sp is None else len(sp),
                              None if pt is None else pt.ctypes.data, None if sp is None else sp.ctypes.data,
                              dp(outs.get("shadow")), dp(outs.get("ao")), dp(outs.get("point_shadow")), dp(outs.get("spot_shadow")), dp(outs.get("spec_t")))
    check(ctx.lib.vxl_lighting(ctx.h, shadowVox.h, C.byref(a)), "vxl_lighting")


# ---- on-disk formats: thin wrappers over the host-side readers of libvxl.so ---------------------------------
def asset_guid(path: str) -> int:
    """Assets::Hash (Sources/Asset/Assets.h:207-210): the GUID of an asset path relative to Mods/."""
    lib = capi.load()
    out = C.c_uint64()
    check(lib.vxl_asset_guid(path.encode(), C.byref(out)), "vxl_asset_guid")
    return int(out.value)


def read_vox_file(path: str) -> np.ndarray:
    """A .v volume (VoxAsset::Serialize) -> uint8 (sz, sy, sx) palette indices."""
    lib = capi.load()
    dims = np.zeros(3, np.int32)
    check(lib.vxl_vox_file_read(path.encode(), _np_ptr(dims), None, 0), "vxl_vox_file_read")
    out = np.zeros((int(dims[2]), int(dims[1]), int(dims[0])), np.uint8)
    check(lib.vxl_vox_file_read(path.encode(), _np_ptr(dims), _np_ptr(out), out.size), "vxl_vox_file_read")
    return out


def read_pallete_file(path: str):
    """A .p palette (PalleteAsset::Serialize) -> (colour[256], material[256]) uint32 texels of one palette row."""
    lib = capi.load()
    color, material = np.zeros(256, np.uint32), np.zeros(256, np.uint32)
    check(lib.vxl_pallete_file_read(path.encode(), _np_ptr(color), _np_ptr(material)), "vxl_pallete_file_read")
    return color, material


def read_prefab_file(path: str) -> np.ndarray:
    """A .pf scene (PrefabAsset::Spawn + TransformSystem) -> array of scenes.PREFAB_ENTITY_DTYPE in file order."""
    from .scenes import PREFAB_ENTITY_DTYPE
    lib = capi.load()
    n = C.c_int()
    check(lib.vxl_prefab_file_read(path.encode(), None, 0, C.byref(n)), "vxl_prefab_file_read")
    out = np.zeros(n.value, PREFAB_ENTITY_DTYPE)
    check(lib.vxl_prefab_file_read(path.encode(), _np_ptr(out), n.value, C.byref(n)), "vxl_prefab_file_read")
    return out


def load_scene(mods_dir: str, prefab_path: str) -> np.ndarray:
    """vxl_scene_load: a .pf scene relative to a Mods directory with nested prefab instances expanded."""
    from .scenes import PREFAB_ENTITY_DTYPE
    lib = capi.load()
    n = C.c_int()
    check(lib.vxl_scene_load(mods_dir.encode(), prefab_path.encode(), None, 0, C.byref(n)), "vxl_scene_load")
    out = np.zeros(n.value, PREFAB_ENTITY_DTYPE)
    check(lib.vxl_scene_load(mods_dir.encode(), prefab_path.encode(), _np_ptr(out), n.value, C.byref(n)), "vxl_scene_load")
    return out


VOX_IMPORT_ENTITY_DTYPE = np.dtype([("parent", "<i4"), ("model", "<i4"), ("position", "<f4", 3), ("name", "S64")])


class VoxImporter:
    """Sources/Editor/Importer/VoxImporter.cpp: a MagicaVoxel .vox file -> models (.v), palette (.p), entity tree (.pf)."""

    def __init__(self, source):
        self.lib = capi.load()
        self.h = C.c_void_p()
        if isinstance(source, (bytes, bytearray)):
            buf = (C.c_char * len(source)).from_buffer_copy(bytes(source))
            check(self.lib.vxl_vox_import_memory(C.addressof(buf), len(source), C.byref(self.h)), "vxl_vox_import_memory")
        else:
            check(self.lib.vxl_vox_import(str(source).encode(), C.byref(self.h)), "vxl_vox_import")
        ne, nm = C.c_int(), C.c_int()
        check(self.lib.vxl_vox_scene_counts(self.h, C.byref(ne), C.byref(nm)), "vxl_vox_scene_counts")
        self.n_entities, self.n_models = ne.value, nm.value

    def entities(self) -> np.ndarray:
        out = np.zeros(self.n_entities, VOX_IMPORT_ENTITY_DTYPE)
        check(self.lib.vxl_vox_scene_entities(self.h, _np_ptr(out), len(out)), "vxl_vox_scene_entities")
        return out

    def model(self, i: int):
        """-> (name, uint8 (sz, sy, sx))"""
        dims = np.zeros(3, np.int32)
        name = C.create_string_buffer(64)
        check(self.lib.vxl_vox_scene_model(self.h, i, _np_ptr(dims), C.addressof(name), None, 0), "vxl_vox_scene_model")
        out = np.zeros((int(dims[2]), int(dims[1]), int(dims[0])), np.uint8)
        check(self.lib.vxl_vox_scene_model(self.h, i, _np_ptr(dims), C.addressof(name), _np_ptr(out), out.size), "vxl_vox_scene_model")
        return name.value.decode("latin-1"), out

    def pallete_records(self) -> np.ndarray:
        out = np.zeros((256, 7), np.uint8)
        check(self.lib.vxl_vox_scene_pallete(self.h, _np_ptr(out)), "vxl_vox_scene_pallete")
        return out

    def write(self, mods_dir: str, path: str, file_name: str):
        """VoxImporter::Import's asset files under <mods_dir>/<path>/."""
        check(self.lib.vxl_vox_scene_write(self.h, mods_dir.encode(), path.encode(), file_name.encode()), "vxl_vox_scene_write")

    def close(self):
        if self.h:
            self.lib.vxl_vox_scene_free(self.h)
            self.h = C.c_void_p()
