"""The BASELINE.json configs as concrete device-resident workloads.

"N^3 volume" = N^3 voxels = (N/2)^3 bytes in the reference's 2x2x2-bits-per-byte layout.
A workload owns the volume, the (tile-sharded) G-buffer, the packed output tensor and runs one
"step" = one frame of the light passes over this rank's tiles.
"""
from __future__ import annotations

import numpy as np

from . import engine as E
from . import scenes as S

CONFIGS = {
    1: dict(name="cfg1: 64^3 .vox-style model, 256x256, 1 sun shadow + 4 AO", texels=(32, 32, 32), res=(256, 256),
            scene="house", n_ao=4, n_point=0, spec=False),
    2: dict(name="cfg2: 512^3 FastNoise terrain, 1080p, 1 sun shadow + 8 AO", texels=(256, 256, 256), res=(1920, 1080),
            scene="terrain", n_ao=8, n_point=0, spec=False),
    3: dict(name="cfg3: 1024^3 terrain + 200 props, 4K, sun shadow + 16 AO + 4 point-light shadows + spec occlusion",
            texels=(512, 512, 512), res=(3840, 2160), scene="terrain+props", n_props=200, n_ao=16, n_point=4, spec=True),
    4: dict(name="cfg4: 1024^3, 1000 moving 16^3 entities re-voxelised per frame, 1080p sun shadow + 1 AO",
            texels=(512, 512, 512), res=(1920, 1080), scene="dynamic", n_entities=1000, n_ao=1, n_point=0, spec=False),
    5: dict(name="cfg5: 2048^3 terrain + 200 props, 8K, sun shadow + 16 AO + 4 point-light shadows + spec occlusion",
            texels=(1024, 1024, 1024), res=(7680, 4320), scene="terrain+props", n_props=200, n_ao=16, n_point=4, spec=True),
}


class Workload:
    def __init__(self, config: int, rank: int = 0, world: int = 1, device: int | None = None, tile=(128, 128),
                 scale: float = 1.0, frame_index: int = 0, gather: str = "auto"):
        """scale < 1 shrinks volume and resolution proportionally (tests only; the bench uses 1.0).
        gather: how the output tiles of world > 1 are exchanged -- "fused": the pass kernels store every output value into all
        ranks' stacks over peer memory (tiles.PeerStack); "nccl": one all-gather after the passes; "auto": fused when
        torch.distributed runs on NCCL, else nccl (also: no process group, e.g. shards run one after another on one GPU)."""
        import torch
        self.torch = torch
        cfg = dict(CONFIGS[config])
        self.cfg, self.config, self.rank, self.world = cfg, config, rank, world
        tex = tuple(max(16, int(round(t * scale))) for t in cfg["texels"])
        res = tuple(max(32, int(round(r * scale))) for r in cfg["res"])
        self.texels, self.res = tex, res
        self.ctx = E.Context(rank if device is None else device)
        self.vol = E.ShadowVoxSystem(self.ctx, tex)
        W, H = res
        if world == 1:
            self.gb = E.GeometryBuffer(self.ctx, W, H)
        else:
            self.gb = E.GeometryBuffer(self.ctx, W, H, tile[0], tile[1], rank=rank, world=world)
        self.host_volume = None
        self.entities = None
        self.props = None
        self._build_scene(scale)
        self.view = S.default_camera(tex, W, H, frame_index) if cfg["scene"] != "house" else self._house_view(W, H, frame_index)
        self.gb.set_noise(S.blue_noise(4))
        self.gb.synthesize(self.vol, self.view)
        self.n_ao, self.n_point, self.spec = cfg["n_ao"], cfg["n_point"], cfg["spec"]
        self.lights = S.quarter_point_lights(self.host_volume, self.n_point) if self.n_point else None
        # packed outputs: planes [shadow, ao, spec_t, point_0..point_{n-1}], every rank padded to the same tile count
        total_tiles = self.gb.tiles_x * self.gb.tiles_y
        self.tiles_padded = -(-total_tiles // world)
        self.n_planes = 3 + self.n_point
        self.gathered_main = self.gathered_point = None      # (world, 3, tiles, th, tw) / (world, n_point, tiles, th, tw)
        self.stack = None
        import os
        self.concurrent = os.environ.get("VXL_CONCURRENT", "1") != "0"
        auto = gather == "auto"
        if auto:
            import torch.distributed as dist
            gather = "fused" if world > 1 and dist.is_available() and dist.is_initialized() and dist.get_backend() == "nccl" else "nccl"
        if world > 1 and gather == "fused":
            from .tiles import PeerStack, PeerStackUnavailable
            try:
                self.stack = PeerStack(self.ctx, (self.n_planes, self.tiles_padded, self.gb.tile_h, self.gb.tile_w), rank, world)
            except PeerStackUnavailable:
                if not auto:
                    raise
                gather = "nccl"                                   # every rank reaches the same decision (PeerStack agrees on it collectively)
        self.gather_mode = gather if world > 1 else "none"
        if self.stack is not None:
            self._bind_stack(self.stack.tensor)
        else:
            self.out = torch.zeros((self.n_planes, self.tiles_padded, self.gb.tile_h, self.gb.tile_w), dtype=torch.float32,
                                   device=self.ctx.torch_device)
        self._side = torch.cuda.Stream(device=self.ctx.torch_device) if world > 1 else None
        self.ctx.sync()

    def _bind_stack(self, t):
        self.out = t[self.rank]                                   # this rank's slot of its own copy; the mirrors fill the peers' copies
        self.gathered_main, self.gathered_point = t[:, :3], t[:, 3:]

    # -- scene -------------------------------------------------------------------------------------
    def _house_view(self, W, H, frame):
        ext = 2 * self.texels[0] * 0.1
        return S.make_view((-ext * 0.2, ext * 0.9, -ext * 0.25), 3.927, -0.5, W, H, frame)

    def _build_scene(self, scale):
        cfg, vol = self.cfg, self.vol
        sx, sy, sz = self.texels
        if cfg["scene"] == "house":
            msz = max(8, int(round(40 * scale)))
            mid = vol.add_model(S.house_model(msz, seed=1))
            e = S.entities(1)
            off = (2 * sx - msz) // 2
            e[0]["model"] = mid
            e[0]["cur"] = S.transform_matrix((off * 0.1, 0.2, off * 0.1))
            vol.OnUpdate(e, want_regions=False)
            self.host_volume = vol.download()
            return
        vol.gen_terrain()
        self.host_volume = vol.download()
        if cfg["scene"] == "terrain+props":
            msz = max(8, int(round(40 * scale)))
            mid = vol.add_model(S.house_model(msz, seed=1))
            e = S.prop_entities(self.host_volume, n=cfg["n_props"], model_size=msz, seed=2, model=mid)
            self.props = e                                       # the instanced models, also the draw list of the geometry pass
            vol.OnUpdate(e, want_regions=False)
            self.host_volume = vol.download()
        elif cfg["scene"] == "dynamic":
            mid = vol.add_model(S.shell_cube_model(16))
            self.entities, pos, yaw = S.dynamic_entities(self.texels, cfg["n_entities"], seed=3, model=mid)
            # The motion is INPUT, not part of the measured path: the 60 frames of transforms are generated
            # once here; a step walks them forwards then backwards so that every frame's `prev` is the previous frame's `cur`.
            frames, e = [self.entities["cur"].copy()], self.entities
            for _ in range(59):
                e, pos, yaw = S.advance_entities(e, pos, yaw)
                frames.append(e["cur"].copy())
            self._frames, self._t = np.stack(frames), 0
            vol.OnUpdate(self.entities, want_regions=False)     # first frame: prev = identity (reference quirk)
            self.host_volume = None                              # changes every frame; download on demand

    # -- one frame ---------------------------------------------------------------------------------
    def planes(self):
        n = self.gb.n_tiles
        o = self.out
        return dict(shadow=o[0, :n], ao=o[1, :n], spec_t=o[2, :n], point=o[3:, :n])

    def advance(self):
        """Dynamic scene: move the entities, re-voxelise them (ShadowVoxSystem::OnUpdate) and rebuild the occupancy levels."""
        n = len(self._frames)
        tri = lambda t: (n - 1) - abs((t % (2 * n - 2)) - (n - 1))     # 0, 1, ..., n-1, n-2, ..., 1, 0, 1, ...
        self._t += 1
        self.entities["prev"] = self._frames[tri(self._t - 1)]
        self.entities["cur"] = self._frames[tri(self._t)]
        self.vol.OnUpdate(self.entities, want_regions=False)
        self.vol.build_occupancy()

    def step(self, gather: bool = True, advance: bool = True):
        """One frame over this rank's tiles; with world > 1, all-gather the packed output tiles."""
        if advance and self.cfg["scene"] == "dynamic":
            self.advance()
        n = self.gb.n_tiles
        o = self.out
        torch = self.torch
        gather = gather and self.world > 1
        if self.stack is not None:
            # fused gather: the passes write this rank's tiles into every rank's stack (peer-to-peer stores); one fence ends the frame.
            # Mirrors and the padded plane stride are on only here, where every output pointer lies inside the stack.
            # Frames alternate between two stacks (tiles.PeerStack): whoever consumes this frame's stack queues that on the context's
            # stream before the next step().
            self._bind_stack(self.stack.begin())
            o = self.out
            self.ctx.set_light_plane_stride(self.tiles_padded * self.gb.tile_h * self.gb.tile_w)
            try:
                self._passes(o, n)
            finally:
                self.stack.end()
                self.ctx.set_light_plane_stride(0)
            if gather:
                self.stack.fence()
            return self.out
        if self.world == 1 and self.concurrent:
            self._passes(o, n)                                   # one GPU: the three pass kernels side by side (vxl_lighting)
            return self.out
        if gather:
            from .tiles import gather_tiles
        if self.n_point:
            # point planes are [n_point][tiles_padded] inside `out`; the C ABI wants consecutive planes of the
            # shard's own size, which holds when tiles_padded == n_tiles; otherwise stage through a view copy
            if n == self.tiles_padded:
                self._point(o[3:])
            else:
                tmp = self.ctx.empty((self.n_point, n, self.gb.tile_h, self.gb.tile_w), torch.float32)
                self._point(tmp)
                o[3:, :n].copy_(tmp)
            if gather:
                # the local-light planes are the largest output: their all-gather runs on a side stream under the
                # ambient and reflection passes (the one collective of the path, SURVEY 8e, in two pieces)
                self._side.wait_stream(torch.cuda.current_stream())
                with torch.cuda.stream(self._side):
                    self.gathered_point = gather_tiles(o[3:], out=self.gathered_point)
        E.LightAmbientPipeline.Get().Use(self.view, self.gb, self.vol, n_ao=self.n_ao, out_shadow=o[0, :n], out_ao=o[1, :n])
        if self.spec:
            E.LightReflectionPipeline.Get().Use(self.view, self.gb, self.vol, out_spec_t=o[2, :n])
        if gather:
            self.gathered_main = gather_tiles(o[:3], out=self.gathered_main)
            torch.cuda.current_stream().wait_stream(self._side)
        return self.out

    def _passes(self, o, n):
        """All passes of this rank's shard.  Default: vxl_lighting, the three kernels on concurrent streams joined on the context's
        stream; VXL_CONCURRENT=0: one pass after the other (the A/B baseline)."""
        if self.concurrent:
            outs = dict(shadow=o[0, :n], ao=o[1, :n])
            if self.spec:
                outs["spec_t"] = o[2, :n]
            if self.n_point:
                outs["point_shadow"] = o[3:]
            E.lighting(self.ctx, self.vol, self.view, self.gb, outs, n_ao=self.n_ao, point=self.lights)
            return
        if self.n_point:
            self._point(o[3:])
        E.LightAmbientPipeline.Get().Use(self.view, self.gb, self.vol, n_ao=self.n_ao, out_shadow=o[0, :n], out_ao=o[1, :n])
        if self.spec:
            E.LightReflectionPipeline.Get().Use(self.view, self.gb, self.vol, out_spec_t=o[2, :n])

    @property
    def gathered(self):
        """(world, n_planes, tiles_padded, tile_h, tile_w): every rank's packed output tiles after a gathering step."""
        if self.gathered_main is None:
            return None
        if self.gathered_point is None:
            return self.gathered_main
        return self.torch.cat([self.gathered_main, self.gathered_point], dim=1)

    def _point(self, out):
        L = self.lights
        E.LightPointPipeline.Get().Use(self.view, self.gb, self.vol,
                                       lambda p: [p.DrawLight(l["Position"], l["Range"], l["Color"], l["Attenuation"]) for l in L],
                                       out_shadow=out)

    # -- accounting --------------------------------------------------------------------------------
    def count(self):
        """(rays, probes, lit pixels) of one frame on this rank, counted by the kernels themselves."""
        self.ctx.stats_reset()
        self.step(gather=False)
        return self.ctx.stats()

    def per_pass_counts(self):
        """stats per pass (ambient, point, reflection) of one frame on this rank."""
        n = self.gb.n_tiles
        o = self.out
        res = {}
        self.ctx.stats_reset()
        E.LightAmbientPipeline.Get().Use(self.view, self.gb, self.vol, n_ao=self.n_ao, out_shadow=o[0, :n], out_ao=o[1, :n])
        res["ambient"] = self.ctx.stats()
        if self.n_point:
            self.ctx.stats_reset()
            tmp = self.ctx.empty((self.n_point, n, self.gb.tile_h, self.gb.tile_w), self.torch.float32)
            self._point(tmp)
            res["point"] = self.ctx.stats()
        if self.spec:
            self.ctx.stats_reset()
            E.LightReflectionPipeline.Get().Use(self.view, self.gb, self.vol, out_spec_t=o[2, :n])
            res["reflection"] = self.ctx.stats()
        return res

    def assemble(self, gathered: np.ndarray | None = None):
        """Full-frame float planes (n_planes, H, W) from the gathered tile-compact tensor (or, world == 1, from out)."""
        W, H = self.res
        full = np.zeros((self.n_planes, H, W), np.float32)
        if self.world == 1:
            full[:] = self.out.cpu().numpy()[:, 0, :H, :W]
            return full
        g = self.gathered.cpu().numpy() if gathered is None else gathered
        return self.gb.layout.assemble(g)

    def close(self):
        if self.stack is not None:
            self.out = self.gathered_main = self.gathered_point = None
            self.stack.close()
            self.stack = None
        self.vol.close()
        self.ctx.close()
