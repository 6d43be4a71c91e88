"""The fused output-tile gather: mirrored, row-major output stores of the light-pass kernels.

On several GPUs every rank's copy of the gathered tile stack is mapped into every process (CUDA IPC) and the pass kernels repeat each
output store at address + delta (vxl_ctx_set_output_mirrors); bench.py verifies that path against an NCCL all-gather on 2 / 4 / 8
GPUs.  Here the same kernels run on ONE GPU with the 'peer' stacks in the same allocation: three copies of a padded stack, the passes
write slot `rank` of copy 0 and must leave identical bytes in copies 1 and 2 -- and the planes must still equal the oracle's."""
import numpy as np
import pytest

import scene_util as U
from voxelengine_b200 import scenes as S


def test_mirror_deltas():
    from voxelengine_b200.tiles import mirror_deltas
    assert mirror_deltas([1000, 5000, 200], 0) == [4000, -800]
    assert mirror_deltas([1000, 5000, 200], 1) == [-4000, -4800]
    assert mirror_deltas([64], 0) == []


@pytest.mark.gpu
@pytest.mark.parametrize("rank,world", [(0, 1), (1, 3), (2, 3)])
def test_mirrored_stores_fill_every_copy(gpu_ctx, oracle, rank, world):
    import torch
    from voxelengine_b200 import engine as E
    sc = U.terrain_scene(oracle)
    sz, sy, sx = sc["volume"].shape
    h, w = sc["gb"]["depth24"].shape
    vol = E.ShadowVoxSystem(gpu_ctx, (sx, sy, sz))
    vol.upload(sc["volume"])
    gb = E.GeometryBuffer(gpu_ctx, w, h, 64, 32, rank=rank, world=world)            # 160x90 frame: ragged edge tiles
    gb.set_noise(sc["gb"]["noise"])
    gb.set_planes(sc["gb"]["depth24"], sc["gb"]["normal"], sc["gb"]["material"])
    ext = np.array([2 * sx, 2 * sy, 2 * sz], np.float32) * 0.1
    lights = S.point_lights([(ext[0] * 0.3, ext[1] * 0.8, ext[2] * 0.4), (ext[0] * 0.7, ext[1] * 0.7, ext[2] * 0.6), (ext[0] * 0.5, ext[1] * 0.9, ext[2] * 0.5)],
                            float(ext[0]) * 0.5)
    n, padded = gb.n_tiles, gb.n_tiles + 2                                           # a stack padded beyond this shard's tile count
    n_planes = 3 + len(lights)
    copies = torch.zeros((3, world, n_planes, padded, gb.tile_h, gb.tile_w), dtype=torch.float32, device=gpu_ctx.torch_device)
    own = copies[0, rank]
    deltas = [copies[c].data_ptr() - copies[0].data_ptr() for c in (1, 2)]
    gpu_ctx.set_output_mirrors(deltas)
    gpu_ctx.set_light_plane_stride(padded * gb.tile_h * gb.tile_w)
    try:
        E.LightPointPipeline.Get().Use(sc["view"], gb, vol, lambda p: [p.DrawLight(l["Position"], l["Range"], l["Color"], l["Attenuation"]) for l in lights],
                                       out_shadow=own[3:])
        E.LightAmbientPipeline.Get().Use(sc["view"], gb, vol, n_ao=2, out_shadow=own[0, :n], out_ao=own[1, :n])
        E.LightReflectionPipeline.Get().Use(sc["view"], gb, vol, out_spec_t=own[2, :n])
    finally:
        gpu_ctx.set_output_mirrors([])
        gpu_ctx.set_light_plane_stride(0)
    torch.cuda.synchronize()
    got = copies.cpu().numpy()
    assert np.array_equal(got[1].view(np.uint32), got[0].view(np.uint32)) and np.array_equal(got[2].view(np.uint32), got[0].view(np.uint32))
    others = np.delete(got[0], rank, axis=0)
    assert not others.any() and not got[0][rank][:, n:].any()                         # nothing outside this rank's slot / tiles was touched
    wsh, wao, _ = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], 2)
    wpt, _ = oracle.pass_point(sc["volume"], sc["view"], sc["gb"], lights)
    wt, _ = oracle.pass_reflection(sc["volume"], sc["view"], sc["gb"])
    want = np.concatenate([np.stack([wsh, wao, wt]), wpt])
    for c in range(3):
        for p in range(n_planes):
            full = np.zeros((h, w), np.float32)
            ref = np.zeros((h, w), np.float32)
            gb.from_tiles(got[c][rank][p, :n], full)
            gb.from_tiles(gb.layout.to_tiles(want[p]), ref)                            # the oracle's plane restricted to this shard's tiles
            assert np.array_equal(full.view(np.uint32), ref.view(np.uint32)), (c, p)
    # with the mirrors off again the kernels write one copy only
    copies.zero_()
    E.LightReflectionPipeline.Get().Use(sc["view"], gb, vol, out_spec_t=own[2, :n])
    torch.cuda.synchronize()
    assert copies[0].any() and not copies[1].any() and not copies[2].any()
    vol.close()


@pytest.mark.gpu
def test_group_two_members_on_one_gpu(oracle):
    """vxl_group (the C side of the sharded frame): two members in ONE process on one GPU, each with its own context and stream, connected by
    plain pointers.  Each runs the passes over its tile shard into its slot of the frame's stack; the mirrors fill the other member's copy,
    the flag fence closes the frame without a collective, two frames alternate between the two stacks -- and both copies assemble to the
    oracle's planes."""
    import torch
    from voxelengine_b200 import engine as E
    sc = U.terrain_scene(oracle)
    sz, sy, sx = sc["volume"].shape
    h, w = sc["gb"]["depth24"].shape
    world = 2
    ctxs = [E.Context(0, use_torch_stream=False) for _ in range(world)]             # own streams: the two fence kernels wait for each other
    vols, gbs, groups = [], [], []
    for r, c in enumerate(ctxs):
        v = E.ShadowVoxSystem(c, (sx, sy, sz)); v.upload(sc["volume"]); vols.append(v)
        gb = E.GeometryBuffer(c, w, h, 64, 32, rank=r, world=world)
        gb.set_noise(sc["gb"]["noise"]); gb.set_planes(sc["gb"]["depth24"], sc["gb"]["normal"], sc["gb"]["material"])
        gbs.append(gb)
    padded = max(g.n_tiles for g in gbs)
    shape = (world, 3, padded, gbs[0].tile_h, gbs[0].tile_w)
    nbytes = int(np.prod(shape)) * 4
    groups = [c.group_create(r, world, nbytes, 2) for r, c in enumerate(ctxs)]
    bases = [g.base() for g in groups]
    for g in groups:
        g.connect_pointers(bases)
    wsh, wao, _ = oracle.pass_ambient(sc["volume"], sc["view"], sc["gb"], 2)
    wt, _ = oracle.pass_reflection(sc["volume"], sc["view"], sc["gb"])
    want = np.stack([wsh, wao, wt])
    torch.cuda.synchronize()
    for frame in range(3):
        stacks = []
        for r, (c, g, gb, v) in enumerate(zip(ctxs, groups, gbs, vols)):
            p = g.begin_frame(frame)
            assert p == g.stack(frame % 2)
            t = c.tensor_view(p, shape)
            stacks.append(t)
            n = gb.n_tiles
            own = t[r]
            E.LightAmbientPipeline.Get().Use(sc["view"], gb, v, n_ao=2, out_shadow=own[0, :n], out_ao=own[1, :n])
            E.LightReflectionPipeline.Get().Use(sc["view"], gb, v, out_spec_t=own[2, :n])
            g.end_frame()
            g.fence()
        for c, g in zip(ctxs, groups):
            c.sync()
            assert g.status() == 0
        a, b = stacks[0].cpu().numpy(), stacks[1].cpu().numpy()
        assert np.array_equal(a.view(np.uint32), b.view(np.uint32)), frame             # both members hold the same gathered stack
        full = np.zeros((3, h, w), np.float32)
        for r, gb in enumerate(gbs):
            for p in range(3):
                gb.from_tiles(a[r][p, :gb.n_tiles], full[p])
        assert np.array_equal(full.view(np.uint32), want.view(np.uint32)), frame
    for g in groups:
        g.destroy()
    for v in vols:
        v.close()
