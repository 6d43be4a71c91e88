"""CPU: screen-tile partitioning and the output-tile gather (the only collective of the path),
including a world_size-2 gloo run of the same gather code the NCCL path uses."""
import os
import socket
import sys

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.mark.parametrize("world", [1, 2, 3, 8])
@pytest.mark.parametrize("res,tile", [((160, 90), (64, 32)), ((3840, 2160), (128, 128)), ((33, 17), (16, 16))])
def test_partition_roundtrip(world, res, tile):
    from voxelengine_b200.tiles import TileLayout
    W, H = res
    rs = np.random.RandomState(0)
    plane = rs.rand(2, H, W).astype(np.float32)
    seen = np.zeros((H, W), np.int32)
    stacks = []
    for r in range(world):
        L = TileLayout(W, H, tile[0], tile[1], r, world)
        t = L.to_tiles(plane)
        assert t.shape == (2, L.n_tiles, tile[1], tile[0])
        for g in L.ids():
            y0, x0, h, w = L.rect(g)
            seen[y0:y0 + h, x0:x0 + w] += 1
        pad = np.zeros((2, L.tiles_padded, tile[1], tile[0]), np.float32)
        pad[:, :L.n_tiles] = t
        stacks.append(pad)
    assert (seen == 1).all()                                  # every pixel owned by exactly one rank
    L0 = TileLayout(W, H, tile[0], tile[1], 0, world)
    assert sum(TileLayout(W, H, tile[0], tile[1], r, world).n_tiles for r in range(world)) == L0.total
    full = L0.assemble(np.stack(stacks))
    assert np.array_equal(full, plane)


def test_single_tile_is_plain_row_major():
    from voxelengine_b200.tiles import TileLayout
    L = TileLayout(37, 11)
    assert (L.n_tiles, L.tile_w, L.tile_h, L.tiles_padded) == (1, 37, 11, 1)
    p = np.arange(37 * 11, dtype=np.float32).reshape(11, 37)
    assert np.array_equal(L.to_tiles(p)[0], p)


def _worker(rank, world, port, q):
    sys.path.insert(0, ROOT)
    import torch
    import torch.distributed as dist
    from voxelengine_b200.tiles import TileLayout, gather_tiles
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    W, H, tw, th = 200, 120, 64, 48
    rs = np.random.RandomState(5)
    frame = rs.rand(3, H, W).astype(np.float32)               # the "outputs" every rank would compute for its tiles
    L = TileLayout(W, H, tw, th, rank, world)
    local = np.zeros((3, L.tiles_padded, th, tw), np.float32)
    local[:, :L.n_tiles] = L.to_tiles(frame)
    g = gather_tiles(torch.from_numpy(local))
    full = L.assemble(g.numpy())
    q.put((rank, bool(np.array_equal(full, frame))))
    dist.destroy_process_group()


def test_gather_world2_gloo():
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_worker, args=(r, 2, port, q)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    assert res == [(0, True), (1, True)]
