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


# ---- the fused gather's host logic (tiles.PeerStack) on two gloo ranks with a stand-in context -------------------------------------
class _FakeGroup:
    """What PeerStack needs from engine.Group (vxl_group), on host memory: the allocation is a numpy buffer, a handle is
    (pid, address), and connecting fails when asked to -- the peer-mapping failure of one rank."""

    def __init__(self, ctx, rank, world, stack_bytes, n_stacks):
        self.ctx, self.rank, self.world, self.n_stacks, self.stack_bytes = ctx, rank, world, n_stacks, (stack_bytes + 255) & ~255
        self.buf = np.zeros(self.stack_bytes * n_stacks + 256, np.uint8)
        self.connected, self.destroyed, self.frames, self.fences, self.mirrors_on = False, False, [], 0, False

    def handle(self):
        return (str(os.getpid()) + ":" + str(self.buf.ctypes.data)).encode().ljust(64, b"\0")

    def connect(self, handles):
        assert len(handles) == self.world and all(len(h) == 64 for h in handles)
        if self.ctx.fail:
            raise RuntimeError("peer access is not supported between these two devices")
        self.connected = True

    def stack(self, which):
        return self.buf.ctypes.data + which * self.stack_bytes

    def begin_frame(self, frame):
        self.frames.append(frame); self.mirrors_on = True
        return self.stack(frame % self.n_stacks)

    def end_frame(self):
        self.mirrors_on = False

    def fence(self):
        self.fences += 1

    def destroy(self):
        self.destroyed = True


class _FakeCtx:
    def __init__(self, rank, fail=False):
        import torch
        self.rank, self.fail, self.torch_device = rank, fail, torch.device("cpu")
        self.groups = []

    def group_create(self, rank, world, stack_bytes, n_stacks=2):
        g = _FakeGroup(self, rank, world, stack_bytes, n_stacks)
        self.groups.append(g)
        return g

    def sync(self):
        pass

    def tensor_view(self, p, shape, dtype="float32"):
        import torch
        g = self.groups[-1]
        off = p - g.buf.ctypes.data
        n = int(np.prod(shape)) * 4
        return torch.from_numpy(g.buf[off:off + n].view(np.float32).reshape(shape))


def _stack_worker(rank, world, port, q, failing_rank):
    sys.path.insert(0, ROOT)
    import torch.distributed as dist
    from voxelengine_b200.tiles import PeerStack, PeerStackUnavailable
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    ctx = _FakeCtx(rank, fail=(rank == failing_rank))
    try:
        st = PeerStack(ctx, (3, 2, 4, 4), rank, world)
        g = ctx.groups[-1]
        ok = st.tensor.shape == (world, 3, 2, 4, 4) and g.connected and len(st.tensors) == 2
        t0 = st.begin()
        ok = ok and g.mirrors_on and t0.data_ptr() == st.tensors[0].data_ptr()
        st.end(); st.fence()
        t1 = st.begin()                                               # frames alternate between the two stacks
        ok = ok and t1.data_ptr() == st.tensors[1].data_ptr() and st.tensor.data_ptr() == t1.data_ptr() and g.frames == [0, 1]
        st.end(); st.fence()
        ok = ok and st.begin().data_ptr() == t0.data_ptr()
        st.end()
        st.close()
        ok = ok and not g.mirrors_on and g.fences == 2 and g.destroyed
        q.put((rank, "stack" if ok else "broken"))
    except PeerStackUnavailable:
        q.put((rank, "unavailable" if ctx.groups and ctx.groups[-1].destroyed else "leaked"))
    dist.destroy_process_group()


@pytest.mark.parametrize("failing_rank", [-1, 1])
def test_peer_stack_ranks_agree_world2_gloo(failing_rank):
    """Both ranks build the stack, or -- when one rank cannot map its peer -- both give it up together (no rank is left waiting in a
    collective), close what they had opened and free their allocation: Workload then falls back to the NCCL gather on every rank."""
    import torch.multiprocessing as mp
    s = socket.socket(); s.bind(("127.0.0.1", 0)); port = s.getsockname()[1]; s.close()
    ctx = mp.get_context("spawn")
    q = ctx.Queue()
    ps = [ctx.Process(target=_stack_worker, args=(r, 2, port, q, failing_rank)) for r in range(2)]
    for p in ps:
        p.start()
    res = sorted(q.get(timeout=120) for _ in ps)
    for p in ps:
        p.join(timeout=60)
    want = "stack" if failing_rank < 0 else "unavailable"
    assert res == [(0, want), (1, want)]
