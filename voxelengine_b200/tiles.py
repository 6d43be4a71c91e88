"""Screen-tile partitioning of a frame across ranks: the tile grid is laid row-major
over the frame, tile g goes to rank g % world ("round-robin" for load balance between sky and dense
regions), each rank stores its tiles compactly as [n_local][tile_h][tile_w] (zero padded at the frame
edge) and, for the gather, pads its tile count to ceil(total / world).  Pure host logic (numpy /
torch.distributed); the kernels consume the same description through vxl_frame."""
from __future__ import annotations

import numpy as np


class TileLayout:
    def __init__(self, width, height, tile_w=None, tile_h=None, rank=0, world=1):
        self.width, self.height = int(width), int(height)
        self.tile_w, self.tile_h = int(tile_w or width), int(tile_h or height)
        self.tiles_x = -(-self.width // self.tile_w)
        self.tiles_y = -(-self.height // self.tile_h)
        self.total = self.tiles_x * self.tiles_y
        self.rank, self.world = int(rank), int(world)
        if not (0 <= self.rank < self.world):
            raise ValueError("rank out of range")
        self.tile_first, self.tile_stride = self.rank, self.world
        self.n_tiles = len(range(self.tile_first, self.total, self.tile_stride))
        self.tiles_padded = -(-self.total // self.world)

    def ids(self, rank=None):
        r = self.rank if rank is None else rank
        return list(range(r, self.total, self.world))

    def rect(self, g):
        ty, tx = divmod(g, self.tiles_x)
        y0, x0 = ty * self.tile_h, tx * self.tile_w
        return y0, x0, min(self.tile_h, self.height - y0), min(self.tile_w, self.width - x0)

    def to_tiles(self, plane: np.ndarray, rank=None) -> np.ndarray:
        """(..., H, W) -> (..., n_local, tile_h, tile_w) for `rank` (default: own), zero padded."""
        ids = self.ids(rank)
        out = np.zeros(plane.shape[:-2] + (len(ids), self.tile_h, self.tile_w), dtype=plane.dtype)
        for i, g in enumerate(ids):
            y0, x0, h, w = self.rect(g)
            out[..., i, :h, :w] = plane[..., y0:y0 + h, x0:x0 + w]
        return out

    def from_tiles(self, tiles: np.ndarray, out: np.ndarray, rank=None) -> np.ndarray:
        """Scatter (..., >=n_local, tile_h, tile_w) of `rank` into the full (..., H, W) plane `out`."""
        for i, g in enumerate(self.ids(rank)):
            y0, x0, h, w = self.rect(g)
            out[..., y0:y0 + h, x0:x0 + w] = tiles[..., i, :h, :w]
        return out

    def assemble(self, gathered: np.ndarray) -> np.ndarray:
        """gathered: (world, planes, tiles_padded, tile_h, tile_w) as produced by all_gather of every
        rank's padded tile stack -> (planes, H, W)."""
        full = np.zeros((gathered.shape[1], self.height, self.width), gathered.dtype)
        for r in range(self.world):
            self.from_tiles(gathered[r], full, rank=r)
        return full


def gather_tiles(local, group=None, out=None):
    """all-gather every rank's padded tile stack (torch tensor, same shape on all ranks) ->
    (world, *local.shape).  Works with NCCL (device tensors) and gloo (CPU tensors)."""
    import torch
    import torch.distributed as dist
    world = dist.get_world_size(group)
    if out is None:
        out = torch.empty((world,) + tuple(local.shape), dtype=local.dtype, device=local.device)
    dist.all_gather_into_tensor(out.view(-1), local.contiguous().view(-1), group=group)
    return out


def mirror_deltas(bases, rank):
    """Byte deltas from rank's own copy of the gathered stack to every peer's copy, as mapped in THIS process: a store at
    own_base + off repeated at own_base + off + delta lands at the same offset of the peer's stack."""
    return [int(b) - int(bases[rank]) for r, b in enumerate(bases) if r != rank]


class PeerStackUnavailable(RuntimeError):
    pass


class PeerStack:
    """The gathered tile stack (world, planes, tiles_padded, tile_h, tile_w), one copy per rank, every copy mapped into every process
    (vxl_group: CUDA IPC).  Inside begin() ... end() the light-pass kernels store each output value into all copies -- the rank's own
    slot of its own stack and, by peer-to-peer stores over NVLink, the same slot of every peer's -- so the all-gather of SURVEY 8e
    happens inside the passes; fence() closes the frame with peer-written arrival flags (one tiny kernel, no collective).
    Two stacks alternate frame by frame: a rank may still read frame N (queued on the context's stream behind fence N) while its
    peers already store frame N + 1; consumers of frame N must be queued before the passes of frame N + 1."""

    def __init__(self, ctx, shape_per_rank, rank, world, group=None, n_stacks=2):
        import torch
        import torch.distributed as dist
        self.ctx, self.rank, self.world, self.group = ctx, int(rank), int(world), group
        self.shape = (world,) + tuple(int(v) for v in shape_per_rank)
        self.nbytes = int(np.prod(self.shape)) * 4
        self.n_stacks, self.frame, self.cur = int(n_stacks), 0, 0
        self.g = ctx.group_create(self.rank, self.world, self.nbytes, self.n_stacks)
        handles = [None] * world
        dist.all_gather_object(handles, self.g.handle(), group=group)
        err = None
        try:
            self.g.connect(handles)
        except Exception as e:                                        # no peer access between two of the GPUs, IPC disabled, ...
            err = e
        ok = torch.tensor([0.0 if err else 1.0], device=ctx.torch_device)
        dist.all_reduce(ok, op=dist.ReduceOp.MIN, group=group)        # all ranks take the same decision
        if float(ok.item()) == 0.0:
            dist.barrier(group=group)
            self.g.destroy()                                          # closes whatever it had opened, frees the allocation
            raise PeerStackUnavailable(f"peer mapping failed on at least one rank ({err})")
        self.tensors = [ctx.tensor_view(self.g.stack(w), self.shape) for w in range(self.n_stacks)]   # (world, planes, tiles_padded, th, tw)
        dist.barrier(group=group)                                     # every copy is zeroed and mapped before anyone stores into it

    @property
    def tensor(self):
        """The stack of the current frame (the one begin() last pointed the mirrors at)."""
        return self.tensors[self.cur]

    def begin(self):
        """Mirrors on, pointed at this frame's stack of every peer; returns this rank's own copy of that stack."""
        self.cur = self.frame % self.n_stacks
        self.g.begin_frame(self.frame)
        return self.tensors[self.cur]

    def end(self):
        self.g.end_frame()

    def fence(self):
        """Stream-ordered: returns (on the stream) once every rank's passes of this frame, and with them their peer stores, have
        completed; the next begin() moves on to the other stack."""
        self.g.fence()
        self.frame += 1

    def close(self):
        import torch.distributed as dist
        self.ctx.sync()
        dist.barrier(group=self.group)                                # nobody is still storing into a copy that is about to go
        self.tensors = None
        self.g.destroy()                                              # closes the peer mappings, then frees the allocation
        dist.barrier(group=self.group)
