"""Per-shard kernel times of config 3 split into `world` screen-tile shards, run one after another on ONE GPU: separates load
imbalance between ranks from per-rank inefficiency.  usage: python tools/exp/shard_times.py [world] [tile_w] [tile_h]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__)))))
from voxelengine_b200.workloads import Workload  # noqa: E402

world = int(sys.argv[1]) if len(sys.argv) > 1 else 8
tile = (int(sys.argv[2]), int(sys.argv[3])) if len(sys.argv) > 3 else (128, 128)
flush = torch.empty(256 << 20, dtype=torch.uint8, device="cuda")
tot = []
for rank in range(world):
    wl = Workload(3, rank=rank, world=world, device=0, tile=tile)
    for _ in range(2):
        wl.step(gather=False)
    torch.cuda.synchronize()
    ts = []
    for _ in range(5):
        flush.fill_(1)
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record(); wl.step(gather=False); b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    st = wl.count()
    tot.append(sum(ts) / len(ts))
    print(f"rank {rank}: {tot[-1]:.3f} ms, tiles {wl.gb.n_tiles}, rays {st['rays']}, probes {st['steps']}", flush=True)
    wl.close()
print(f"tile {tile} world {world}: max {max(tot):.3f} mean {sum(tot) / len(tot):.3f} min {min(tot):.3f}  (sum {sum(tot):.3f})")
