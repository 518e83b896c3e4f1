#!/usr/bin/env python3
"""Dense call against the block-sparse form (p3d_mc_extract_sparse) on a level set that touches few tiles: a sphere in
an N^3 float32 grid (the reference's examples/sphere.py field), the tile list computed once beforehand.
  python tools/prof_sparse.py [N]"""
import json
import os
import statistics
import sys

import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from primitive3d_b200 import capi  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 1024
dev = torch.device("cuda", 0)
ax = torch.arange(n, device=dev, dtype=torch.float32) - n / 2
grid = (n * 0.4) ** 2 - (ax[:, None, None] ** 2 + ax[None, :, None] ** 2 + ax[None, None, :] ** 2)   # > 0 inside the sphere
tiles = capi.active_tiles(grid, 0.0)
total = -(-n // 8) * -(-n // 8) * -(-n // 128)


def timed(fn, reps=10):
    for _ in range(3):
        out = fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        out = fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts), out


t_dense, (v0, f0) = timed(lambda: capi.marching_cubes(grid, 0.0))
caps = (v0.shape[0] + v0.shape[0] // 16, f0.shape[0] + f0.shape[0] // 16)
t_sparse, (v1, f1) = timed(lambda: capi.marching_cubes_sparse(grid, 0.0, tiles, vertex_capacity=caps[0], face_capacity=caps[1]))
assert v1.shape == v0.shape and f1.shape == f0.shape
print(json.dumps({"grid": n, "tiles_listed": int(tiles.numel()), "tiles_total": total, "V": v0.shape[0], "F": f0.shape[0],
                  "dense_ms": t_dense, "sparse_ms": t_sparse, "speedup": t_dense / t_sparse}))
