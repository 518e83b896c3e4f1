#!/usr/bin/env python3
"""Small grids (the reference's own example sizes): time per extraction through prim3d.libPrim3D.marching_cubes
(CUDA events around back-to-back calls, so launch overhead and the one host wait are inside), and a batch of 64
bunny-sized grids in one launch (p3d_mc_extract_batch) against the same grids one by one.

  python tools/prof_small.py                             # defaults (single grids of up to 2^20 samples in one launch)
  P3D_MC_SMALL_SINGLE_MAX=0 python tools/prof_small.py   # single grids through the tiled passes, for comparison
"""
import json
import os
import statistics
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import prim3d  # noqa: E402
from primitive3d_b200 import capi, workloads  # noqa: E402


def timed(fn, reps=30, warm=5):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    ts = []
    for _ in range(reps):
        a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        a.record()
        fn()
        b.record()
        torch.cuda.synchronize()
        ts.append(a.elapsed_time(b))
    return statistics.median(ts) * 1e3  # microseconds


def main():
    dev = torch.device("cuda", 0)
    bunny = np.load(os.path.join(ROOT, "tests", "golden", "mc_bunny66.npz"))["grid"]
    grids = {"bunny66": torch.from_numpy(bunny).to(dev),
             "sphere128": torch.from_numpy(workloads.sphere_int64(128).astype(np.float32)).to(dev),
             "sphere200": torch.from_numpy(workloads.sphere_int64(200).astype(np.float32)).to(dev),
             "gyroid128": torch.from_numpy(workloads.gyroid(128)).to(dev),
             "bunny256": torch.from_numpy(workloads.upsample_trilinear(bunny, 256)).to(dev)}
    out = {"small_max": os.environ.get("P3D_MC_SMALL_MAX", "default")}
    for name, g in grids.items():
        box = [float(s) for s in g.shape]
        v, f = prim3d._C.marching_cubes(g, 0.0, [0, 0, 0], box)
        out[name] = {"us_per_call": timed(lambda: prim3d._C.marching_cubes(g, 0.0, [0, 0, 0], box)), "V": v.shape[0], "F": f.shape[0]}
    batch = [grids["bunny66"] * (1.0 + 0.01 * i) for i in range(64)]
    res = capi.marching_cubes_batch(batch, 0.0)
    t_batch = timed(lambda: capi.marching_cubes_batch(batch, 0.0), reps=10, warm=2)
    t_each = timed(lambda: [prim3d._C.marching_cubes(g, 0.0, [0, 0, 0], [66.0] * 3) for g in batch], reps=10, warm=2)
    out["batch64_bunny66"] = {"one_launch_us": t_batch, "one_by_one_us": t_each, "speedup": t_each / t_batch,
                              "F_total": int(sum(r[1].shape[0] for r in res))}
    print(json.dumps(out))


if __name__ == "__main__":
    main()
