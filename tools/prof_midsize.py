#!/usr/bin/env python3
"""Mid-size grids through the tiled passes (bunny 256^3 = BASELINE configs[1], gyroid 128^3, sphere 128^3): a few calls
each, for an ncu launch list (`ncu --metrics gpu__time_duration.sum`), or wall-clock per call with `time`."""
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import prim3d  # noqa: E402
from primitive3d_b200 import workloads  # noqa: E402


def main():
    dev = torch.device("cuda", 0)
    bunny = np.load(os.path.join(ROOT, "tests", "golden", "mc_bunny66.npz"))["grid"]
    grids = {"sphere128": torch.from_numpy(workloads.sphere_int64(128).astype(np.float32)).to(dev),
             "gyroid128": torch.from_numpy(workloads.gyroid(128)).to(dev),
             "bunny256": torch.from_numpy(workloads.upsample_trilinear(bunny, 256)).to(dev)}
    reps = int(sys.argv[1]) if len(sys.argv) > 1 else 3
    for name, g in grids.items():
        box = [float(s) for s in g.shape]
        for _ in range(3):
            prim3d._C.marching_cubes(g, 0.0, [0, 0, 0], box)
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            prim3d._C.marching_cubes(g, 0.0, [0, 0, 0], box)
        torch.cuda.synchronize()
        print(name, "wall us per call", (time.perf_counter() - t0) / reps * 1e6, flush=True)


if __name__ == "__main__":
    main()
