#!/usr/bin/env python3
"""Inputs for ncu captures of k_small and for the call-overhead floor.

  python tools/prof_small_kernel.py single   # bunny 66^3 through p3d_mc_extract, 3 calls
  python tools/prof_small_kernel.py batch    # 64 bunny-sized grids in one launch, 2 calls
  python tools/prof_small_kernel.py floor    # us per call of a 4^3 grid (launch + host wait + wrapper) and of bunny 66^3
  python tools/prof_small_kernel.py both     # single, single, batch, single, batch: `ncu --launch-skip 3 -c 2` captures
                                             # one warm launch of each kind (tools/collect_profiles.sh)
"""
import json
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import prim3d  # noqa: E402
from primitive3d_b200 import capi  # noqa: E402
from tools.prof_small import timed  # noqa: E402


def main():
    mode = sys.argv[1] if len(sys.argv) > 1 else "single"
    dev = torch.device("cuda", 0)
    bunny = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "mc_bunny66.npz"))["grid"]).to(dev)
    if mode == "single":
        for _ in range(3):
            prim3d._C.marching_cubes(bunny, 0.0, [0, 0, 0], [66.0] * 3)
    elif mode == "sphere128":   # with P3D_MC_SMALL_SINGLE_MAX=4194304: the single-launch kernel at 128^3
        from primitive3d_b200 import workloads
        g = torch.from_numpy(workloads.sphere_int64(128).astype(np.float32)).to(dev)
        for _ in range(3):
            prim3d._C.marching_cubes(g, 0.0, [0, 0, 0], [128.0] * 3)
    elif mode == "both":
        batch = [bunny * (1.0 + 0.01 * i) for i in range(64)]
        for kind in "ssbsb":
            if kind == "s":
                prim3d._C.marching_cubes(bunny, 0.0, [0, 0, 0], [66.0] * 3)
            else:
                capi.marching_cubes_batch(batch, 0.0)
    elif mode == "batch":
        batch = [bunny * (1.0 + 0.01 * i) for i in range(64)]
        for _ in range(2):
            capi.marching_cubes_batch(batch, 0.0)
    else:
        tiny = torch.from_numpy(np.random.default_rng(0).standard_normal((4, 4, 4)).astype(np.float32)).to(dev)
        out = {}
        for name, g in (("tiny4", tiny), ("bunny66", bunny)):
            box = [float(s) for s in g.shape]
            out[name + "_pybind_us"] = timed(lambda: prim3d._C.marching_cubes(g, 0.0, [0, 0, 0], box), reps=200, warm=20)
            desc = capi.McDesc.make(g.shape, 0.0)
            out[name + "_capi_us"] = timed(lambda: capi.mc_extract(desc, g), reps=200, warm=20)
        print(json.dumps(out))
    torch.cuda.synchronize()


if __name__ == "__main__":
    main()
