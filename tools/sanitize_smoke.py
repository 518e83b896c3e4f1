#!/usr/bin/env python3
"""Small end-to-end run of every marching-cubes entry point, meant to be run under compute-sanitizer:

  compute-sanitizer --tool memcheck  python tools/sanitize_smoke.py
  compute-sanitizer --tool racecheck python tools/sanitize_smoke.py --tiny

Checks the outputs against the staged path so a wrong answer is not mistaken for a clean run."""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from primitive3d_b200 import workloads as inputs  # noqa: E402  (input generators only)
from primitive3d_b200 import capi  # noqa: E402


def main():
    tiny = "--tiny" in sys.argv
    shapes = [(20, 17, 140)] if tiny else [(33, 33, 33), (20, 17, 140), (9, 9, 384), (40, 24, 260)]
    for i, shape in enumerate(shapes):
        for kind in ("noise", "smooth"):
            g = inputs.noise(shape, 50 + i) if kind == "noise" else inputs.gyroid(max(shape), periods=2)[:shape[0], :shape[1], :shape[2]].copy()
            dev = torch.from_numpy(np.ascontiguousarray(g)).cuda()
            v0, f0 = capi.marching_cubes(dev, 0.0)
            desc = capi.McDesc.make(dev.shape, 0.0)
            v1, f1, V, F = capi.mc_extract(desc, dev)   # small grids: the single-launch kernel (its own vertex numbering)
            same = lambda va, fa, vb, fb: fa.shape == fb.shape and torch.equal(va[fa.long()], vb[fb.long()])  # triangle by triangle
            assert v0.shape == v1.shape and same(v0, f0, v1, f1)
            v2, f2 = capi.marching_cubes(dev.double(), 0.0)
            assert torch.equal(v0, v2) and torch.equal(f0, f2)
            out = capi.marching_cubes_batch([dev, dev[:, :, : shape[2] // 2].contiguous()], 0.0)
            assert out[0][0].shape == v0.shape and same(out[0][0], out[0][1], v0, f0)
            hv, hf = capi.marching_cubes_host(torch.from_numpy(np.ascontiguousarray(g)).pin_memory(), 0.0, slab_planes=8)
            assert hv.shape == v0.shape and hf.shape == f0.shape
            torch.cuda.synchronize()
            print(f"{shape} {kind}: V={V} F={F} ok", flush=True)


if __name__ == "__main__":
    main()
