#!/usr/bin/env python3
"""Where the ~55 us of a small-grid call go: wall-clock per call of (a) nothing between two events, (b) the three
allocations, (c) the raw C call p3d_mc_extract into preallocated buffers (memset + one launch + one stream wait),
(d) a bare launch + stream wait of torch's own (x.add_(1); synchronize), (e) the pybind call."""
import ctypes
import json
import os
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
import prim3d  # noqa: E402
from primitive3d_b200 import capi  # noqa: E402


def wall(fn, reps=2000, warm=100):
    for _ in range(warm):
        fn()
    torch.cuda.synchronize()
    t0 = time.perf_counter()
    for _ in range(reps):
        fn()
    torch.cuda.synchronize()
    return (time.perf_counter() - t0) / reps * 1e6


def main():
    dev = torch.device("cuda", 0)
    out = {}
    for name in ("tiny4", "bunny66"):
        if name == "tiny4":
            g = torch.from_numpy(np.random.default_rng(0).standard_normal((4, 4, 4)).astype(np.float32)).to(dev)
        else:
            g = torch.from_numpy(np.load(os.path.join(ROOT, "tests", "golden", "mc_bunny66.npz"))["grid"]).to(dev)
        box = [float(s) for s in g.shape]
        desc = capi.McDesc.make(g.shape, 0.0)
        ws_bytes, hint = capi._desc_sizes(desc)
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
        vbuf = torch.empty((hint, 3), dtype=torch.float32, device=dev)
        fbuf = torch.empty((2 * hint, 3), dtype=torch.int32, device=dev)
        counts = (ctypes.c_int64 * 2)()
        lib = capi.lib()
        stream = capi._stream()
        args = (ctypes.byref(desc), g.data_ptr(), 0, ws.data_ptr(), ws.numel(), vbuf.data_ptr(), hint, fbuf.data_ptr(), 2 * hint,
                counts, stream)
        x = torch.zeros(16, device=dev)
        r = {}
        r["raw_c_call_us"] = wall(lambda: lib.p3d_mc_extract(*args))
        r["pybind_us"] = wall(lambda: prim3d._C.marching_cubes(g, 0.0, [0, 0, 0], box))
        r["three_allocs_us"] = wall(lambda: (torch.empty(ws_bytes, dtype=torch.uint8, device=dev),
                                             torch.empty((hint, 3), dtype=torch.float32, device=dev),
                                             torch.empty((2 * hint, 3), dtype=torch.int32, device=dev)))
        r["torch_launch_and_sync_us"] = wall(lambda: (x.add_(1), torch.cuda.current_stream().synchronize()))
        r["V"], r["F"] = counts[0], counts[1]
        out[name] = r
    print(json.dumps(out))


if __name__ == "__main__":
    main()
