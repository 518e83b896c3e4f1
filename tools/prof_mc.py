#!/usr/bin/env python3
"""Profiling driver for the marching-cubes kernels (run under gpurun, optionally under ncu).

  python tools/prof_mc.py [--size 1024] [--reps 5]

Runs the full extraction once (warm-up + known-answer check), then times each kernel with CUDA
events: the tile pass in mode 0 (with look-back) and in mode 1 (vertices only: no look-back wait, no
side-product stores), and the face pass (chunk scan included).  Under `ncu -k regex:k_` the same
launches are what gets captured.
"""
import argparse
import ctypes
import json
import os
import statistics
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--size", type=int, default=1024)
    ap.add_argument("--reps", type=int, default=5)
    ap.add_argument("--noise", action="store_true", help="uniform-noise grid instead of the gyroid")
    ap.add_argument("--planes", type=int, default=0, help="only the first PLANES dim-0 planes of the gyroid (a multi-GPU slab)")
    args = ap.parse_args()
    import torch
    from bench import gyroid_cuda
    from primitive3d_b200 import capi
    n = args.size
    dev = torch.device("cuda", 0)
    if args.noise:
        g = torch.rand((n, n, n), device=dev) - 0.5
    else:
        g = gyroid_cuda(n, 0, args.planes or n, dev)
    desc = capi.McDesc.make(g.shape, 0.0, [0.0, 0.0, 0.0], [float(n)] * 3, global_rx=n)
    V, F, ws, vbuf = capi.mc_count(desc, g)
    if vbuf.shape[0] < V:
        V, F, ws, vbuf = capi.mc_count(desc, g, vertex_capacity=V)
    faces = capi.mc_faces(desc, ws, F)
    torch.cuda.synchronize()
    L = capi.lib()
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    def timed(fn, before=None):
        ts = []
        for _ in range(args.reps + 1):
            if before:
                before()
            a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
            a.record()
            fn()
            b.record()
            torch.cuda.synchronize()
            ts.append(a.elapsed_time(b))
        return statistics.median(ts[1:] or ts)

    stage = lambda k: capi.check(L.p3d_mc_debug_stage(ctypes.byref(desc), g.data_ptr(), ws.data_ptr(), k,
                                                      vbuf.data_ptr(), vbuf.shape[0], stream))
    vert_only = lambda: capi.check(L.p3d_mc_vertices(ctypes.byref(desc), g.data_ptr(), ws.data_ptr(), vbuf.data_ptr(),
                                                     vbuf.shape[0], stream))
    do_faces = lambda: capi.check(L.p3d_mc_faces(ctypes.byref(desc), ws.data_ptr(), faces.data_ptr(), 0, stream))
    def checksum(t):  # order-sensitive
        flat = t.reshape(-1).view(torch.int32).to(torch.int64)
        wts = torch.arange(flat.numel(), device=dev, dtype=torch.int64) % 65521 + 1
        return int((flat * wts).sum().item() & 0xffffffffffff)

    out = {"size": n, "planes": int(g.shape[0]), "V": V, "F": F, "vsum": checksum(vbuf[:V]), "fsum": checksum(faces)}
    out["tile_pass_ms"] = timed(lambda: stage(1), before=lambda: stage(0))
    stage(0), stage(1)
    out["tile_vertices_only_ms"] = timed(vert_only)
    out["faces_ms"] = timed(do_faces)
    print(json.dumps(out))


if __name__ == "__main__":
    main()
