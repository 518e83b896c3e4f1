"""Times prim3d.marching_tetrahedras on the Kuhn grid (BASELINE configs[3] at n = 128).
  python tools/prof_mt.py [n] [calls]      calls = 0: ONE call, for an ncu capture"""
import os
import sys
import time

import torch

sys.path.insert(0, '.')
import prim3d  # noqa: E402
from primitive3d_b200 import workloads as inputs  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 128
calls = int(sys.argv[2]) if len(sys.argv) > 2 else 20
if len(sys.argv) > 3 and sys.argv[3] == "capi":   # through ctypes (P3D_CORE_LIB selects a build variant), capacities as the pybind layer keeps them
    from primitive3d_b200 import capi

    class prim3d:  # noqa: F811
        caps = None

        @staticmethod
        def marching_tetrahedras(P, T, S):
            v, f, ti, e = capi.marching_tetrahedra(P, T, S, capacities=prim3d.caps)
            if prim3d.caps is None:
                n12, V, F = f.shape[0], v.shape[0], f.shape[0]
                prim3d.caps = (n12 + n12 // 16 + 1024, 6 * V + 1024, V + V // 16 + 1024, F + F // 16 + 1024)
            return v, f
pts, tets, sdf = inputs.kuhn_tet_grid(n)
P, T, S = torch.from_numpy(pts).cuda(), torch.from_numpy(tets).cuda(), torch.from_numpy(sdf).cuda()
if calls == 0:   # two calls: the second one runs with capacities remembered from the first (ncu: --launch-skip)
    v, f = prim3d.marching_tetrahedras(P, T.clone(), S)
    v, f = prim3d.marching_tetrahedras(P, T.clone(), S)
    torch.cuda.synchronize()
    print(v.shape, f.shape)
    sys.exit(0)
if os.environ.get("MT_PREORIENTED"):   # every call sees tets that are oriented already: no flips, no write traffic
    prim3d.marching_tetrahedras(P, T, S)
clones = [T.clone() for _ in range(calls + 3)]
for i in range(3):
    v, f = prim3d.marching_tetrahedras(P, clones[i], S)
torch.cuda.synchronize()
ev = [torch.cuda.Event(enable_timing=True) for _ in range(2)]
t = time.perf_counter()
ev[0].record()
for i in range(calls):
    v, f = prim3d.marching_tetrahedras(P, clones[3 + i], S)
ev[1].record()
torch.cuda.synchronize()
print("ms per call: wall %.4f, device %.4f" % ((time.perf_counter() - t) / calls * 1e3, ev[0].elapsed_time(ev[1]) / calls), tuple(v.shape), tuple(f.shape))
