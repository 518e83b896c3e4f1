import sys, torch, time
sys.path.insert(0, '.')
import prim3d
from primitive3d_b200 import workloads as inputs
pts, tets, sdf = inputs.kuhn_tet_grid(128)
P, T, S = torch.from_numpy(pts).cuda(), torch.from_numpy(tets).cuda(), torch.from_numpy(sdf).cuda()
for _ in range(3):
    v, f = prim3d.marching_tetrahedras(P, T.clone(), S)
torch.cuda.synchronize()
t = time.perf_counter()
for _ in range(10):
    v, f = prim3d.marching_tetrahedras(P, T.clone(), S)
torch.cuda.synchronize()
print("ms per call incl clone", (time.perf_counter() - t) / 10 * 1e3, v.shape, f.shape)
