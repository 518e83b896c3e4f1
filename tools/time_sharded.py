"""Times the multi-GPU drivers of the sharded extraction on the bench workload (gyroid N^3, dim-0 slabs):
sharded.marching_cubes_slab (python, torch.distributed all-gather), sharded.marching_cubes_slab_c (the single C entry
p3d_mc_sharded_extract over a raw NCCL communicator) and sharded.marching_cubes_slab_p2p (the same with the exchange over
peer memory, p3d_mc_sharded_extract_p2p).  Launch with torch.distributed.run, one rank per GPU.
  python -m torch.distributed.run --nproc-per-node N ... tools/time_sharded.py [size] [steps]"""
import os
import sys

import torch
import torch.distributed as dist

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from primitive3d_b200 import sharded  # noqa: E402
from bench import gyroid_cuda  # noqa: E402

n = int(sys.argv[1]) if len(sys.argv) > 1 else 2048
steps = int(sys.argv[2]) if len(sys.argv) > 2 else 10
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dev = torch.device("cuda", local)
dist.init_process_group("nccl", device_id=dev)
x0, x1h = sharded.slab_with_halo(n, world, rank)
slab = gyroid_cuda(n, x0, x1h, dev)
comm = sharded.nccl_comm_init()
out = sharded.marching_cubes_slab(slab, 0.0, x0, n)
caps = (out.vertices.shape[0] + out.vertices.shape[0] // 16, out.faces.shape[0] + out.faces.shape[0] // 16)
del out


def timed(fn):
    for _ in range(3):
        fn()
    torch.cuda.synchronize()
    dist.barrier()
    torch.cuda.synchronize()
    a, b = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    a.record()
    for _ in range(steps):
        fn()
    b.record()
    torch.cuda.synchronize()
    t = torch.tensor([a.elapsed_time(b) / steps], device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


t_py = timed(lambda: sharded.marching_cubes_slab(slab, 0.0, x0, n))
t_c = timed(lambda: sharded.marching_cubes_slab_c(slab, 0.0, x0, n, comm, rank, world, caps[0], caps[1]))
peer = sharded.PeerExchange(slab.shape[1], slab.shape[2])
t_p = timed(lambda: sharded.marching_cubes_slab_p2p(slab, 0.0, x0, n, peer, caps[0], caps[1]))
t_c2 = timed(lambda: sharded.marching_cubes_slab_c(slab, 0.0, x0, n, comm, rank, world, caps[0], caps[1]))
t_p2 = timed(lambda: sharded.marching_cubes_slab_p2p(slab, 0.0, x0, n, peer, caps[0], caps[1]))
if rank == 0:
    print(f"world={world} n={n}: python driver {t_py:.4f} ms/step, C entry (ncclAllGather) {t_c:.4f} / {t_c2:.4f} ms/step, "
          f"C entry (peer memory) {t_p:.4f} / {t_p2:.4f} ms/step")
peer.close()
sharded.nccl_comm_destroy(comm)
dist.destroy_process_group()
