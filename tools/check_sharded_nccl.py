#!/usr/bin/env python3
"""Real multi-GPU check of primitive3d_b200.sharded (run under torchrun on a box with >= 2 GPUs):

  python -m torch.distributed.run --nnodes=1 --nproc-per-node N --master-addr 127.0.0.1 \
      --master-port 29533 tools/check_sharded_nccl.py [size]

Every rank extracts its dim-0 slab of a gyroid with NCCL count/table exchange; rank 0 gathers the
shards, concatenates them in rank order and compares with its own single-GPU extraction of the
full grid: the arrays must be identical (same vertex numbering, global face ids)."""
import os
import sys

import torch
import torch.distributed as dist

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from primitive3d_b200 import capi, sharded  # noqa: E402
from bench import gyroid_cuda  # noqa: E402


def main():
    n = int(sys.argv[1]) if len(sys.argv) > 1 else 256
    rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
    torch.cuda.set_device(local)
    dev = torch.device("cuda", local)
    dist.init_process_group("nccl", device_id=dev)
    x0, x1h = sharded.slab_with_halo(n, world, rank)
    slab = gyroid_cuda(n, x0, x1h, dev)
    out = sharded.marching_cubes_slab(slab, 0.0, x0, n)
    torch.cuda.synchronize()
    # the same shard through the single C entry p3d_mc_sharded_extract over a raw NCCL communicator
    comm = sharded.nccl_comm_init()
    out_c = sharded.marching_cubes_slab_c(slab, 0.0, x0, n, comm, rank, world)
    torch.cuda.synchronize()
    assert (out_c.v_offset, out_c.f_offset, out_c.num_vertices_total, out_c.num_faces_total) == \
        (out.v_offset, out.f_offset, out.num_vertices_total, out.num_faces_total)
    assert torch.equal(out_c.vertices.view(torch.int32), out.vertices.view(torch.int32)) and torch.equal(out_c.faces, out.faces)
    sharded.nccl_comm_destroy(comm)
    # ... and with the exchange over peer memory (p3d_mc_sharded_extract_p2p: no collective between the passes),
    # several calls in a row (the mailbox alternates between two buffers)
    peer = sharded.PeerExchange(n, n)
    for it in range(5):
        out_p = sharded.marching_cubes_slab_p2p(slab, 0.0, x0, n, peer)
        torch.cuda.synchronize()
        assert (out_p.v_offset, out_p.f_offset, out_p.num_vertices_total, out_p.num_faces_total) == \
            (out.v_offset, out.f_offset, out.num_vertices_total, out.num_faces_total), f"p2p call {it}: offsets differ"
        assert torch.equal(out_p.vertices.view(torch.int32), out.vertices.view(torch.int32)) and torch.equal(out_p.faces, out.faces), \
            f"p2p call {it}: mesh differs"
    peer.close()
    # numbering-independent checksums of the sharded mesh (primitive3d_b200/verify.py), taken shard-wise
    from primitive3d_b200 import verify
    sums = verify.mesh_checksums(out.vertices, out.faces, out.v_offset, out.f_offset, float(x0))
    # the shard from HOST memory, uploaded in parts overlapped with their tile passes: the same mesh, numbered part by part
    host = slab.cpu().pin_memory()
    hosted = sharded.marching_cubes_slab_host(host, 0.0, x0, n, parts=4)
    assert (hosted.v_offset, hosted.f_offset, hosted.num_vertices_total, hosted.num_faces_total) == \
        (out.v_offset, out.f_offset, out.num_vertices_total, out.num_faces_total)
    assert verify.mesh_checksums(hosted.vertices, hosted.faces, hosted.v_offset, hosted.f_offset, float(x0)) == sums, \
        "checksums of the host-pipelined shards differ"
    shards = [None] * world
    dist.all_gather_object(shards, (out.vertices.cpu(), out.faces.cpu(), out.v_offset, out.f_offset))
    if rank == 0:
        full = gyroid_cuda(n, 0, n, dev)
        v, f = capi.marching_cubes(full, 0.0)
        vs = torch.cat([s[0] for s in shards])
        fs = torch.cat([s[1] for s in shards])
        assert shards[1][2] == shards[0][0].shape[0] and shards[1][3] == shards[0][1].shape[0]
        assert (out.num_vertices_total, out.num_faces_total) == (v.shape[0], f.shape[0])
        assert torch.equal(vs.view(torch.int32), v.cpu().view(torch.int32)), "vertices differ"
        assert torch.equal(fs, f.cpu()), "faces differ"
        assert verify.mesh_checksums(v, f, single=True) == sums, "checksums of the sharded and the single-GPU mesh differ"
        print(f"sharded NCCL check OK: world={world} n={n} V={v.shape[0]} F={f.shape[0]} "
              f"(python driver == p3d_mc_sharded_extract == p3d_mc_sharded_extract_p2p == host-pipelined shards == single GPU; checksums {sums[0]:016x} {sums[1]:016x})")
    dist.barrier()
    dist.destroy_process_group()


if __name__ == "__main__":
    main()
