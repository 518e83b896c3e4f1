"""Checksums of a (possibly sharded) mesh that do not depend on how the vertices are numbered.

Used by bench.py and tools/check_sharded_nccl.py to compare the multi-GPU extraction of a grid with the
single-GPU extraction of the same grid at sizes where the oracle does not run: the vertex numbering of the two
differs (shard by shard against one tile order), the geometry must not.

  vertex checksum    sum over all vertices of a 64-bit hash of the vertex's three float32 bit patterns (a multiset
                     property: independent of order and numbering)
  triangle checksum  sum over all faces k, in their GLOBAL output order, of weight(k) * (h0 + 3 h1 + 7 h2), h_i the
                     hash of the face's i-th corner: sensitive to the order of the faces and of the corners in a face
                     (faces are emitted voxel-major, marching_cubes.cu:194-208), independent of the numbering

All arithmetic is int64 with wrap-around.  torch is used for device memory and the collectives only.
"""
import torch
import torch.distributed as dist

_M1, _M2, _M3 = -7046029254386353131, -4658895280553007687, -7723592293110705685  # odd 64-bit constants (as int64)


def vertex_hashes(vertices):
    """int64 [V]: hash of each vertex row's float32 bits."""
    b = vertices.contiguous().view(torch.int32).to(torch.int64) & 0xFFFFFFFF
    h = (b[:, 0] * _M1 + b[:, 1]) * _M2 + b[:, 2]
    h = (h ^ (h >> 29)) * _M3
    return h ^ (h >> 32)


def _wrap_sum(t):
    return int(t.sum(dtype=torch.int64).item())


def mesh_checksums(vertices, faces, v_offset=0, f_offset=0, first_plane_x=None, group=None, single=False, chunk=1 << 24):
    """-> (vertex_checksum, triangle_checksum) of the whole mesh, identical on every rank.

    vertices / faces: this rank's shard (faces hold GLOBAL vertex ids, this rank's vertices are ids
    [v_offset, v_offset + V_r)); f_offset: global index of faces[0]; first_plane_x: x coordinate of this rank's first
    plane (its in-plane vertices are the only ones a lower rank's faces may name).  A whole mesh on one device:
    leave the defaults; single=True does that inside a multi-rank job (no collective is entered)."""
    world = dist.get_world_size(group) if dist.is_initialized() and not single else 1
    rank = dist.get_rank(group) if world > 1 else 0
    dev = vertices.device
    vh = vertex_hashes(vertices)
    V = vh.shape[0]
    rid = rh = None
    if world > 1:
        # every rank publishes (global id, hash) of the vertices on its first plane; rank r reads rank r + 1's
        sel = torch.nonzero(vertices[:, 0] == float(first_plane_x)).reshape(-1) if rank > 0 else torch.empty(0, dtype=torch.int64, device=dev)
        n = torch.tensor([sel.numel()], dtype=torch.int64, device=dev)
        ns = torch.empty(world, dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(ns, n, group=group)
        cap = int(ns.max().item())
        mine = torch.zeros((2, max(cap, 1)), dtype=torch.int64, device=dev)
        mine[0, :sel.numel()] = sel + v_offset
        mine[1, :sel.numel()] = vh[sel]
        allp = torch.empty(world * 2 * max(cap, 1), dtype=torch.int64, device=dev)
        dist.all_gather_into_tensor(allp, mine.view(-1), group=group)
        allp = allp.view(world, 2, max(cap, 1))
        if rank + 1 < world:
            m = int(ns[rank + 1].item())
            order = torch.argsort(allp[rank + 1, 0, :m])
            rid, rh = allp[rank + 1, 0, :m][order], allp[rank + 1, 1, :m][order]
    tsum = 0
    F = faces.shape[0]
    for a in range(0, F, chunk):
        ids = faces[a:a + chunk].to(torch.int64)
        loc = ids - v_offset
        own = (loc >= 0) & (loc < V)
        h = torch.zeros_like(ids)
        h[own] = vh[loc[own]]
        if not bool(own.all()):
            if rid is None or rid.numel() == 0:
                raise AssertionError("a face names a vertex of another shard that is not on the next shard's first plane")
            far = ids[~own]
            pos = torch.searchsorted(rid, far).clamp_(max=rid.numel() - 1)
            if not bool((rid[pos] == far).all()):
                raise AssertionError("a face names a vertex of another shard that is not on the next shard's first plane")
            h[~own] = rh[pos]
        k = torch.arange(a, a + ids.shape[0], device=dev, dtype=torch.int64) + f_offset
        w = k % 65521 + 1
        tsum += _wrap_sum((h[:, 0] + h[:, 1] * 3 + h[:, 2] * 7) * w)
    sums = torch.tensor([_wrap_sum(vh), _to_i64(tsum)], dtype=torch.int64, device=dev)
    if world > 1:
        dist.all_reduce(sums, group=group)
    return int(sums[0].item()) & 0xFFFFFFFFFFFFFFFF, int(sums[1].item()) & 0xFFFFFFFFFFFFFFFF


def _to_i64(x):
    x &= 0xFFFFFFFFFFFFFFFF
    return x - (1 << 64) if x >= (1 << 63) else x
