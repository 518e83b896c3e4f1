"""Multi-GPU marching cubes: dim-0 slab decomposition, one process per GPU.

The reference is single-GPU; this is the sharding the path
admits naturally:

  * tensor dim 0 (the slowest axis of idx = i*(Ry*Rz) + j*Rz + k, marching_cubes.cu:20) is split
    into `world` contiguous plane ranges; rank r holds its planes plus ONE halo plane (the first
    plane of rank r+1), because the cells and +x edges of its last plane read it;
  * every rank runs the same kernels on its slab (no data-path collective);
  * the only exchange is tiny: an all-gather of {V_r, F_r} (16 bytes per rank) whose exclusive
    prefix gives each rank the global id of its first vertex / face, and an all-gather of each
    rank's first-plane piece table (16 bytes per 128 samples of one plane) so the cells next to
    a slab boundary can name the vertices the next rank owns;
  * outputs stay sharded: rank r returns its vertices and its faces, the faces holding GLOBAL
    vertex ids.  Concatenating the shards in rank order gives the single-GPU mesh: the same
    triangles in the same (voxel-major) order, over a vertex array numbered shard by shard.
"""
import ctypes
from dataclasses import dataclass

import torch
import torch.distributed as dist

from . import capi


def slab_range(global_rx, world, rank):
    """Planes [x0, x1) owned by `rank`: contiguous, sizes differ by at most one."""
    base, rem = divmod(int(global_rx), int(world))
    x0 = rank * base + min(rank, rem)
    return x0, x0 + base + (1 if rank < rem else 0)


def slab_with_halo(global_rx, world, rank):
    """Planes [x0, x1_halo) a rank must hold in memory: its own plus one halo plane (none for the last)."""
    x0, x1 = slab_range(global_rx, world, rank)
    return x0, min(x1 + 1, int(global_rx))


def exclusive_offsets(counts, rank):
    """counts: sequence of (V_r, F_r) per rank -> (v_offset, f_offset, V_total, F_total) for `rank`."""
    v_off = sum(int(c[0]) for c in counts[:rank])
    f_off = sum(int(c[1]) for c in counts[:rank])
    return v_off, f_off, sum(int(c[0]) for c in counts), sum(int(c[1]) for c in counts)


def gather_counts(V, F, device, group=None):
    """All-gather of the per-rank {V, F} (int64 x 2).  Works on NCCL (CUDA tensors) and gloo (CPU)."""
    world = dist.get_world_size(group)
    mine = torch.tensor([int(V), int(F)], dtype=torch.int64, device=device)
    out = torch.empty(world * 2, dtype=torch.int64, device=device)
    dist.all_gather_into_tensor(out, mine, group=group)
    return [tuple(r) for r in out.view(world, 2).cpu().tolist()]


def exchange_counts_and_tables(V, F, table, group=None):
    """ONE all-gather per extraction: every rank contributes its first-plane piece table (int32 [words], on the
    compute device for NCCL, CPU for gloo) with its {V, F} appended as two int64.  Returns (counts as a list of
    (V_r, F_r), tables int32 [world, words]).  Reading the counts back is the only synchronisation."""
    world = dist.get_world_size(group)
    words = table.numel()
    mine = torch.empty(words + 4, dtype=torch.int32, device=table.device)
    mine[:words] = table
    mine[words:] = torch.tensor([int(V), int(F)], dtype=torch.int64).view(torch.int32).to(table.device, non_blocking=True)
    out = torch.empty(world * (words + 4), dtype=torch.int32, device=table.device)
    dist.all_gather_into_tensor(out, mine, group=group)
    out = out.view(world, words + 4)
    counts = out[:, words:].contiguous().view(torch.int64).cpu().tolist()
    return [tuple(c) for c in counts], out[:, :words]


def unpack_counts(gathered, world, words):
    """(V_r, F_r) of every rank from the gathered exchange payloads (int32 [world * words], the last four words
    of each payload are the two int64 counts).  The device-to-host copy is the extraction's only synchronisation."""
    tail = gathered.view(world, words)[:, words - 4:].contiguous().view(torch.int64).cpu().tolist()
    return [tuple(c) for c in tail]


_last_vertex_count = {}   # slab shape -> V of its last extraction (sizes the speculative vertex buffer)
_last_face_count = {}     # slab shape -> F of its last extraction (single GPU: sizes the speculative face buffer)


def capacity_for(shape):
    """Vertex capacity for the counting pass: the previous V of this slab shape plus 1/16, or None (= the
    library's hint, samples / 16) the first time.  Too small only costs a second vertices-only pass."""
    v = _last_vertex_count.get(tuple(int(s) for s in shape))
    return None if v is None else min(v + v // 16 + 4096, 2 ** 31 - 1)


@dataclass
class SlabMesh:
    vertices: torch.Tensor      # float32 [V_r, 3], this rank's vertices
    faces: torch.Tensor         # int32 [F_r, 3], GLOBAL vertex ids
    v_offset: int               # global id of vertices[0]
    f_offset: int               # global index of faces[0]
    num_vertices_total: int
    num_faces_total: int


def marching_cubes_slab(slab, thresh, x_begin, global_rx, lower=None, upper=None, group=None, vertex_capacity=None):
    """Extract this rank's shard.

    slab: contiguous float32 CUDA tensor holding planes [x_begin, x_begin + slab.shape[0]) of the
    global grid, i.e. the rank's own planes plus one halo plane unless it is the last rank.
    """
    world = dist.get_world_size(group) if dist.is_initialized() else 1
    rank = dist.get_rank(group) if dist.is_initialized() else 0
    x0, x1 = slab_range(global_rx, world, rank)
    if x0 != x_begin or slab.shape[0] != min(x1 + 1, global_rx) - x0:
        raise ValueError(f"rank {rank}: slab must hold planes [{x0}, {min(x1 + 1, global_rx)})")
    owned = x1 - x0
    ry, rz = slab.shape[1], slab.shape[2]
    lower = [0.0, 0.0, 0.0] if lower is None else lower
    upper = [float(global_rx), float(ry), float(rz)] if upper is None else upper
    desc = capi.McDesc.make(slab.shape, thresh, lower, upper, owned_x=owned, x_origin=x0, global_rx=global_rx)

    if vertex_capacity is None:
        vertex_capacity = capacity_for(slab.shape)
    key = tuple(int(s) for s in slab.shape)
    if world == 1:
        # one GPU: nothing to exchange, so both passes are queued at once and the host waits a single time
        f_prev = _last_face_count.get(key)
        f_cap = None if f_prev is None else f_prev + f_prev // 16 + 4096
        verts, faces, V, F = capi.mc_extract(desc, slab, vertex_capacity, f_cap)
        if len(_last_vertex_count) > 64:
            _last_vertex_count.clear()
            _last_face_count.clear()
        _last_vertex_count[key], _last_face_count[key] = V, F
        return SlabMesh(verts, faces, 0, 0, V, F)
    # Several GPUs: tile pass, exchange and face pass are all queued; the host waits once, for the gathered counts.
    # The exchange payload (first-plane numbering table + {V, F}) is written, gathered and consumed on the device:
    # the vertex id base and the halo-plane numbering never pass through the host.
    L = capi.lib()
    dtype = capi._grid_ok(slab)
    dev = slab.device
    ws_bytes, hint = capi._desc_sizes(desc)
    if vertex_capacity is None:
        vertex_capacity = hint
    f_prev = _last_face_count.get(key)
    face_capacity = 2 * int(vertex_capacity) if f_prev is None else f_prev + f_prev // 16 + 4096
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    vbuf = torch.empty((int(vertex_capacity), 3), dtype=torch.float32, device=dev)
    fbuf = torch.empty((int(face_capacity), 3), dtype=torch.int32, device=dev)
    words = L.p3d_mc_exchange_words(ctypes.byref(desc))
    mine = torch.empty(words, dtype=torch.int32, device=dev)
    gathered = torch.empty(world * words, dtype=torch.int32, device=dev)
    with capi._on_device(dev):
        stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
        capi.check(L.p3d_mc_tile_async(ctypes.byref(desc), slab.data_ptr(), dtype, ws.data_ptr(), ws.numel(),
                                       vbuf.data_ptr() if vertex_capacity else None, int(vertex_capacity), stream))
        capi.check(L.p3d_mc_export_exchange(ctypes.byref(desc), ws.data_ptr(), mine.data_ptr(), stream))
        dist.all_gather_into_tensor(gathered, mine, group=group)
        capi.check(L.p3d_mc_faces_exchanged(ctypes.byref(desc), ws.data_ptr(), gathered.data_ptr(), rank, world,
                                            fbuf.data_ptr() if face_capacity else None, int(face_capacity), stream))
    counts = unpack_counts(gathered, world, words)          # the only synchronisation
    V, F = counts[rank]
    v_off, f_off, v_tot, f_tot = exclusive_offsets(counts, rank)
    if v_tot > 2 ** 31 - 1:
        raise OverflowError("global vertex count exceeds the int32 face-index contract")
    if len(_last_vertex_count) > 64:
        _last_vertex_count.clear()
        _last_face_count.clear()
    _last_vertex_count[key], _last_face_count[key] = V, F
    verts = capi.mc_vertices(desc, slab, ws, V, vbuf)       # exact-size second pass only if the guess was too small
    faces = fbuf[:F] if F <= face_capacity else capi.mc_faces(desc, ws, F, v_off)   # halo numbering is installed
    return SlabMesh(verts, faces, v_off, f_off, v_tot, f_tot)


# ---------------------------------------------------------------------------------------------
# A shard that lives in HOST memory: uploaded part by part while the parts already on the device are extracted.
# ---------------------------------------------------------------------------------------------
_part_counts = {}   # (slab shape, parts) -> [V_k] of the last extraction (sizes the parts' speculative vertex segments)


def _part_bounds(owned, parts):
    """Plane ranges [a, b) of the parts of a shard: equal, multiples of 8 planes (the tile depth) but the last."""
    step = max(8, -(-owned // max(1, parts)) + 7 & ~7)
    edges = list(range(0, owned, step)) + [owned]
    if len(edges) > 2 and edges[-1] - edges[-2] < 8:
        del edges[-2]          # a sliver at the end joins the part before it (a part needs two planes at least)
    return list(zip(edges[:-1], edges[1:]))


def marching_cubes_slab_host(host_slab, thresh, x_begin, global_rx, lower=None, upper=None, group=None, parts=None,
                             out_vertices=None, out_faces=None, device=None, distributed=True):
    """marching_cubes_slab for a shard in HOST memory (pinned for full PCIe speed), upload overlapped with extraction.

    The shard is cut into `parts` plane ranges that behave as consecutive shards of a world of `world * parts`:
    while part k + 1 uploads (a second stream, two device buffers), part k runs its tile pass and exports its exchange
    payload; ONE all-gather then carries every part's payload, and the face passes of all parts run against it (they
    read the workspaces only, the grid data has left the device by then).  The host waits once, for the gathered
    counts.  Numbering: part by part within the rank, rank by rank (concatenated in rank order the shards are the
    single-GPU mesh of the whole grid over a vertex array numbered part by part).

    out_vertices / out_faces: optional pinned CPU tensors ([>= V_r, 3] float32, [>= F_r, 3] int32) that receive the
    shard; the returned SlabMesh then holds views of them (complete on return).  Otherwise device tensors.
    distributed=False: the slab is the whole grid whatever process group exists (prim3d.marching_cubes on a CPU tensor)."""
    world = dist.get_world_size(group) if distributed and dist.is_initialized() else 1
    rank = dist.get_rank(group) if distributed and dist.is_initialized() else 0
    if host_slab.is_cuda or not host_slab.is_contiguous() or host_slab.dim() != 3 or host_slab.dtype not in capi.GRID_DTYPES:
        raise ValueError("host_slab must be a contiguous CPU tensor [planes, Ry, Rz] of a supported dtype")
    x0, x1 = slab_range(global_rx, world, rank)
    if x0 != x_begin or host_slab.shape[0] != min(x1 + 1, global_rx) - x0:
        raise ValueError(f"rank {rank}: slab must hold planes [{x0}, {min(x1 + 1, global_rx)})")
    dev = torch.device("cuda", torch.cuda.current_device()) if device is None else torch.device(device)
    owned, planes = x1 - x0, host_slab.shape[0]
    ry, rz = int(host_slab.shape[1]), int(host_slab.shape[2])
    lower = [0.0, 0.0, 0.0] if lower is None else lower
    upper = [float(global_rx), float(ry), float(rz)] if upper is None else upper
    if parts is None:
        parts = max(1, min(16, owned // 32))
    bounds = _part_bounds(owned, parts)
    parts = len(bounds)
    L, dtype = capi.lib(), capi.GRID_DTYPES[host_slab.dtype]
    descs, sizes = [], []
    for a, b in bounds:
        held = min(b + 1, planes) - a                      # the part's planes plus the halo plane, if there is one
        d = capi.McDesc.make((held, ry, rz), thresh, lower, upper, owned_x=b - a, x_origin=x0 + a, global_rx=global_rx)
        descs.append(d)
        sizes.append(capi._desc_sizes(d))
    key = (tuple(int(v) for v in host_slab.shape), parts)
    for attempt in range(2):
        prev = _part_counts.get(key)
        caps = [min(v + v // 16 + 4096, 2 ** 31 - 1) for v in prev] if prev else [sz[1] for sz in sizes]
        with torch.cuda.device(dev):
            cur = torch.cuda.current_stream()
            up = torch.cuda.Stream()
            depth = max(min(b + 1, planes) - a for a, b in bounds)
            bufs = [torch.empty((depth, ry, rz), dtype=host_slab.dtype, device=dev) for _ in range(min(2, parts))]
            wss = [torch.empty(sz[0], dtype=torch.uint8, device=dev) for sz in sizes]
            seg = [0]
            for c in caps:
                seg.append(seg[-1] + int(c))
            vbuf = torch.empty((seg[-1], 3), dtype=torch.float32, device=dev)
            words = L.p3d_mc_exchange_words(ctypes.byref(descs[0]))
            payload = torch.empty(parts * words, dtype=torch.int32, device=dev)
            stream = ctypes.c_void_p(cur.cuda_stream)
            up.wait_stream(cur)
            released = [None] * len(bufs)
            for k, (a, b) in enumerate(bounds):
                held = min(b + 1, planes) - a
                buf = bufs[k % len(bufs)][:held]
                with torch.cuda.stream(up):
                    if released[k % len(bufs)] is not None:
                        up.wait_event(released[k % len(bufs)])   # the tile pass that read this buffer is over
                    buf.copy_(host_slab[a:a + held], non_blocking=True)
                    arrived = torch.cuda.Event()
                    arrived.record(up)
                cur.wait_event(arrived)
                capi.check(L.p3d_mc_tile_async(ctypes.byref(descs[k]), buf.data_ptr(), dtype, wss[k].data_ptr(), wss[k].numel(),
                                               vbuf[seg[k]:].data_ptr() if caps[k] else None, int(caps[k]), stream))
                capi.check(L.p3d_mc_export_exchange(ctypes.byref(descs[k]), wss[k].data_ptr(),
                                                    payload[k * words:].data_ptr(), stream))
                released[k % len(bufs)] = torch.cuda.Event()
                released[k % len(bufs)].record(cur)
            if world > 1:
                gathered = torch.empty(world * parts * words, dtype=torch.int32, device=dev)
                dist.all_gather_into_tensor(gathered, payload, group=group)
            else:
                gathered = payload
            counts = unpack_counts(gathered, world * parts, words)      # the only synchronisation
            mine = counts[rank * parts:(rank + 1) * parts]
            if len(_part_counts) > 64:
                _part_counts.clear()
            _part_counts[key] = [int(c[0]) for c in mine]
            if any(int(c[0]) > cap for c, cap in zip(mine, caps)):
                if attempt == 0:
                    continue      # a segment was too small (first call on a busy field): once more with the counts known
                raise capi.P3DError(capi.P3D_ERR_INVALID, "marching_cubes_slab_host: counts changed between two passes")
            v_off, f_off, v_tot, f_tot = exclusive_offsets(counts, rank * parts)
            if v_tot > 2 ** 31 - 1:
                raise OverflowError("global vertex count exceeds the int32 face-index contract")
            V_r, F_r = sum(int(c[0]) for c in mine), sum(int(c[1]) for c in mine)
            fbuf = torch.empty((F_r, 3), dtype=torch.int32, device=dev)
            f_at = 0
            for k in range(parts):
                F_k = int(mine[k][1])
                capi.check(L.p3d_mc_faces_exchanged(ctypes.byref(descs[k]), wss[k].data_ptr(), gathered.data_ptr(), rank * parts + k,
                                                    world * parts, fbuf[f_at:].data_ptr() if F_k else None, F_k, stream))
                f_at += F_k
            pieces = [vbuf[seg[k]:seg[k] + int(mine[k][0])] for k in range(parts)]
            if out_vertices is not None or out_faces is not None:
                if out_vertices is None or out_faces is None or out_vertices.shape[0] < V_r or out_faces.shape[0] < F_r:
                    raise ValueError("out_vertices / out_faces must both be given and hold the shard")
                at = 0
                for piece in pieces:
                    out_vertices[at:at + piece.shape[0]].copy_(piece, non_blocking=True)
                    at += piece.shape[0]
                out_faces[:F_r].copy_(fbuf, non_blocking=True)
                cur.synchronize()
                return SlabMesh(out_vertices[:V_r], out_faces[:F_r], v_off, f_off, v_tot, f_tot)
            verts = torch.cat(pieces) if parts > 1 else pieces[0]
            return SlabMesh(verts, fbuf, v_off, f_off, v_tot, f_tot)


# ---------------------------------------------------------------------------------------------
# The same extraction through the single C entry p3d_mc_sharded_extract, over a raw NCCL communicator
# (what a C / C++ host would do; torch.distributed only carries the unique id to the other ranks).
# ---------------------------------------------------------------------------------------------
class _NcclUniqueId(ctypes.Structure):
    _fields_ = [("internal", ctypes.c_byte * 128)]


def _nccl():
    """The libnccl.so.2 torch itself uses (already loaded once torch.distributed's NCCL backend is up)."""
    for name in ("libnccl.so.2",):
        try:
            return ctypes.CDLL(name, mode=ctypes.RTLD_GLOBAL)
        except OSError:
            pass
    import glob
    import os
    hits = glob.glob(os.path.join(os.path.dirname(torch.__file__), "..", "nvidia", "nccl", "lib", "libnccl.so.2"))
    if not hits:
        raise ImportError("libnccl.so.2 not found")
    return ctypes.CDLL(hits[0], mode=ctypes.RTLD_GLOBAL)


def nccl_comm_init(group=None):
    """A raw ncclComm_t (as an int) spanning the ranks of `group`, on the current CUDA device."""
    nccl = _nccl()
    rank, world = dist.get_rank(group), dist.get_world_size(group)
    uid = _NcclUniqueId()
    if rank == 0:
        if nccl.ncclGetUniqueId(ctypes.byref(uid)) != 0:
            raise RuntimeError("ncclGetUniqueId failed")
    box = [bytes(uid.internal)]
    dist.broadcast_object_list(box, src=0, group=group)
    ctypes.memmove(ctypes.byref(uid), box[0], 128)
    comm = ctypes.c_void_p()
    nccl.ncclCommInitRank.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_int, _NcclUniqueId, ctypes.c_int]
    if nccl.ncclCommInitRank(ctypes.byref(comm), world, uid, rank) != 0:
        raise RuntimeError("ncclCommInitRank failed")
    return comm.value


def nccl_comm_destroy(comm):
    nccl = _nccl()
    nccl.ncclCommDestroy.argtypes = [ctypes.c_void_p]
    nccl.ncclCommDestroy(ctypes.c_void_p(comm))


class PeerExchange:
    """This rank's mailbox for the shard-boundary exchange over peer memory (p3d_mc_peer_*): created collectively by
    every rank of `group` for shards whose planes are ry x rz; the IPC handles travel through one all-gather of
    torch.distributed (set-up only: the extraction calls that follow make no collective call).  close() is collective."""

    def __init__(self, ry, rz, thresh=0.0, group=None):
        self.group = group
        self.rank, self.world = dist.get_rank(group), dist.get_world_size(group)
        L = capi.lib()
        desc = capi.McDesc.make((8, int(ry), int(rz)), thresh)   # only the plane matters
        nbytes = L.p3d_mc_peer_handle_bytes()
        mine = (ctypes.c_ubyte * nbytes)()
        self.ptr = ctypes.c_void_p()
        dev = torch.device("cuda", torch.cuda.current_device())
        with capi._on_device(dev):
            capi.check(L.p3d_mc_peer_create(ctypes.byref(desc), self.rank, self.world, ctypes.byref(self.ptr), mine))
        send = torch.frombuffer(bytearray(bytes(mine)), dtype=torch.uint8).to(dev)
        recv = torch.empty(self.world * nbytes, dtype=torch.uint8, device=dev)
        dist.all_gather_into_tensor(recv, send, group=group)
        handles = recv.cpu().numpy().tobytes()
        with capi._on_device(dev):
            capi.check(L.p3d_mc_peer_connect(self.ptr, handles))
        dist.barrier(group=group)   # every rank has mapped every mailbox before anybody stores into one
        self.plane = (int(ry), int(rz))

    def close(self):
        if self.ptr:
            torch.cuda.synchronize()
            dist.barrier(group=self.group)   # nobody is inside a call any more
            capi.lib().p3d_mc_peer_disconnect(self.ptr)
            torch.cuda.synchronize()
            dist.barrier(group=self.group)   # nobody maps anybody's mailbox any more: now they can be freed
            capi.lib().p3d_mc_peer_destroy(self.ptr)
            self.ptr = ctypes.c_void_p()


def marching_cubes_slab_p2p(slab, thresh, x_begin, global_rx, peer, vertex_capacity=None, face_capacity=None):
    """marching_cubes_slab through ONE C call with the exchange over peer memory (p3d_mc_sharded_extract_p2p): tile
    pass, NVLink stores into the neighbours' mailboxes, a flag wait, face pass; no NCCL call.  Collective over the
    ranks of `peer` (a PeerExchange for this plane size)."""
    rank, world = peer.rank, peer.world
    x0, x1 = slab_range(global_rx, world, rank)
    if x0 != x_begin or slab.shape[0] != min(x1 + 1, global_rx) - x0:
        raise ValueError(f"rank {rank}: slab must hold planes [{x0}, {min(x1 + 1, global_rx)})")
    ry, rz = slab.shape[1], slab.shape[2]
    if (ry, rz) != peer.plane:
        raise ValueError("the PeerExchange was made for another plane size")
    desc = capi.McDesc.make(slab.shape, thresh, [0.0, 0.0, 0.0], [float(global_rx), float(ry), float(rz)], owned_x=x1 - x0,
                            x_origin=x0, global_rx=global_rx)
    L = capi.lib()
    dev = slab.device
    ws_bytes, hint = capi._desc_sizes(desc)
    key = tuple(int(s) for s in slab.shape)
    if vertex_capacity is None:
        vertex_capacity = capacity_for(slab.shape)
    vertex_capacity = hint if vertex_capacity is None else int(vertex_capacity)
    if face_capacity is None:
        f_prev = _last_face_count.get(key)
        face_capacity = 2 * vertex_capacity if f_prev is None else f_prev + f_prev // 16 + 4096
    face_capacity = int(face_capacity)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    vbuf = torch.empty((vertex_capacity, 3), dtype=torch.float32, device=dev)
    fbuf = torch.empty((face_capacity, 3), dtype=torch.int32, device=dev)
    counts = (ctypes.c_int64 * (2 * world))()
    with capi._on_device(dev):
        capi.check(L.p3d_mc_sharded_extract_p2p(ctypes.byref(desc), slab.data_ptr(), capi._grid_ok(slab), ws.data_ptr(), ws.numel(),
                                                peer.ptr, vbuf.data_ptr(), vertex_capacity, fbuf.data_ptr(), face_capacity, counts,
                                                ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    pairs = [(counts[2 * r], counts[2 * r + 1]) for r in range(world)]
    V, F = pairs[rank]
    if len(_last_vertex_count) > 64:
        _last_vertex_count.clear()
        _last_face_count.clear()
    _last_vertex_count[key], _last_face_count[key] = V, F
    v_off, f_off, v_tot, f_tot = exclusive_offsets(pairs, rank)
    verts = capi.mc_vertices(desc, slab, ws, V, vbuf)
    faces = fbuf[:F] if F <= face_capacity else capi.mc_faces(desc, ws, F, v_off)
    return SlabMesh(verts, faces, v_off, f_off, v_tot, f_tot)


def marching_cubes_slab_c(slab, thresh, x_begin, global_rx, comm, rank, world, vertex_capacity=None, face_capacity=None):
    """marching_cubes_slab through ONE C call (p3d_mc_sharded_extract) over the raw communicator `comm`."""
    x0, x1 = slab_range(global_rx, world, rank)
    if x0 != x_begin or slab.shape[0] != min(x1 + 1, global_rx) - x0:
        raise ValueError(f"rank {rank}: slab must hold planes [{x0}, {min(x1 + 1, global_rx)})")
    ry, rz = slab.shape[1], slab.shape[2]
    desc = capi.McDesc.make(slab.shape, thresh, [0.0, 0.0, 0.0], [float(global_rx), float(ry), float(rz)], owned_x=x1 - x0,
                            x_origin=x0, global_rx=global_rx)
    L = capi.lib()
    dev = slab.device
    ws_bytes, hint = capi._desc_sizes(desc)
    key = tuple(int(s) for s in slab.shape)
    if vertex_capacity is None:
        vertex_capacity = capacity_for(slab.shape)
    vertex_capacity = hint if vertex_capacity is None else int(vertex_capacity)
    if face_capacity is None:
        f_prev = _last_face_count.get(key)
        face_capacity = 2 * vertex_capacity if f_prev is None else f_prev + f_prev // 16 + 4096
    face_capacity = int(face_capacity)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=dev)
    vbuf = torch.empty((vertex_capacity, 3), dtype=torch.float32, device=dev)
    fbuf = torch.empty((face_capacity, 3), dtype=torch.int32, device=dev)
    words = L.p3d_mc_exchange_words(ctypes.byref(desc))
    send = torch.empty(words, dtype=torch.int32, device=dev)
    recv = torch.empty(world * words, dtype=torch.int32, device=dev)
    counts = (ctypes.c_int64 * (2 * world))()
    with capi._on_device(dev):
        capi.check(L.p3d_mc_sharded_extract(ctypes.byref(desc), slab.data_ptr(), capi._grid_ok(slab), ws.data_ptr(), ws.numel(),
                                            ctypes.c_void_p(comm), rank, world, send.data_ptr(), recv.data_ptr(),
                                            vbuf.data_ptr(), vertex_capacity, fbuf.data_ptr(), face_capacity, counts,
                                            ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)))
    pairs = [(counts[2 * r], counts[2 * r + 1]) for r in range(world)]
    V, F = pairs[rank]
    if len(_last_vertex_count) > 64:
        _last_vertex_count.clear()
        _last_face_count.clear()
    _last_vertex_count[key], _last_face_count[key] = V, F
    v_off, f_off, v_tot, f_tot = exclusive_offsets(pairs, rank)
    verts = capi.mc_vertices(desc, slab, ws, V, vbuf)
    faces = fbuf[:F] if F <= face_capacity else capi.mc_faces(desc, ws, F, v_off)
    return SlabMesh(verts, faces, v_off, f_off, v_tot, f_tot)
