"""ctypes bindings of include/prim3d_b200.h over torch CUDA tensors.

This is the C-ABI call path the parity tests use (`tests/ -m gpu`) and the one the multi-GPU
driver (sharded.py) is built on.  torch is used for device memory and streams only.
"""
import ctypes
import os

import torch

_HERE = os.path.dirname(os.path.abspath(__file__))
# P3D_CORE_LIB: tuning runs (tools/variants.py) point the ctypes view at another build of the same library
LIB_PATH = os.environ.get("P3D_CORE_LIB") or os.path.join(_HERE, "libprim3d_b200.so")
_lib = None

P3D_OK, P3D_ERR_INVALID, P3D_ERR_CUDA, P3D_ERR_OVERFLOW, P3D_ERR_WORKSPACE = range(5)


class McDesc(ctypes.Structure):
    """struct p3d_mc_desc (include/prim3d_b200.h)."""
    _fields_ = [("rx", ctypes.c_int64), ("ry", ctypes.c_int64), ("rz", ctypes.c_int64),
                ("owned_x", ctypes.c_int64), ("x_origin", ctypes.c_int64), ("global_rx", ctypes.c_int64),
                ("thresh", ctypes.c_float), ("lower", ctypes.c_float * 3), ("upper", ctypes.c_float * 3)]

    @classmethod
    def make(cls, shape, thresh, lower=None, upper=None, owned_x=None, x_origin=0, global_rx=None):
        rx, ry, rz = (int(s) for s in shape)
        global_rx = rx if global_rx is None else int(global_rx)
        lower = [0.0, 0.0, 0.0] if lower is None else lower
        upper = [float(global_rx), float(ry), float(rz)] if upper is None else upper
        return cls(rx, ry, rz, rx if owned_x is None else int(owned_x), int(x_origin), global_rx,
                   float(thresh), (ctypes.c_float * 3)(*[float(v) for v in lower]),
                   (ctypes.c_float * 3)(*[float(v) for v in upper]))


class P3DError(RuntimeError):
    def __init__(self, status, message):
        super().__init__(f"prim3d_b200 status {status}: {message}")
        self.status = status


def lib():
    """Load libprim3d_b200.so; there is no fallback if it is missing."""
    global _lib
    if _lib is None:
        if not os.path.exists(LIB_PATH):
            raise ImportError(f"{LIB_PATH} is not built; run `python -m primitive3d_b200.build`")
        L = ctypes.CDLL(LIB_PATH)
        vp, i64, sz = ctypes.c_void_p, ctypes.c_int64, ctypes.c_size_t
        dp = ctypes.POINTER(McDesc)
        L.p3d_abi_version.restype = ctypes.c_int
        L.p3d_last_error.restype = ctypes.c_char_p
        L.p3d_mc_workspace_bytes.restype = sz
        L.p3d_mc_workspace_bytes.argtypes = [dp]
        L.p3d_mc_vertex_capacity_hint.restype = i64
        L.p3d_mc_vertex_capacity_hint.argtypes = [dp]
        L.p3d_mc_plane_table_words.restype = i64
        L.p3d_mc_plane_table_words.argtypes = [dp]
        L.p3d_mc_count.restype = ctypes.c_int
        L.p3d_mc_count.argtypes = [dp, vp, vp, sz, vp, i64, ctypes.POINTER(i64), vp]
        L.p3d_mc_vertices.restype = ctypes.c_int
        L.p3d_mc_vertices.argtypes = [dp, vp, vp, vp, i64, vp]
        L.p3d_mc_count_typed.restype = ctypes.c_int
        L.p3d_mc_count_typed.argtypes = [dp, vp, ctypes.c_int, vp, sz, vp, i64, ctypes.POINTER(i64), vp]
        L.p3d_mc_vertices_typed.restype = ctypes.c_int
        L.p3d_mc_vertices_typed.argtypes = [dp, vp, ctypes.c_int, vp, vp, i64, vp]
        L.p3d_mc_extract.restype = ctypes.c_int
        L.p3d_mc_extract.argtypes = [dp, vp, ctypes.c_int, vp, sz, vp, i64, vp, i64, ctypes.POINTER(i64), vp]
        L.p3d_mc_tile_async.restype = ctypes.c_int
        L.p3d_mc_tile_async.argtypes = [dp, vp, ctypes.c_int, vp, sz, vp, i64, vp]
        L.p3d_mc_exchange_words.restype = i64
        L.p3d_mc_exchange_words.argtypes = [dp]
        L.p3d_mc_export_exchange.restype = ctypes.c_int
        L.p3d_mc_export_exchange.argtypes = [dp, vp, vp, vp]
        L.p3d_mc_faces_exchanged.restype = ctypes.c_int
        L.p3d_mc_faces_exchanged.argtypes = [dp, vp, vp, ctypes.c_int, ctypes.c_int, vp, i64, vp]
        L.p3d_mc_sharded_extract.restype = ctypes.c_int
        L.p3d_mc_sharded_extract.argtypes = [dp, vp, ctypes.c_int, vp, sz, vp, ctypes.c_int, ctypes.c_int, vp, vp, vp, i64, vp, i64,
                                             ctypes.POINTER(i64), vp]
        L.p3d_mc_peer_handle_bytes.restype = sz
        L.p3d_mc_peer_create.restype = ctypes.c_int
        L.p3d_mc_peer_create.argtypes = [dp, ctypes.c_int, ctypes.c_int, ctypes.POINTER(vp), vp]
        L.p3d_mc_peer_connect.restype = ctypes.c_int
        L.p3d_mc_peer_connect.argtypes = [vp, vp]
        L.p3d_mc_peer_disconnect.restype = None
        L.p3d_mc_peer_disconnect.argtypes = [vp]
        L.p3d_mc_peer_destroy.restype = None
        L.p3d_mc_peer_destroy.argtypes = [vp]
        L.p3d_mc_sharded_extract_p2p.restype = ctypes.c_int
        L.p3d_mc_sharded_extract_p2p.argtypes = [dp, vp, ctypes.c_int, vp, sz, vp, vp, i64, vp, i64, ctypes.POINTER(i64), vp]
        L.p3d_mc_extract_sparse.restype = ctypes.c_int
        L.p3d_mc_extract_sparse.argtypes = [dp, vp, ctypes.c_int, vp, i64, vp, sz, vp, i64, vp, i64, ctypes.POINTER(i64), vp]
        L.p3d_mc_extract_host.restype = ctypes.c_int
        L.p3d_mc_extract_host.argtypes = [dp, vp, ctypes.c_int, i64, vp, i64, vp, i64, ctypes.POINTER(i64), vp, sz]
        L.p3d_mc_extract_host_arena_bytes.restype = sz
        L.p3d_mc_extract_host_arena_bytes.argtypes = [dp, ctypes.c_int, i64]
        L.p3d_mc_single_launch.restype = ctypes.c_int
        L.p3d_mc_single_launch.argtypes = [dp, ctypes.c_int]
        L.p3d_mc_batch_workspace_bytes.restype = sz
        L.p3d_mc_batch_workspace_bytes.argtypes = [i64, vp]
        L.p3d_mc_extract_batch.restype = ctypes.c_int
        L.p3d_mc_extract_batch.argtypes = [i64, vp, vp, ctypes.c_int, vp, sz, vp, vp, vp, vp, vp, vp]
        L.p3d_mc_faces.restype = ctypes.c_int
        L.p3d_mc_faces.argtypes = [dp, vp, vp, i64, vp]
        L.p3d_mc_debug_stage.restype = ctypes.c_int
        L.p3d_mc_debug_stage.argtypes = [dp, vp, vp, ctypes.c_int, vp, i64, vp]
        L.p3d_mc_export_first_plane.restype = ctypes.c_int
        L.p3d_mc_export_first_plane.argtypes = [dp, vp, vp, vp]
        L.p3d_mc_import_halo_plane.restype = ctypes.c_int
        L.p3d_mc_import_halo_plane.argtypes = [dp, vp, vp, i64, vp]
        _bind_mt(L, vp, i64, sz)
        _lib = L
    return _lib


def _bind_mt(L, vp, i64, sz):
    pi64 = ctypes.POINTER(i64)
    L.p3d_mt_codes_bytes.restype = sz
    L.p3d_mt_codes_bytes.argtypes = [i64]
    L.p3d_mt_classify.restype = ctypes.c_int
    L.p3d_mt_classify.argtypes = [vp, i64, vp, i64, vp, ctypes.c_int, vp, pi64, vp]
    L.p3d_mt_workspace_bytes.restype = sz
    L.p3d_mt_workspace_bytes.argtypes = [i64, i64, i64, i64]
    L.p3d_mt_index.restype = ctypes.c_int
    L.p3d_mt_index.argtypes = [vp, i64, i64, vp, vp, i64, i64, i64, vp, sz, pi64, vp]
    L.p3d_mt_emit.restype = ctypes.c_int
    L.p3d_mt_emit.argtypes = [vp, vp, i64, vp, vp, i64, i64, i64, i64, vp, vp, vp, vp, vp, vp]
    L.p3d_mt_extract_workspace_bytes.restype = sz
    L.p3d_mt_extract_workspace_bytes.argtypes = [i64, i64, i64, i64]
    L.p3d_mt_extract.restype = ctypes.c_int
    L.p3d_mt_extract.argtypes = [vp, i64, vp, i64, vp, ctypes.c_int, vp, sz, i64, i64, vp, vp, i64, vp, vp, i64, pi64, vp]
    L.p3d_mt_backward.restype = ctypes.c_int
    L.p3d_mt_backward.argtypes = [vp, vp, vp, i64, vp, vp, vp, vp]


def abi_version():
    return lib().p3d_abi_version()


def check(status):
    if status != P3D_OK:
        raise P3DError(status, lib().p3d_last_error().decode())


def _stream():
    return ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)


# torch dtype -> p3d_dtype (include/prim3d_b200.h): element types the tile pass reads directly
GRID_DTYPES = {torch.float32: 0, torch.float16: 1, torch.bfloat16: 2, torch.float64: 3, torch.int64: 4, torch.int32: 5,
               torch.int16: 6, torch.uint8: 7}


def _grid_ok(grid):
    if not (grid.is_cuda and grid.is_contiguous() and grid.dtype in GRID_DTYPES and grid.dim() == 3):
        raise ValueError("grid must be a contiguous CUDA tensor [Rx,Ry,Rz] of a supported dtype")
    return GRID_DTYPES[grid.dtype]


_sizes = {}   # descriptor bytes -> (workspace bytes, vertex capacity hint): both depend on the shape fields only


def _desc_sizes(desc):
    key = bytes(desc)
    hit = _sizes.get(key)
    if hit is None:
        n = lib().p3d_mc_workspace_bytes(ctypes.byref(desc))
        if n == 0:
            raise P3DError(P3D_ERR_INVALID, "invalid descriptor")
        if len(_sizes) > 256:
            _sizes.clear()
        hit = _sizes[key] = (n, lib().p3d_mc_vertex_capacity_hint(ctypes.byref(desc)))
    return hit


def mc_workspace_bytes(desc):
    return _desc_sizes(desc)[0]


class _on_device:
    """torch.cuda.device(...) only when the tensor's device is not already current (the context manager costs
    several microseconds per extraction)."""

    def __init__(self, device):
        self.ctx = None if device.index == torch.cuda.current_device() else torch.cuda.device(device)

    def __enter__(self):
        if self.ctx is not None:
            self.ctx.__enter__()

    def __exit__(self, *exc):
        if self.ctx is not None:
            self.ctx.__exit__(*exc)


def mc_count(desc, grid, workspace=None, vertex_capacity=None):
    """One pass over the grid -> (V, F, workspace, vbuf).

    vbuf is a float32 [vertex_capacity, 3] tensor holding every vertex with id < vertex_capacity
    (all of them when V <= vertex_capacity).  vertex_capacity=None asks the library for its hint,
    0 writes no vertices.  Synchronises the current stream."""
    dtype = _grid_ok(grid)
    nbytes = mc_workspace_bytes(desc)
    if workspace is None:
        workspace = torch.empty(nbytes, dtype=torch.uint8, device=grid.device)
    if vertex_capacity is None:
        vertex_capacity = lib().p3d_mc_vertex_capacity_hint(ctypes.byref(desc))
    vbuf = torch.empty((int(vertex_capacity), 3), dtype=torch.float32, device=grid.device)
    counts = (ctypes.c_int64 * 2)()
    with torch.cuda.device(grid.device):
        check(lib().p3d_mc_count_typed(ctypes.byref(desc), grid.data_ptr(), dtype, workspace.data_ptr(), workspace.numel(),
                                       vbuf.data_ptr() if vertex_capacity else None, int(vertex_capacity), counts,
                                       _stream()))
    return counts[0], counts[1], workspace, vbuf


def mc_vertices(desc, grid, workspace, V, vbuf=None):
    """-> vertices float32 [V,3]: the speculative buffer of mc_count when it was large enough, otherwise an
    exact-size buffer filled by the vertices-only second pass (asynchronous)."""
    if vbuf is not None and vbuf.shape[0] >= V:
        return vbuf[:V]
    verts = torch.empty((V, 3), dtype=torch.float32, device=grid.device)
    with torch.cuda.device(grid.device):
        check(lib().p3d_mc_vertices_typed(ctypes.byref(desc), grid.data_ptr(), _grid_ok(grid), workspace.data_ptr(),
                                          verts.data_ptr(), int(V), _stream()))
    return verts


def mc_faces(desc, workspace, F, vertex_id_base=0):
    """-> faces int32 [F,3] on the workspace's device (asynchronous)."""
    faces = torch.empty((F, 3), dtype=torch.int32, device=workspace.device)
    with torch.cuda.device(workspace.device):
        check(lib().p3d_mc_faces(ctypes.byref(desc), workspace.data_ptr(), faces.data_ptr(), int(vertex_id_base), _stream()))
    return faces


def mc_extract(desc, grid, vertex_capacity=None, face_capacity=None):
    """Whole single-GPU extraction with one host synchronisation (p3d_mc_extract): both passes are queued into
    buffers of speculative capacity (None = the library's vertex hint, twice that many faces); whatever did not
    fit is redone into an exact buffer.  -> (vertices f32 [V,3], faces i32 [F,3], V, F)."""
    dtype = _grid_ok(grid)
    ws_bytes, hint = _desc_sizes(desc)
    ws = torch.empty(ws_bytes, dtype=torch.uint8, device=grid.device)
    if vertex_capacity is None:
        vertex_capacity = hint
    if face_capacity is None:
        face_capacity = 2 * int(vertex_capacity)
    vbuf = torch.empty((int(vertex_capacity), 3), dtype=torch.float32, device=grid.device)
    fbuf = torch.empty((int(face_capacity), 3), dtype=torch.int32, device=grid.device)
    counts = (ctypes.c_int64 * 2)()
    with _on_device(grid.device):
        check(lib().p3d_mc_extract(ctypes.byref(desc), grid.data_ptr(), dtype, ws.data_ptr(), ws.numel(),
                                   vbuf.data_ptr() if vertex_capacity else None, int(vertex_capacity),
                                   fbuf.data_ptr() if face_capacity else None, int(face_capacity), counts, _stream()))
    V, F = counts[0], counts[1]
    if lib().p3d_mc_single_launch(ctypes.byref(desc), dtype) and (V > vertex_capacity or F > face_capacity):
        # small grid (single-launch path): what did not fit is redone by the same call with an exact buffer
        rv, rf = V > vertex_capacity, F > face_capacity
        verts = torch.empty((V, 3), dtype=torch.float32, device=grid.device) if rv else vbuf[:V]
        faces = torch.empty((F, 3), dtype=torch.int32, device=grid.device) if rf else fbuf[:F]
        with _on_device(grid.device):
            check(lib().p3d_mc_extract(ctypes.byref(desc), grid.data_ptr(), dtype, ws.data_ptr(), ws.numel(),
                                       verts.data_ptr() if rv else None, V if rv else 0,
                                       faces.data_ptr() if rf else None, F if rf else 0, counts, _stream()))
        return verts, faces, V, F
    verts = mc_vertices(desc, grid, ws, V, vbuf)
    faces = fbuf[:F] if F <= face_capacity else mc_faces(desc, ws, F)
    return verts, faces, V, F


def marching_cubes_batch(grids, thresh, lower=None, upper=None):
    """Many (small) grids, one host synchronisation (p3d_mc_extract_batch): the extractions are queued back to
    back and share one workspace.  grids: sequence of contiguous CUDA tensors [Rx,Ry,Rz] of ONE supported dtype on
    one device (shapes may differ); lower / upper: None (each grid's own [0, shape] box) or one box for all.
    -> list of (vertices float32 [V,3], faces int32 [F,3]).  A grid whose speculative buffers were too small is
    redone on its own."""
    grids = list(grids)
    if not grids:
        return []
    dtype = _grid_ok(grids[0])
    dev = grids[0].device
    if any(_grid_ok(g) != dtype or g.device != dev for g in grids):
        raise ValueError("a batch must have one dtype and one device")
    n = len(grids)
    # one descriptor and one size query per distinct shape (the host side of a batch of small grids costs as much as
    # its one kernel launch: everything per grid is kept to a few Python operations)
    by_shape = {}
    for g in grids:
        if g.shape not in by_shape:
            d = McDesc.make(g.shape, thresh, lower, upper)
            by_shape[g.shape] = (d, _desc_sizes(d))
    descs = (McDesc * n)(*[by_shape[g.shape][0] for g in grids])
    sizes = [by_shape[g.shape][1] for g in grids]
    ws = torch.empty(max(max(s[0] for s in sizes), lib().p3d_mc_batch_workspace_bytes(n, descs)), dtype=torch.uint8, device=dev)
    vcaps, fcaps = [s[1] for s in sizes], [2 * s[1] for s in sizes]
    # one allocation per output kind, carved into per-grid buffers (256-byte aligned starts)
    pad = lambda rows: (rows * 12 + 255) // 256 * 256
    voff, foff = [0], [0]
    for v, f in zip(vcaps, fcaps):
        voff.append(voff[-1] + pad(v))
        foff.append(foff[-1] + pad(f))
    vall = torch.empty(voff[-1], dtype=torch.uint8, device=dev)
    fall = torch.empty(foff[-1], dtype=torch.uint8, device=dev)
    ptrs = lambda vals: (ctypes.c_void_p * n)(*vals)
    i64s = lambda vals: (ctypes.c_int64 * n)(*vals)
    counts = (ctypes.c_int64 * (2 * n))()
    with _on_device(dev):
        check(lib().p3d_mc_extract_batch(n, descs, ptrs([g.data_ptr() for g in grids]), dtype, ws.data_ptr(), ws.numel(),
                                         ptrs([vall.data_ptr() + o for o in voff[:-1]]), i64s(vcaps),
                                         ptrs([fall.data_ptr() + o for o in foff[:-1]]), i64s(fcaps), counts, _stream()))
    out = []
    vall32, fall32 = vall.view(torch.float32), fall.view(torch.int32)
    counts = list(counts)
    for i, g in enumerate(grids):
        V, F = counts[2 * i], counts[2 * i + 1]
        if V <= vcaps[i] and F <= fcaps[i]:
            v = torch.as_strided(vall32, (V, 3), (3, 1), voff[i] // 4)
            f = torch.as_strided(fall32, (F, 3), (3, 1), foff[i] // 4)
        else:
            v, f, _, _ = mc_extract(descs[i], g, V, F)
        out.append((v, f))
    return out


def marching_cubes_host(grid, thresh, lower=None, upper=None, slab_planes=0, vertices_out=None, faces_out=None,
                        device=None):
    """Marching cubes of a grid in HOST memory, slab-pipelined on one GPU (p3d_mc_extract_host): upload, extraction
    and download overlap, and the grid may be larger than device memory.

    grid: contiguous CPU tensor [Rx,Ry,Rz] of a supported dtype (pin it for full PCIe speed).  vertices_out /
    faces_out: optional CPU tensors (float32 [cap,3] / int32 [cap,3], ideally pinned) to write into; otherwise
    pinned buffers are allocated from the library's hint.  Returns (vertices float32 [V,3], faces int32 [F,3]) as
    CPU tensors (views of the buffers); a too small buffer is replaced by one of the exact size and the call
    repeated.  Vertex numbering is slab by slab, faces are voxel-major with global ids."""
    if grid.is_cuda or not grid.is_contiguous() or grid.dtype not in GRID_DTYPES or grid.dim() != 3:
        raise ValueError("grid must be a contiguous CPU tensor [Rx,Ry,Rz] of a supported dtype")
    if not torch.cuda.is_available():
        raise RuntimeError("marching_cubes_host needs a CUDA device (there is no CPU fallback)")
    desc = McDesc.make(grid.shape, thresh, lower, upper)
    hint = _desc_sizes(desc)[1]
    pin = lambda n, dt: torch.empty((int(n), 3), dtype=dt).pin_memory()
    if vertices_out is None:
        vertices_out = pin(hint, torch.float32)
    if faces_out is None:
        faces_out = pin(2 * hint, torch.int32)
    for _ in range(2):
        for t, dt in ((vertices_out, torch.float32), (faces_out, torch.int32)):
            if t.is_cuda or not t.is_contiguous() or t.dtype != dt or t.dim() != 2 or t.shape[1] != 3:
                raise ValueError("output buffers must be contiguous CPU tensors [cap,3] (float32 vertices, int32 faces)")
        counts = (ctypes.c_int64 * 2)()
        with torch.cuda.device(torch.cuda.current_device() if device is None else device):
            # device scratch from torch's caching allocator: cudaMalloc / cudaFree would cost more than a slab
            nbytes = lib().p3d_mc_extract_host_arena_bytes(ctypes.byref(desc), GRID_DTYPES[grid.dtype], int(slab_planes))
            arena = torch.empty(nbytes, dtype=torch.uint8, device="cuda")
            # the call runs on its own streams: whatever the current stream still has queued (a copy into `grid`,
            # the previous owner of the arena's block) must be over first (include/prim3d_b200.h)
            torch.cuda.current_stream().synchronize()
            check(lib().p3d_mc_extract_host(ctypes.byref(desc), grid.data_ptr(), GRID_DTYPES[grid.dtype], int(slab_planes),
                                            vertices_out.data_ptr(), vertices_out.shape[0], faces_out.data_ptr(),
                                            faces_out.shape[0], counts, arena.data_ptr(), arena.numel()))
            del arena   # the call is synchronous: nothing is in flight
        V, F = counts[0], counts[1]
        if V <= vertices_out.shape[0] and F <= faces_out.shape[0]:
            return vertices_out[:V], faces_out[:F]
        if V > vertices_out.shape[0]:
            vertices_out = pin(V, torch.float32)
        if F > faces_out.shape[0]:
            faces_out = pin(F, torch.int32)
    raise P3DError(P3D_ERR_INVALID, "marching_cubes_host: counts changed between two passes over the same grid")


def marching_cubes(grid, thresh, lower=None, upper=None, vertex_capacity=None):
    """Single-GPU extraction through the C ABI (same outputs as prim3d.libPrim3D.marching_cubes)."""
    desc = McDesc.make(grid.shape, thresh, lower, upper)
    V, F, ws, vbuf = mc_count(desc, grid, vertex_capacity=vertex_capacity)
    return mc_vertices(desc, grid, ws, V, vbuf), mc_faces(desc, ws, F)


def active_tiles(grid, thresh):
    """The tile list p3d_mc_extract_sparse needs for the CUDA tensor `grid`, as a sorted int32 tensor: every tile
    (8 x 8 rows x 128 samples, id = (x // 8 * ceil(ry / 8) + y // 8) * ceil(rz / 128) + z // 128) that holds a sample
    whose +x / +y / +z edge is crossed or a cell with mixed corners.  Plain torch ops over the whole grid: a helper for
    tests and for callers without a cheaper source of the list (the point of the sparse form is NOT to look at the
    whole grid; a coarse pass or the previous frame usually says where the surface is)."""
    inside = grid.to(torch.float32) > thresh
    rx, ry, rz = inside.shape
    ex, ey, ez = inside[:-1] ^ inside[1:], inside[:, :-1] ^ inside[:, 1:], inside[:, :, :-1] ^ inside[:, :, 1:]
    need = torch.zeros_like(inside)
    need[:-1] |= ex
    need[:, :-1] |= ey
    need[:, :, :-1] |= ez
    if min(rx, ry, rz) > 1:
        cell = ex[:, :-1, :-1] | ex[:, 1:, :-1] | ex[:, :-1, 1:] | ex[:, 1:, 1:]
        cell |= ey[:-1, :, :-1] | ey[1:, :, :-1] | ey[:-1, :, 1:] | ey[1:, :, 1:]
        cell |= ez[:-1, :-1, :] | ez[1:, :-1, :] | ez[:-1, 1:, :] | ez[1:, 1:, :]
        need[:-1, :-1, :-1] |= cell
    nxb, nyb, npz = -(-rx // 8), -(-ry // 8), -(-rz // 128)
    padded = torch.zeros((nxb * 8, nyb * 8, npz * 128), dtype=torch.bool, device=grid.device)
    padded[:rx, :ry, :rz] = need
    tiles = padded.view(nxb, 8, nyb, 8, npz, 128).any(dim=5).any(dim=3).any(dim=1)
    return tiles.reshape(-1).nonzero().reshape(-1).to(torch.int32)


def marching_cubes_sparse(grid, thresh, tiles, lower=None, upper=None, vertex_capacity=None, face_capacity=None):
    """p3d_mc_extract_sparse: marching cubes of the listed tiles of a dense CUDA grid (see include/prim3d_b200.h);
    tiles = int32 CUDA tensor of distinct tile ids, e.g. active_tiles(grid, thresh) -> (vertices, faces)."""
    dtype = _grid_ok(grid)
    if grid.dtype != torch.float32:
        raise ValueError("the block-sparse form takes float32 grids")
    if not (tiles.is_cuda and tiles.dtype == torch.int32 and tiles.is_contiguous() and tiles.dim() == 1):
        raise ValueError("tiles must be a contiguous int32 CUDA vector")
    desc = McDesc.make(grid.shape, thresh, lower, upper)
    ws_bytes, hint = _desc_sizes(desc)
    n = int(tiles.numel())
    vcap = min(hint, n * 24576 + 1) if vertex_capacity is None else int(vertex_capacity)   # a tile has at most 24576 vertices
    fcap = 2 * vcap if face_capacity is None else int(face_capacity)
    with _on_device(grid.device):
        ws = torch.empty(ws_bytes, dtype=torch.uint8, device=grid.device)
        for _ in range(2):
            verts = torch.empty((vcap, 3), dtype=torch.float32, device=grid.device)
            faces = torch.empty((fcap, 3), dtype=torch.int32, device=grid.device)
            counts = (ctypes.c_int64 * 2)()
            check(lib().p3d_mc_extract_sparse(ctypes.byref(desc), grid.data_ptr(), dtype, tiles.data_ptr(), n, ws.data_ptr(), ws.numel(),
                                              verts.data_ptr(), vcap, faces.data_ptr(), fcap, counts, _stream()))
            V, F = counts[0], counts[1]
            if V <= vcap and F <= fcap:
                return verts[:V], faces[:F]
            vcap, fcap = max(V, 1), max(F, 1)
    raise P3DError(P3D_ERR_INVALID, "marching_cubes_sparse: counts changed between two passes over the same grid")


def marching_tetrahedra(points, tets, sdf, staged=False, capacities=None):
    """Marching tetrahedra through the C ABI: points f32 [P,3], tets i64 [T,4] (mutated in place),
    sdf f32 [P], all contiguous CUDA tensors -> (verts f32 [V,3], faces i64 [F,3], tet_idx i64 [F],
    edges i64 [V,2]).

    One call of p3d_mt_extract (a second one with exact capacities if the guess `capacities` = (slot, key, vertex,
    face) was too small); the staged calls p3d_mt_classify / p3d_mt_index / p3d_mt_emit when `staged` or when the
    library reports that the input does not suit the one-call layout."""
    if not (points.is_cuda and tets.is_cuda and sdf.is_cuda and points.is_contiguous() and tets.is_contiguous()
            and sdf.is_contiguous() and points.dtype == torch.float32 and tets.dtype == torch.int64
            and sdf.dtype == torch.float32):
        raise ValueError("points/sdf must be contiguous float32 and tets contiguous int64 CUDA tensors")
    L, dev = lib(), points.device
    P, T = points.shape[0], tets.shape[0]
    oriented = 0
    if not staged:
        slot, key, vcap, fcap = capacities if capacities is not None else (min(T, T // 16 + 4096), T // 4 + 4096, T // 4 + 4096,
                                                                           3 * min(T, T // 16 + 4096))
        with torch.cuda.device(dev):
            for _ in range(2):
                ws = torch.empty(L.p3d_mt_extract_workspace_bytes(T, P, slot, key), dtype=torch.uint8, device=dev)
                verts = torch.empty((vcap, 3), dtype=torch.float32, device=dev)
                edges = torch.empty((vcap, 2), dtype=torch.int64, device=dev)
                faces = torch.empty((fcap, 3), dtype=torch.int64, device=dev)
                tet_idx = torch.empty((fcap,), dtype=torch.int64, device=dev)
                c = (ctypes.c_int64 * 5)()
                check(L.p3d_mt_extract(points.data_ptr(), P, tets.data_ptr(), T, sdf.data_ptr(), oriented, ws.data_ptr(), ws.numel(),
                                       slot, key, verts.data_ptr(), edges.data_ptr(), vcap, faces.data_ptr(), tet_idx.data_ptr(), fcap,
                                       c, _stream()))
                marching_tetrahedra.last_state = c[4]
                if c[4] == 3:
                    break        # nothing ran
                oriented = 1     # the tets are fixed now: a later pass must not fix them again
                # too small a guess (1), or buckets sized for fewer entries than there are (2 with counts above the
                # guesses; V is unknown then, ne bounds it): one more run with capacities from the counts
                if c[4] == 0 or (c[4] == 2 and c[0] + c[1] <= slot and c[2] <= key):
                    break
                slot, key, vcap, fcap = c[0] + c[1], c[2], (c[3] if c[4] == 1 else c[2]), c[0] + 2 * c[1]
        if c[4] == 0:
            V, F = c[3], c[0] + 2 * c[1]
            return verts[:V], faces[:F], tet_idx[:F], edges[:V]
    with torch.cuda.device(dev):
        codes = torch.empty(L.p3d_mt_codes_bytes(T), dtype=torch.uint8, device=dev)
        c = (ctypes.c_int64 * 3)()
        check(L.p3d_mt_classify(points.data_ptr(), P, tets.data_ptr(), T, sdf.data_ptr(), oriented, codes.data_ptr(), c, _stream()))
        n1, n2, ne = c[0], c[1], c[2]
        ws = torch.empty(L.p3d_mt_workspace_bytes(T, n1, n2, ne), dtype=torch.uint8, device=dev)
        v = (ctypes.c_int64 * 1)()
        check(L.p3d_mt_index(tets.data_ptr(), T, P, sdf.data_ptr(), codes.data_ptr(), n1, n2, ne, ws.data_ptr(),
                             ws.numel(), v, _stream()))
        V, F = v[0], n1 + 2 * n2
        verts = torch.empty((V, 3), dtype=torch.float32, device=dev)
        edges = torch.empty((V, 2), dtype=torch.int64, device=dev)
        faces = torch.empty((F, 3), dtype=torch.int64, device=dev)
        tet_idx = torch.empty((F,), dtype=torch.int64, device=dev)
        check(L.p3d_mt_emit(points.data_ptr(), tets.data_ptr(), T, sdf.data_ptr(), codes.data_ptr(), n1, n2, ne, V,
                            ws.data_ptr(), verts.data_ptr(), edges.data_ptr(), faces.data_ptr(), tet_idx.data_ptr(),
                            _stream()))
    return verts, faces, tet_idx, edges
