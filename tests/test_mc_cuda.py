"""GPU parity tests of the marching-cubes path: CUDA kernels (through the C ABI and through
prim3d.libPrim3D) against the CPU oracle on identical inputs.

Bar: V and F equal; vertex positions bit-identical as a multiset; triangles bit-identical
(9 corner coordinates each) and, because both sides emit faces in voxel-major order, identical
triangle by triangle."""
import os

import numpy as np
import pytest
import torch

from canonical import assert_same_mesh, triangle_soup
from oracle import inputs, mc

pytestmark = pytest.mark.gpu
HERE = os.path.dirname(os.path.abspath(__file__))


def bunny():
    return np.load(os.path.join(HERE, "golden", "mc_bunny66.npz"))["grid"]


def nan_inf_grid():
    g = inputs.noise((12, 13, 40), seed=11)
    rng = np.random.default_rng(12)
    flat = g.reshape(-1)
    flat[rng.integers(0, flat.size, 200)] = np.nan
    flat[rng.integers(0, flat.size, 200)] = np.inf
    flat[rng.integers(0, flat.size, 200)] = -np.inf
    return g


CASES = {
    # name: (grid factory, thresh, lower, upper)
    "sphere128": (lambda: inputs.sphere_int64(128).astype(np.float32), 0.0, None, None),
    "sphere200": (lambda: inputs.sphere_int64(200).astype(np.float32), 0.0, None, None),
    "bunny66": (bunny, 0.0, None, None),
    "bunny256": (lambda: inputs.upsample_trilinear(bunny(), 256), 0.0, None, None),  # BASELINE configs[1]
    "gyroid128": (lambda: inputs.gyroid(128), 0.0, None, None),
    "gyroid256": (lambda: inputs.gyroid(256), 0.0, None, None),
    "noise33_s0": (lambda: inputs.noise((33, 33, 33), 0), 0.0, None, None),
    "noise33_s1": (lambda: inputs.noise((33, 33, 33), 1), 0.0, None, None),
    "noise65_s2": (lambda: inputs.noise((65, 65, 65), 2), 0.0, None, None),
    "noise64_flat": (lambda: inputs.noise((64, 64, 64), 3), 0.0, None, None),
    "noise_flat_tail": (lambda: inputs.noise((3, 5, 96), 4), 0.0, None, None),      # one partial piece, TMA path
    "noise_long_rows": (lambda: inputs.noise((3, 4, 2100), 5), 0.0, None, None),    # rows span 17 pieces, TMA path
    "noise_long_rows_odd": (lambda: inputs.noise((3, 4, 2101), 5), 0.0, None, None),  # same, generic loader
    "noise_long_rows_flat": (lambda: inputs.noise((2, 3, 2048), 6), 0.0, None, None),
    # rows of 17..128 bit words: the row-streaming face pass (mc_faces_rows.cu), 1 / 2 / 4 words per lane
    "noise_rows_np8": (lambda: inputs.noise((6, 7, 1024), 20), 0.0, None, None),
    "noise_rows_np6": (lambda: inputs.noise((5, 9, 700), 21), 0.0, None, None),
    "noise_rows_np5_odd": (lambda: inputs.noise((4, 5, 611), 22), 0.0, None, None),
    "noise_rows_np12": (lambda: inputs.noise((3, 5, 1500), 23), 0.0, None, None),
    "noise_rows_np32": (lambda: inputs.noise((3, 3, 4096), 24), 0.0, None, None),
    "waves_rows_np8": (lambda: inputs.waves((40, 150, 1000)), 0.0, None, None),
    "waves_rows_np16": (lambda: inputs.waves((10, 70, 2048)), 0.1, None, None),
    "noise_piece_edges": (lambda: inputs.noise((10, 11, 257), 15), 0.0, None, None),  # piece boundary at 128, 256
    "noise_blocks": (lambda: inputs.noise((17, 25, 132), 16), 0.0, None, None),      # partial x/y blocks
    "noise_dense_tile": (lambda: inputs.noise((9, 9, 384), 17), 0.0, None, None),    # > 2048 vertices per tile
    "noncubic": (lambda: inputs.noise((17, 33, 65), 7), 0.0, None, None),
    "ties": (lambda: inputs.ties((24, 24, 24), 8), 0.0, None, None),
    "nan_inf": (nan_inf_grid, 0.0, None, None),
    "min222": (lambda: inputs.noise((2, 2, 2), 9), 0.0, None, None),
    "thin_x": (lambda: inputs.noise((2, 40, 40), 10), 0.0, None, None),
    "thin_y": (lambda: inputs.noise((40, 2, 40), 11), 0.0, None, None),
    "thin_z": (lambda: inputs.noise((40, 40, 2), 12), 0.0, None, None),
    "thresh_nonzero": (lambda: inputs.noise((31, 32, 33), 13), 0.37, None, None),
    "bounds_asym": (lambda: inputs.noise((20, 24, 28), 14), -0.1, [-1.0, -2.0, -3.0], [1.0, 5.0, 3.5]),
    "empty": (lambda: np.full((16, 16, 16), -1.0, np.float32), 0.0, None, None),
    "full": (lambda: np.full((16, 16, 16), 1.0, np.float32), 0.0, None, None),
}


def run_capi(grid_np, thresh, lower, upper):
    from primitive3d_b200 import capi
    g = torch.from_numpy(np.ascontiguousarray(grid_np)).cuda()
    v, f = capi.marching_cubes(g, thresh, lower, upper)
    torch.cuda.synchronize()
    return v.cpu().numpy(), f.cpu().numpy()


@pytest.mark.parametrize("name", list(CASES))
def test_capi_matches_oracle(name):
    make, thresh, lower, upper = CASES[name]
    grid = make()
    v, f = run_capi(grid, thresh, lower, upper)
    ov, of = mc.marching_cubes(grid, thresh, lower, upper)
    assert v.dtype == np.float32 and f.dtype == np.int32
    assert v.shape == ov.shape and f.shape == of.shape, (v.shape, ov.shape, f.shape, of.shape)
    assert_same_mesh(v, f, ov, of, ordered_faces=True)


@pytest.mark.parametrize("name", ["bunny66", "sphere128", "noise33_s0", "bounds_asym", "empty"])
def test_pybind_module_matches_capi(name):
    import prim3d
    make, thresh, lower, upper = CASES[name]
    grid = make()
    lo = [0.0, 0.0, 0.0] if lower is None else lower
    up = [float(s) for s in grid.shape] if upper is None else upper
    out = prim3d._C.marching_cubes(torch.from_numpy(grid).cuda(), thresh, lo, up)
    assert isinstance(out, list) and len(out) == 2
    v, f = out
    assert v.is_cuda and f.is_cuda and v.dtype == torch.float32 and f.dtype == torch.int32
    assert v.shape[1:] == (3,) and f.shape[1:] == (3,)
    # the module goes through p3d_mc_extract, which takes the single-launch path for small grids (its own vertex
    # numbering): same mesh as the staged calls, triangle by triangle
    cv, cf = run_capi(grid, thresh, lower, upper)
    assert_same_mesh(v.cpu().numpy(), f.cpu().numpy(), cv, cf, ordered_faces=True)


def test_python_wrapper_contract():
    """prim3d.marching_cubes: numpy / int64 input, scale forms, ValueError on thin grids
    (reference prim3d/utility/marching_cubes.py:34-98)."""
    import prim3d
    grid64 = inputs.sphere_int64(64)
    v, f = prim3d.marching_cubes(grid64, 0)                      # numpy int64 in
    ov, of = mc.marching_cubes(grid64.astype(np.float32), 0.0)
    assert_same_mesh(v.cpu().numpy(), f.cpu().numpy(), ov, of, ordered_faces=True)
    v2, f2 = prim3d.marching_cubes(torch.tensor(grid64).cuda(), 0, scale=2.0)
    ov2, _ = mc.marching_cubes(grid64.astype(np.float32), 0.0, [0, 0, 0], [2.0, 2.0, 2.0])
    assert_same_mesh(v2.cpu().numpy(), f2.cpu().numpy(), ov2, of, ordered_faces=True)
    v3, _ = prim3d.marching_cubes(torch.tensor(grid64).cuda(), 0, scale=[[-1.0, -1.0, -1.0], [1.0, 1.0, 1.0]])
    ov3, _ = mc.marching_cubes(grid64.astype(np.float32), 0.0, [-1, -1, -1], [1, 1, 1])
    assert np.array_equal(np.sort(v3.cpu().numpy().view(np.uint32), 0), np.sort(ov3.view(np.uint32), 0))
    with pytest.raises(ValueError):
        prim3d.marching_cubes(torch.zeros(1, 8, 8), 0)
    with pytest.raises(TypeError):
        prim3d.marching_cubes(torch.zeros(8, 8, 8), 0, scale="x")
    with pytest.raises(RuntimeError, match="must be a CUDA tensor"):
        prim3d._C.marching_cubes(torch.zeros(8, 8, 8), 0.0, [0, 0, 0], [8, 8, 8])
    with pytest.raises(RuntimeError, match="must be contiguous"):
        prim3d._C.marching_cubes(torch.zeros(8, 8, 16).cuda()[:, :, ::2], 0.0, [0, 0, 0], [8, 8, 8])


DTYPE_GRIDS = {
    # dtype: integer-valued or narrow-range grids whose float32 cast is not the identity where that is possible
    "float16": lambda: (inputs.noise((19, 21, 140), 21) * 4).astype(np.float16),
    "bfloat16": lambda: inputs.noise((19, 21, 140), 22),          # rounded to bf16 on the device below
    "float64": lambda: inputs.noise((19, 21, 140), 23).astype(np.float64) * (1 + 2.0 ** -30) + 1e-9,
    "int64": lambda: inputs.sphere_int64(72) * (2 ** 33 + 12345) + 1,  # |values| > 2^24: the cast rounds
    "int32": lambda: (inputs.sphere_int64(40) * 1000003 + 7).astype(np.int32),
    "int16": lambda: (inputs.noise((9, 40, 200), 24) * 3000).astype(np.int16),
    "uint8": lambda: ((inputs.noise((33, 9, 130), 25) + 1) * 127).astype(np.uint8),
}


@pytest.mark.parametrize("name", list(DTYPE_GRIDS))
def test_fused_dtype_ingest_equals_cast_then_run(name):
    """Grids of other element types are converted to float32 on chip (p3d_mc_count_typed): every output must be
    bit-identical to casting first (the reference wrapper's `.to(torch.float32)`, marching_cubes.py:86-87) --
    through the C ABI, through prim3d.libPrim3D and through prim3d.marching_cubes, and equal to the oracle on
    the cast grid."""
    import prim3d
    from primitive3d_b200 import capi
    g = torch.from_numpy(np.ascontiguousarray(DTYPE_GRIDS[name]())).cuda()
    if name == "bfloat16":
        g = g.to(torch.bfloat16)
    thresh = 100.0 if name == "uint8" else 0.0
    cast = g.to(torch.float32)
    v0, f0 = capi.marching_cubes(cast, thresh)
    v1, f1 = capi.marching_cubes(g, thresh)
    assert v1.dtype == torch.float32 and f1.dtype == torch.int32
    assert torch.equal(v0.view(torch.int32), v1.view(torch.int32)) and torch.equal(f0, f1)
    up = [float(s) for s in g.shape]
    v2, f2 = prim3d._C.marching_cubes(g, thresh, [0.0, 0.0, 0.0], up)
    assert torch.equal(v0.view(torch.int32), v2.view(torch.int32)) and torch.equal(f0, f2)
    v3, f3 = prim3d.marching_cubes(g.cpu(), thresh)   # host tensor of the native dtype: transferred as it is
    assert torch.equal(v0.view(torch.int32), v3.view(torch.int32)) and torch.equal(f0, f3)
    # vertices-only second pass (capacity too small) reads the typed grid too
    desc = capi.McDesc.make(g.shape, thresh)
    V, F, ws, vbuf = capi.mc_count(desc, g, vertex_capacity=7)
    v4 = capi.mc_vertices(desc, g, ws, V, vbuf)
    assert torch.equal(v0.view(torch.int32), v4.view(torch.int32))
    ov, of = mc.marching_cubes(cast.cpu().numpy(), thresh)
    assert_same_mesh(v1.cpu().numpy(), f1.cpu().numpy(), ov, of, ordered_faces=True)
    assert v1.shape[0] > 0


@pytest.mark.parametrize("name", ["gyroid128", "noise65_s2", "bounds_asym", "empty", "min222"])
def test_single_sync_extraction_equals_staged_calls(name):
    """p3d_mc_extract queues both passes into speculative buffers and waits once; whatever the capacities, the
    outputs equal those of p3d_mc_count + p3d_mc_vertices + p3d_mc_faces."""
    from primitive3d_b200 import capi
    factory, thresh, lower, upper = CASES[name]
    g = torch.from_numpy(np.ascontiguousarray(factory())).cuda()
    v0, f0 = capi.marching_cubes(g, thresh, lower, upper)
    desc = capi.McDesc.make(g.shape, thresh, lower, upper)
    V0, F0 = v0.shape[0], f0.shape[0]
    for vcap, fcap in [(None, None), (V0, F0), (V0 + 9, F0 + 9), (max(V0 - 1, 0), F0), (V0, max(F0 - 1, 0)), (3, 5), (0, 0)]:
        v, f, V, F = capi.mc_extract(desc, g, vcap, fcap)
        assert (V, F) == (V0, F0)
        # small grids take the single-launch path in p3d_mc_extract (its own vertex numbering): same mesh, triangle by
        # triangle; whatever did not fit the speculative buffers was redone by the staged calls
        assert_same_mesh(v.cpu().numpy(), f.cpu().numpy(), v0.cpu().numpy(), f0.cpu().numpy(), ordered_faces=True)


@pytest.mark.parametrize("nv,nf", [(0, 0), (1, 0), (3, 1), (1001, 333), (4096, 8190)])
def test_ply_assembled_on_the_device_has_the_same_bytes(tmp_path, nv, nf):
    """save_mesh on CUDA tensors packs the 15-byte vertex and 16-byte face records on the device (p3d_ply_pack);
    the file must equal the one written from host tensors, i.e. the reference's layout (marching_cubes.cu:307-352)."""
    import prim3d
    rng = np.random.default_rng(nv * 7 + nf)
    v = torch.from_numpy(rng.standard_normal((nv, 3)).astype(np.float32))
    f = torch.from_numpy(rng.integers(0, max(nv, 1), (nf, 3)).astype(np.int32))
    c = torch.from_numpy(rng.integers(0, 256, (nv, 3)).astype(np.uint8))
    a, b = str(tmp_path / "host.ply"), str(tmp_path / "device.ply")
    prim3d.save_mesh(v, f, c, filename=a)
    prim3d.save_mesh(v.cuda(), f.cuda(), c.cuda(), filename=b)
    raw = open(a, "rb").read()
    assert raw == open(b, "rb").read()
    body = raw[raw.index(b"end_header\n") + 11:]
    assert len(body) == 15 * nv + 16 * nf
    rec = np.frombuffer(body[:15 * nv], np.uint8).reshape(nv, 15)
    assert np.array_equal(rec[:, :12].copy().view(np.float32).reshape(nv, 3), v.numpy()) and np.array_equal(rec[:, 12:], c.numpy())
    fr = np.frombuffer(body[15 * nv:], np.int32).reshape(nf, 4)
    assert (fr[:, 0] == 3).all() and np.array_equal(fr[:, 1:], f.numpy())
    # default colours (127) with a mesh that is still on the device
    prim3d.save_mesh(v.cuda(), f.cuda().long(), filename=b)
    prim3d.save_mesh(v, f.long(), filename=a)
    assert open(a, "rb").read() == open(b, "rb").read()


@pytest.mark.parametrize("shape,planes,dtype", [((40, 24, 140), 8, np.float32), ((37, 20, 64), 16, np.float32),
                                                 ((16, 16, 16), 0, np.float32), ((50, 12, 33), 8, np.float64),
                                                 ((9, 8, 8), 8, np.int64)])
def test_host_streamed_extraction_equals_single_shot(shape, planes, dtype):
    """p3d_mc_extract_host (grid and mesh in host memory, slabs pipelined through the device) gives the mesh of the
    single-shot call: same faces in the same voxel-major order over a vertex array numbered slab by slab."""
    from primitive3d_b200 import capi
    g = inputs.noise(shape, sum(shape))
    g = (g * 1000).astype(dtype) if dtype == np.int64 else g.astype(dtype)
    host = torch.from_numpy(np.ascontiguousarray(g)).pin_memory()
    lower, upper = [-1.0, 0.5, 2.0], [3.0, 4.5, 2.5]
    v0, f0 = capi.marching_cubes(host.cuda(), 0.05, lower, upper)
    v, f = capi.marching_cubes_host(host, 0.05, lower, upper, slab_planes=planes)
    assert not v.is_cuda and v.dtype == torch.float32 and f.dtype == torch.int32
    assert_same_mesh(v.numpy(), f.numpy(), v0.cpu().numpy(), f0.cpu().numpy(), ordered_faces=True)
    # too small output buffers: the counts come back and the call is repeated with exact ones
    v2, f2 = capi.marching_cubes_host(host, 0.05, lower, upper, slab_planes=planes,
                                      vertices_out=torch.empty((5, 3)), faces_out=torch.empty((7, 3), dtype=torch.int32))
    assert torch.equal(v2, v) and torch.equal(f2, f)


@pytest.mark.parametrize("shape,parts,dtype", [((72, 24, 140), 4, np.float32), ((37, 20, 64), 3, np.float32),
                                                ((16, 16, 16), 1, np.float32), ((50, 12, 33), 6, np.float64),
                                                ((130, 9, 40), 16, np.float16)])
def test_host_shard_uploaded_in_overlapped_parts(shape, parts, dtype):
    """sharded.marching_cubes_slab_host on one GPU (the shard is the whole grid): the parts behave as consecutive
    shards, faces in voxel-major order over a vertex array numbered part by part; device or pinned-host outputs; a
    noise field overflows the first call's speculative vertex segments and is redone with the counts known."""
    from primitive3d_b200 import capi, sharded
    g = inputs.noise(shape, sum(shape)).astype(dtype)
    host = torch.from_numpy(np.ascontiguousarray(g)).pin_memory()
    lower, upper = [-1.0, 0.5, 2.0], [3.0, 4.5, 2.5]
    v0, f0 = capi.marching_cubes(host.cuda(), 0.05, lower, upper)
    sharded._part_counts.clear()
    for _ in range(2):      # without, then with remembered counts
        m = sharded.marching_cubes_slab_host(host, 0.05, 0, shape[0], lower, upper, parts=parts, distributed=False)
        torch.cuda.synchronize()
        assert m.vertices.is_cuda and (m.v_offset, m.f_offset) == (0, 0)
        assert (m.num_vertices_total, m.num_faces_total) == (v0.shape[0], f0.shape[0])
        assert_same_mesh(m.vertices.cpu().numpy(), m.faces.cpu().numpy(), v0.cpu().numpy(), f0.cpu().numpy(), ordered_faces=True)
    hv = torch.empty((v0.shape[0] + 3, 3)).pin_memory()
    hf = torch.empty((f0.shape[0], 3), dtype=torch.int32).pin_memory()
    m2 = sharded.marching_cubes_slab_host(host, 0.05, 0, shape[0], lower, upper, parts=parts, out_vertices=hv, out_faces=hf,
                                          distributed=False)
    assert not m2.vertices.is_cuda and torch.equal(m2.vertices, m.vertices.cpu()) and torch.equal(m2.faces, m.faces.cpu())


@pytest.mark.parametrize("name", ["bunny66", "sphere128", "gyroid128", "noise33_s0", "noise_long_rows", "noncubic", "thin_z",
                                  "bounds_asym", "waves_rows_np8", "empty", "full"])
def test_block_sparse_form_equals_the_dense_call(name):
    """p3d_mc_extract_sparse over the tiles that hold the surface gives the dense call's mesh (faces in the same
    voxel-major order; vertices numbered in list order): with the minimal list, with the list reversed, with every
    tile of the grid listed, and with an id past the grid thrown in."""
    from primitive3d_b200 import capi
    make, thresh, lower, upper = CASES[name]
    grid_np = np.ascontiguousarray(make())
    g = torch.from_numpy(grid_np).cuda()
    v0, f0 = capi.marching_cubes(g, thresh, lower, upper)
    v0, f0 = v0.cpu().numpy(), f0.cpu().numpy()
    tiles = capi.active_tiles(g, thresh)
    nxb, nyb, npz = -(-g.shape[0] // 8), -(-g.shape[1] // 8), -(-g.shape[2] // 128)
    every = torch.arange(nxb * nyb * npz, dtype=torch.int32, device="cuda")
    assert tiles.numel() <= every.numel()
    if name in ("empty", "full"):
        assert tiles.numel() == 0
    lists = [tiles, tiles.flip(0).contiguous(), every]
    if tiles.numel() < every.numel():
        lists.append(torch.cat([tiles, torch.tensor([nxb * nyb * npz + 5], dtype=torch.int32, device="cuda")]))
    for tl in lists:
        v, f = capi.marching_cubes_sparse(g, thresh, tl, lower, upper)
        torch.cuda.synchronize()
        assert_same_mesh(v.cpu().numpy(), f.cpu().numpy(), v0, f0, ordered_faces=True)
    # capacities that are too small: the counts come back and the wrapper calls again
    v, f = capi.marching_cubes_sparse(g, thresh, tiles, lower, upper, vertex_capacity=3, face_capacity=2)
    assert_same_mesh(v.cpu().numpy(), f.cpu().numpy(), v0, f0, ordered_faces=True)


def test_block_sparse_form_reads_only_the_listed_tiles():
    """Samples outside the listed tiles (and their one-sample halo) are never looked at: poison them."""
    from primitive3d_b200 import capi
    grid_np = inputs.sphere_int64(200).astype(np.float32)
    g = torch.from_numpy(np.ascontiguousarray(grid_np)).cuda()
    v0, f0 = capi.marching_cubes(g, 0.0)
    tiles = capi.active_tiles(g, 0.0)
    nxb, nyb, npz = -(-200 // 8), -(-200 // 8), -(-200 // 128)
    keep = torch.zeros((nxb, nyb, npz), dtype=torch.bool, device="cuda").view(-1)
    keep[tiles.long()] = True
    keep = keep.view(nxb, 1, nyb, 1, npz, 1).expand(nxb, 8, nyb, 8, npz, 128).reshape(nxb * 8, nyb * 8, npz * 128)
    near = keep.clone()                      # listed tiles plus one sample towards +x, +y, +z
    near[1:] |= keep[:-1]
    near[:, 1:] |= near[:, :-1].clone()
    near[:, :, 1:] |= near[:, :, :-1].clone()
    poisoned = torch.where(near[:200, :200, :200], g, torch.full_like(g, float("nan")))
    assert tiles.numel() < 0.5 * nxb * nyb * npz
    v, f = capi.marching_cubes_sparse(poisoned, 0.0, tiles)
    assert_same_mesh(v.cpu().numpy(), f.cpu().numpy(), v0.cpu().numpy(), f0.cpu().numpy(), ordered_faces=True)


def test_batched_small_grids_equal_one_by_one():
    """p3d_mc_extract_batch: many small grids queued back to back, one host wait; every mesh equals the one the
    single call gives (including a grid dense enough to overflow its speculative buffers, and an empty one)."""
    from primitive3d_b200 import capi
    shapes = [(66, 66, 66), (17, 33, 65), (2, 2, 2), (40, 8, 130), (16, 16, 16), (9, 9, 384)]
    grids = [torch.from_numpy(inputs.noise(s, 40 + i)).cuda() for i, s in enumerate(shapes)]
    grids[4] = torch.full((16, 16, 16), -1.0, device="cuda")                     # empty mesh
    grids[0] = torch.from_numpy(np.ascontiguousarray(bunny())).cuda()             # smooth: fits the speculative buffers
    out = capi.marching_cubes_batch(grids, 0.0)
    assert len(out) == len(grids)
    same = lambda a, b: assert_same_mesh(a[0].cpu().numpy(), a[1].cpu().numpy(), b[0].cpu().numpy(), b[1].cpu().numpy(),
                                         ordered_faces=True)
    for g, (v, f) in zip(grids, out):
        same((v, f), capi.marching_cubes(g, 0.0))
    assert out[4][0].shape == (0, 3) and out[4][1].shape == (0, 3)
    one_box = capi.marching_cubes_batch(grids[:2], 0.1, [-1.0, -1.0, -1.0], [1.0, 2.0, 3.0])
    same(one_box[1], capi.marching_cubes(grids[1], 0.1, [-1.0, -1.0, -1.0], [1.0, 2.0, 3.0]))
    assert capi.marching_cubes_batch([], 0.0) == []


@pytest.mark.parametrize("name", list(CASES))
def test_one_call_extraction_matches_oracle(name):
    """p3d_mc_extract, the entry prim3d.libPrim3D.marching_cubes sits on (the tiled passes, or the single-launch kernel
    of mc_small.cu for grids of up to P3D_MC_SMALL_SINGLE_MAX samples; test_single_launch_kernel_matches_oracle runs every case through that kernel).
    Against the oracle, triangle by triangle."""
    from primitive3d_b200 import capi
    make, thresh, lower, upper = CASES[name]
    grid = make()
    g = torch.from_numpy(np.ascontiguousarray(grid)).cuda()
    v, f, V, F = capi.mc_extract(capi.McDesc.make(g.shape, thresh, lower, upper), g)
    ov, of = mc.marching_cubes(grid, thresh, lower, upper)
    assert (V, F) == (ov.shape[0], of.shape[0])
    assert_same_mesh(v.cpu().numpy(), f.cpu().numpy(), ov, of, ordered_faces=True)


@pytest.mark.parametrize("name", ["sphere128", "bunny66", "gyroid128", "noise33_s0", "noise65_s2", "noise_flat_tail", "noncubic",
                                  "ties", "nan_inf", "min222", "thin_x", "thin_y", "thin_z", "thresh_nonzero", "bounds_asym",
                                  "empty", "full", "noise_piece_edges"])
def test_single_launch_kernel_matches_oracle(name):
    """The single-launch kernel on its own (a batch of one grid goes through it): against the oracle, triangle by
    triangle."""
    from primitive3d_b200 import capi
    make, thresh, lower, upper = CASES[name]
    grid = make()
    (v, f), = capi.marching_cubes_batch([torch.from_numpy(np.ascontiguousarray(grid)).cuda()], thresh, lower, upper)
    ov, of = mc.marching_cubes(grid, thresh, lower, upper)
    assert_same_mesh(v.cpu().numpy(), f.cpu().numpy(), ov, of, ordered_faces=True)


def test_one_launch_for_a_batch_of_small_grids():
    """64 bunny-sized grids in ONE kernel launch (p3d_mc_extract_batch on small float32 grids): every mesh equals the
    oracle's, numbering restarts at every grid."""
    from primitive3d_b200 import capi
    base = bunny()
    grids, want = [], []
    for i in range(64):
        g = np.ascontiguousarray(np.roll(base, i, axis=i % 3) * np.float32(1.0 + 0.01 * i))
        grids.append(torch.from_numpy(g).cuda())
        if i % 9 == 0:
            want.append((i, mc.marching_cubes(g, 0.0)))
    out = capi.marching_cubes_batch(grids, 0.0)
    assert len(out) == 64
    for i, (ov, of) in want:
        assert_same_mesh(out[i][0].cpu().numpy(), out[i][1].cpu().numpy(), ov, of, ordered_faces=True)
    for v, f in out:
        assert f.shape[0] == 0 or (int(f.min()) >= 0 and int(f.max()) < v.shape[0])


def test_small_path_switch(monkeypatch):
    """P3D_MC_SMALL_SINGLE_MAX decides which single grids go through the single-launch kernel (default: up to 2^20
    samples; 0: none): same mesh as the tiled passes give, another vertex numbering."""
    import subprocess
    import sys
    code = ("import numpy as np, torch, sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from oracle import inputs; from primitive3d_b200 import capi;"
            "g = torch.from_numpy(inputs.noise((21, 30, 70), 77)).cuda();"
            "v, f, V, F = capi.mc_extract(capi.McDesc.make(g.shape, 0.0), g); torch.cuda.synchronize();"
            "np.savez(sys.argv[1], v=v.cpu().numpy(), f=f.cpu().numpy())") % (os.path.dirname(HERE), HERE)
    outs = []
    for limit in ["0", "4194304"]:
        path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"p3d_small_{limit}_{os.getpid()}.npz")
        subprocess.check_call([sys.executable, "-c", code, path], env=dict(os.environ, P3D_MC_SMALL_SINGLE_MAX=limit))
        outs.append(np.load(path))
        os.remove(path)
    assert not np.array_equal(outs[0]["f"], outs[1]["f"])       # two numberings ...
    assert_same_mesh(outs[0]["v"], outs[0]["f"], outs[1]["v"], outs[1]["f"], ordered_faces=True)   # ... of one mesh


def test_small_grids_from_two_host_threads_at_once():
    """Two host threads extract small grids on their own streams at the same time: the single-launch kernel spans the
    device with a barrier inside, so its launches are cooperative (two plain launches could each hold half of the SMs
    and wait for the other half) and its barrier words are per host thread.  Run in a child process under a timeout:
    a deadlock must fail this test, not hang the suite."""
    import subprocess
    import sys
    code = """
import sys, threading
import numpy as np, torch
sys.path.insert(0, %r); sys.path.insert(0, %r)
from oracle import inputs
from primitive3d_b200 import capi
grids = [torch.from_numpy(inputs.noise((40, 36, 70), 5 + i)).cuda() for i in range(2)]
want = [capi.mc_extract(capi.McDesc.make(g.shape, 0.0), g) for g in grids]
torch.cuda.synchronize()
bad = []
def work(i):
    with torch.cuda.stream(torch.cuda.Stream()):
        desc = capi.McDesc.make(grids[i].shape, 0.0)
        for _ in range(300):
            v, f, V, F = capi.mc_extract(desc, grids[i])
            if (V, F) != want[i][2:] or not torch.equal(f, want[i][1]) or not torch.equal(v.view(torch.int32), want[i][0].view(torch.int32)):
                bad.append(i)
                return
ts = [threading.Thread(target=work, args=(i,)) for i in range(2)]
[t.start() for t in ts]; [t.join() for t in ts]
torch.cuda.synchronize()
assert not bad, bad
print("ok")
""" % (os.path.dirname(HERE), HERE)
    out = subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=120)
    assert out.returncode == 0 and "ok" in out.stdout, out.stderr[-2000:]


def test_unsupported_dtype_is_cast_by_the_wrapper():
    import prim3d
    g = torch.from_numpy(inputs.noise((12, 12, 12), 26)).cuda()
    v0, f0 = prim3d.marching_cubes(g, 0.0)
    v1, f1 = prim3d.marching_cubes((g > 0).to(torch.int8) * 2 - 1, 0.0)   # int8 is not read directly
    v2, f2 = prim3d.marching_cubes(((g > 0).to(torch.float32) * 2 - 1), 0.0)
    assert torch.equal(v1, v2) and torch.equal(f1, f2) and f0.shape == f1.shape
    with pytest.raises(RuntimeError, match="expected scalar type Float"):
        prim3d._C.marching_cubes((g > 0).to(torch.int8), 0.0, [0, 0, 0], [12, 12, 12])


def test_deterministic_and_misaligned_input():
    grid = inputs.noise((40, 48, 64), 21)
    a = run_capi(grid, 0.0, None, None)
    b = run_capi(grid, 0.0, None, None)
    assert np.array_equal(a[0].view(np.uint32), b[0].view(np.uint32)) and np.array_equal(a[1], b[1])
    # a grid whose base pointer is only 4-byte aligned takes the generic (non-TMA) staging path
    from primitive3d_b200 import capi
    buf = torch.empty(grid.size + 1, dtype=torch.float32, device="cuda")
    g = buf[1:].view(grid.shape)
    g.copy_(torch.from_numpy(grid))
    assert g.data_ptr() % 16 != 0
    v, f = capi.marching_cubes(g, 0.0)
    assert np.array_equal(v.cpu().numpy().view(np.uint32), a[0].view(np.uint32))
    assert np.array_equal(f.cpu().numpy(), a[1])


def test_workspace_reuse_and_errors():
    from primitive3d_b200 import capi
    grid = torch.from_numpy(inputs.noise((20, 20, 20), 3)).cuda()
    desc = capi.McDesc.make(grid.shape, 0.0)
    V, F, ws, _ = capi.mc_count(desc, grid)
    V2, F2, _, _ = capi.mc_count(desc, grid, ws)  # same workspace, second run
    assert (V, F) == (V2, F2)
    small = torch.empty(16, dtype=torch.uint8, device="cuda")
    with pytest.raises(capi.P3DError) as e:
        capi.mc_count(desc, grid, small)
    assert e.value.status == capi.P3D_ERR_WORKSPACE


@pytest.mark.parametrize("name", ["bunny66", "noise33_s0", "gyroid128", "bounds_asym", "empty"])
def test_one_shot_c_entry_with_a_cudamalloc_callback(name):
    """p3d_mc_run (include/prim3d_b200.h) the way a C caller uses it: no torch on the path, device memory from a
    callback that calls cudaMalloc, results read back with cudaMemcpy."""
    import ctypes
    from primitive3d_b200 import capi
    make, thresh, lower, upper = CASES[name]
    grid = np.ascontiguousarray(make(), dtype=np.float32)
    rt = ctypes.CDLL("libcudart.so.12")
    rt.cudaMalloc.argtypes = [ctypes.POINTER(ctypes.c_void_p), ctypes.c_size_t]
    rt.cudaMemcpy.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t, ctypes.c_int]
    rt.cudaFree.argtypes = [ctypes.c_void_p]
    torch.cuda.init()
    torch.zeros(1, device="cuda")  # a context on device 0
    blocks = []

    @ctypes.CFUNCTYPE(ctypes.c_void_p, ctypes.c_void_p, ctypes.c_size_t)
    def alloc(_ctx, nbytes):
        ptr = ctypes.c_void_p()
        if rt.cudaMalloc(ctypes.byref(ptr), max(nbytes, 1)) != 0:
            return None
        blocks.append(ptr.value)
        return ptr.value

    L = capi.lib()
    L.p3d_mc_run.restype = ctypes.c_int
    L.p3d_mc_run.argtypes = [ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p, ctypes.c_void_p,
                             ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_void_p),
                             ctypes.POINTER(ctypes.c_int64), ctypes.POINTER(ctypes.c_int64), ctypes.c_void_p]
    try:
        gptr = ctypes.c_void_p(alloc(None, grid.nbytes))
        assert rt.cudaMemcpy(gptr, grid.ctypes.data, grid.nbytes, 1) == 0
        desc = capi.McDesc.make(grid.shape, thresh, lower, upper)
        vp, fp = ctypes.c_void_p(), ctypes.c_void_p()
        nv, nf = ctypes.c_int64(-1), ctypes.c_int64(-1)
        capi.check(L.p3d_mc_run(ctypes.byref(desc), gptr, ctypes.cast(alloc, ctypes.c_void_p), None, ctypes.byref(vp),
                                ctypes.byref(fp), ctypes.byref(nv), ctypes.byref(nf), None))
        v = np.empty((nv.value, 3), np.float32)
        f = np.empty((nf.value, 3), np.int32)
        if v.size:
            assert rt.cudaMemcpy(v.ctypes.data, vp, v.nbytes, 2) == 0
        if f.size:
            assert rt.cudaMemcpy(f.ctypes.data, fp, f.nbytes, 2) == 0
    finally:
        for b in blocks:
            rt.cudaFree(ctypes.c_void_p(b))
    ov, of = mc.marching_cubes(grid, thresh, lower, upper)
    assert v.shape == ov.shape and f.shape == of.shape
    assert_same_mesh(v, f, ov, of, ordered_faces=True)


@pytest.mark.parametrize("name", ["noise33_s0", "gyroid128", "noise_long_rows", "bounds_asym"])
def test_vertex_capacity_paths_agree(name):
    """The speculative vertex buffer of the counting pass, a too-small one (vertices-only second pass) and
    no buffer at all give bit-identical vertices; a partial buffer holds exactly the ids below its capacity."""
    from primitive3d_b200 import capi
    make, thresh, lower, upper = CASES[name]
    g = torch.from_numpy(np.ascontiguousarray(make())).cuda()
    desc = capi.McDesc.make(g.shape, thresh, lower, upper)
    V, F, ws, vbuf = capi.mc_count(desc, g, vertex_capacity=None)
    big = capi.mc_count(desc, g, vertex_capacity=V + 100)
    assert (big[0], big[1]) == (V, F)
    want = big[3][:V].clone()
    faces = capi.mc_faces(desc, big[2], F)
    half = max(V // 2, 1)
    V3, F3, ws3, part = capi.mc_count(desc, g, vertex_capacity=half)
    assert (V3, F3) == (V, F) and part.shape[0] == half
    assert torch.equal(part.view(torch.int32), want[:half].view(torch.int32))
    full = capi.mc_vertices(desc, g, ws3, V, part)          # V > capacity: second pass
    assert full.shape[0] == V and torch.equal(full.view(torch.int32), want.view(torch.int32))
    assert torch.equal(capi.mc_faces(desc, ws3, F), faces)
    V4, F4, ws4, none = capi.mc_count(desc, g, vertex_capacity=0)
    assert (V4, F4) == (V, F) and none.shape[0] == 0
    assert torch.equal(capi.mc_vertices(desc, g, ws4, V, none).view(torch.int32), want.view(torch.int32))
    if vbuf.shape[0] >= V:
        assert torch.equal(vbuf[:V].view(torch.int32), want.view(torch.int32))


def test_generic_loader_matches_tma(monkeypatch):
    """rz % 4 != 0 or a misaligned base take the non-TMA staging path of the tile kernel; same shapes through
    both loaders must agree (the TMA path needs rz % 4 == 0, so compare on such a grid in a subprocess with
    P3D_MC_LOADER=generic)."""
    import subprocess
    import sys
    code = ("import numpy as np, torch, sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from oracle import inputs; from primitive3d_b200 import capi;"
            "g = torch.from_numpy(inputs.noise((19, 21, 260), 5)).cuda();"
            "v, f = capi.marching_cubes(g, 0.0); torch.cuda.synchronize();"
            "np.savez(sys.argv[1], v=v.cpu().numpy(), f=f.cpu().numpy())") % (os.path.dirname(HERE), HERE)
    outs = []
    for mode in ["tma", "generic"]:
        path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"p3d_loader_{mode}_{os.getpid()}.npz")
        env = dict(os.environ, P3D_MC_LOADER=mode)
        subprocess.check_call([sys.executable, "-c", code, path], env=env)
        outs.append(np.load(path))
        os.remove(path)
    assert np.array_equal(outs[0]["v"].view(np.uint32), outs[1]["v"].view(np.uint32))
    assert np.array_equal(outs[0]["f"], outs[1]["f"])


@pytest.mark.parametrize("shape", [(6, 7, 1024), (5, 9, 700), (3, 4, 2100), (9, 40, 2048), (3, 3, 4992)])
def test_both_forms_of_the_face_pass_agree(shape):
    """The face pass exists in a row-streaming form (mc_faces_rows.cu; the default for rows of up to 1024 samples)
    and in the chunk form (k_faces); P3D_MC_FACES forces one of them where both apply.  Same vertex numbering, so
    the face arrays are identical -- and equal to the oracle's triangles."""
    import subprocess
    import sys
    code = ("import numpy as np, torch, sys; sys.path.insert(0, %r); sys.path.insert(0, %r);"
            "from oracle import inputs; from primitive3d_b200 import capi;"
            "g = torch.from_numpy(inputs.noise(%r, 31)).cuda();"
            "v, f = capi.marching_cubes(g, 0.0); torch.cuda.synchronize();"
            "np.savez(sys.argv[1], v=v.cpu().numpy(), f=f.cpu().numpy())") % (os.path.dirname(HERE), HERE, tuple(shape))
    outs = []
    for mode in ["rows", "chunks"]:
        path = os.path.join(os.environ.get("TMPDIR", "/tmp"), f"p3d_faces_{mode}_{os.getpid()}.npz")
        subprocess.check_call([sys.executable, "-c", code, path], env=dict(os.environ, P3D_MC_FACES=mode))
        outs.append(np.load(path))
        os.remove(path)
    assert np.array_equal(outs[0]["v"].view(np.uint32), outs[1]["v"].view(np.uint32))
    assert np.array_equal(outs[0]["f"], outs[1]["f"])
    ov, of = mc.marching_cubes(inputs.noise(tuple(shape), 31), 0.0)
    assert_same_mesh(outs[0]["v"], outs[0]["f"], ov, of, ordered_faces=True)


@pytest.mark.parametrize("n,V,F", [(512, 10111488, 20157724)])
def test_gyroid_known_counts_large(n, V, F):
    """SURVEY.md Appendix B known answers at a size the oracle still checks in seconds."""
    from primitive3d_b200 import capi
    g = torch.from_numpy(inputs.gyroid(n)).cuda()
    desc = capi.McDesc.make(g.shape, 0.0)
    v, f, ws, vbuf = capi.mc_count(desc, g)
    assert (v, f) == (V, F)
    verts, faces = capi.mc_vertices(desc, g, ws, v, vbuf), capi.mc_faces(desc, ws, f)
    ov, of = mc.marching_cubes(g.cpu().numpy(), 0.0)
    assert_same_mesh(verts.cpu().numpy(), faces.cpu().numpy(), ov, of, ordered_faces=True)


def _gyroid_cuda(n, x0=0, x1=None):
    s, c = (torch.from_numpy(t).cuda() for t in inputs.gyroid_tables(n))
    x1 = n if x1 is None else x1
    g = s[x0:x1, None, None] * c[None, :, None]
    g = g + s[None, :, None] * c[None, None, :]
    g = g + s[None, None, :] * c[x0:x1, None, None]
    return g.contiguous()


def test_gyroid_generator_is_bit_identical_on_gpu():
    assert np.array_equal(_gyroid_cuda(64).cpu().numpy().view(np.uint32), inputs.gyroid(64).view(np.uint32))


def _sorted_rows_u32(rows):
    """[N, C] int64 keys (32 bits each) sorted lexicographically, on the GPU (stable sorts, last column first)."""
    order = torch.arange(rows.shape[0], device=rows.device)
    for c in range(rows.shape[1] - 1, -1, -1):
        order = order[torch.sort(rows[order, c], stable=True).indices]
    return rows[order]


def test_gyroid1024_full_size_against_the_oracle():
    """BASELINE config 3 at FULL size against the 64-bit oracle: known-answer counts, the vertex multiset bit for
    bit, and every triangle (9 corner coordinates, emitted corner order) in the oracle's order -- both sides emit
    faces voxel-major (marching_cubes.cu:194-208), so no sorting is involved.  Plus the size-independent properties:
    every index used, every vertex on exactly one grid edge, every triangle inside one cell."""
    from primitive3d_b200 import capi
    n = 1024
    g = _gyroid_cuda(n)
    desc = capi.McDesc.make(g.shape, 0.0)
    V, F, ws, vbuf = capi.mc_count(desc, g)
    assert (V, F) == (40621056, 81103132)
    verts, faces = capi.mc_vertices(desc, g, ws, V, vbuf), capi.mc_faces(desc, ws, F)
    del vbuf, ws
    assert int(faces.min()) == 0 and int(faces.max()) == V - 1
    used = torch.zeros(V, dtype=torch.bool, device="cuda")
    used[faces.reshape(-1).long()] = True
    assert bool(used.all())
    frac = verts - torch.floor(verts)
    assert int(((frac != 0).sum(1) > 1).sum()) == 0
    del used, frac
    # the oracle on the same grid (OpenMP, 64-bit indices: a few seconds)
    ov, of = mc.marching_cubes(g.cpu().numpy(), 0.0)
    del g
    assert ov.shape == (V, 3) and of.shape == (F, 3)
    ov, of = torch.from_numpy(ov).cuda(), torch.from_numpy(of).cuda()
    # triangles, in order: corner coordinates as raw bits, chunked to bound the temporaries
    step = 1 << 23
    for a in range(0, F, step):
        ours = verts[faces[a:a + step].reshape(-1).long()].view(torch.int32)
        want = ov[of[a:a + step].reshape(-1).long()].view(torch.int32)
        assert torch.equal(ours, want), f"triangles [{a}, {a + step}) differ from the oracle's"
        tri = verts[faces[a:a + step].reshape(-1).long()].reshape(-1, 3, 3)
        assert float((tri.max(1).values - tri.min(1).values).max()) <= 1.0
    del ours, want, tri
    # vertex multisets, bit for bit
    mine = _sorted_rows_u32(verts.view(torch.int32).long() & 0xffffffff)
    theirs = _sorted_rows_u32(ov.view(torch.int32).long() & 0xffffffff)
    assert torch.equal(mine, theirs)
