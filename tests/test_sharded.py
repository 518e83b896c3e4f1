"""Dim-0 slab sharding (primitive3d_b200/sharded.py).

CPU: the host-side logic (plane ranges, count exchange over a real world_size-2 gloo group,
offsets).  GPU: the shard kernels with halo planes and the first-plane table exchange, run as
"virtual ranks" one after another on one GPU -- concatenating the shards must reproduce the
single-GPU mesh: the same triangles in the same voxel-major order (corner coordinates bit-identical)
and the same vertex multiset (vertex ids are numbered shard by shard, so the arrays are permuted)."""
import ctypes
import os

import numpy as np
import pytest
import torch

from oracle import inputs


def test_slab_ranges_partition_the_grid():
    from primitive3d_b200.sharded import slab_range, slab_with_halo
    for n, w in [(1024, 8), (2048, 8), (10, 3), (7, 7), (5, 2)]:
        nxt = 0
        for r in range(w):
            x0, x1 = slab_range(n, w, r)
            assert x0 == nxt and x1 > x0
            nxt = x1
            h0, h1 = slab_with_halo(n, w, r)
            assert h0 == x0 and h1 == (x1 + 1 if r + 1 < w else n)
        assert nxt == n
        sizes = [slab_range(n, w, r)[1] - slab_range(n, w, r)[0] for r in range(w)]
        assert max(sizes) - min(sizes) <= 1


def test_exclusive_offsets():
    from primitive3d_b200.sharded import exclusive_offsets
    counts = [(5, 9), (0, 0), (7, 13)]
    assert exclusive_offsets(counts, 0) == (0, 0, 12, 22)
    assert exclusive_offsets(counts, 1) == (5, 9, 12, 22)
    assert exclusive_offsets(counts, 2) == (5, 9, 12, 22)


def test_unpack_counts_reads_the_payload_tails():
    """Host side of the device-driven exchange: every payload is [table words ... , V lo, V hi, F lo, F hi]."""
    from primitive3d_b200.sharded import exclusive_offsets, unpack_counts
    world, words = 3, 12 + 4
    gathered = torch.zeros(world * words, dtype=torch.int32)
    expect = [(5, 9), (2 ** 33 + 1, 2 ** 35 + 7), (0, 0)]
    for r, (v, f) in enumerate(expect):
        gathered.view(world, words)[r, :12] = 100 * r + torch.arange(12, dtype=torch.int32)
        gathered.view(world, words)[r, 12:] = torch.tensor([v, f], dtype=torch.int64).view(torch.int32)
    counts = unpack_counts(gathered, world, words)
    assert counts == expect
    assert exclusive_offsets(counts, 2) == (5 + 2 ** 33 + 1, 9 + 2 ** 35 + 7, 5 + 2 ** 33 + 1, 9 + 2 ** 35 + 7)


def test_count_exchange_world_size_2_gloo():
    import json
    import socket
    import subprocess
    import sys
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    worker = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_gloo_worker.py")
    procs = [subprocess.Popen([sys.executable, worker, str(r), "2", str(port)], stdout=subprocess.PIPE,
                              stderr=subprocess.PIPE, text=True) for r in range(2)]
    outs = []
    for p in procs:
        out, err = p.communicate(timeout=180)
        assert p.returncode == 0, err[-2000:]
        outs.append(json.loads(out.strip().splitlines()[-1]))
    outs.sort(key=lambda o: o["rank"])
    assert outs[0]["counts"] == outs[1]["counts"] == [[100, 1000], [101, 1010]]
    assert outs[0]["offsets"] == [0, 0, 201, 2010] and outs[1]["offsets"] == [100, 1000, 201, 2010]
    # the fused exchange: tables of every rank + 64-bit counts in one all-gather
    for o in outs:
        assert o["counts2"] == [[2 ** 33, 2 ** 40], [2 ** 33 + 1, 2 ** 40 + 10]]
        assert o["tables"] == [list(range(24)), [1000 + i for i in range(24)]]


@pytest.mark.gpu
@pytest.mark.parametrize("shape,world,seed", [((12, 20, 40), 2, 0), ((13, 9, 70), 3, 1), ((8, 33, 1056), 4, 2),
                                               ((64, 64, 64), 8, 3)])
def test_virtual_shards_reproduce_single_gpu_result(shape, world, seed):
    from primitive3d_b200 import capi
    from primitive3d_b200.sharded import exclusive_offsets, slab_range, slab_with_halo
    L = capi.lib()
    grid = torch.from_numpy(inputs.noise(shape, seed)).cuda()
    n = shape[0]
    lower, upper = [-1.0, 0.0, 1.0], [2.0, 3.0, 5.0]
    ref_v, ref_f = capi.marching_cubes(grid, 0.1, lower, upper)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)

    shards = []
    for r in range(world):
        x0, x1 = slab_range(n, world, r)
        _, xh = slab_with_halo(n, world, r)
        slab = grid[x0:xh].contiguous()
        desc = capi.McDesc.make(slab.shape, 0.1, lower, upper, owned_x=x1 - x0, x_origin=x0, global_rx=n)
        V, F, ws, vbuf = capi.mc_count(desc, slab)
        table = torch.empty(L.p3d_mc_plane_table_words(ctypes.byref(desc)), dtype=torch.int32, device="cuda")
        capi.check(L.p3d_mc_export_first_plane(ctypes.byref(desc), ws.data_ptr(), table.data_ptr(), stream))
        shards.append(dict(slab=slab, desc=desc, V=V, F=F, ws=ws, table=table, vbuf=vbuf))
    counts = [(s["V"], s["F"]) for s in shards]
    assert sum(c[0] for c in counts) == ref_v.shape[0] and sum(c[1] for c in counts) == ref_f.shape[0]
    vs, fs = [], []
    for r, s in enumerate(shards):
        v_off = exclusive_offsets(counts, r)[0]
        if r + 1 < world:
            capi.check(L.p3d_mc_import_halo_plane(ctypes.byref(s["desc"]), s["ws"].data_ptr(),
                                                  shards[r + 1]["table"].data_ptr(), s["V"], stream))
        vs.append(capi.mc_vertices(s["desc"], s["slab"], s["ws"], s["V"], s["vbuf"]))
        fs.append(capi.mc_faces(s["desc"], s["ws"], s["F"], v_off))
    torch.cuda.synchronize()
    from canonical import assert_same_mesh
    assert_same_mesh(torch.cat(vs).cpu().numpy(), torch.cat(fs).cpu().numpy(), ref_v.cpu().numpy(), ref_f.cpu().numpy(),
                     ordered_faces=True)


@pytest.mark.gpu
@pytest.mark.parametrize("shape,world,seed", [((12, 20, 40), 2, 4), ((13, 9, 270), 3, 5), ((64, 64, 64), 8, 6)])
def test_virtual_shards_device_side_exchange(shape, world, seed):
    """The multi-GPU fast path (p3d_mc_tile_async -> p3d_mc_export_exchange -> all-gather -> p3d_mc_faces_exchanged)
    with the all-gather played by a concatenation on one GPU: vertex id bases and halo numbering are computed on the
    device; the concatenated shards must equal the single-GPU arrays, whatever the face capacity guess."""
    from primitive3d_b200 import capi
    from primitive3d_b200.sharded import slab_range, slab_with_halo, unpack_counts
    L = capi.lib()
    grid = torch.from_numpy(inputs.noise(shape, seed)).cuda()
    n = shape[0]
    ref_v, ref_f = capi.marching_cubes(grid, -0.05)
    stream = ctypes.c_void_p(torch.cuda.current_stream().cuda_stream)
    shards, payloads = [], []
    for r in range(world):
        x0, x1 = slab_range(n, world, r)
        slab = grid[x0:slab_with_halo(n, world, r)[1]].contiguous()
        desc = capi.McDesc.make(slab.shape, -0.05, owned_x=x1 - x0, x_origin=x0, global_rx=n,
                                upper=[float(s) for s in shape])
        ws = torch.empty(capi.mc_workspace_bytes(desc), dtype=torch.uint8, device="cuda")
        vbuf = torch.empty((slab.numel() * 3, 3), dtype=torch.float32, device="cuda")
        capi.check(L.p3d_mc_tile_async(ctypes.byref(desc), slab.data_ptr(), 0, ws.data_ptr(), ws.numel(), vbuf.data_ptr(),
                                       vbuf.shape[0], stream))
        words = L.p3d_mc_exchange_words(ctypes.byref(desc))
        mine = torch.empty(words, dtype=torch.int32, device="cuda")
        capi.check(L.p3d_mc_export_exchange(ctypes.byref(desc), ws.data_ptr(), mine.data_ptr(), stream))
        shards.append(dict(desc=desc, ws=ws, vbuf=vbuf, slab=slab))
        payloads.append(mine)
    gathered = torch.cat(payloads)
    counts = unpack_counts(gathered, world, words)
    assert sum(c[0] for c in counts) == ref_v.shape[0] and sum(c[1] for c in counts) == ref_f.shape[0]
    vs, fs = [], []
    for r, s in enumerate(shards):
        V, F = counts[r]
        for fcap in (F, F + 3):
            fbuf = torch.full((fcap, 3), -7, dtype=torch.int32, device="cuda")
            capi.check(L.p3d_mc_faces_exchanged(ctypes.byref(s["desc"]), s["ws"].data_ptr(), gathered.data_ptr(), r, world,
                                                fbuf.data_ptr(), fcap, stream))
        if F:   # too small a guess: nothing is written
            small = torch.full((F - 1 if F > 1 else 1, 3), -7, dtype=torch.int32, device="cuda")
            capi.check(L.p3d_mc_faces_exchanged(ctypes.byref(s["desc"]), s["ws"].data_ptr(), gathered.data_ptr(), r, world,
                                                small.data_ptr(), F - 1, stream))
            assert bool((small == -7).all())
        vs.append(s["vbuf"][:V])
        fs.append(fbuf[:F])
        assert bool((fbuf[F:] == -7).all())
    torch.cuda.synchronize()
    from canonical import assert_same_mesh
    assert_same_mesh(torch.cat(vs).cpu().numpy(), torch.cat(fs).cpu().numpy(), ref_v.cpu().numpy(), ref_f.cpu().numpy(),
                     ordered_faces=True)
