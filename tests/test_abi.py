"""CPU-side checks of the drop-in boundary: the C-ABI library loads and exports every symbol
include/prim3d_b200.h declares, the reference-facing module exposes the reference's names, and
the host-side Python mirror behaves like the reference's wrappers (no compute calls here)."""
import ctypes
import os
import re

import numpy as np
import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def declared_symbols():
    text = open(os.path.join(ROOT, "include", "prim3d_b200.h")).read()
    text = re.sub(r"/\*.*?\*/", "", text, flags=re.S)
    return sorted(set(re.findall(r"\b(p3d_[a-z0-9_]+)\s*\(", text)) - {"p3d_alloc_fn"})


def test_header_declares_the_expected_entry_points():
    syms = declared_symbols()
    for s in ["p3d_mc_workspace_bytes", "p3d_mc_vertex_capacity_hint", "p3d_mc_count", "p3d_mc_vertices", "p3d_mc_faces",
              "p3d_mc_count_typed", "p3d_mc_vertices_typed", "p3d_mc_extract", "p3d_mc_extract_sparse", "p3d_mc_extract_batch", "p3d_mc_batch_workspace_bytes", "p3d_mc_extract_host",
              "p3d_mc_extract_host_arena_bytes", "p3d_mc_tile_async", "p3d_mc_exchange_words", "p3d_mc_export_exchange",
              "p3d_mc_faces_exchanged", "p3d_mc_sharded_extract", "p3d_mc_sharded_extract_p2p", "p3d_mc_peer_create",
              "p3d_mc_peer_connect", "p3d_mc_peer_destroy", "p3d_mc_peer_handle_bytes", "p3d_ply_pack",
              "p3d_mc_run", "p3d_mc_plane_table_words", "p3d_mc_export_first_plane",
              "p3d_mc_import_halo_plane", "p3d_mt_classify", "p3d_mt_index", "p3d_mt_emit", "p3d_mt_backward", "p3d_mt_extract",
              "p3d_mt_extract_workspace_bytes",
              "p3d_last_error", "p3d_abi_version"]:
        assert s in syms


def test_library_exports_every_declared_symbol():
    from primitive3d_b200 import capi
    lib = ctypes.CDLL(capi.LIB_PATH)
    for s in declared_symbols():
        assert hasattr(lib, s), f"{s} declared in include/prim3d_b200.h but not exported"
    assert capi.abi_version() == 5


def test_workspace_size_is_a_few_bits_per_sample():
    from primitive3d_b200 import capi
    d = capi.McDesc.make((1024, 1024, 1024), 0.0)
    n = capi.mc_workspace_bytes(d)
    # 1 bit per sample + 20 bytes per 128-sample piece + its face prefix + scan state: < 0.30 B/sample
    # (the reference's vertex_grids alone is 12 B/sample, marching_cubes.cu:257-259)
    assert 1024 ** 3 // 8 < n < int(0.30 * 1024 ** 3)
    assert capi.lib().p3d_mc_plane_table_words(ctypes.byref(d)) == 1024 * 8 * 4
    assert capi.lib().p3d_mc_vertex_capacity_hint(ctypes.byref(d)) == 1024 ** 3 // 16 + 4096
    tiny = capi.McDesc.make((2, 2, 2), 0.0)
    assert capi.lib().p3d_mc_vertex_capacity_hint(ctypes.byref(tiny)) == 24
    bad = capi.McDesc.make((0, 4, 4), 0.0)
    assert capi.lib().p3d_mc_workspace_bytes(ctypes.byref(bad)) == 0


def test_module_surface_matches_reference():
    # reference: src/pybind/bindings.cpp:16-31, prim3d/__init__.py:13-16
    import prim3d
    for name in ["enable_optix", "test", "RayCaster", "create_raycaster", "marching_cubes", "save_mesh_as_ply"]:
        assert hasattr(prim3d._C, name)
    assert prim3d._C.enable_optix is False and prim3d.ENABLE_OPTIX is False
    assert prim3d.__version__ == "0.0.1"
    for name in ["__version__", "ENABLE_OPTIX", "Timer", "create_raycaster", "marching_cubes", "save_mesh",
                 "marching_tetrahedras"]:
        assert name in prim3d.__all__ and hasattr(prim3d, name)


def test_scale_to_bound_forms():
    # reference: prim3d/utility/marching_cubes.py:10-31
    from prim3d.utility.marching_cubes import scale_to_bound
    assert scale_to_bound(2.0) == ([0.0, 0.0, 0.0], [2.0, 2.0, 2.0])
    assert scale_to_bound([1.0, 2.0, 3.0]) == ([0.0, 0.0, 0.0], [1.0, 2.0, 3.0])
    assert scale_to_bound((-1.0, 1.0)) == ([-1.0] * 3, [1.0] * 3)
    assert scale_to_bound([[0.0, 1.0, 2.0], [3.0, 4.0, 5.0]]) == ([0.0, 1.0, 2.0], [3.0, 4.0, 5.0])
    lo, up = scale_to_bound(np.array([1.0, 2.0, 3.0]))
    assert up == [1.0, 2.0, 3.0]
    for bad in (3, "x", [1.0], [1.0, 2.0, 3.0, 4.0]):
        with pytest.raises(TypeError):
            scale_to_bound(bad)


def test_timer_templates(capsys):
    # reference: prim3d/misc/utils.py:63-71,84-86
    import prim3d
    with prim3d.Timer("cpu:"):
        pass
    with prim3d.Timer("took {:.6f}s"):
        pass
    with prim3d.Timer():
        pass
    out = capsys.readouterr().out.strip().splitlines()
    assert re.fullmatch(r"cpu: \d+\.\d{3}", out[0])
    assert re.fullmatch(r"took \d+\.\d{6}s", out[1])
    assert re.fullmatch(r"\d+\.\d{3}", out[2])
    t = prim3d.Timer(start=False)
    with pytest.raises(Exception):
        t.since_start()
    t.start()
    assert t.is_running and t.since_start() >= 0 and t.since_last_check() >= 0


def test_cpu_mode_wraps_mcubes_and_no_silent_fallback():
    """cpu=True is the reference's mcubes wrapper (marching_cubes.py:66-81: float64 vertices,
    int64 faces, vertices / scale + offset); without cpu=True a missing CUDA device is an error."""
    import sys
    import torch
    import prim3d
    sys.path.insert(0, os.path.join(ROOT, "oracle", "pymcubes_compat"))
    import mcubes
    from oracle import inputs
    g = inputs.sphere_int64(40)
    v, f = prim3d.marching_cubes(torch.tensor(g), 0, cpu=True)
    mv, mf = mcubes.marching_cubes(g, 0)
    assert v.dtype == torch.float64 and f.dtype == torch.int64
    assert (v.numpy() == mv).all() and (f.numpy() == mf).all()     # examples/sphere.py:29-30
    if not torch.cuda.is_available():
        with pytest.raises(RuntimeError, match="CUDA"):
            prim3d.marching_cubes(torch.tensor(g), 0)
        with pytest.raises(RuntimeError, match="CUDA"):
            prim3d.marching_tetrahedras(torch.zeros(4, 3), torch.zeros(1, 4, dtype=torch.long), torch.zeros(4))
        from primitive3d_b200 import capi
        with pytest.raises(RuntimeError, match="CUDA"):       # host-resident grids still need the device
            capi.marching_cubes_host(torch.zeros(8, 8, 8), 0.0)
        with pytest.raises(ValueError):                       # and the batch / staged calls want CUDA tensors
            capi.marching_cubes_batch([torch.zeros(8, 8, 8)], 0.0)


def test_host_and_batch_entry_points_validate_their_arguments():
    import torch
    from primitive3d_b200 import capi
    with pytest.raises(ValueError):
        capi.marching_cubes_host(torch.zeros(8, 8, 16)[:, :, ::2], 0.0)      # not contiguous
    with pytest.raises(ValueError):
        capi.marching_cubes_host(torch.zeros(8, 8), 0.0)                      # not 3-D
    with pytest.raises(ValueError):
        capi.marching_cubes_host(torch.zeros(8, 8, 8, dtype=torch.int8), 0.0)  # element type the kernels do not read
    assert capi.marching_cubes_batch([], 0.0) == []
    # descriptor-only helpers work without a device
    d = capi.McDesc.make((1024, 1024, 1024), 0.0)
    L = capi.lib()
    import ctypes
    arena = L.p3d_mc_extract_host_arena_bytes(ctypes.byref(d), 0, 0)
    assert 2 * 65 * 4 * 1024 ** 2 < arena < 2 * 1024 ** 3          # two 65-plane slabs + workspaces + output buffers
    assert L.p3d_mc_exchange_words(ctypes.byref(d)) == 1024 * 8 * 4 + 4


def test_save_mesh_ply_bytes(tmp_path):
    """save_mesh writes the reference's binary PLY (marching_cubes.cu:307-352): same header text,
    15-byte vertex records, faces as int32 [3,a,b,c]."""
    import torch
    import prim3d
    v = torch.tensor([[0.0, 0.5, 1.0], [1.0, 2.0, 3.0], [4.0, 5.0, 6.0]])
    f = torch.tensor([[0, 1, 2]], dtype=torch.int64)
    path = str(tmp_path / "m.ply")
    prim3d.save_mesh(v, f, filename=path)
    raw = open(path, "rb").read()
    head = (b"ply\nformat binary_little_endian 1.0\nelement vertex 3\nproperty float x\nproperty float y\n"
            b"property float z\nproperty uchar red\nproperty uchar green\nproperty uchar blue\n"
            b"element face 1\nproperty list int int vertex_index\nend_header\n")
    assert raw.startswith(head)
    body = raw[len(head):]
    assert len(body) == 3 * 15 + 16
    rec = np.frombuffer(body[:45], dtype=np.uint8).reshape(3, 15)
    assert np.array_equal(rec[:, :12].copy().view(np.float32), v.numpy())
    assert (rec[:, 12:] == 127).all()
    assert np.frombuffer(body[45:], dtype=np.int32).tolist() == [3, 0, 1, 2]
    with pytest.raises(NotImplementedError):
        prim3d.save_mesh(v, f, filename=str(tmp_path / "m.obj"))


def test_marching_tets_packed_tables_match_reference_tables():
    """The nibble-packed tables in mt_common.cuh against the reference's tables
    (marching_tetrahedras.py:7-43) as restated in oracle/mt.py."""
    from oracle import mt
    text = open(os.path.join(ROOT, "primitive3d_b200/csrc/mt_common.cuh")).read()
    rows = re.search(r"c_tri_rows\[16\] = \{(.*?)\};", text, re.S).group(1)
    rows = [int(x, 16) for x in re.findall(r"0x([0-9a-f]+)", rows)]
    assert len(rows) == 16
    for code, w in enumerate(rows):
        got = [(w >> (4 * k)) & 15 for k in range(6)]
        want = [v if v >= 0 else 15 for v in mt.TRIANGLE_TABLE[code].tolist()]
        assert got == want, code
    nt = re.search(r"const int nt\[16\] = \{(.*?)\};", text).group(1)
    assert [int(x) for x in nt.split(",")] == mt.NUM_TRIANGLES.tolist()
