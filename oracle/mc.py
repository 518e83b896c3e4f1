"""ctypes front end of oracle/mc_oracle.c -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

`marching_cubes(grid, thresh, lower, upper)` restates
/root/reference/src/prim3d/Utility/marching_cubes.cu:212-305 on the CPU with a
fixed output order (vertex ids in (x, y, z, axis) order, faces in cell order).
"""
import ctypes
import os

import numpy as np

_LIB_PATH = os.path.join(os.path.dirname(os.path.abspath(__file__)), "_build", "libp3d_oracle.so")
_lib = None


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing; run `make -C oracle` (or __graft_entry__.build())")
        # libgomp reads these once, when it is first loaded; without a binding policy the sandbox
        # kernels used here were seen to stack every OpenMP thread on one core (no speed-up at all)
        os.environ.setdefault("OMP_PROC_BIND", "spread")
        os.environ.setdefault("OMP_PLACES", "cores")
        _lib = ctypes.CDLL(_LIB_PATH)
        i64, f32, vp = ctypes.c_int64, ctypes.c_float, ctypes.c_void_p
        _lib.p3d_oracle_mc_count.restype = ctypes.c_int
        _lib.p3d_oracle_mc_count.argtypes = [vp, i64, i64, i64, f32, ctypes.POINTER(i64),
                                             ctypes.POINTER(i64), ctypes.c_int]
        _lib.p3d_oracle_mc_extract.restype = ctypes.c_int
        _lib.p3d_oracle_mc_extract.argtypes = [vp, i64, i64, i64, f32, vp, vp, vp, vp, ctypes.c_int]
        _lib.p3d_oracle_mc_table.argtypes = [vp]
    return _lib


def triangle_table():
    """The expanded int8[256][16] Bourke table (marching_cubes.h:21-277)."""
    out = np.empty((256, 16), np.int8)
    lib().p3d_oracle_mc_table(out.ctypes.data)
    return out


def _as_grid(grid):
    g = np.ascontiguousarray(grid, dtype=np.float32)
    if g.ndim != 3:
        raise ValueError("grid must be 3-D")
    return g


def count(grid, thresh, threads=0):
    g = _as_grid(grid)
    V, F = ctypes.c_int64(), ctypes.c_int64()
    rc = lib().p3d_oracle_mc_count(g.ctypes.data, *g.shape, float(thresh), ctypes.byref(V),
                                   ctypes.byref(F), threads)
    if rc:
        raise RuntimeError(f"p3d_oracle_mc_count failed rc={rc}")
    return V.value, F.value


def marching_cubes(grid, thresh, lower=None, upper=None, threads=0):
    """-> (vertices float32 [V,3], faces int32 [F,3]); bounds default to the wrapper's
    scale=None case, lower=0 and upper=shape (prim3d/utility/marching_cubes.py:59-62)."""
    g = _as_grid(grid)
    lo = np.asarray([0.0, 0.0, 0.0] if lower is None else lower, dtype=np.float32)
    up = np.asarray(list(g.shape) if upper is None else upper, dtype=np.float32)
    V, F = count(g, thresh, threads)
    verts = np.empty((V, 3), np.float32)
    faces = np.empty((F, 3), np.int32)
    rc = lib().p3d_oracle_mc_extract(g.ctypes.data, *g.shape, float(thresh), lo.ctypes.data,
                                     up.ctypes.data, verts.ctypes.data, faces.ctypes.data, threads)
    if rc:
        raise RuntimeError(f"p3d_oracle_mc_extract failed rc={rc}")
    return verts, faces
