"""PyMCubes-compatible `mcubes` stand-in -- TEST INFRASTRUCTURE, NOT PRODUCT CODE.

The reference's examples and its `cpu=True` path import the third-party package
`mcubes` (examples/sphere.py:2, examples/bunny_sdf.py:4,
prim3d/utility/marching_cubes.py:66-81).  PyMCubes is not installable in this
image, so tests and bench.py put this directory on sys.path instead.  It exposes
the two calls the reference makes, `marching_cubes(volume, isovalue)` and
`export_obj(vertices, triangles, filename)`, backed by the single-threaded C
restatement in oracle/pymcubes_compat.c (parity with real PyMCubes is UNPINNED;
see that file's header).
"""
import ctypes
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "..", "..", "_build", "libp3d_pymcubes_compat.so")
_lib = None


def _load():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            raise ImportError(f"{_LIB_PATH} is missing; run `make -C oracle` (or __graft_entry__.build())")
        _lib = ctypes.CDLL(_LIB_PATH)
        _lib.p3d_mcubes_marching_cubes.restype = ctypes.c_int
        _lib.p3d_mcubes_marching_cubes.argtypes = [
            ctypes.c_void_p, ctypes.c_int64, ctypes.c_int64, ctypes.c_int64, ctypes.c_double,
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64),
            ctypes.POINTER(ctypes.c_void_p), ctypes.POINTER(ctypes.c_int64)]
        _lib.p3d_mcubes_free.argtypes = [ctypes.c_void_p]
    return _lib


def marching_cubes(volume, isovalue):
    """(vertices float64 [V,3] in index coordinates, triangles uint64 [F,3])."""
    lib = _load()
    vol = np.ascontiguousarray(volume, dtype=np.float64)
    if vol.ndim != 3:
        raise ValueError("Only three-dimensional arrays are supported.")
    vp, tp = ctypes.c_void_p(), ctypes.c_void_p()
    nv, nt = ctypes.c_int64(), ctypes.c_int64()
    rc = lib.p3d_mcubes_marching_cubes(vol.ctypes.data, vol.shape[0], vol.shape[1], vol.shape[2],
                                       float(isovalue), ctypes.byref(vp), ctypes.byref(nv),
                                       ctypes.byref(tp), ctypes.byref(nt))
    if rc:
        raise MemoryError("marching_cubes: out of memory")
    try:
        if nv.value:
            verts = np.ctypeslib.as_array(ctypes.cast(vp, ctypes.POINTER(ctypes.c_double)),
                                          shape=(nv.value, 3)).copy()
        else:
            verts = np.zeros((0, 3), np.float64)
        if nt.value:
            tris = np.ctypeslib.as_array(ctypes.cast(tp, ctypes.POINTER(ctypes.c_uint64)),
                                         shape=(nt.value, 3)).copy()
        else:
            tris = np.zeros((0, 3), np.uint64)
    finally:
        lib.p3d_mcubes_free(vp)
        lib.p3d_mcubes_free(tp)
    return verts, tris


def export_obj(vertices, triangles, filename):
    """Wavefront OBJ, 1-based face indices."""
    with open(filename, "w") as fh:
        for v in np.asarray(vertices):
            fh.write("v {} {} {}\n".format(*v))
        for f in np.asarray(triangles):
            fh.write("f {} {} {}\n".format(*(int(i) + 1 for i in f)))
