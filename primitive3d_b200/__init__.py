"""primitive3d_b200 -- B200-native (sm_100a) kernels for the one hot path of lzhnb/Primitive3D this
repository replaces: dense-grid marching cubes and marching tetrahedra.

  csrc/            hand-written CUDA kernels + the torch-free C ABI (include/prim3d_b200.h)
  capi.py          ctypes view of the C ABI over torch device memory (tests, sharded driver)
  sharded.py       dim-0 slab decomposition across GPUs (one process per GPU, NCCL counts)
  build.py         in-tree build of libprim3d_b200.so and prim3d/libPrim3D.so

The reference-facing API is the sibling package `prim3d` (same names as the reference).
There is no CPU implementation in this package; loading fails loudly if the library is unbuilt.
"""
from .capi import (McDesc, abi_version, lib, marching_cubes_batch, marching_cubes_host, mc_count, mc_extract, mc_faces, mc_vertices,  # noqa: F401
                   mc_workspace_bytes)

__all__ = ["McDesc", "abi_version", "lib", "marching_cubes_batch", "marching_cubes_host", "mc_count", "mc_extract", "mc_faces", "mc_vertices",
           "mc_workspace_bytes"]
