"""`prim3d.create_raycaster` -- import-surface only.

Ray casting (reference: prim3d/utility/ray_cast.py, src/prim3d/Utility/ray_cast.cu, OptiX/BVH)
is outside this repository's scope: SURVEY.md section 2 rows 11-14.  The name is kept because the
reference's package exports it; calling it raises.
"""
import prim3d.libPrim3D as _C


def create_raycaster(vertices, faces) -> "_C.RayCaster":
    return _C.create_raycaster(vertices, faces)
