"""prim3d -- drop-in for lzhnb/Primitive3D's marching cubes / marching tetrahedra API on B200.

Public surface of the reference's prim3d/__init__.py:4-16: `marching_cubes`, `save_mesh`,
`marching_tetrahedras`, `create_raycaster`, `Timer`, `ENABLE_OPTIX`, `__version__`.
The compiled module `prim3d.libPrim3D` is a hard requirement exactly as in the reference
(prim3d/__init__.py:2): if it has not been built, importing this package fails -- there is
no Python or CPU fallback for the CUDA path.
"""
import torch  # noqa: F401  -- loads libtorch/libc10 before the extension module resolves them

try:
    import prim3d.libPrim3D as _C
except ImportError as exc:  # pragma: no cover - exercised only on an unbuilt tree
    raise ImportError(
        "prim3d.libPrim3D is not built; run `python -m primitive3d_b200.build` "
        "(or __graft_entry__.build()) first") from exc

from .version import __version__
from .misc import Timer
from .utility import create_raycaster, marching_cubes, marching_tetrahedras, save_mesh

ENABLE_OPTIX = _C.enable_optix

__all__ = ["__version__", "ENABLE_OPTIX", "Timer", "create_raycaster", "marching_cubes", "save_mesh",
           "marching_tetrahedras"]
