from .marching_cubes import marching_cubes, save_mesh
from .marching_tetrahedras import marching_tetrahedras
from .ray_cast import create_raycaster

__all__ = ["create_raycaster", "marching_cubes", "save_mesh", "marching_tetrahedras"]
