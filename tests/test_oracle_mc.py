"""The marching-cubes oracle against the known answers of SURVEY.md Appendix B and its own
invariants.  (Its comparison with the compiled reference runs on the GPU box:
tests/test_reference_cuda.py.)"""
import os
import sys

import numpy as np
import pytest

from oracle import inputs, mc

HERE = os.path.dirname(os.path.abspath(__file__))


def bunny():
    return np.load(os.path.join(HERE, "golden", "mc_bunny66.npz"))["grid"]


def bunny256():
    """BASELINE configs[1]: the bunny grid at 256^3."""
    return inputs.upsample_trilinear(bunny(), 256)


KNOWN = [  # (name, grid factory, V, F)  -- SURVEY.md Appendix B
    ("sphere200", lambda: inputs.sphere_int64(200).astype(np.float32), 11766, 23528),
    ("sphere128", lambda: inputs.sphere_int64(128).astype(np.float32), 11766, 23528),
    ("bunny66", bunny, 13282, 26560),
    ("gyroid128", lambda: inputs.gyroid(128), 635904, 1261852),
    ("gyroid256", lambda: inputs.gyroid(256), 2500608, 4972828),
    ("bunny256", bunny256, 204670, 409336),  # frozen from the oracle (pinned to the compiled reference on the GPU box)
]


@pytest.mark.parametrize("name,make,V,F", KNOWN, ids=[k[0] for k in KNOWN])
def test_known_counts(name, make, V, F):
    assert mc.count(make(), 0.0) == (V, F)


def test_bunny256_grid_is_reproducible():
    """The 256^3 bunny is defined by separately rounded fp32 numpy operations: same bits everywhere."""
    import hashlib
    g = bunny256()
    assert g.shape == (256, 256, 256) and g.dtype == np.float32
    assert hashlib.sha256(g.tobytes()).hexdigest().startswith("fcbb771afab4f064")
    v, f = mc.marching_cubes(g, 0.0)
    assert f.shape[0] == 2 * v.shape[0] - 4  # closed, genus 0, like the 66^3 original


def test_closed_surface_invariant():
    # closed genus-0 surfaces: F = 2V - 4
    v, f = mc.marching_cubes(bunny(), 0.0)
    assert f.shape[0] == 2 * v.shape[0] - 4
    assert f.min() == 0 and f.max() == v.shape[0] - 1


def test_positions_follow_reference_arithmetic():
    g = inputs.noise((9, 7, 5), seed=3)
    v, f = mc.marching_cubes(g, 0.1)
    # default bounds are lower=0, upper=shape; with the reference's y term (upper[2]-lower[1])/Ry
    # a non-cubic grid gets y scaled by Rz/Ry even with scale=None (marching_cubes.cu:295)
    scale = np.array([9 / np.float32(9), np.float32(5) / np.float32(7), 5 / np.float32(5)], np.float32)
    # recompute one axis by hand: vertices are numbered in (x, y, z, axis) order
    t = np.float32(0.1)
    expect = []
    for x in range(9):
        for y in range(7):
            for z in range(5):
                d0 = g[x, y, z]
                for axis, (dx, dy, dz) in enumerate([(1, 0, 0), (0, 1, 0), (0, 0, 1)]):
                    if x + dx >= 9 or y + dy >= 7 or z + dz >= 5:
                        continue
                    d1 = g[x + dx, y + dy, z + dz]
                    if (d0 > t) != (d1 > t):
                        p = np.array([x, y, z], np.float32)
                        p[axis] = p[axis] + (t - d0) / (d1 - d0)
                        expect.append(p * scale + np.float32(0))
    assert np.array_equal(np.array(expect, np.float32), v)


def test_bbox_transform_including_reference_y_term():
    # marching_cubes.cu:294-296: y scale uses upper[2], not upper[1]
    g = inputs.noise((8, 8, 8), seed=1)  # cubic: default bounds are then the identity transform
    v0, f0 = mc.marching_cubes(g, 0.0)
    lower, upper = [-1.0, -2.0, -3.0], [1.0, 5.0, 3.0]
    v1, f1 = mc.marching_cubes(g, 0.0, lower, upper)
    scale = np.array([(1.0 - -1.0) / 8, (3.0 - -2.0) / 8, (3.0 - -3.0) / 8], np.float32)
    assert np.array_equal(v1, (v0 * scale).astype(np.float32) + np.array(lower, np.float32))
    assert np.array_equal(f0, f1)


def test_ties_nan_and_inf_are_outside_or_handled():
    g = inputs.ties((8, 8, 8), seed=0)
    V, F = mc.count(g, 0.0)
    assert V > 0 and F > 0
    g2 = g.copy()
    g2[g2 == 0] = -1.0  # ==thresh is "outside" (marching_cubes.cu:25): same topology as a negative
    assert mc.count(g2, 0.0) == (V, F)
    g3 = g.copy()
    g3[g3 == 0] = np.nan
    assert mc.count(g3, 0.0) == (V, F)


def test_threads_do_not_change_output():
    g = inputs.noise((12, 11, 10), seed=5)
    a = mc.marching_cubes(g, 0.0, threads=1)
    b = mc.marching_cubes(g, 0.0, threads=4)
    assert np.array_equal(a[0], b[0]) and np.array_equal(a[1], b[1])


def test_empty_and_minimum_grids():
    v, f = mc.marching_cubes(np.zeros((2, 2, 2), np.float32), 0.5)
    assert v.shape == (0, 3) and f.shape == (0, 3)
    g = np.zeros((2, 2, 2), np.float32)
    g[0, 0, 0] = 1.0
    v, f = mc.marching_cubes(g, 0.5)
    assert v.shape == (3, 3) and f.shape == (1, 3)


def test_pymcubes_compat_counts_match_on_reference_examples():
    """examples/sphere.py:27-28 and bunny_sdf.py:28-29 assert exactly this."""
    sys.path.insert(0, os.path.join(os.path.dirname(HERE), "oracle", "pymcubes_compat"))
    import mcubes
    for grid in (inputs.sphere_int64(200), bunny()):
        v, f = mcubes.marching_cubes(grid, 0)
        V, F = mc.count(grid.astype(np.float32), 0.0)
        assert (v.shape[0], f.shape[0]) == (V, F)
        assert v.dtype == np.float64 and f.dtype == np.uint64
