"""Synthetic workloads named by BASELINE.json / SURVEY.md Appendix B (numpy, CPU).

Deterministic INPUT generators only (no extraction code): bench.py, the profiling tools and the tests all
build their grids here, so the CUDA path, the oracle and the reference see bit-identical inputs.  No RNG
unless a seed is given.  `oracle/inputs.py` re-exports this module for the tests.
"""
import itertools

import numpy as np


def gyroid_tables(n, periods=8):
    """The two length-n fp32 tables the gyroid is built from (Appendix B)."""
    t = (2.0 * np.pi * periods / n) * (np.arange(n, dtype=np.float64) + 0.5)
    return np.sin(t).astype(np.float32), np.cos(t).astype(np.float32)


def gyroid(n, periods=8, x0=0, x1=None):
    """Gyroid(N, P) fp32 grid, planes [x0, x1) of dim 0; separately rounded fp32 ops:
    g = s[i]*c[j]; g = g + s[j]*c[k]; g = g + s[k]*c[i]."""
    s, c = gyroid_tables(n, periods)
    x1 = n if x1 is None else x1
    si, ci = s[x0:x1, None, None], c[x0:x1, None, None]
    sj, cj = s[None, :, None], c[None, :, None]
    sk, ck = s[None, None, :], c[None, None, :]
    g = (si * cj).astype(np.float32)
    g = (g + (sj * ck).astype(np.float32)).astype(np.float32)
    g = (g + (sk * ci).astype(np.float32)).astype(np.float32)
    return np.ascontiguousarray(g)


def sphere_int64(n=200):
    """examples/sphere.py:8-9 -- int64 grid, centre 50, radius 25."""
    X, Y, Z = np.mgrid[:n, :n, :n]
    return (X - 50) ** 2 + (Y - 50) ** 2 + (Z - 50) ** 2 - 25 ** 2


def upsample_trilinear(grid, n):
    """`grid` resampled to n^3 by separable linear interpolation with corner alignment (sample i of the
    output sits at i * (n_in - 1) / (n - 1) of the input), every operation a separately rounded fp32
    one in a fixed order: reproducible bit for bit wherever numpy runs.  BASELINE's "bunny SDF at 256^3"
    is upsample_trilinear(bunny66, 256) (the mesh the reference sampled its grid from is not available;
    SURVEY.md section 8d config 2)."""
    g = np.ascontiguousarray(grid, dtype=np.float32)
    for axis in range(3):
        m = g.shape[axis]
        pos = (np.arange(n, dtype=np.float32) * np.float32(m - 1)) / np.float32(n - 1)
        i0 = np.minimum(np.floor(pos).astype(np.int64), m - 2)
        w = (pos - i0.astype(np.float32)).astype(np.float32)
        shape = [1, 1, 1]
        shape[axis] = n
        w = w.reshape(shape)
        a, b = np.take(g, i0, axis=axis), np.take(g, i0 + 1, axis=axis)
        g = ((a * (np.float32(1.0) - w)).astype(np.float32) + (b * w).astype(np.float32)).astype(np.float32)
    return np.ascontiguousarray(g)


def noise(shape, seed):
    rng = np.random.default_rng(seed)
    return rng.uniform(-1.0, 1.0, size=shape).astype(np.float32)


def waves(shape, periods=(1.5, 2.5, 3.5)):
    """Smooth non-cubic field (a sum of three sines, fp32): sheet-like surfaces at any shape, no RNG."""
    ax = [np.sin((2.0 * np.pi * p / n) * (np.arange(n, dtype=np.float64) + 0.25)).astype(np.float32)
          for n, p in zip(shape, periods)]
    g = (ax[0][:, None, None] + ax[1][None, :, None]).astype(np.float32)
    return np.ascontiguousarray((g + ax[2][None, None, :]).astype(np.float32))


def ties(shape, seed):
    """Small-integer-valued grid: many samples exactly equal to thresh=0."""
    rng = np.random.default_rng(seed)
    return rng.integers(-2, 3, size=shape).astype(np.float32)


def kuhn_tet_grid(n):
    """Kuhn 6-tet grid on an n^3 lattice over [-1,1]^3, sdf = |p| - 0.5 (Appendix B)."""
    lin = np.linspace(-1.0, 1.0, n, dtype=np.float32)
    pts = np.stack(np.meshgrid(lin, lin, lin, indexing="ij"), -1).reshape(-1, 3)
    idx = np.arange(n ** 3, dtype=np.int64).reshape(n, n, n)
    v0 = idx[:-1, :-1, :-1].reshape(-1)
    step = np.array([n * n, n, 1], dtype=np.int64)
    tets = []
    for perm in itertools.permutations(range(3)):
        a = v0 + step[perm[0]]
        b = a + step[perm[1]]
        c = b + step[perm[2]]
        tets.append(np.stack([v0, a, b, c], -1))
    tets = np.stack(tets, 1).reshape(-1, 4)
    sdf = (np.sqrt((pts ** 2).sum(-1, dtype=np.float32)) - np.float32(0.5)).astype(np.float32)
    return np.ascontiguousarray(pts), np.ascontiguousarray(tets), sdf
