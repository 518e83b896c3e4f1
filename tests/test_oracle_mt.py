"""The marching-tetrahedra oracle against outputs of the REAL reference
(prim3d/utility/marching_tetrahedras.py run on CPU, stored by tests/golden/make_golden.py)."""
import hashlib
import os

import numpy as np
import pytest

from oracle import inputs, mt

HERE = os.path.dirname(os.path.abspath(__file__))
CASES = ["kat", "fixture", "random0", "random1", "kuhn8", "kuhn8_noise"]


def load(name):
    return np.load(os.path.join(HERE, "golden", f"mt_{name}.npz"))


@pytest.mark.parametrize("name", CASES)
def test_bit_exact_against_reference_outputs(name):
    g = load(name)
    tets = g["tets"].copy()
    v, f, ti = mt.marching_tetrahedras(g["points"], tets, g["sdf"], return_tet_idx=True)
    assert np.array_equal(tets, g["tets_after"])          # in-place flip, :148
    assert v.dtype == np.float32 and f.dtype == np.int64 and ti.dtype == np.int64
    assert np.array_equal(v.view(np.uint32), g["verts"].view(np.uint32))
    assert np.array_equal(f, g["faces"])
    assert np.array_equal(ti, g["tet_idx"])


def test_docstring_known_answer():
    # marching_tetrahedras.py:119-136
    g = load("kat")
    v, f = mt.marching_tetrahedras(g["points"], g["tets"].copy(), g["sdf"])
    assert np.allclose(v, [[0, 2 / 3, 0], [0, 0, 2 / 3], [1 / 3, 2 / 3, 0], [1 / 3, 0, 2 / 3]], atol=1e-6)
    assert f.tolist() == [[3, 0, 1], [3, 2, 0]]


def test_kuhn32_digests():
    d = np.load(os.path.join(HERE, "golden", "mt_digests.npz"))
    pts, tets, sdf = inputs.kuhn_tet_grid(32)
    v, f, ti = mt.marching_tetrahedras(pts, tets, sdf, True)
    sha = lambda a: hashlib.sha256(np.ascontiguousarray(a).tobytes()).hexdigest()
    assert (len(v), len(f)) == (int(d["kuhn32.V"]), int(d["kuhn32.F"])) == (3314, 6624)
    assert sha(v) == str(d["kuhn32.verts"]) and sha(f) == str(d["kuhn32.faces"])
    assert sha(ti) == str(d["kuhn32.tet_idx"]) and sha(tets) == str(d["kuhn32.tets_after"])


def test_no_valid_tets():
    pts = np.array([[0, 0, 0], [1, 0, 0], [0, 1, 0], [0, 0, 1]], np.float32)
    v, f, ti = mt.marching_tetrahedras(pts, np.array([[0, 1, 2, 3]], np.int64), np.ones(4, np.float32), True)
    assert v.shape == (0, 3) and f.shape == (0, 3) and ti.shape == (0,)
