"""examples/ holds the reference's example scripts and fixtures unchanged (CPU check, where the reference tree is
mounted); tests/test_examples.py runs them on the GPU box."""
import os

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "examples")
FILES = ["sphere.py", "bunny_sdf.py", "sphere_tetrahedra.py", "data/bunny.npy", "data/tetrahedra/points.npy",
         "data/tetrahedra/sdfs.npy", "data/tetrahedra/tetrahedras.npy"]


def test_examples_are_present():
    for rel in FILES:
        assert os.path.getsize(os.path.join(EXAMPLES, rel)) > 0, rel


def test_examples_are_the_reference_files():
    ref = os.environ.get("P3D_REFERENCE_DIR", "/root/reference")
    if not os.path.isdir(os.path.join(ref, "examples")):
        pytest.skip("reference tree not mounted here: nothing to compare with")
    import filecmp
    for rel in FILES:
        assert filecmp.cmp(os.path.join(EXAMPLES, rel), os.path.join(ref, "examples", rel), shallow=False), rel
