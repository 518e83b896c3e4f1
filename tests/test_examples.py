"""The reference's own example scripts, byte-unchanged (examples/ in this tree: the acceptance
fixtures BASELINE.json names), must run against this repository's `prim3d`.
Their asserts are the only tests the reference has (examples/sphere.py:27-30,
examples/bunny_sdf.py:28-31)."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
EXAMPLES = os.path.join(ROOT, "examples")


@pytest.mark.parametrize("script", ["sphere.py", "bunny_sdf.py", "sphere_tetrahedra.py"])
def test_reference_example_runs_unchanged(script, tmp_path):
    path = os.path.join(EXAMPLES, script)
    assert os.path.exists(path)
    env = dict(os.environ)
    # `mcubes` (PyMCubes) is a third-party dependency of the examples that is not installable
    # here; the PyMCubes-compatible stand-in is put on the path for them.
    env["PYTHONPATH"] = os.pathsep.join([ROOT, os.path.join(ROOT, "oracle", "pymcubes_compat"),
                                         env.get("PYTHONPATH", "")])
    out = subprocess.run([sys.executable, path], cwd=tmp_path, env=env, capture_output=True, text=True, timeout=600)
    assert out.returncode == 0, out.stdout + out.stderr
    if script != "sphere_tetrahedra.py":
        assert "#vertices=" in out.stdout and "#triangles=" in out.stdout
    produced = {"sphere.py": "sphere.ply", "bunny_sdf.py": "bunny.ply", "sphere_tetrahedra.py": "sphere_tetrahedra.ply"}
    assert os.path.getsize(os.path.join(tmp_path, produced[script])) > 1000

