import os
import sys

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
if ROOT not in sys.path:
    sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))

GOLDEN = os.path.join(ROOT, "tests", "golden")


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (run on the B200 box with -m gpu)")


@pytest.fixture(scope="session")
def golden_dir():
    return GOLDEN


@pytest.fixture(scope="session", autouse=True)
def _built_checkers():
    """The oracle libraries are built by __graft_entry__.build(); build them on demand so a
    bare `pytest` in a fresh checkout works too (gcc only, a second or two)."""
    import subprocess
    if not os.path.exists(os.path.join(ROOT, "oracle", "_build", "libp3d_oracle.so")):
        subprocess.check_call(["make", "-C", os.path.join(ROOT, "oracle")])
