"""The synthetic inputs the tests feed to the oracle and to the CUDA path: re-exported from
primitive3d_b200/workloads.py (input generators only, shared with bench.py so that nothing on the measured path
imports this package)."""
from primitive3d_b200.workloads import *  # noqa: F401,F403
from primitive3d_b200.workloads import gyroid, gyroid_tables, kuhn_tet_grid, noise, sphere_int64, ties  # noqa: F401
