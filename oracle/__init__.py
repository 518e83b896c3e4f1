"""oracle/ -- CPU restatements of the reference's algorithms.  TEST INFRASTRUCTURE.

Only tests/, __graft_entry__.smoke() and bench.py's cpu_baseline / --impl reference
legs may import this package.  Nothing under primitive3d_b200/ or prim3d/ does.
"""
