# Same value as the reference's prim3d/version.py:1 (part of the public surface).
__version__ = "0.0.1"
