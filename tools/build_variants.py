#!/usr/bin/env python3
"""Build tuning variants of libprim3d_b200.so HERE (nvcc cross-compiles), so that the GPU box only runs them.

  python tools/build_variants.py name1="-DP3D_X=1 -DP3D_Y=2" name2="..."

Outputs build/variants/<name>.so (git-ignored, shipped by gpurun).  Time them on the box with
  for v in build/variants/*.so; do P3D_CORE_LIB=$PWD/$v python tools/prof_mc.py; done
"""
import os
import subprocess
import sys
from concurrent.futures import ThreadPoolExecutor

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)
from primitive3d_b200 import build as b  # noqa: E402


def one(item):
    name, flags = item
    out = os.path.join(ROOT, "build", "variants", name + ".so")
    srcs = [os.path.join(b.CSRC, n) for n in b.CORE_SOURCES]
    cmd = [b.NVCC] + b.NVCC_FLAGS + flags.split() + ["-shared", "-o", out] + srcs
    r = subprocess.run(cmd, capture_output=True, text=True)
    return name, r.returncode, r.stderr[-2000:]


def main():
    os.makedirs(os.path.join(ROOT, "build", "variants"), exist_ok=True)
    items = [a.split("=", 1) for a in sys.argv[1:]]
    with ThreadPoolExecutor(8) as ex:
        for name, rc, err in ex.map(one, items):
            print(name, "ok" if rc == 0 else "FAILED\n" + err)


if __name__ == "__main__":
    main()
