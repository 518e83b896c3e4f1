#!/bin/bash
# usage: tools/variants.sh "<nvcc extra flags A>" "<nvcc extra flags B>" ...   (run on the GPU box)
# Rebuilds the core library with each flag set and prints the per-kernel times at 1024^3.
for v in "$@"; do
  P3D_NVCC_EXTRA="$v" python -c "from primitive3d_b200 import build; build.build_core(force=True)" > /dev/null 2>&1
  echo "variant [$v]: $(timeout 120 python tools/prof_mc.py --size 1024 2>&1 | tail -1)"
done
