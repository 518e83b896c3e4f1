"""In-tree build of the native pieces (no JIT cache: the .so files travel with the tree).

  libprim3d_b200.so        primitive3d_b200/   nvcc, sm_100a, torch-free C ABI (include/prim3d_b200.h)
  libPrim3D.so             prim3d/             g++, pybind11 + torch headers, links the C ABI library

Run as `python -m primitive3d_b200.build` or through __graft_entry__.build().
"""
import os
import subprocess
import sys
import sysconfig
import time

PKG = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(PKG)
CSRC = os.path.join(PKG, "csrc")
CORE_SO = os.path.join(PKG, "libprim3d_b200.so")
BIND_SO = os.path.join(ROOT, "prim3d", "libPrim3D.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
CXX = "/usr/bin/g++" if os.path.exists("/usr/bin/g++") else "g++"

CORE_SOURCES = ["prim3d_b200.cu", "mc_kernels.cu", "mc_faces_rows.cu", "mc_small.cu", "mc_peer.cu", "mt_kernels.cu", "mt_extract.cu", "ply_kernels.cu", "mc_host_stream.cu"]
NVCC_FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "--threads", "4",
              "-Xcompiler", "-fPIC", "-ccbin", CXX]


def _stale(target, sources):
    if not os.path.exists(target):
        return True
    t = os.path.getmtime(target)
    return any(os.path.getmtime(s) > t for s in sources)


def _deps(names):
    hdrs = [os.path.join(CSRC, f) for f in os.listdir(CSRC) if f.endswith((".h", ".cuh", ".inc"))]
    return [os.path.join(CSRC, n) for n in names] + hdrs + [os.path.join(ROOT, "include", "prim3d_b200.h"),
                                                             os.path.abspath(__file__)]


def build_core(force=False, verbose=False):
    srcs = [os.path.join(CSRC, n) for n in CORE_SOURCES if os.path.exists(os.path.join(CSRC, n))]
    if not force and not _stale(CORE_SO, _deps(CORE_SOURCES[:0]) + srcs):
        return CORE_SO
    t0 = time.time()
    extra = os.environ.get("P3D_NVCC_EXTRA", "").split()  # e.g. -DP3D_STRIP_MINBLOCKS=2 for tuning runs
    cmd = [NVCC] + NVCC_FLAGS + extra + (["-Xptxas", "-v"] if verbose else []) + ["-shared", "-o", CORE_SO] + srcs
    subprocess.check_call(cmd)
    print(f"[build] {os.path.relpath(CORE_SO, ROOT)} in {time.time() - t0:.1f}s")
    return CORE_SO


def build_bindings(force=False):
    src = os.path.join(CSRC, "bindings.cpp")
    if not force and not _stale(BIND_SO, _deps(["bindings.cpp"])):
        return BIND_SO
    import torch
    from torch.utils import cpp_extension as ce
    t0 = time.time()
    tl = os.path.join(os.path.dirname(torch.__file__), "lib")
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}"]
    cmd = [CXX, "-shared", "-fPIC", "-O2", "-std=c++17", src, "-o", BIND_SO,
           "-DTORCH_EXTENSION_NAME=libPrim3D", "-DTORCH_API_INCLUDE_EXTENSION_H",
           f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}"] + inc + [
           f"-L{PKG}", "-lprim3d_b200", "-Wl,-rpath,$ORIGIN/../primitive3d_b200",
           f"-L{tl}", f"-Wl,-rpath,{tl}", "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch",
           "-ltorch_python", "-L/usr/local/cuda/lib64", "-lcudart"]
    subprocess.check_call(cmd)
    print(f"[build] {os.path.relpath(BIND_SO, ROOT)} in {time.time() - t0:.1f}s")
    return BIND_SO


def build_all(force=False, verbose=False):
    build_core(force, verbose)
    build_bindings(force)


if __name__ == "__main__":
    build_all(force="--force" in sys.argv, verbose="-v" in sys.argv)
