#!/usr/bin/env python3
"""Build the REAL reference into oracle/_ref/ (git-ignored; travels to the GPU box).

TEST INFRASTRUCTURE.  Compiles /root/reference/src/prim3d/Utility/marching_cubes.cu where it
lies, unmodified, with nvcc for sm_100a (the reference's own CMake does not configure with
CMake 4 / C++17 / torch 2.11), plus oracle/ref_stub.cpp, into
oracle/_ref/libPrim3D_ref.so.  Also stages, unmodified, the files the GPU box needs to run
the reference side of the comparison because /root/reference does not exist there:
  _ref/ref_marching_tetrahedras.py   <- prim3d/utility/marching_tetrahedras.py
  _ref/examples/                     <- examples/*.py and examples/data/
Nothing under _ref/ is committed and nothing in the product imports it.
"""
import os
import shutil
import subprocess
import sys
import sysconfig
import time

HERE = os.path.dirname(os.path.abspath(__file__))
REF = os.environ.get("P3D_REFERENCE_DIR", "/root/reference")
OUT = os.path.join(HERE, "_ref")


def newer(target, sources):
    if not os.path.exists(target):
        return False
    t = os.path.getmtime(target)
    return all(os.path.getmtime(s) <= t for s in sources)


def main():
    if not os.path.isdir(os.path.join(REF, "src/prim3d")):
        print(f"[build_ref] {REF} not present: keeping whatever is already in {OUT}")
        return 0
    os.makedirs(OUT, exist_ok=True)

    # unmodified python reference + examples (for the GPU box)
    shutil.copyfile(os.path.join(REF, "prim3d/utility/marching_tetrahedras.py"),
                    os.path.join(OUT, "ref_marching_tetrahedras.py"))
    ex = os.path.join(OUT, "examples")
    if os.path.isdir(ex):
        shutil.rmtree(ex)
    shutil.copytree(os.path.join(REF, "examples"), ex)

    import torch
    from torch.utils import cpp_extension as ce
    cu = os.path.join(REF, "src/prim3d/Utility/marching_cubes.cu")
    stub = os.path.join(HERE, "ref_stub.cpp")
    so = os.path.join(OUT, "libPrim3D_ref.so")
    if newer(so, [cu, stub, __file__]):
        print("[build_ref] up to date")
        return 0
    inc = [f"-I{p}" for p in ce.include_paths("cuda")] + [f"-I{sysconfig.get_paths()['include']}",
                                                           f"-I{os.path.join(REF, 'src/prim3d')}"]
    common = ["-DTORCH_EXTENSION_NAME=libPrim3D_ref", "-DTORCH_API_INCLUDE_EXTENSION_H",
              f"-D_GLIBCXX_USE_CXX11_ABI={int(torch._C._GLIBCXX_USE_CXX11_ABI)}", "-O2", "-std=c++17"]
    t0 = time.time()
    o_cu, o_stub = os.path.join(OUT, "marching_cubes.o"), os.path.join(OUT, "ref_stub.o")
    jobs = [
        subprocess.Popen(["nvcc", "-c", cu, "-o", o_cu, "-gencode", "arch=compute_100a,code=sm_100a",
                          "--expt-relaxed-constexpr", "--expt-extended-lambda", "-Xcompiler", "-fPIC",
                          "-ccbin", "/usr/bin/g++"] + common + inc),
        subprocess.Popen(["/usr/bin/g++", "-c", stub, "-o", o_stub, "-fPIC"] + common + inc),
    ]
    if any(j.wait() for j in jobs):
        print("[build_ref] compile failed")
        return 1
    tl = os.path.join(os.path.dirname(torch.__file__), "lib")
    subprocess.check_call(["/usr/bin/g++", "-shared", "-o", so, o_cu, o_stub, f"-L{tl}", f"-Wl,-rpath,{tl}",
                           "-lc10", "-lc10_cuda", "-ltorch_cpu", "-ltorch_cuda", "-ltorch", "-ltorch_python",
                           "-L/usr/local/cuda/lib64", "-lcudart"])
    os.remove(o_cu)
    os.remove(o_stub)
    print(f"[build_ref] built {so} in {time.time() - t0:.0f}s")
    return 0


if __name__ == "__main__":
    sys.exit(main())
