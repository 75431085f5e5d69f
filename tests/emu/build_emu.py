"""TEST INFRASTRUCTURE ONLY: builds tests/emu/libipddp_emu.so = the product's CUDA sources compiled with
g++ against the fiber-based SIMT emulator (cpu_simt.h).  Used by tests/test_emu_*.py to check kernel
logic bit-for-bit against the oracle in a container without a GPU.  Never loaded by the package."""
import glob
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "interiorpointddp.jl_b200", "csrc")
LIB = os.path.join(HERE, "libipddp_emu.so")
CXX = "/usr/bin/g++"
FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-mfma", "-fno-fast-math", "-x", "c++",
         "-include", os.path.join(HERE, "cpu_simt.h"), "-w"]


def build(force=False):
    srcs = [os.path.join(CSRC, "ipddp_api.cu")] + sorted(glob.glob(os.path.join(CSRC, "models", "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(CSRC, "models_gen", "*.cuh")) + glob.glob(os.path.join(HERE, "cpu_simt.*")) + \
        [os.path.join(ROOT, "include", "ipddp_b200.h")]
    if not force and os.path.exists(LIB) and os.path.getmtime(LIB) >= max(os.path.getmtime(d) for d in deps):
        return LIB
    objdir = os.path.join(HERE, "_build")
    os.makedirs(objdir, exist_ok=True)
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        procs.append(subprocess.Popen([CXX] + FLAGS + ["-c", s, "-o", o]))
    o = os.path.join(objdir, "cpu_simt.o")
    procs.append(subprocess.Popen([CXX, "-O2", "-std=c++17", "-fPIC", "-fopenmp", "-c",
                                   os.path.join(HERE, "cpu_simt.cpp"), "-o", o]))
    objs.append(o)
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("emulator build failed")
    subprocess.check_call([CXX, "-shared", "-fopenmp", "-o", LIB] + objs + ["-ldl", "-lm"])
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
