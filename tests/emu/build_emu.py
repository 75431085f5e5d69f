"""TEST INFRASTRUCTURE ONLY: builds tests/emu/libipddp_emu.so = the product's CUDA sources compiled with
g++ against the fiber-based SIMT emulator (cpu_simt.h).  Used by tests/test_emu_*.py to check kernel
logic bit-for-bit against the oracle in a container without a GPU.  Never loaded by the package."""
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))
CSRC = os.path.join(ROOT, "interiorpointddp.jl_b200", "csrc")
LIB = os.path.join(HERE, "libipddp_emu.so")
CXX = "/usr/bin/g++"
FLAGS = ["-O2", "-std=c++17", "-fPIC", "-fopenmp", "-ffp-contract=off", "-mfma", "-fno-fast-math", "-fno-gnu-unique", "-fvisibility-inlines-hidden", "-x", "c++",
         "-include", os.path.join(HERE, "cpu_simt.h"), "-w"]
FLAGS += os.environ.get("IPDDP_EMU_DEFS", "").split()     # e.g. "-DIPDDP_SWAP_LOOP=1" to check an experiment macro
LDFLAGS = os.environ.get("IPDDP_EMU_LDFLAGS", "").split()  # e.g. "-fsanitize=address" (with the same in IPDDP_EMU_DEFS)
if os.environ.get("IPDDP_EMU_LIB"):                        # build a differently-flagged library next to the default one
    LIB = os.path.join(HERE, os.environ["IPDDP_EMU_LIB"])


def build(force=False):
    srcs = [os.path.join(CSRC, "ipddp_api.cu")] + sorted(glob.glob(os.path.join(CSRC, "models", "*.cu")))
    deps = srcs + glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) + \
        glob.glob(os.path.join(CSRC, "models_gen", "*.cuh")) + glob.glob(os.path.join(HERE, "cpu_simt.*")) + \
        [os.path.join(ROOT, "include", "ipddp_b200.h")]
    objdir = os.path.join(HERE, "_build" + ("_" + os.path.basename(LIB) if os.environ.get("IPDDP_EMU_LIB") else ""))
    os.makedirs(objdir, exist_ok=True)
    h = hashlib.sha256(" ".join(FLAGS + LDFLAGS).encode())
    for d in sorted(deps):
        h.update(os.path.basename(d).encode())
        with open(d, "rb") as fh:
            h.update(fh.read())
    stamp, stamp_file = h.hexdigest(), os.path.join(objdir, "lib.hash")
    if not force and os.path.exists(LIB) and os.path.exists(stamp_file) and open(stamp_file).read() == stamp:
        return LIB
    procs = []
    objs = []
    for s in srcs:
        o = os.path.join(objdir, os.path.basename(s) + ".o")
        objs.append(o)
        procs.append(subprocess.Popen([CXX] + FLAGS + ["-c", s, "-o", o]))
    o = os.path.join(objdir, "cpu_simt.o")
    procs.append(subprocess.Popen([CXX, "-O2", "-std=c++17", "-fPIC", "-fopenmp"] + LDFLAGS + ["-c",
                                   os.path.join(HERE, "cpu_simt.cpp"), "-o", o]))
    objs.append(o)
    for p in procs:
        if p.wait() != 0:
            raise RuntimeError("emulator build failed")
    subprocess.check_call([CXX, "-shared", "-fopenmp"] + LDFLAGS + ["-o", LIB] + objs + ["-ldl", "-lm"])
    with open(stamp_file, "w") as fh:
        fh.write(stamp)
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv))
