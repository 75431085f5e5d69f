"""In-tree build of libipddp_b200.so for sm_100a (nvcc cross-compiles without a GPU).

    python interiorpointddp.jl_b200/build.py [--force] [--verbose]

-fmad=false: no implicit FMA contraction.  FMAs are written explicitly (IPDDP_FMA) where the algorithm
wants them, so results do not depend on the compiler's contraction choices (bit parity with the test
oracle).
"""
from __future__ import annotations

import concurrent.futures as cf
import glob
import hashlib
import os
import subprocess
import sys

HERE = os.path.dirname(os.path.abspath(__file__))
CSRC = os.path.join(HERE, "csrc")
OBJ = os.path.join(HERE, "_build")
LIB = os.path.join(HERE, "libipddp_b200.so")
NVCC = os.environ.get("NVCC", "/usr/local/cuda/bin/nvcc")
FLAGS = ["-gencode", "arch=compute_100a,code=sm_100a", "-lineinfo", "-O3", "-std=c++17", "-fmad=false",
         "-Xcompiler", "-fPIC", "-Xcompiler", "-fno-gnu-unique", "-ccbin", "/usr/bin/g++"]


def sources():
    return [os.path.join(CSRC, "ipddp_api.cu")] + sorted(glob.glob(os.path.join(CSRC, "models", "*.cu")))


def _headers():
    return sorted(glob.glob(os.path.join(CSRC, "*.cuh")) + glob.glob(os.path.join(CSRC, "*.h")) +
                  glob.glob(os.path.join(CSRC, "models_gen", "*.cuh")) + [os.path.join(HERE, "..", "include", "ipddp_b200.h")])


def content_hash(files, extra=""):
    """sha256 over file contents (not mtimes: checkouts and container moves reset those)."""
    h = hashlib.sha256(extra.encode())
    for f in files:
        h.update(os.path.basename(f).encode())
        with open(f, "rb") as fh:
            h.update(fh.read())
    return h.hexdigest()


def _compile(src, verbose):
    obj = os.path.join(OBJ, os.path.basename(src).replace(".cu", ".o"))
    cmd = [NVCC] + FLAGS + (["-Xptxas", "-v"] if verbose else []) + ["-c", src, "-o", obj]
    r = subprocess.run(cmd, capture_output=True, text=True)
    return src, obj, r.returncode, r.stdout + r.stderr


def build_variant(name: str, extra_flags, verbose: bool = False) -> str:
    """A/B experiments: libipddp_b200_<name>.so compiled with extra nvcc flags (e.g. -DIPDDP_FW_MINBLOCKS=3).
    Never loaded by the package; pass its path to tools/phase_bench.py."""
    objdir = os.path.join(OBJ, name)
    os.makedirs(objdir, exist_ok=True)
    objs = []
    for s in sources():
        o = os.path.join(objdir, os.path.basename(s).replace(".cu", ".o"))
        cmd = [NVCC] + FLAGS + list(extra_flags) + (["-Xptxas", "-v"] if verbose else []) + ["-c", s, "-o", o]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if verbose or r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
        if r.returncode != 0:
            raise RuntimeError(f"nvcc failed on {s}")
        objs.append(o)
    lib = os.path.join(HERE, f"libipddp_b200_{name}.so")
    subprocess.check_call([NVCC, "-shared", "-o", lib] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin",
                                                                 "/usr/bin/g++", "-Xcompiler", "-fPIC", "-ldl"])
    return lib


def build(force: bool = False, verbose: bool = False) -> str:
    os.makedirs(OBJ, exist_ok=True)
    srcs = sources()
    hdrs = _headers()
    todo, stamps = [], {}
    for s in srcs:
        o = os.path.join(OBJ, os.path.basename(s).replace(".cu", ".o"))
        stamps[s] = content_hash([s] + hdrs, " ".join(FLAGS))
        old = open(o + ".hash").read() if os.path.exists(o + ".hash") else ""
        if force or not os.path.exists(o) or old != stamps[s]:
            todo.append(s)
    if todo:
        with cf.ThreadPoolExecutor(max_workers=min(8, len(todo))) as ex:
            for src, obj, rc, out in ex.map(lambda s: _compile(s, verbose), todo):
                if verbose or rc != 0:
                    sys.stderr.write(out)
                if rc != 0:
                    raise RuntimeError(f"nvcc failed on {src}")
                with open(obj + ".hash", "w") as fh:
                    fh.write(stamps[src])
    objs = [os.path.join(OBJ, os.path.basename(s).replace(".cu", ".o")) for s in srcs]
    if todo or not os.path.exists(LIB):
        cmd = [NVCC, "-shared", "-o", LIB] + objs + ["-gencode", "arch=compute_100a,code=sm_100a", "-ccbin", "/usr/bin/g++",
                                                     "-Xcompiler", "-fPIC", "-ldl"]
        r = subprocess.run(cmd, capture_output=True, text=True)
        if r.returncode != 0:
            sys.stderr.write(r.stdout + r.stderr)
            raise RuntimeError("link failed")
    return LIB


if __name__ == "__main__":
    print(build(force="--force" in sys.argv, verbose="--verbose" in sys.argv))
