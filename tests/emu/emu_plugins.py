"""TEST INFRASTRUCTURE ONLY: lets the reference-facing host API (api.Solver: trace -> emit -> compile -> ipddp_model_load
-> solve) run in a container without a GPU.  The emitted model header -- the very text the product hands to nvcc -- is
compiled with g++ against the SIMT emulator (cpu_simt.h) into a plugin for tests/emu/libipddp_emu.so, and `_lib.load()` is
pointed at that library.  Never imported by the package."""
import hashlib
import os
import subprocess

import build_emu

HERE = os.path.dirname(os.path.abspath(__file__))
PLUGINS = os.path.join(HERE, "_build_plugins")


def compile_plugin(name, src, force=False):
    from ipddp_b200 import api
    os.makedirs(PLUGINS, exist_ok=True)
    csrc = os.path.join(api.HERE, "csrc")
    src = src.replace('#include "../model_common.cuh"', f'#include "{os.path.join(csrc, "model_common.cuh")}"')
    tag = hashlib.sha256((src + " ".join(build_emu.FLAGS)).encode()).hexdigest()[:12]
    so = os.path.join(PLUGINS, f"{name}_{tag}.so")
    cuh, cu = so[:-3] + ".cuh", so[:-3] + ".cu"
    with open(cuh, "w") as fh:
        fh.write(src)
    with open(cu, "w") as fh:
        fh.write(f'#include "{cuh}"\n#include "{os.path.join(csrc, "model_register.cuh")}"\n'
                 f'IPDDP_REGISTER_MODEL(Model_{name}, ipddp_plugin_vtable)\n')
    # always rebuilt: the kernel templates the plugin embeds change with the sources under test.  The plugin is linked
    # against the emulator library for the SIMT runtime (emu::launch ...): resolved through its own dependency list, so
    # the library is never loaded RTLD_GLOBAL (that would interpose the kernels of the differently-flagged emulator builds
    # other tests load into the same process).
    emu_dir, emu_name = os.path.split(build_emu.LIB)
    subprocess.check_call([build_emu.CXX] + build_emu.FLAGS + ["-shared", cu, "-o", so, f"-L{emu_dir}", f"-l:{emu_name}",
                                                                f"-Wl,-rpath,{emu_dir}"])
    return so


def install(monkeypatch):
    """Points the package at the emulator for the duration of a test; returns the emulator's Lib."""
    from ipddp_b200 import _lib, api
    path = build_emu.build()
    lib = _lib.Lib(path)
    monkeypatch.setattr(_lib, "load", lambda *a, **k: lib)
    monkeypatch.setattr(api, "_compile_plugin", compile_plugin)
    monkeypatch.setattr(api, "_loaded_models", {})
    return lib
