"""TEST INFRASTRUCTURE ONLY: parity of an ARBITRARY model (a codegen ModelDef that is not one of the built-in workloads)
between the kernel sources (emulator plugin compiled from the emitted device header, tests/emu/emu_plugins.py) and the
oracle (a scratch copy of oracle/ built with the model's emitted oracle header added to its table).  Both headers come
from one trace of the same closures, exactly as for the built-in models.  Never imported by the package."""
import importlib.util
import os
import re
import shutil
import subprocess
import sys

import build_emu
import emu_plugins

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(os.path.dirname(HERE))


def scratch_oracle(tmpdir, models):
    """A private copy of the oracle whose model table also holds `models` (a list of (ModelDef, bundles)); returns the
    loaded binding module."""
    from ipddp_b200.codegen import generate
    d = os.path.join(str(tmpdir), "oracle_scratch")
    shutil.copytree(os.path.join(ROOT, "oracle"), d, ignore=shutil.ignore_patterns("*.so", "*.hash", "__pycache__"))
    p = os.path.join(d, "ipddp_oracle.cpp")
    s = open(p).read()
    for md, bundles in models:
        with open(os.path.join(d, "models_gen", md.name + ".h"), "w") as fh:
            fh.write(generate.emit_oracle(md, bundles))
        s = s.replace('#include "models_gen/ragged.h"', f'#include "models_gen/ragged.h"\n#include "models_gen/{md.name}.h"', 1)
        s, n = re.subn(r"(const OracleModel\* kModels\[\] = \{)", rf"\1&gen_{md.name}::model, ", s, count=1)
        assert n == 1
    open(p, "w").write(s)
    mk = os.path.join(d, "Makefile")       # -O1: a third of the compile time; without fast-math / contraction the optimisation
    text = open(mk).read().replace("-O3", "-O1")                 # level cannot change a floating-point result
    open(mk, "w").write(text)
    subprocess.check_call(["make", "-C", d, "-s", "-B"])
    spec = importlib.util.spec_from_file_location("oracle_scratch", os.path.join(d, "oracle.py"))
    mod = importlib.util.module_from_spec(spec)
    spec.loader.exec_module(mod)
    mod.build = lambda force=False: mod.LIB_PATH        # already built; LIB_PATH points into the scratch copy
    return mod


def emulator_with_models(models):
    """The emulator library with `models` registered through ipddp_model_load (plugins compiled from the emitted headers)."""
    from ipddp_b200 import _lib
    from ipddp_b200.codegen import generate
    import concurrent.futures as cf
    lib = _lib.Lib(build_emu.build())
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(models)))) as ex:      # g++ runs outside the GIL
        sos = list(ex.map(lambda m: emu_plugins.compile_plugin(m[0].name, generate.emit_device(m[0], m[1])), models))
    for so in sos:
        lib.check(lib.L.ipddp_model_load(so.encode()), "ipddp_model_load")
    return lib
