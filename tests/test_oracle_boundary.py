"""The oracle is test infrastructure: only tests/, __graft_entry__.smoke() and bench.py (its cpu_baseline leg, the
--impl reference arm and the bitwise check of that sample) may import, link or execute anything under oracle/.  The product
package, its C-ABI library and the tools/ scripts must not, and the package must fail loudly without its CUDA library."""
import ast
import glob
import os
import re
import subprocess

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "interiorpointddp.jl_b200")


def _imports(path):
    tree = ast.parse(open(path).read())
    names = set()
    for n in ast.walk(tree):
        if isinstance(n, ast.Import):
            names |= {a.name.split(".")[0] for a in n.names}
        elif isinstance(n, ast.ImportFrom) and n.module:
            names.add(n.module.split(".")[0])
    return names


def test_product_python_never_imports_the_oracle():
    files = glob.glob(os.path.join(PKG, "**", "*.py"), recursive=True) + glob.glob(os.path.join(ROOT, "tools", "*.py"))
    assert len(files) > 10
    for f in files:
        assert "oracle" not in _imports(f), f
        src = open(f).read()
        assert not re.search(r"libipddp_oracle|oracle/_ref|ctypes\.CDLL\([^)]*oracle", src), f


def test_product_sources_never_include_the_oracle():
    srcs = [f for pat in ("*.cu", "*.cuh", "*.h") for f in glob.glob(os.path.join(PKG, "csrc", "**", pat), recursive=True)]
    srcs += glob.glob(os.path.join(ROOT, "include", "*.h"))
    assert len(srcs) > 10
    for f in srcs:
        for line in open(f):
            if line.lstrip().startswith("#include"):
                assert "oracle" not in line, (f, line)


def test_c_abi_library_does_not_link_the_oracle():
    lib = os.path.join(PKG, "libipddp_b200.so")
    if not os.path.exists(lib):
        pytest.skip("library not built")
    out = subprocess.run(["readelf", "-d", lib], capture_output=True, text=True).stdout
    needed = re.findall(r"\(NEEDED\).*\[(.*)\]", out)
    assert needed and not any("oracle" in n for n in needed), needed


def test_bench_touches_the_oracle_only_in_its_baseline_and_reference_legs():
    src = open(os.path.join(ROOT, "bench.py")).read()
    tree = ast.parse(src)
    users = set()
    for fn in [n for n in ast.walk(tree) if isinstance(n, ast.FunctionDef)]:
        seg = ast.get_source_segment(src, fn) or ""
        if re.search(r"\boracle\b", seg):
            users.add(fn.name)
    top = [n for n in tree.body if isinstance(n, (ast.Import, ast.ImportFrom))]
    assert not any("oracle" in ast.get_source_segment(src, n) for n in top), "bench.py imports the oracle at module level"
    print(sorted(users))
    assert users, "expected the cpu_baseline leg to use the oracle"
    for u in users:
        assert re.search(r"cpu|baseline|reference|oracle|parity|config|main", u), u


def test_package_fails_loudly_without_its_library(tmp_path):
    code = ("import sys; sys.path.insert(0, %r)\n"
            "import ipddp_b200\nfrom ipddp_b200 import _lib\n"
            "try:\n    _lib.Lib(%r)\nexcept Exception as e:\n    print('RAISED', type(e).__name__)\n" % (ROOT, str(tmp_path / "missing.so")))
    r = subprocess.run(["python", "-c", code], capture_output=True, text=True)
    assert "RAISED" in r.stdout, r.stdout + r.stderr
