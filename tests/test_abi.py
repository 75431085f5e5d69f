"""The C-ABI library builds for sm_100a, loads, and exports every symbol include/ipddp_b200.h declares.
No compute calls (no GPU in the build container)."""
import ctypes as C
import os
import re

import pytest

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


@pytest.fixture(scope="module")
def product_lib():
    import ipddp_b200  # noqa: F401
    from ipddp_b200 import build
    path = build.build()
    return C.CDLL(path)


def test_header_symbols_exported(product_lib):
    hdr = open(os.path.join(ROOT, "include", "ipddp_b200.h")).read()
    names = sorted(set(re.findall(r"\b(ipddp_[a-z0-9_]+)\s*\(", hdr)))
    assert len(names) >= 30
    missing = [n for n in names if not hasattr(product_lib, n)]
    assert not missing, f"declared but not exported: {missing}"


def test_binding_list_matches_header():
    import ipddp_b200  # noqa: F401
    from ipddp_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "ipddp_b200.h")).read()
    names = set(re.findall(r"\b(ipddp_[a-z0-9_]+)\s*\(", hdr))
    assert names == set(_lib.EXPORTS)


def test_options_layout_and_defaults(product_lib):
    """ipddp_options mirrors reference src/options.jl:1-38 (31 fields, defaults)."""
    from ipddp_b200 import _lib
    lib = _lib.Lib(_lib.LIB_PATH)
    o = lib.default_options()
    assert len(_lib.Options._fields_) == 31
    assert (o.quasi_newton, o.max_iterations, o.reset_cache, o.verbose, o.print_frequency) == (0, 1000, 1, 0, 10)
    ref = dict(optimality_tolerance=1e-8, mu_init=1.0, ineq_dual_init=1.0, kappa_1=0.01, kappa_2=0.01, reg_1=1e-4,
               reg_min=1e-20, reg_max=1e40, kappa_bar_w_p=100.0, kappa_w_p=8.0, kappa_w_m=1.0 / 3.0, kappa_c=0.25,
               delta_c=1e-8, kappa_eps=10.0, kappa_mu=0.2, theta_mu=1.2, tau_min=0.99, s_max=100.0, eta_L=1e-4, s_L=2.3,
               delta=1.0, s_theta=1.1, gamma_alpha=0.05, gamma_theta=1e-5, gamma_L=1e-5, kappa_Sigma=1e10)
    for k, v in ref.items():
        assert getattr(o, k) == v, k
    assert product_lib.ipddp_abi_version() == 2


def test_model_registry(product_lib):
    from ipddp_b200 import _lib
    lib = _lib.Lib(_lib.LIB_PATH)
    assert set(lib.models()) == {"cartpole", "acrobot", "concar", "concar_quad", "pushing", "double_integrator"}
    assert lib.model_dims("cartpole")[:4] == (4, 21, 14, 5)
    assert lib.model_dims("acrobot")[:4] == (4, 9, 6, 8)
    assert lib.model_dims("concar")[:4] == (4, 10, 4, 14)
    assert lib.model_dims("pushing")[:4] == (4, 11, 6, 8)
    assert lib.model_dims("double_integrator")[:4] == (2, 3, 1, 0)


def test_sass_is_sm100a(product_lib):
    """the shipped library carries sm_100a SASS with FP64 FMA in the backward kernel"""
    import subprocess
    from ipddp_b200 import _lib
    out = subprocess.run(["/usr/local/cuda/bin/cuobjdump", "-lelf", _lib.LIB_PATH], capture_output=True, text=True).stdout
    assert "sm_100a" in out


def test_julia_stub_struct_layouts_match_the_header():
    """julia/InteriorPointDDPB200.jl cannot be executed here; at least its Options / Stats mirrors must have the C
    structs' field count and type sequence (they are passed by reference through ccall), and every function it binds
    must be declared in include/ipddp_b200.h."""
    import ctypes as C
    import re
    from ipddp_b200 import _lib
    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    jl = open(os.path.join(root, "julia", "InteriorPointDDPB200.jl"), encoding="utf-8").read()

    def fields(name):
        body = re.search(r"mutable struct %s\n(.*?)\nend" % name, jl, re.S).group(1)
        return re.findall(r"::(Cint|Cdouble|Clonglong)\b", body)

    ctype = {C.c_int: "Cint", C.c_double: "Cdouble", C.c_longlong: "Clonglong"}
    assert fields("Options") == [ctype[t] for _, t in _lib.Options._fields_]
    assert fields("Stats") == [ctype[t] for _, t in _lib.Stats._fields_]
    header = open(os.path.join(root, "include", "ipddp_b200.h")).read()
    bound = set(re.findall(r"\(:(ipddp_[a-z_0-9]+), LIB\)", jl))
    assert bound and all(re.search(r"\b%s\(" % f, header) for f in bound), sorted(bound)
