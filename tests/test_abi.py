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
    assert set(lib.models()) == {"cartpole", "acrobot", "concar", "concar_quad", "pushing", "double_integrator", "ragged"}
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


JL = os.path.join(ROOT, "julia", "InteriorPointDDPB200.jl")


def _c_prototypes():
    """name -> (return type, [argument types]) of every function include/ipddp_b200.h declares (comments stripped)."""
    hdr = open(os.path.join(ROOT, "include", "ipddp_b200.h")).read()
    hdr = re.sub(r"/\*.*?\*/", " ", hdr, flags=re.S)
    protos = {}
    for m in re.finditer(r"([A-Za-z_][A-Za-z_0-9 \*]*?)\b(ipddp_[a-z0-9_]+)\s*\(([^)]*)\)\s*;", hdr):
        ret, name, args = m.group(1).strip(), m.group(2), m.group(3).strip()
        if ret.startswith("typedef") or ret == "":
            continue
        alist = []
        if args and args != "void":
            for a in args.split(","):
                a = a.strip()
                t = re.sub(r"\b[A-Za-z_][A-Za-z_0-9]*$", "", a).strip() if not a.endswith("*") else a   # drop the parameter name
                alist.append(re.sub(r"\s+", " ", t).replace(" *", "*"))
        protos[name] = (re.sub(r"\s+", " ", ret).replace(" *", "*"), alist)
    return protos


# what a Julia ccall may declare for a C parameter type
_JL_OK = {
    "int": {"Cint"}, "double": {"Cdouble"}, "long long": {"Clonglong"},
    "const char*": {"Cstring"}, "void*": {"Ptr{Cvoid}"},
    "int*": {"Ptr{Cint}", "Ref{Cint}"}, "const int*": {"Ptr{Cint}", "Ref{Cint}"},
    "double*": {"Ptr{Cdouble}", "Ref{Cdouble}"}, "const double*": {"Ptr{Cdouble}", "Ref{Cdouble}"},
    "long long*": {"Ptr{Clonglong}", "Ref{Clonglong}"},
    "ipddp_problem*": {"Ptr{Cvoid}"}, "ipddp_problem**": {"Ref{Ptr{Cvoid}}", "Ptr{Ptr{Cvoid}}"},
    "const ipddp_options*": {"Ref{COptions}"}, "ipddp_options*": {"Ref{COptions}"},
    "ipddp_stats*": {"Ref{Stats}"}, "const ipddp_queue*": {"Ref{Queue}"},
}
_JL_RET = {"int": "Cint", "double": "Cdouble", "long long": "Clonglong", "const char*": "Cstring", "void": "Cvoid",
           "void*": "Ptr{Cvoid}"}


def _split_types(tup: str):
    out, depth, cur = [], 0, ""
    for ch in tup:
        if ch == "{":
            depth += 1
        if ch == "}":
            depth -= 1
        if ch == "," and depth == 0:
            out.append(cur.strip()); cur = ""
        else:
            cur += ch
    if cur.strip():
        out.append(cur.strip())
    return out


def test_julia_ccall_signatures_match_the_header():
    """julia/InteriorPointDDPB200.jl cannot be executed here (no Julia in the image).  Checked instead, for EVERY ccall
    in the file: the function is declared in include/ipddp_b200.h, the declared return type and each argument type are
    what the C prototype requires (count, order, pointer-ness, struct mirrors by name)."""
    jl = open(JL, encoding="utf-8").read()
    protos = _c_prototypes()
    sigvars = dict(re.findall(r"\n\s*(sig)\s*=\s*\(([^\n]*)\)\n", jl))
    calls = re.findall(r"ccall\(\(:(ipddp_[a-z_0-9]+), LIB\),\s*([A-Za-z{}]+),\s*(\((?:[^()]|\([^()]*\))*\)|sig)\s*[,)]", jl, flags=re.S)
    assert len(calls) >= 18, len(calls)
    seen = set()
    for name, ret, tup in calls:
        assert name in protos, f"{name} is not declared in the header"
        cret, cargs = protos[name]
        assert ret == _JL_RET[cret], f"{name}: return {ret} vs C {cret}"
        body = sigvars["sig"] if tup == "sig" else tup[1:-1]
        jargs = _split_types(body.replace("\n", " "))
        assert len(jargs) == len(cargs), f"{name}: {len(jargs)} ccall arguments vs {len(cargs)} in the prototype"
        for k, (ja, ca) in enumerate(zip(jargs, cargs)):
            assert ja in _JL_OK[ca], f"{name} argument {k + 1}: Julia {ja} vs C `{ca}`"
        seen.add(name)
    # the binding covers the calls the reference-facing layer needs
    need = {"ipddp_problem_create", "ipddp_problem_destroy", "ipddp_set_inputs", "ipddp_solve", "ipddp_solve_queue",
            "ipddp_get_results", "ipddp_get_trajectory", "ipddp_get_stats", "ipddp_model_load", "ipddp_model_dims",
            "ipddp_last_error", "ipddp_set_tuning", "ipddp_set_stream", "ipddp_solve_many", "ipddp_get_duals", "ipddp_get_trace"}
    assert need <= seen, sorted(need - seen)


def test_julia_struct_mirrors_match_the_c_structs():
    """The structs passed by reference through ccall (COptions, Stats, Queue) have the C structs' field count and type
    sequence; Options{T} keeps the reference's 31 field names (src/options.jl:1-38) in the order of the C struct."""
    from ipddp_b200 import _lib
    jl = open(JL, encoding="utf-8").read()

    def fields(name):
        body = re.search(r"mutable struct %s\n(.*?)\nend" % re.escape(name), jl, re.S).group(1)
        return re.findall(r"^\s*(\S+?)::(\S+)", body, re.M)

    ctype = {C.c_int: "Cint", C.c_double: "Cdouble", C.c_longlong: "Clonglong"}
    assert [t for _, t in fields("COptions")] == [ctype[t] for _, t in _lib.Options._fields_]
    assert [t for _, t in fields("Stats")] == [ctype[t] for _, t in _lib.Stats._fields_]
    q = fields("Queue")
    assert [n for n, _ in q] == [n for n, _ in _lib.Queue._fields_]
    for (n, jt), (_, ct) in zip(q, _lib.Queue._fields_):
        assert (jt == "Cint") if ct is C.c_int else jt.startswith("Ptr{"), (n, jt)
    # Options{T}: the reference's names, same order as the C mirror
    ref_names = re.findall(r"^\s*([^\s:#=]+)(?:::\S+)?\s*=", re.search(r"mutable struct Options\{T\}\n(.*?)\nend", jl, re.S).group(1), re.M)
    assert ref_names == [n for n, _ in fields("COptions")] and len(ref_names) == 31
    upstream = "/root/reference/src/options.jl"
    if os.path.exists(upstream):     # only in the build container
        up = re.findall(r"^\s*([^\s:#=]+)(?:::\S+)?\s*=", open(upstream, encoding="utf-8").read(), re.M)
        assert up == ref_names


def test_julia_emitter_writes_the_names_the_kernels_read():
    """julia/codegen.jl emits the same `Model_<name>` header layout as codegen/generate.py: every member the CUDA kernel
    templates read from the model struct (M::...) is written by the Julia emitter too."""
    cj = open(os.path.join(ROOT, "julia", "codegen.jl"), encoding="utf-8").read()
    used = set()
    csrc = os.path.join(ROOT, "interiorpointddp.jl_b200", "csrc")
    for f in os.listdir(csrc):
        if f.endswith(".cuh") and f != "ldlt_warp.cuh":     # (there S is the LDLT scratch layout, not a stage type)
            # M = the (chain) model, S = a stage type, T = the terminal stage type
            used |= set(re.findall(r"\b[MST]::([A-Za-z_][A-Za-z_0-9]*)", open(os.path.join(csrc, f)).read()))
    assert len(used) > 30
    used.discard("template")          # `M::template Stage<I>`: a keyword, not a member
    for name in sorted(used):
        if re.fullmatch(r"(D|VF|DN)_[a-z]+_(OFF|N)", name):
            prefix, mat, _ = name.split("_")
            assert f'"{mat}"' in cj or f'"N{mat}"' in cj, name        # the matrix is one of the emitter's groups
            assert "$(prefix)_$(shown)_OFF" in cj and "$(prefix)_$(shown)_N" in cj
        else:
            assert re.search(r"\b%s\b" % name, cj), f"the Julia emitter never writes M::{name}"


def test_stage_chain_model_tables(product_lib):
    """The built-in stage chain `ragged` (state / control sizes that change along the horizon) is registered with its
    stage types; ipddp_model_dims reports the maxima that stride the arrays."""
    from ipddp_b200 import _lib
    lib = _lib.Lib(_lib.LIB_PATH)
    assert "ragged" in lib.models()
    nstage, stages, nxt = lib.model_stages("ragged")
    assert (nstage, nxt) == (3, 3) and stages == [(2, 3, 1, 2), (2, 3, 1, 3), (3, 2, 0, 3)]
    assert lib.model_dims("ragged")[:3] == (3, 3, 1)
    assert lib.model_stages("cartpole") == (1, [(4, 21, 14, 4)], 4)
