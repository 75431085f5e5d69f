"""SymPy -> straight-line C / CUDA code generator for IPDDP2 model functions.

This is the build's counterpart of the reference's Symbolics.jl tracing + `build_function`
(reference src/dynamics.jl:15-47, src/objectives.jl:12-33, src/constraints.jl:16-50): user closures
are traced with symbolic x, u, the Jacobians / Hessians / Hessian contractions are formed
symbolically, and in-place straight-line functions are emitted.  Two flavours are written from ONE
intermediate representation, so both evaluate the same floating-point expression trees in the same
order (both are compiled with FP contraction disabled):

  * oracle flavour (plain C, dense column-major outputs incl. structural zeros) -> oracle/models_gen/<name>.h
  * device flavour (CUDA `__device__`, only structurally non-constant entries go through HBM as a
    compact "tile"; scatter tables tell the backward kernel where each entry lives) ->
    interiorpointddp.jl_b200/csrc/models_gen/<name>.cuh

Bundles (one CSE pass each):
  dyn     f(x,u,p)                         reference Dynamics.evaluate
  cost    l(x,u,p), costN  l_N(x,p)        reference Objective.evaluate
  con     c(x,u,p)                         reference Constraint.evaluate
  derivs  fx fu lx lu lxx luu lux cx cu vcxx vcux vcuu   (x,u,phi,p)   reference src/derivatives.jl:1-35
  vf      vfxx vfux vfuu (x,u,lam,p)       reference src/dynamics.jl:63-70 (needs the costate of the sweep)
  derivsN lx lxx (x,p)                     terminal stage
"""
from __future__ import annotations

import hashlib
import os
import sys
from dataclasses import dataclass
from typing import Dict, List, Tuple

import sympy as sp
from sympy.printing.c import C99CodePrinter

HERE = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, HERE)
import workloads  # noqa: E402


class _Printer(C99CodePrinter):
    """Prints doubles with 17 significant digits, integer powers as products and
    transcendental calls through the deterministic DM_* math layer."""

    def _print_Float(self, expr):
        return repr(float(expr))

    def _print_Integer(self, expr):
        return f"{int(expr)}.0"

    def _print_Rational(self, expr):
        return f"({int(expr.p)}.0/{int(expr.q)}.0)"

    def _print_Pow(self, expr):
        b, e = expr.base, expr.exp
        if e.is_Integer:
            n = int(e)
            bs = self.parenthesize(b, 1000)
            if n > 0 and n <= 4:
                return "(" + "*".join([bs] * n) + ")"
            if n < 0 and n >= -4:
                return "(1.0/(" + "*".join([bs] * (-n)) + "))"
        if e == sp.Rational(1, 2):
            return f"sqrt({self._print(b)})"
        return f"DM_POW({self._print(b)}, {self._print(e)})"

    def _print_sin(self, expr):
        return f"DM_SIN({self._print(expr.args[0])})"

    def _print_cos(self, expr):
        return f"DM_COS({self._print(expr.args[0])})"

    def _print_tan(self, expr):
        return f"DM_TAN({self._print(expr.args[0])})"

    def _print_log(self, expr):
        return f"DM_LOG({self._print(expr.args[0])})"

    def _print_exp(self, expr):
        return f"DM_EXP({self._print(expr.args[0])})"


_P = _Printer()


def _sym(prefix, n):
    return [sp.Symbol(f"{prefix}[{i}]", real=True) for i in range(n)]


@dataclass
class Entry:
    mat: str
    i: int
    j: int
    kind: str          # 'zero' | 'const' | 'dyn'
    text: str          # C expression (for 'const' a literal, for 'dyn' an expression over temps/inputs)
    value: float = 0.0


@dataclass
class Bundle:
    name: str
    inputs: List[str]                       # names of input arrays in call order
    outputs: List[Tuple[str, int, int]]     # (name, rows, cols) column-major
    temps: List[Tuple[str, str]]            # (tmp name, expression text)
    entries: List[Entry]


def _make_bundle(name, inputs, mats: List[Tuple[str, sp.Matrix]]) -> Bundle:
    flat = []
    for mname, M in mats:
        M = sp.Matrix(M)
        for j in range(M.cols):
            for i in range(M.rows):
                flat.append((mname, i, j, M[i, j]))
    exprs = [sp.nsimplify(e, rational=False) if False else e for (_, _, _, e) in flat]
    repl, red = sp.cse(exprs, symbols=sp.numbered_symbols("w"), order="canonical")
    temps = [(str(s), _P.doprint(e)) for s, e in repl]
    entries = []
    for (mname, i, j, _), e in zip(flat, red):
        e = sp.sympify(e)
        if e == 0:
            entries.append(Entry(mname, i, j, "zero", "0.0", 0.0))
        elif e.is_number:
            entries.append(Entry(mname, i, j, "const", repr(float(e)), float(e)))
        else:
            entries.append(Entry(mname, i, j, "dyn", _P.doprint(e)))
    outs = [(mname, sp.Matrix(M).rows, sp.Matrix(M).cols) for mname, M in mats]
    return Bundle(name, inputs, outs, temps, entries)


def trace(md: workloads.ModelDef) -> Dict[str, Bundle]:
    nx, nu, npar = md.nx, md.nu, md.np_
    x, u, p = _sym("x", nx), _sym("u", nu), _sym("p", max(npar, 1))
    fx_ = sp.Matrix(md.f(x, u, p))
    nxn = fx_.rows
    c_ = sp.Matrix(md.c(x, u, p)) if md.nc > 0 else sp.zeros(0, 1)
    nc = c_.rows
    l_ = sp.sympify(md.stage_cost(x, u, p))
    nxt = int(getattr(md, "nx_term", None) or nx)      # the terminal cost may see a state of another size (stage chains)
    xt = x if nxt == nx else _sym("x", nxt)
    lN_ = sp.sympify(md.term_cost(xt, p))
    v = _sym("v", max(nc, 1))
    lam = _sym("v", nxn)
    X, U = sp.Matrix(x), sp.Matrix(u)

    b: Dict[str, Bundle] = {}
    b["dyn"] = _make_bundle("dyn", ["x", "u", "p"], [("f", fx_)])
    b["cost"] = _make_bundle("cost", ["x", "u", "p"], [("l", sp.Matrix([l_]))])
    b["costN"] = _make_bundle("costN", ["x", "p"], [("l", sp.Matrix([lN_]))])
    b["con"] = _make_bundle("con", ["x", "u", "p"], [("c", c_)])

    ud = getattr(md, "user_derivs", None) or {}

    def user(name, rows, cols, *args):
        """matrix from a user-provided derivative closure (shape checked), or None"""
        fn = ud.get(name)
        if fn is None:
            return None
        if getattr(fn, "_ipddp_inplace", False):     # the reference's in-place form: fn!(out, args...)
            M = sp.zeros(rows, cols)
            fn(M, *args)
        else:
            M = sp.Matrix(fn(*args))
        if M.shape != (rows, cols):
            if rows * cols == len(M):
                M = M.reshape(rows, cols)
            else:
                raise ValueError(f"user-provided {name}: expected {rows}x{cols}, got {M.shape}")
        return M

    def pick(name, rows, cols, args, auto, group_is_user):
        M = user(name, rows, cols, *args)
        if M is not None:
            return M
        if group_is_user and name.startswith("v"):   # contraction not provided: stays zero (reference semantics)
            return sp.zeros(rows, cols)
        return auto()

    udyn, ucon = bool(getattr(md, "user_dynamics", False)), bool(getattr(md, "user_constraint", False))
    qnd, qnc = bool(getattr(md, "qn_dynamics", False)), bool(getattr(md, "qn_constraint", False))
    Jfx = pick("fx", nxn, nx, (x, u, p), lambda: fx_.jacobian(X), udyn)
    Jfu = pick("fu", nxn, nu, (x, u, p), lambda: fx_.jacobian(U), udyn)
    lx = sp.Matrix([l_]).jacobian(X).T
    lu = sp.Matrix([l_]).jacobian(U).T
    lxx, luu, lux = lx.jacobian(X), lu.jacobian(U), lu.jacobian(X)
    cx = pick("cx", nc, nx, (x, u, p), lambda: c_.jacobian(X), ucon)
    cu = pick("cu", nc, nu, (x, u, p), lambda: c_.jacobian(U), ucon)
    vv = sp.Matrix(v[:nc])
    if qnc or nc == 0:   # quasi-Newton constraint object (or no constraints): the contraction caches stay zero
        vcxx, vcux, vcuu = sp.zeros(nx, nx), sp.zeros(nu, nx), sp.zeros(nu, nu)
    else:
        vcxx = pick("vcxx", nx, nx, (x, u, v[:nc], p), lambda: (cx.T * vv).jacobian(X), ucon)
        vcux = pick("vcux", nu, nx, (x, u, v[:nc], p), lambda: (cu.T * vv).jacobian(X), ucon)
        vcuu = pick("vcuu", nu, nu, (x, u, v[:nc], p), lambda: (cu.T * vv).jacobian(U), ucon)
    b["derivs"] = _make_bundle("derivs", ["x", "u", "v", "p"], [
        ("fx", Jfx), ("fu", Jfu), ("lx", lx), ("lu", lu), ("lxx", lxx), ("luu", luu), ("lux", lux),
        ("cx", cx), ("cu", cu), ("vcxx", vcxx), ("vcux", vcux), ("vcuu", vcuu)])

    lv = sp.Matrix(lam)
    if qnd:   # quasi-Newton dynamics object
        vfxx, vfux, vfuu = sp.zeros(nx, nx), sp.zeros(nu, nx), sp.zeros(nu, nu)
    else:
        vfxx = pick("vfxx", nx, nx, (x, u, lam, p), lambda: (Jfx.T * lv).jacobian(X), udyn)
        vfux = pick("vfux", nu, nx, (x, u, lam, p), lambda: (Jfu.T * lv).jacobian(X), udyn)
        vfuu = pick("vfuu", nu, nu, (x, u, lam, p), lambda: (Jfu.T * lv).jacobian(U), udyn)
    b["vf"] = _make_bundle("vf", ["x", "u", "v", "p"], [("vfxx", vfxx), ("vfux", vfux), ("vfuu", vfuu)])

    XT = sp.Matrix(xt)
    lNx = sp.Matrix([lN_]).jacobian(XT).T
    b["derivsN"] = _make_bundle("derivsN", ["x", "p"], [("lx", lNx), ("lxx", lNx.jacobian(XT))])
    return b


# ----------------------------------------------------------------------------------------
# oracle flavour
# ----------------------------------------------------------------------------------------
def emit_oracle(md, bundles) -> str:
    n = md.name
    nc = bundles["con"].outputs[0][1]
    nxn = bundles["dyn"].outputs[0][1]
    o = []
    o.append(f"// GENERATED by interiorpointddp.jl_b200/codegen/generate.py -- do not edit.\n"
             f"// Oracle (CPU, dense) model functions for workload '{n}'.  {md.doc}\n"
             f"// Dense column-major outputs; every entry is written (as the reference's generated closures do,\n"
             f"// reference src/dynamics.jl:26-34).\n#pragma once\n#include \"../oracle_model.h\"\n")
    o.append(f"namespace gen_{n} {{\n")
    for bname, b in bundles.items():
        args = ", ".join(f"const double* {a}" for a in b.inputs)
        outs = ", ".join(f"double* {m}" for (m, _, _) in b.outputs)
        o.append(f"static void {bname}({args}, {outs}) {{\n")
        for a in b.inputs:
            o.append(f"  (void){a};\n")
        for t, e in b.temps:
            o.append(f"  const double {t} = {e};\n")
        rows = {m: r for (m, r, _) in b.outputs}
        for en in b.entries:
            o.append(f"  {en.mat}[{en.i + en.j * rows[en.mat]}] = {en.text};\n")
        o.append("}\n")
    nxt = bundles["derivsN"].outputs[0][1]
    o.append(f"static const OracleModel model = {{\"{n}\", {md.nx}, {md.nu}, {nc}, {nxn}, {md.np_}, "
             f"dyn, cost, costN, con, derivs, vf, derivsN, {nxt}}};\n")
    o.append("}\n")
    return "".join(o)


# ----------------------------------------------------------------------------------------
# device flavour
# ----------------------------------------------------------------------------------------
_UPPER_ONLY = {"luu", "vcuu", "vfuu"}     # only the upper triangle of H-type blocks is ever read (sytrf 'U')


def _device_entries(b: Bundle):
    """Entries the device keeps: structural zeros dropped, strictly-lower part of H-type blocks dropped.
    Returns list of (Entry, slot) with slot >= 0 for tile entries, -1-constidx for constants."""
    out, consts, k = [], [], 0
    for en in b.entries:
        if en.kind == "zero":
            continue
        if en.mat in _UPPER_ONLY and en.i > en.j:
            continue
        if en.kind == "dyn":
            out.append((en, k)); k += 1
        else:
            out.append((en, -1 - len(consts))); consts.append(en.value)
    return out, consts, k


def emit_device(md, bundles) -> str:
    n = md.name
    nc = bundles["con"].outputs[0][1]
    nxn = bundles["dyn"].outputs[0][1]
    nx, nu = md.nx, md.nu
    o = []
    o.append(f"// GENERATED by interiorpointddp.jl_b200/codegen/generate.py -- do not edit.\n"
             f"// Device model functions for workload '{n}'.  {md.doc}\n"
             f"// Same temporaries / expression trees as the dense CPU flavour; only structurally non-constant\n"
             f"// derivative entries are stored (compact tile), constants live in the scatter tables.\n"
             f"#pragma once\n#include \"../model_common.cuh\"\n")
    o.append(f"namespace gen_{n} {{\n")

    tbl_defs = []   # (prefix, mat, offset, count)
    all_entries = []
    all_consts = []
    nslots = {}
    for bname, prefix in (("derivs", "D"), ("vf", "VF"), ("derivsN", "DN")):
        ents, consts, k = _device_entries(bundles[bname])
        nslots[prefix] = k
        cbase = len(all_consts)
        all_consts += consts
        for (m, _, _) in bundles[bname].outputs:
            lst = [(en, s) for (en, s) in ents if en.mat == m]
            tbl_defs.append((prefix, m, len(all_entries), len(lst)))
            for en, s in lst:
                slot = s if s >= 0 else -1 - (cbase + (-1 - s))
                all_entries.append((slot, en.i, en.j, f"{prefix}.{m}"))
    o.append(f"IPDDP_TABLE MEntry TBL[{max(len(all_entries), 1)}] = {{\n")
    if not all_entries:
        o.append("  {0, 0, 0}\n")
    for (slot, i, j, tag) in all_entries:
        o.append(f"  {{{slot}, {i}, {j}}},  // {tag}\n")
    o.append("};\n")
    o.append(f"IPDDP_TABLE double CONSTS[{max(len(all_consts), 1)}] = {{"
             + (", ".join(repr(float(c)) for c in all_consts) if all_consts else "0.0") + "};\n")
    # controls the dynamics depend on: the columns of fu that are not structurally zero (ascending) and the inverse map
    d_ents, _, _ = _device_entries(bundles["derivs"])
    fu_cols = sorted({en.j for (en, _s) in d_ents if en.mat == "fu"})
    fu_idx = [fu_cols.index(j) if j in fu_cols else 255 for j in range(nu)]
    o.append(f"IPDDP_TABLE unsigned char FUCOL[{max(len(fu_cols), 1)}] = {{" + (", ".join(str(c) for c in fu_cols) or "0") + "};\n")
    o.append(f"IPDDP_TABLE unsigned char FUIDX[{max(nu, 1)}] = {{" + (", ".join(str(c) for c in fu_idx) or "255") + "};\n")
    o.append("}\n")

    o.append(f"struct Model_{n} {{\n")
    o.append(f"  static constexpr const char* NAME = \"{n}\";\n")
    nxt = bundles["derivsN"].outputs[0][1]
    o.append(f"  static constexpr int NX = {nx}, NU = {nu}, NC = {nc}, NXN = {nxn}, NP = {md.np_};\n")
    o.append(f"  static constexpr int NXT = {nxt};   // state size the terminal cost (costN / derivsN) is evaluated on\n")
    o.append(f"  // a plain model is a chain of one stage type (see ipk::for_stage, model_common.cuh)\n")
    o.append(f"  static constexpr int NSTAGE = 1;\n")
    o.append(f"  template <int I> using Stage = Model_{n};\n")
    o.append(f"  using Terminal = Model_{n};\n")
    o.append(f"  static constexpr int NTBL = {len(all_entries)}, NCONST = {len(all_consts)};\n")
    o.append(f"  static constexpr int D_NSLOT = {nslots['D']}, VF_NSLOT = {nslots['VF']}, DN_NSLOT = {nslots['DN']};\n")
    for (prefix, m, off, cnt) in tbl_defs:
        o.append(f"  static constexpr int {prefix}_{m}_OFF = {off}, {prefix}_{m}_N = {cnt};\n")
    o.append(f"  static constexpr int FU_NC = {len(fu_cols)};   // controls with a structurally non-zero column in fu\n")
    o.append(f"  static IPDDP_D int fu_col(int c) {{ return gen_{n}::FUCOL[c]; }}\n")
    o.append(f"  static IPDDP_D int fu_idx(int u) {{ return gen_{n}::FUIDX[u]; }}   // compact column of control u, 255 = none\n")
    o.append(f"  static IPDDP_D const MEntry* tbl() {{ return gen_{n}::TBL; }}\n")
    o.append(f"  static IPDDP_D const double* consts() {{ return gen_{n}::CONSTS; }}\n")

    def fn(bname, compact):
        b = bundles[bname]
        args = ", ".join(f"const double* __restrict__ {a}" for a in b.inputs)
        if not compact:
            outs = ", ".join(f"double* __restrict__ {m}" for (m, _, _) in b.outputs)
            o.append(f"  static IPDDP_D void {bname}({args}, {outs}) {{\n")
        else:
            o.append(f"  template <class Store> static IPDDP_D void {bname}({args}, Store st) {{\n")
        for a in b.inputs:
            o.append(f"    (void){a};\n")
        if compact:
            ents, _, _ = _device_entries(b)
            needed = [en for (en, s) in ents if s >= 0]
        else:
            needed = b.entries
        # prune temporaries that no kept output needs (same expression text for the kept ones)
        used = set()
        import re
        tok = re.compile(r"\bw\d+\b")
        for en in needed:
            used.update(tok.findall(en.text))
        tdict = dict(b.temps)
        order = [t for t, _ in b.temps]
        stack = list(used)
        while stack:
            t = stack.pop()
            for d in tok.findall(tdict[t]):
                if d not in used:
                    used.add(d); stack.append(d)
        for t in order:
            if t in used:
                o.append(f"    const double {t} = {tdict[t]};\n")
        rows = {m: r for (m, r, _) in b.outputs}
        if not compact:
            for en in b.entries:
                o.append(f"    {en.mat}[{en.i + en.j * rows[en.mat]}] = {en.text};\n")
        else:
            for (en, s) in ents:
                if s >= 0:
                    o.append(f"    st({s}, {en.text});  // {en.mat}[{en.i},{en.j}]\n")
        o.append("  }\n")

    fn("dyn", False); fn("cost", False); fn("costN", False); fn("con", False)
    fn("derivs", True); fn("vf", True); fn("derivsN", True)
    o.append("};\n")
    return "".join(o)


def emit_device_chain(chain, stage_sources) -> str:
    """One header for a chain of stage types (state / control sizes that change along the horizon): the stage structs
    `Model_<chain>_s<i>` as emit_device writes them, plus the composite `Model_<chain>` the kernels are instantiated with.
    The composite carries the MAXIMA over the stage types under the usual names (NX, NU, NC, D_NSLOT, ...): they size every
    buffer and stride; the per-knot arithmetic is instantiated per stage type (ipk::for_stage)."""
    n = chain.name
    o = [f"// GENERATED by interiorpointddp.jl_b200/codegen/generate.py -- do not edit.\n"
         f"// Stage chain '{n}': {chain.doc}\n#pragma once\n"]
    for src in stage_sources:
        o.append(src.replace("#pragma once\n", ""))
    S = [f"Model_{md.name}" for md in chain.stages]
    mx = lambda field: "ipk::cmax(" + ", ".join(f"{s}::{field}" for s in S) + ")"
    o.append(f"struct Model_{n} {{\n")
    o.append(f"  static constexpr const char* NAME = \"{n}\";\n")
    o.append(f"  static constexpr int NSTAGE = {len(S)};\n")
    o.append(f"  template <int I> using Stage = typename ipk::TypeAt<I, {', '.join(S)}>::type;\n")
    o.append(f"  using Terminal = {S[-1]};\n")
    for f_ in ("NX", "NU", "NC", "NXN", "NP", "D_NSLOT", "VF_NSLOT", "FU_NC"):
        o.append(f"  static constexpr int {f_} = {mx(f_)};\n")
    o.append(f"  static constexpr int NXT = Terminal::NXT, DN_NSLOT = Terminal::DN_NSLOT;\n")
    o.append("};\n")
    return "".join(o)


def emit_oracle_chain(chain, stage_sources) -> str:
    n = chain.name
    o = [f"// GENERATED by interiorpointddp.jl_b200/codegen/generate.py -- do not edit.\n"
         f"// Oracle models of the stage types of chain '{n}'.\n#pragma once\n"]
    for src in stage_sources:
        o.append(src.replace("#pragma once\n", ""))
    return "".join(o)


def generate_chain(chain, oracle_dir, device_dir):
    bundles = [trace(md) for md in chain.stages]
    with open(os.path.join(oracle_dir, f"{chain.name}.h"), "w") as fh:
        fh.write(emit_oracle_chain(chain, [emit_oracle(md, b) for md, b in zip(chain.stages, bundles)]))
    with open(os.path.join(device_dir, f"{chain.name}.cuh"), "w") as fh:
        fh.write(emit_device_chain(chain, [emit_device(md, b) for md, b in zip(chain.stages, bundles)]))
    print(f"chain {chain.name}: " + ", ".join(f"{md.name}(nx={md.nx} nu={md.nu} nc={b['con'].outputs[0][1]} -> {b['dyn'].outputs[0][1]})"
                                              for md, b in zip(chain.stages, bundles)))


def generate_all(names=None, oracle_dir=None, device_dir=None):
    root = os.path.dirname(os.path.dirname(HERE))
    oracle_dir = oracle_dir or os.path.join(root, "oracle", "models_gen")
    device_dir = device_dir or os.path.join(HERE, "..", "csrc", "models_gen")
    os.makedirs(oracle_dir, exist_ok=True)
    os.makedirs(device_dir, exist_ok=True)
    names = names or (list(workloads.WORKLOADS) + list(workloads.CHAINS))
    for nm in names:
        if nm in workloads.CHAINS:
            generate_chain(workloads.get_chain(nm), oracle_dir, device_dir)
            continue
        md = workloads.get(nm)
        b = trace(md)
        with open(os.path.join(oracle_dir, f"{nm}.h"), "w") as fh:
            fh.write(emit_oracle(md, b))
        with open(os.path.join(device_dir, f"{nm}.cuh"), "w") as fh:
            fh.write(emit_device(md, b))
        _e, _c, nd = _device_entries(b["derivs"])
        ncst = len(_c)
        nvf = sum(1 for e in b["vf"].entries if e.kind != "zero")
        print(f"{nm}: nx={md.nx} nu={md.nu} nc={b['con'].outputs[0][1]} tile slots={nd} consts={ncst} "
              f"dense={len(b['derivs'].entries)} vf_nonzero={nvf} temps={len(b['derivs'].temps)}")


if __name__ == "__main__":
    generate_all(sys.argv[1:] or None)
