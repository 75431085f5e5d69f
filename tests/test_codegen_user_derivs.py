"""User-provided derivative constructors (reference src/dynamics.jl:58-61, src/constraints.jl:60-64) through the code
generator: closures that return the true derivatives give the same generated device code as symbolic differentiation,
and omitted contractions stay zero as in the reference.  CPU only (no compilation)."""
import dataclasses

import sympy as sp

import ipddp_b200  # noqa: F401
from ipddp_b200.codegen import generate, workloads


def _jac(fn, wrt):
    """closure (x, u, p) -> Jacobian of fn(x, u, p) w.r.t. x (wrt = 0) or u (wrt = 1)"""
    return lambda x, u, p: sp.Matrix(fn(x, u, p)).jacobian(sp.Matrix([x, u][wrt]))


def test_user_derivatives_reproduce_symbolic_ones():
    md = workloads.get("concar")          # nonlinear dynamics (RK2), state and control constraints
    ud = {"fx": _jac(md.f, 0), "fu": _jac(md.f, 1), "cx": _jac(md.c, 0), "cu": _jac(md.c, 1)}

    def contr(fn, first, second):          # (x, u, v, p) -> d/d second of (d fn / d first)' v
        def g(x, u, v, p):
            J = sp.Matrix(fn(x, u, p)).jacobian(sp.Matrix([x, u][first]))
            return (J.T * sp.Matrix(v)).jacobian(sp.Matrix([x, u][second]))
        return g
    ud.update(vfxx=contr(md.f, 0, 0), vfux=contr(md.f, 1, 0), vfuu=contr(md.f, 1, 1),
              vcxx=contr(md.c, 0, 0), vcux=contr(md.c, 1, 0), vcuu=contr(md.c, 1, 1))
    md_user = dataclasses.replace(md, user_derivs=ud, user_dynamics=True, user_constraint=True)
    auto, user = generate.trace(md), generate.trace(md_user)
    assert generate.emit_device(md, auto) == generate.emit_device(md_user, user)
    assert generate.emit_oracle(md, auto) == generate.emit_oracle(md_user, user)


def test_omitted_contractions_are_zero():
    md = workloads.get("concar")
    md_user = dataclasses.replace(md, user_derivs={"fx": _jac(md.f, 0), "fu": _jac(md.f, 1)}, user_dynamics=True)
    b = generate.trace(md_user)
    assert all(e.kind == "zero" for e in b["vf"].entries)               # vfxx = vfux = vfuu = nothing in the reference
    assert any(e.kind != "zero" for e in generate.trace(md)["vf"].entries)
    # the constraint group was not user-provided: its contractions are still differentiated symbolically
    assert [e.text for e in b["derivs"].entries if e.mat == "vcuu"] == \
        [e.text for e in generate.trace(md)["derivs"].entries if e.mat == "vcuu"]


def test_shape_check():
    md = workloads.get("double_integrator")
    bad = dataclasses.replace(md, user_derivs={"fx": lambda x, u, p: [[1.0, 0.0, 0.0]]}, user_dynamics=True)
    try:
        generate.trace(bad)
    except ValueError as e:
        assert "fx" in str(e)
    else:
        raise AssertionError("shape mismatch not detected")
