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


def test_per_object_quasi_newton_only_drops_that_objects_contractions():
    """Dynamics(...; quasi_newton=true) leaves vfxx/vfux/vfuu at zero, Constraint(...; quasi_newton=true) leaves
    vcxx/vcux/vcuu at zero (reference src/dynamics.jl:27-37,63-70, src/constraints.jl:76-83); the other object's terms
    stay (src/backward_pass.jl:101-114 only looks at Options.quasi_newton)."""
    md = workloads.get("concar")
    full = generate.trace(md)
    qd = generate.trace(dataclasses.replace(md, qn_dynamics=True))
    qc = generate.trace(dataclasses.replace(md, qn_constraint=True))

    def nonzero(bundle, names):
        return {en.mat for (en, _slot) in generate._device_entries(bundle)[0] if en.mat in names}
    vf, vc = {"vfxx", "vfux", "vfuu"}, {"vcxx", "vcux", "vcuu"}
    assert nonzero(full["vf"], vf) and nonzero(full["derivs"], vc)
    assert not nonzero(qd["vf"], vf) and nonzero(qd["derivs"], vc) == nonzero(full["derivs"], vc)
    assert not nonzero(qc["derivs"], vc) and nonzero(qc["vf"], vf) == nonzero(full["vf"], vf)


def test_model_identity_is_the_traced_source():
    """Two Solvers share a compiled model exactly when their emitted device code agrees: a module-level global that a
    closure reads is folded into the code at trace time, so changing it changes the model's digest (api.model_digest),
    while a differently named but identical model has the same digest."""
    from ipddp_b200 import api
    md = workloads.get("double_integrator")
    d0 = api.model_digest(md)
    assert d0 == api.model_digest(dataclasses.replace(md, name="something_else"))
    scale = {"dt": 0.01}
    f1 = lambda x, u, p: [x[0] + scale["dt"] * x[1], x[1] + scale["dt"] * u[0]]
    m1 = dataclasses.replace(md, f=f1)
    a = api.model_digest(m1)
    assert a == d0                                     # same code as the built-in closure
    scale["dt"] = 0.02
    assert api.model_digest(m1) != a                   # same closure object, different traced code
    assert api.model_digest(dataclasses.replace(md, indices_compl=[0])) != d0


def test_inplace_user_closures_as_in_the_reference():
    """The reference's user-derivative constructors take IN-PLACE closures (`f!(y, x, u)`, `fx!(J, x, u)`,
    `vfxx!(H, x, u, v)`: src/dynamics.jl:49-61, src/constraints.jl:60-64).  `inplace=True` accepts that form; the model
    it traces is the one the value-returning form gives."""
    from ipddp_b200 import api
    dt = 0.05
    f = lambda x, u: [x[0] + dt * x[1], x[1] + dt * sp.sin(u[0]) * x[0]]
    fx = lambda x, u: [[1, dt], [dt * sp.sin(u[0]), 1]]
    fu = lambda x, u: [[0, 0], [dt * sp.cos(u[0]) * x[0], 0]]
    vfux = lambda x, u, v: [[v[1] * dt * sp.cos(u[0]), 0], [0, 0]]
    c = lambda x, u: [u[1] - x[0] * u[0]]
    cx = lambda x, u: [[-u[0], 0]]
    cu = lambda x, u: [[-x[0], 1]]

    def f_ip(y, x, u):
        y[0] = x[0] + dt * x[1]
        y[1] = x[1] + dt * sp.sin(u[0]) * x[0]

    def fx_ip(J, x, u):
        J[0, 0] = 1; J[0, 1] = dt; J[1, 0] = dt * sp.sin(u[0]); J[1, 1] = 1

    def fu_ip(J, x, u):
        J[1, 0] = dt * sp.cos(u[0]) * x[0]

    def vfux_ip(H, x, u, v):
        H[0, 0] = v[1] * dt * sp.cos(u[0])

    def c_ip(out, x, u):
        out[0] = u[1] - x[0] * u[0]

    def cx_ip(J, x, u):
        J[0, 0] = -u[0]

    def cu_ip(J, x, u):
        J[0, 0] = -x[0]; J[0, 1] = 1

    dv = api.Dynamics(f, fx, fu, 2, 2, 2, vfux=vfux)
    di = api.Dynamics(f_ip, fx_ip, fu_ip, 2, 2, 2, vfux=vfux_ip, inplace=True)
    cv = api.Constraint(c, cx, cu, 1, 2, 2)
    ci = api.Constraint(c_ip, cx_ip, cu_ip, 1, 2, 2, inplace=True)

    def model(d, cc):
        return workloads.ModelDef(name="m", nx=2, nu=2, np_=0, f=d.f, stage_cost=lambda x, u, p: u[0] ** 2 + x[1] ** 2,
                                  term_cost=lambda x, p: x[0] ** 2, c=cc.c, lower=lambda p: [-1, -1], upper=lambda p: [1, 1],
                                  u_init=[0.0, 0.0], dt=dt, user_derivs={**d.user_derivs, **cc.user_derivs},
                                  user_dynamics=True, user_constraint=True)
    mv, mi = model(dv, cv), model(di, ci)
    assert generate.emit_device(mv, generate.trace(mv)) == generate.emit_device(mi, generate.trace(mi))
    assert api.model_digest(mv) == api.model_digest(mi)
