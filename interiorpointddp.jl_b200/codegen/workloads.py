"""The six benchmark OCP classes of the reference, restated as traceable Python closures.

Each workload is a `ModelDef`: dynamics f(x,u,p), stage cost, terminal cost and the equality
constraint c(x,u,p), written against *symbolic* x, u and a runtime parameter vector p (the
reference folds the parameters into the closure at trace time; batching needs them
at run time).  The generator in `generate.py` differentiates these with SymPy, exactly where the
reference differentiates with Symbolics.jl (reference src/dynamics.jl:15-47,
src/objectives.jl:12-33, src/constraints.jl:16-50).

Sources restated (all paths relative to the reference checkout):
  cartpole_friction  experiments/ipddp2/cartpole_friction.jl:37-105, experiments/models/cartpole.jl:19-55,94-131
  acrobot_contact    experiments/ipddp2/acrobot_contact.jl:36-113,   experiments/models/acrobot.jl:36-97,119-138
  concar/concar_quad experiments/ipddp2/concar.jl:31-131, concar_quad.jl:75
  pushing_1_obs      experiments/ipddp2/pushing_1_obs.jl:36-139
  double_integrator  experiments/ipddp2/double_integrator.jl:27-63
"""
from __future__ import annotations

import math
from dataclasses import dataclass, field
from typing import Callable, List, Sequence

import sympy as sp

INF = float("inf")


@dataclass
class ModelDef:
    name: str
    nx: int
    nu: int
    np_: int                       # number of runtime parameters per instance
    f: Callable                    # (x, u, p) -> list[nx]
    stage_cost: Callable           # (x, u, p) -> scalar
    term_cost: Callable            # (x, p) -> scalar
    c: Callable                    # (x, u, p) -> list[nc]
    lower: Callable                # (p: list[float]) -> list[nu] (may hold -inf)
    upper: Callable                # (p: list[float]) -> list[nu] (may hold +inf)
    u_init: Sequence[float]        # initial control guess for every running stage
    dt: float
    indices_compl: List[int] = field(default_factory=list)   # 0-based
    doc: str = ""
    # User-provided derivatives (reference src/dynamics.jl:58-61, src/constraints.jl:60-64): name -> callable returning
    # the matrix, (x, u, p) for fx, fu, cx, cu and (x, u, v, p) for the contractions vfxx, vfux, vfuu, vcxx, vcux, vcuu.
    # A key that is present replaces symbolic differentiation; in a user-provided group the missing contractions are
    # zero, as in the reference (`tensor_contraction!` is skipped when they are `nothing`).
    user_derivs: dict = field(default_factory=dict)
    user_dynamics: bool = False       # the Dynamics was built from user-provided derivatives
    user_constraint: bool = False     # the Constraint was built from user-provided derivatives
    # Dynamics(...; quasi_newton=true) / Constraint(...; quasi_newton=true): that object's second-order contractions are
    # never evaluated and its caches stay zero (reference src/dynamics.jl:27-37,63-70, src/constraints.jl:76-83); the other
    # object's terms are still added unless Options.quasi_newton is set (src/backward_pass.jl:101-114)
    qn_dynamics: bool = False
    qn_constraint: bool = False
    # dimension of the state the terminal cost is evaluated on (None = nx).  In a chain of stage types the terminal cost
    # lives with the LAST stage type and sees that stage's next state (num_next_state may differ from num_state).
    nx_term: object = None

    @property
    def nc(self) -> int:
        x = sp.symbols(f"x0:{self.nx}")
        u = sp.symbols(f"u0:{self.nu}")
        p = sp.symbols(f"p0:{max(self.np_, 1)}")
        return len(self.c(list(x), list(u), list(p)))


def _dot(a, b):
    return sum(ai * bi for ai, bi in zip(a, b))


# --------------------------------------------------------------------------------------
# cartpole with Coulomb friction (variational integrator + relaxed complementarity)
# --------------------------------------------------------------------------------------
def _cartpole() -> ModelDef:
    dt = 0.05
    g = 9.81
    nq, nF, ncp = 2, 1, 2
    nx = 2 * nq
    nu = nF + nq + 6 * ncp + 6

    def M(p, q):
        mc, mp, l = p[0], p[1], p[2]
        return [[mc + mp, mp * l * sp.cos(q[1])], [mp * l * sp.cos(q[1]), mp * l * l]]

    def Cf(p, q, qd):
        mc, mp, l = p[0], p[1], p[2]
        # C*qd - G  (models/cartpole.jl:37-43)
        c12 = -1.0 * mp * qd[1] * l * sp.sin(q[1])
        return [c12 * qd[1] - 0.0, 0.0 - (-mp * g * l * sp.sin(q[1]))]

    def manip(p, qm, q, qp, F, lam):
        qm_m = [0.5 * (qm[i] + q[i]) for i in range(2)]
        qm_p = [0.5 * (q[i] + qp[i]) for i in range(2)]
        qd_m = [(q[i] - qm[i]) / dt for i in range(2)]
        qd_p = [(qp[i] - q[i]) / dt for i in range(2)]
        Mp, Mm = M(p, qm_p), M(p, qm_m)
        Md = [Mp[i][0] * qd_p[0] + Mp[i][1] * qd_p[1] - (Mm[i][0] * qd_m[0] + Mm[i][1] * qd_m[1]) for i in range(2)]
        Cp, Cm = Cf(p, qm_p, qd_p), Cf(p, qm_m, qd_m)
        Ch = [0.5 * (Cp[i] + Cm[i]) for i in range(2)]
        B = [1.0, 0.0]
        return [Md[i] + dt * (Ch[i] - B[i] * F - lam[i]) for i in range(2)]

    def c(x, u, p):
        mc, mp, l, fr1, fr2 = p[0], p[1], p[2], p[3], p[4]
        qm, q = x[0:2], x[2:4]
        qp = u[nF:nF + nq]
        qd_p = [(qp[i] - q[i]) / dt for i in range(2)]
        F = u[0]
        o = nF + nq
        b1 = u[o:o + 2]
        b2 = u[o + ncp:o + ncp + 2]
        e1 = u[o + 2 * ncp:o + 2 * ncp + 2]
        e2 = u[o + 3 * ncp:o + 3 * ncp + 2]
        psi = u[o + 4 * ncp:o + 4 * ncp + 2]
        s = u[o + 5 * ncp:o + 5 * ncp + 2]
        sc = u[o + 6 * ncp:o + 6 * ncp + 6]
        lam = [b1[0] - b1[1], b2[0] - b2[1]]
        g1 = fr1 * (mp + mc) * g
        g2 = fr2 * mp * g * l
        return (manip(p, qm, q, qp, F, lam)
                + [qd_p[0] + psi[0] - e1[0], -qd_p[0] + psi[0] - e1[1]]
                + [qd_p[1] + psi[1] - e2[0], -qd_p[1] + psi[1] - e2[1]]
                + [g1 - (b1[0] + b1[1]) - s[0], g2 - (b2[0] + b2[1]) - s[1]]
                + [psi[0] * s[0] - sc[0], psi[1] * s[1] - sc[1]]
                + [b1[0] * e1[0] - sc[2], b1[1] * e1[1] - sc[3]]
                + [b2[0] * e2[0] - sc[4], b2[1] * e2[1] - sc[5]])

    def f(x, u, p):
        return [x[2], x[3], u[1], u[2]]

    def stage(x, u, p):
        F = u[0]
        sc = u[nF + nq + 6 * ncp:nF + nq + 6 * ncp + 6]
        return 0.01 * dt * F * F + sum(sc)

    def term(x, p):
        qN = [0.0, math.pi]
        qm, q = x[0:2], x[2:4]
        qd = [(q[i] - qm[i]) / dt for i in range(2)]
        dq = [q[i] - qN[i] for i in range(2)]
        return 200.0 * _dot(qd, qd) + 700.0 * _dot(dq, dq)

    lim = 10.0
    return ModelDef(
        name="cartpole", nx=nx, nu=nu, np_=5, f=f, stage_cost=stage, term_cost=term, c=c,
        lower=lambda p: [-lim] + [-INF] * nq + [0.0] * (6 * ncp) + [0.0] * 6,
        upper=lambda p: [lim] + [INF] * nq + [INF] * (6 * ncp) + [INF] * 6,
        u_init=[0.0] * (nF + nq) + [0.01] * (6 * ncp + 6), dt=dt,
        doc="p = [mc, mp, l, friction1, friction2]")


# --------------------------------------------------------------------------------------
# acrobot with joint-limit contact
# --------------------------------------------------------------------------------------
def _acrobot() -> ModelDef:
    dt = 0.05
    g = 9.81
    nq, nt, ncp = 2, 1, 2
    nx = 2 * nq
    nu = nt + nq + 3 * ncp

    def M(p, q):
        m1, I1, l1, lc1, m2, I2, l2, lc2 = p[0:8]
        a = I1 + I2 + m2 * l1 * l1 + 2.0 * m2 * l1 * lc2 * sp.cos(q[1])
        b = I2 + m2 * l1 * lc2 * sp.cos(q[1])
        return [[a, b], [b, I2]]

    def tau_g(p, q):
        m1, I1, l1, lc1, m2, I2, l2, lc2 = p[0:8]
        a = (-1.0 * m1 * g * lc1 * sp.sin(q[0])
             - m2 * g * (l1 * sp.sin(q[0]) + lc2 * sp.sin(q[0] + q[1])))
        b = -1.0 * m2 * g * lc2 * sp.sin(q[0] + q[1])
        return [a, b]

    def Cf(p, q, qd):
        m1, I1, l1, lc1, m2, I2, l2, lc2 = p[0:8]
        a = -2.0 * m2 * l1 * lc2 * sp.sin(q[1]) * qd[1]
        b = -1.0 * m2 * l1 * lc2 * sp.sin(q[1]) * qd[1]
        cc = m2 * l1 * lc2 * sp.sin(q[1]) * qd[0]
        tg = tau_g(p, q)
        return [a * qd[0] + b * qd[1] - tg[0], cc * qd[0] + 0.0 * qd[1] - tg[1]]

    def manip(p, qm, q, qp, tau, lam):
        qm_m = [0.5 * (qm[i] + q[i]) for i in range(2)]
        qm_p = [0.5 * (q[i] + qp[i]) for i in range(2)]
        qd_m = [(q[i] - qm[i]) / dt for i in range(2)]
        qd_p = [(qp[i] - q[i]) / dt for i in range(2)]
        Mp, Mm = M(p, qm_p), M(p, qm_m)
        Md = [Mp[i][0] * qd_p[0] + Mp[i][1] * qd_p[1] - (Mm[i][0] * qd_m[0] + Mm[i][1] * qd_m[1]) for i in range(2)]
        Cp, Cm = Cf(p, qm_p, qd_p), Cf(p, qm_m, qd_m)
        Ch = [0.5 * (Cp[i] + Cm[i]) for i in range(2)]
        B = [0.0, 1.0]
        # transpose(P) * lam with P = [0 -1; 0 1]
        Ntl = [0.0, -lam[0] + lam[1]]
        return [Md[i] + dt * (Ch[i] - B[i] * tau - Ntl[i] + 0.5 * qd_p[i]) for i in range(2)]

    def c(x, u, p):
        qm, q = x[0:2], x[2:4]
        qp = u[nt:nt + nq]
        tau = u[0]
        lam = u[nt + nq:nt + nq + ncp]
        s = u[nt + nq + ncp:nt + nq + 2 * ncp]
        sc = u[nt + nq + 2 * ncp:nt + nq + 3 * ncp]
        phi = [0.5 * math.pi - qp[1], qp[1] + 0.5 * math.pi]
        return (manip(p, qm, q, qp, tau, lam)
                + [s[0] - phi[0], s[1] - phi[1]]
                + [lam[0] * s[0] - sc[0], lam[1] * s[1] - sc[1]])

    def f(x, u, p):
        return [x[2], x[3], u[1], u[2]]

    def stage(x, u, p):
        tau = u[0]
        sc = u[nt + nq + 2 * ncp:nt + nq + 3 * ncp]
        return 0.01 * dt * tau * tau + 2.0 * sum(sc)

    def term(x, p):
        qN = [math.pi, 0.0]
        qm, q = x[0:2], x[2:4]
        qd = [(q[i] - qm[i]) / dt for i in range(2)]
        dq = [q[i] - qN[i] for i in range(2)]
        return 200.0 * _dot(qd, qd) + 700.0 * _dot(dq, dq)

    lim = 8.0
    return ModelDef(
        name="acrobot", nx=nx, nu=nu, np_=8, f=f, stage_cost=stage, term_cost=term, c=c,
        lower=lambda p: [-lim] + [-INF] * nq + [0.0] * (3 * ncp),
        upper=lambda p: [lim] + [INF] * nq + [INF] * (3 * ncp),
        u_init=[0.0] * (nt + nq) + [0.01] * (3 * ncp), dt=dt,
        doc="p = [m1, I1, l1, lc1, m2, I2, l2, lc2]")


# --------------------------------------------------------------------------------------
# car with 4 circular obstacles (RK2), linear or quadratic slack penalty
# --------------------------------------------------------------------------------------
def _concar(quad: bool) -> ModelDef:
    dt = 0.05
    r_car = 0.02
    nx, ncon, nobs = 4, 2, 4
    nu = ncon + 2 * nobs

    def gdyn(x, u):
        return [x[3] * sp.cos(x[2]), x[3] * sp.sin(x[2]), u[1], u[0]]

    def f(x, u, p):
        k1 = gdyn(x, u)
        xm = [x[i] + dt * 0.5 * k1[i] for i in range(nx)]
        k2 = gdyn(xm, u)
        return [x[i] + dt * k2[i] for i in range(nx)]

    def stage(x, u, p):
        s = u[ncon:ncon + nobs]
        J = dt * (u[0] * 5.0 * u[0] + u[1] * 1.0 * u[1])
        if quad:
            J = J + 1000.0 * _dot(s, s)
        else:
            J = J + 50.0 * sum(s)
        return J

    def term(x, p):
        xN = [1.0, 1.0, math.pi / 4, 0.0]
        d = [x[i] - xN[i] for i in range(nx)]
        return 200.0 * _dot(d, d)

    def c(x, u, p):
        out = []
        for i in range(nobs):
            ox, oy, orad = p[2 + 3 * i], p[3 + 3 * i], p[4 + 3 * i]
            d = [x[0] - ox, x[1] - oy]
            out.append((orad + r_car) * (orad + r_car) - _dot(d, d) - u[ncon + i] + u[ncon + nobs + i])
        return out

    return ModelDef(
        name="concar_quad" if quad else "concar", nx=nx, nu=nu, np_=14, f=f, stage_cost=stage,
        term_cost=term, c=c,
        lower=lambda p: [-p[0], -p[1]] + [0.0] * (2 * nobs),
        upper=lambda p: [p[0], p[1]] + [INF] * (2 * nobs),
        u_init=[0.0] * ncon + [0.01] * (2 * nobs), dt=dt,
        doc="p = [F_lim, tau_lim, obs1(x,y,r), obs2, obs3, obs4]")


# --------------------------------------------------------------------------------------
# planar pushing with one obstacle
# --------------------------------------------------------------------------------------
def _pushing() -> ModelDef:
    dt = 0.04
    nx, nu = 4, 11
    force_lim, vel_lim = 0.3, 3.0

    # p = [zx, zy, c, mu_fric, obs_x, obs_y, obs_r, r_total]; r_total = max(zx,zy) + r_push is
    # computed on the host exactly as the reference does at trace time (pushing_1_obs.jl:65).
    def f(x, u, p):
        zx, cc = p[0], p[2]
        th, ph = x[2], x[3]
        L = [1.0, 1.0, 1.0 / (cc * cc)]
        # Jc(ph)' * u[1:2]  with Jc = [[1 0 zx/2*tan(ph)]; [0 1 -zx/2]]  (2x3)
        Jtu = [u[0], u[1], zx / 2 * sp.tan(ph) * u[0] + (-zx / 2) * u[1]]
        w = [L[i] * Jtu[i] for i in range(3)]
        R = [[sp.cos(th), -sp.sin(th), 0.0], [sp.sin(th), sp.cos(th), 0.0], [0.0, 0.0, 1.0]]
        Rw = [R[i][0] * w[0] + R[i][1] * w[1] + R[i][2] * w[2] for i in range(3)]
        fc = Rw + [u[2] - u[3]]
        return [x[i] + dt * fc[i] for i in range(nx)]

    def stage(x, u, p):
        return 1e-2 * (u[0] * u[0] + u[1] * u[1]) + 2.0 * (u[6] + u[7]) + 2.0 * u[10]

    def term(x, p):
        xN = [0.3, 0.4, 1.5 * math.pi, 0.0]
        d = [x[i] - xN[i] for i in range(nx)]
        return 20.0 * _dot(d, d)

    def c(x, u, p):
        mu = p[3]
        d = [x[0] - p[4], x[1] - p[5]]
        obs = (p[6] + p[7]) * (p[6] + p[7]) - _dot(d, d) + u[9] - u[10]
        return [mu * u[0] - u[1] - u[4],
                mu * u[0] + u[1] - u[5],
                u[4] * u[2] - u[6],
                u[5] * u[3] - u[7],
                x[3] - u[8],
                obs]

    return ModelDef(
        name="pushing", nx=nx, nu=nu, np_=8, f=f, stage_cost=stage, term_cost=term, c=c,
        lower=lambda p: [0.0, -force_lim, 0.0, 0.0, 0.0, 0.0, 0.0, 0.0, -0.9, 0.0, 0.0],
        upper=lambda p: [force_lim, force_lim, vel_lim, vel_lim, INF, INF, INF, INF, 0.9, INF, INF],
        u_init=[0.01] * nu, dt=dt,
        doc="p = [zx, zy, c, mu_fric, obs_x, obs_y, obs_r, r_total]")


# --------------------------------------------------------------------------------------
# double integrator with absolute-work objective
# --------------------------------------------------------------------------------------
def _double_integrator() -> ModelDef:
    dt = 0.01
    nx, nu = 2, 3

    def f(x, u, p):
        return [x[0] + dt * x[1], x[1] + dt * u[0]]

    def stage(x, u, p):
        return dt * (u[1] + u[2])

    def term(x, p):
        xN = [1.0, 0.0]
        d = [x[i] - xN[i] for i in range(nx)]
        return 500.0 * _dot(d, d)

    def c(x, u, p):
        return [u[1] - u[2] - u[0] * x[1]]

    lim = 10.0
    return ModelDef(
        name="double_integrator", nx=nx, nu=nu, np_=0, f=f, stage_cost=stage, term_cost=term, c=c,
        lower=lambda p: [-lim, 0.0, 0.0], upper=lambda p: [lim, INF, INF],
        u_init=[0.01] * nu, dt=dt, doc="no parameters")


# --------------------------------------------------------------------------------------
# A horizon whose state and control sizes change along the way (reference README.md:18, src/data/problem.jl:44-62:
# every buffer is sized per timestep).  Synthetic test chain of three stage types built around the double integrator:
#   s0  (nx 2, nu 3, nc 1) -> 2   the double integrator with its absolute-work slacks
#   s1  (nx 2, nu 3, nc 1) -> 3   same controls, the dynamics adds a third state (running integral of the position)
#   s2  (nx 3, nu 2, nc 0) -> 3   fewer controls, no constraint, a cost on the integral state; carries the terminal cost
# --------------------------------------------------------------------------------------
@dataclass
class ChainDef:
    name: str
    stages: List[ModelDef]            # the stage types; stages[-1] carries the terminal cost (nx_term)
    doc: str = ""

    def stage_types(self, N: int) -> List[int]:
        """stage type of the running stages t = 0..N-2 of a horizon with N knots (the canonical test layout)"""
        n1 = (N - 1) // 2
        return [0] * n1 + [1] + [2] * (N - 2 - n1)


def _ragged_chain() -> ChainDef:
    dt = 0.01
    lim = 10.0

    def f0(x, u, p):
        return [x[0] + dt * x[1], x[1] + dt * u[0]]

    def f1(x, u, p):
        return [x[0] + dt * x[1], x[1] + dt * u[0], dt * x[0]]

    def f2(x, u, p):
        return [x[0] + dt * x[1], x[1] + dt * u[0], x[2] + dt * x[0]]

    def stage01(x, u, p):
        return dt * (u[1] + u[2])

    def stage2(x, u, p):
        return dt * (0.5 * u[0] * u[0] + u[1]) + 0.1 * dt * x[2] * x[2]

    def c01(x, u, p):
        return [u[1] - u[2] - u[0] * x[1]]

    def term(x, p):
        return 500.0 * ((x[0] - 1.0) * (x[0] - 1.0) + x[1] * x[1]) + 5.0 * x[2] * x[2]

    zero_term = lambda x, p: 0.0 * x[0]
    s0 = ModelDef(name="ragged_s0", nx=2, nu=3, np_=0, f=f0, stage_cost=stage01, term_cost=zero_term, c=c01,
                  lower=lambda p: [-lim, 0.0, 0.0], upper=lambda p: [lim, INF, INF], u_init=[0.01] * 3, dt=dt)
    s1 = ModelDef(name="ragged_s1", nx=2, nu=3, np_=0, f=f1, stage_cost=stage01, term_cost=zero_term, c=c01,
                  lower=lambda p: [-lim, 0.0, 0.0], upper=lambda p: [lim, INF, INF], u_init=[0.01] * 3, dt=dt)
    s2 = ModelDef(name="ragged_s2", nx=3, nu=2, np_=0, f=f2, stage_cost=stage2, term_cost=term, c=lambda x, u, p: [],
                  lower=lambda p: [-lim, 0.0], upper=lambda p: [lim, INF], u_init=[0.01] * 2, dt=dt, nx_term=3)
    return ChainDef("ragged", [s0, s1, s2], doc="double integrator whose state grows from 2 to 3 and whose controls shrink from 3 to 2")


CHAINS = {"ragged": _ragged_chain}


def get_chain(name: str) -> ChainDef:
    return CHAINS[name]()


WORKLOADS = {
    "cartpole": _cartpole,
    "acrobot": _acrobot,
    "concar": lambda: _concar(False),
    "concar_quad": lambda: _concar(True),
    "pushing": _pushing,
    "double_integrator": _double_integrator,
}


def get(name: str) -> ModelDef:
    return WORKLOADS[name]()
