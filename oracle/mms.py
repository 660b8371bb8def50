"""Manufactured solutions and MMS forcing terms for the oracle (TEST ORACLE).

Test infrastructure only -- see `oracle/__init__.py`.

The exact solutions are the published definitions of the reference's case
library (src/prob1_mms_cases.py:151-247 SlowlyChangingPeaks, 250-277 Pol,
280-337 ExpSin, 406-511 NonFullySmoothPol); derivatives are taken with SymPy and
lambdified for NumPy exactly as src/prob1base.py:1226-1280 does, including its
DiracDelta rule (|arg| < 1e-13 -> 1).  Forcing terms restate
src/prob1base.py:2313-2378 (fcp, fT, fcl), 3503-3551 (fcd, fcs with H_eta) and the
3x3 Gauss cell average of src/prob1base.py:493-598.
"""

from __future__ import annotations

import numpy as np
import sympy

from .ddoracle import OGrid, OModel, heaviside_reg, reaction_F2

# same assumptions as src/prob1base.py:1164
t_s, x_s, y_s = sympy.symbols("t x y", negative=False, real=True)

CASE_NAMES = ("pol", "expsin", "scp_fast1e1", "nfsp_h1h2", "nfsp_h2h2", "nfsp_h2h3")


def _pack(expr):
    """base/dt/dx/dy/lap callables f(t, X, Y) -> float64 array shaped like X
    (src/prob1base.py:1167-1280)."""
    mods = [{"DiracDelta": lambda a: np.where(abs(a) < 1e-13, 1.0, 0.0)}, "numpy"]
    dx = sympy.diff(expr, x_s)
    dy = sympy.diff(expr, y_s)
    table = {
        "base": expr,
        "dt": sympy.diff(expr, t_s),
        "dx": dx,
        "dy": dy,
        "lap": sympy.diff(dx, x_s) + sympy.diff(dy, y_s),
    }
    out = {}
    for name, e in table.items():
        raw = sympy.lambdify([t_s, x_s, y_s], e, modules=mods)

        def wrapped(t, X, Y, _raw=raw):
            r = np.asarray(_raw(t, X, Y), dtype=np.float64)
            if r.size == 1:
                return np.full(np.shape(X), r.reshape(-1)[0], dtype=np.float64)
            return r.reshape(np.shape(X)).astype(np.float64)

        out[name] = wrapped
    return out


class OCase:
    """Exact solution bundle: cp, T, cl, cd, cs and dt_/dx_/dy_/lap_ of each."""

    def __init__(self, exprs: dict):
        self.exprs = exprs
        for v, e in exprs.items():
            p = _pack(e)
            setattr(self, v, p["base"])
            setattr(self, "dt_" + v, p["dt"])
            setattr(self, "dx_" + v, p["dx"])
            setattr(self, "dy_" + v, p["dy"])
            setattr(self, "lap_" + v, p["lap"])


def _same(e):
    return {v: e for v in ("cp", "T", "cl", "cd", "cs")}


def nonsep_exprs(t, x, y):
    """A manufactured solution none of whose variables is a product f(t) X(x) Y(y); cs changes sign and has a kink
    (the expressions of tests/test_program_codegen.py:nonseparable_exprs, which the device runs from a generated
    forcing program; oracle/make_golden.py feeds the same ones to the reference's MMSCaseSymbolic)."""
    return dict(cp=sympy.exp(-t * x * y) / 2, T=1 + sympy.sin(sympy.pi * x * y + t) / 10,
                cl=sympy.cos(x + y * t) / 3, cd=sympy.exp(-(x - y) ** 2 - t) / 2,
                cs=(sympy.sin(sympy.pi * (x + y * t)) * sympy.exp(-t) - sympy.Rational(1, 5)
                    + sympy.Abs(x - sympy.Rational(1, 2)) ** sympy.Rational(21, 10)))


def make_case(name: str, model: OModel, **kw) -> OCase:
    x, y, t = x_s, y_s, t_s
    if name == "pol":  # src/prob1_mms_cases.py:258-264
        return OCase(_same(x * (1 - x) * y * (1 - y) / (1 + t)))
    if name.startswith("scp"):  # src/prob1_mms_cases.py:183-188, 233-235
        speed = kw.get("evol_speed", 1e1)
        const = kw.get("leading_spatial_const", 1.0)
        W = (x**2 + y**2) ** 3 * (sympy.sin(sympy.pi * x) * sympy.sin(sympy.pi * y)) * const
        return OCase(_same(W * sympy.exp(-speed * t)))
    if name.startswith("nfsp"):  # src/prob1_mms_cases.py:423-489, 507-510
        gam = {"nfsp_h1h2": [1.1, 2.1], "nfsp_h2h2": [2.1, 2.1], "nfsp_h2h3": [2.1, 3.1]}[name]
        gam = kw.get("gamma", gam)
        theta = kw.get("theta", 1 / np.pi)
        common = (1 / (1 + t)) * (x * (1 - x) * y * (1 - y))
        base = sympy.Abs((x - theta) * (y - theta))
        g_of = {"cp": gam[0], "cs": gam[0], "T": gam[1], "cl": gam[1], "cd": gam[1]}
        return OCase({v: common * base ** g_of[v] for v in g_of})
    if name == "nonsep":  # test-local MMSCaseSymbolic of tests/test_program_codegen.py (no library class)
        return OCase(nonsep_exprs(t, x, y))
    if name == "expsin":  # src/prob1_mms_cases.py:296-337
        pi = sympy.pi
        W = sympy.sin(pi * x) * sympy.sin(pi * y)
        T = sympy.exp(-2 * pi**2 * model.DT * t) * W
        cl = -sympy.exp(-t) * W
        cd = -cl
        pcp = sympy.integrate(-model.K1 * (1 + cl) - model.K2 * T, t)
        cp = W * sympy.exp(pcp - pcp.subs(t, 0))
        pcs = sympy.integrate(-model.Kd * (model.Sd - cd) * (1 + cl), t)
        cs = (model.r_sp * W) * sympy.exp(pcs - pcs.subs(t, 0))
        return OCase({"cp": cp, "T": T, "cl": cl, "cd": cd, "cs": cs})
    raise KeyError(name)


class OForcing:
    """ForcingTerms_RegHCsTriple on a fixed grid (src/prob1base.py:3468-3551,
    2296-2378).  Methods take only `t` (the grid is bound at construction)."""

    def __init__(self, case: OCase, model: OModel, eta: float, grid: OGrid):
        self.c, self.m, self.eta, self.g = case, model, float(eta), grid
        # quadrature abscissae of avg_int (src/prob1base.py:508-509, 538-570)
        nodes = np.array([-np.sqrt(3.0 / 5.0), 0, np.sqrt(3.0 / 5.0)])
        self._w = np.array([5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0])
        g = grid
        self._px = [g.xph[0:g.N - 1] + (n + 1.0) * 0.5 * g.hp[1:g.N] for n in nodes]
        self._py = [g.yph[0:g.M - 1] + (n + 1.0) * 0.5 * g.kp[1:g.M] for n in nodes]

    def _fcp_ptwise(self, t, X, Y):
        c, m = self.c, self.m
        cp = c.cp(t, X, Y)
        return c.dt_cp(t, X, Y) - (-cp * (m.K1 * (1 + c.cl(t, X, Y)) + m.K2 * c.T(t, X, Y)))

    def fcp(self, t):
        g = self.g
        acc = np.zeros((g.N - 1, g.M - 1))
        for a in range(3):
            for b in range(3):
                P, Q = np.meshgrid(self._px[a], self._py[b], indexing="ij")
                acc += self._w[a] * self._w[b] * self._fcp_ptwise(t, P, Q)
        out = g.zeros()
        out[1:-1, 1:-1] = 0.25 * acc
        return out

    def fT(self, t):
        c, m, X, Y = self.c, self.m, self.g.xx, self.g.yy
        return c.dt_T(t, X, Y) - (m.DT * c.lap_T(t, X, Y) - m.K3 * c.cp(t, X, Y) * c.T(t, X, Y))

    def fcl(self, t):
        c, m, X, Y = self.c, self.m, self.g.xx, self.g.yy
        cp, T, cl = c.cp(t, X, Y), c.T(t, X, Y), c.cl(t, X, Y)
        dxcl, dycl = c.dx_cl(t, X, Y), c.dy_cl(t, X, Y)
        return c.dt_cl(t, X, Y) - (
            m.dDl(cp) * (c.dx_cp(t, X, Y) * dxcl + c.dy_cp(t, X, Y) * dycl)
            + m.Dl(cp) * c.lap_cl(t, X, Y)
            - m.V1(T) * dxcl
            - (cl + 1) * (m.gamma_T * c.dx_T(t, X, Y))
            - m.K4 * cp * (cl + 1)
        )

    def _cd_parts(self, t):
        c, m, X, Y = self.c, self.m, self.g.xx, self.g.yy
        cp, T = c.cp(t, X, Y), c.T(t, X, Y)
        dcp, dT = m.dDd_dcp(cp, T), m.dDd_dT(cp, T)
        return (c.dt_cd(t, X, Y),
                (dcp * c.dx_cp(t, X, Y) + dT * c.dx_T(t, X, Y)) * c.dx_cd(t, X, Y)
                + (dcp * c.dy_cp(t, X, Y) + dT * c.dy_T(t, X, Y)) * c.dy_cd(t, X, Y)
                + m.Dd(cp, T) * c.lap_cd(t, X, Y))

    def fcd(self, t):
        c, m, X, Y = self.c, self.m, self.g.xx, self.g.yy
        dtcd, diff = self._cd_parts(t)
        H = reaction_F2(c.cs(t, X, Y), m, self.eta)  # H_eta(cs) | cs | (cs > 0): 3503-3551 | 2380-2413 | 3254-3286
        return dtcd - (diff + m.Kd * (m.Sd - c.cd(t, X, Y)) * (c.cl(t, X, Y) + 1) * H)

    def fcs(self, t):
        c, m, X, Y = self.c, self.m, self.g.xx, self.g.yy
        H = reaction_F2(c.cs(t, X, Y), m, self.eta)
        return c.dt_cs(t, X, Y) - (-m.Kd * (1 + c.cl(t, X, Y)) * (m.Sd - c.cd(t, X, Y)) * H)
