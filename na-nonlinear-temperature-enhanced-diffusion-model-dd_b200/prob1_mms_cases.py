"""Manufactured-solution library with device-evaluable descriptions.

Same case names and definitions as the reference's `prob1_mms_cases.py` (cited per
class), written from scratch on top of this package's `prob1base`.  Each case
additionally tells the CUDA path how to evaluate it without host arrays
(`device_spec()`): all cases but ExpSin are products phi(t) X(x) Y(y) per variable,
ExpSin has a closed form coded in the kernels.
"""

from __future__ import annotations

from typing import List

import numpy as np
import sympy

import prob1base as p1
from ddcore import ExpSinSpec, PhiSpec, SeparableSpec

x, y, t = p1.x_sym, p1.y_sym, p1.t_sym


def _spec_from_1d(phis: List[PhiSpec], Xs, Ys):
    """SeparableSpec from per-variable 1-D SymPy factors X_v(x), Y_v(y) (one product per variable)."""
    return p1.build_separable_spec(phis, [[(Xs[v], Ys[v])] for v in range(5)], x, y)


class _SeparableCase(p1.MMSCaseSymbolic):
    """u_v = phi_v(t) X_v(x) Y_v(y); subclasses pass the factors explicitly."""

    def __init__(self, grid, model, *, phi_exprs, phi_specs, Xs, Ys):
        exprs = [phi_exprs[v] * Xs[v] * Ys[v] for v in range(5)]
        super().__init__(grid=grid, model=model, cp_sym_expr=exprs[0], T_sym_expr=exprs[1], cl_sym_expr=exprs[2],
                         cd_sym_expr=exprs[3], cs_sym_expr=exprs[4], t_var=t, x_var=x, y_var=y)
        self._spec = _spec_from_1d(phi_specs, Xs, Ys)


class MMSCaseStiffExpDecay(_SeparableCase):
    """W = x(1-x)y(1-y), exp(-a_v t) with a_cl : a_T : a_cd = a_cs : a_cp = 1 : 1/10 : 1/100 : 1/1000
    (reference src/prob1_mms_cases.py:12-64)."""

    def __init__(self, grid, model, *, a_base: float = 1.0):
        a = [a_base / 1000.0, a_base / 10.0, a_base, a_base / 100.0, a_base / 100.0]  # cp, T, cl, cd, cs
        super().__init__(grid, model, phi_exprs=[sympy.exp(-ai * t) for ai in a],
                         phi_specs=[PhiSpec("exp", (1.0, ai)) for ai in a],
                         Xs=[x * (1 - x)] * 5, Ys=[y * (1 - y)] * 5)


class MMSCasePolWithOscilatingTime(_SeparableCase):
    """ampl (1 + shrink sin(speed t)) x(1-x)y(1-y) (reference :76-135)."""

    def __init__(self, grid, model, *, ampl: float = 1, speed: float = 1, shrink: float = 1):
        phi = ampl * (1 + shrink * sympy.sin(speed * t))
        super().__init__(grid, model, phi_exprs=[phi] * 5,
                         phi_specs=[PhiSpec("osc", (float(ampl), float(shrink), float(speed)))] * 5,
                         Xs=[x * (1 - x)] * 5, Ys=[y * (1 - y)] * 5)


def make_MMSCasePolWithOscilatingTime_cls(*, ampl, speed):
    class the_MMSCasePolWithOscilatingTime(MMSCasePolWithOscilatingTime):
        def __init__(self, grid, model):
            super().__init__(grid=grid, model=model, ampl=ampl, speed=speed)
    return the_MMSCasePolWithOscilatingTime


class MMSCaseSlowlyChangingPeaks(p1.MMSCaseSymbolic):
    """Const (x^2+y^2)^3 sin(pi x) sin(pi y) exp(-a t) for all variables (reference :151-209).

    On the device the spatial profile is the four-term separable sum
    (x^6 + 3 x^4 y^2 + 3 x^2 y^4 + y^6) sin(pi x) sin(pi y)."""

    def __init__(self, grid, model, *, leading_spatial_const=1e1, evol_speed: float = 1e-1):
        W = (x**2 + y**2) ** 3 * (sympy.sin(sympy.pi * x) * sympy.sin(sympy.pi * y)) * leading_spatial_const
        f = W * sympy.exp(-evol_speed * t)
        super().__init__(grid=grid, model=model, cp_sym_expr=f, T_sym_expr=f, cl_sym_expr=f, cd_sym_expr=f,
                         cs_sym_expr=f, t_var=t, x_var=x, y_var=y)
        sx, sy = sympy.sin(sympy.pi * x), sympy.sin(sympy.pi * y)
        c = sympy.Float(leading_spatial_const)
        terms = [(c * x**6 * sx, sy), (3 * c * x**4 * sx, y**2 * sy), (3 * c * x**2 * sx, y**4 * sy),
                 (c * sx, y**6 * sy)]
        self._spec = p1.build_separable_spec([PhiSpec("exp", (1.0, float(evol_speed)))] * 5, [terms] * 5, x, y)


def make_MMSCaseSlowlyChangingPeaks_cls(*, leading_spatial_const, evol_speed):
    class the_MMSCaseSlowlyChangingPeaks(MMSCaseSlowlyChangingPeaks):
        def __init__(self, grid, model):
            super().__init__(grid=grid, model=model, evol_speed=evol_speed,
                             leading_spatial_const=leading_spatial_const)
    return the_MMSCaseSlowlyChangingPeaks


MMSCaseSlowlyChangingPeaks_Slow1e1 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e-1)
MMSCaseSlowlyChangingPeaks_Slow1e2 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e-2)
MMSCaseSlowlyChangingPeaks_Slow1e3 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e-3)
MMSCaseSlowlyChangingPeaks_Slow1e4 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e-4)
MMSCaseSlowlyChangingPeaks_Slow1e8 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e-8)
MMSCaseSlowlyChangingPeaks_Slow1e16 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e-16)
MMSCaseSlowlyChangingPeaks_Fast1e1 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e1)
MMSCaseSlowlyChangingPeaks_Fast1e2 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e2)
MMSCaseSlowlyChangingPeaks_Fast1e3 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e3)
MMSCaseSlowlyChangingPeaks_Fast1e4 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e4)
MMSCaseSlowlyChangingPeaks_Fast1e8 = make_MMSCaseSlowlyChangingPeaks_cls(leading_spatial_const=1.0, evol_speed=1e8)


class MMSCasePol(_SeparableCase):
    """x(1-x)y(1-y)/(1+t) for all variables (reference :250-277)."""

    def __init__(self, grid, *, model):
        super().__init__(grid, model, phi_exprs=[1 / (1 + t)] * 5, phi_specs=[PhiSpec("inv1pt", (1.0,))] * 5,
                         Xs=[x * (1 - x)] * 5, Ys=[y * (1 - y)] * 5)


class MMSCaseExpSin(p1.MMSCaseSymbolic):
    """W = sin(pi x) sin(pi y); T = e^{-2 pi^2 DT t} W, cl = -e^{-t} W, cd = -cl,
    cp = W exp(int_0^t (-K1 (1+cl) - K2 T)), cs = r_sp W exp(int_0^t -Kd (Sd-cd)(1+cl))
    (reference :280-337)."""

    def __init__(self, grid, *, model):
        pi = sympy.pi
        W = sympy.sin(pi * x) * sympy.sin(pi * y)
        T = sympy.exp(-2 * pi**2 * model.DT * t) * W
        cl = -sympy.exp(-t) * W
        cd = -cl
        pcp = sympy.integrate(-model.K1 * (1 + cl) - model.K2 * T, t)
        cp = W * sympy.exp(pcp - pcp.subs(t, 0))
        pcs = sympy.integrate(-model.Kd * (model.Sd - cd) * (1 + cl), t)
        cs = (model.r_sp * W) * sympy.exp(pcs - pcs.subs(t, 0))
        super().__init__(grid=grid, model=model, cp_sym_expr=cp, T_sym_expr=T, cl_sym_expr=cl, cd_sym_expr=cd,
                         cs_sym_expr=cs, t_var=t, x_var=x, y_var=y)
        # the closed form in the kernels reads K1, K2, Kd, Sd, DT, r_sp from the member's model at step time;
        # the SymPy expressions above froze them at construction (as the reference does)
        self._frozen = tuple(float(getattr(model, n)) for n in ("K1", "K2", "Kd", "Sd", "DT", "r_sp"))
        self._spec = ExpSinSpec()

    def device_spec(self):
        now = tuple(float(getattr(self.model, n)) for n in ("K1", "K2", "Kd", "Sd", "DT", "r_sp"))
        return self._spec if now == self._frozen else None


class MMSCaseCsZeroCrossing(_SeparableCase):
    """cp = T = cl = cd = 0, cs = (A - B t) W (reference :341-403)."""

    def __init__(self, grid, model, *, cs_A: float = 0.5, cs_B: float = 1.0,
                 spatial_profile_expr=(x * (1 - x) * y * (1 - y))):
        fac = p1.separable_factors(spatial_profile_expr, t, x, y)
        if fac is None:
            raise ValueError("spatial_profile_expr must factor as X(x) Y(y)")
        c, gx, hy = fac
        ramp = sympy.Float(cs_A) - sympy.Float(cs_B) * t
        zero = sympy.S(0)
        super().__init__(grid, model, phi_exprs=[zero, zero, zero, zero, ramp * c],
                         phi_specs=[PhiSpec("const", (0.0,))] * 4 + [PhiSpec("linear", (float(cs_A * c), float(cs_B * c)))],
                         Xs=[gx] * 5, Ys=[hy] * 5)


class MMSCaseNonFullySmoothPol(_SeparableCase):
    """phi(t) W(x,y) |(x-theta)(y-theta)|^gamma_v, phi = 1/(1+t), W = x(1-x)y(1-y) (reference :406-500).
    gamma: 1 value (all), 2 values ((cp, cs), (T, cl, cd)) or 5 values."""

    def __init__(self, grid, *, model, gamma: List[float], theta: float = 1 / np.pi):
        if not (x.is_real and y.is_real and t.is_real):
            raise ValueError("x_sym, y_sym, and t_sym must be real symbols.")
        if not (x.is_nonnegative and y.is_nonnegative and t.is_nonnegative):
            raise ValueError("x_sym, y_sym, and t_sym must be non-negative symbols.")
        if np.isscalar(gamma):
            gamma = [float(gamma)]
        assert isinstance(gamma, list), "gamma must be a single number or a list of numbers."
        assert len(gamma) in [1, 2, 5], "gamma must have 1, 2 or 5 entries"
        if len(gamma) == 1:
            gamma = [gamma[0]] * 5
        elif len(gamma) == 2:
            gamma = [gamma[0], gamma[1], gamma[1], gamma[1], gamma[0]]
        assert all(gamma[j] > 1 for j in [0, 4]), "Cp's and cs' gamma (0, 4) must be greater than 1."
        assert all(gamma[j] > 2 for j in [1, 2, 3]), "T's, cl's, and cd's gammas (1, 2, 3) must be greater than 2."
        assert 0 < theta < 1, "Theta must be in (0, 1)."
        Xs = [x * (1 - x) * sympy.Abs(x - theta) ** g for g in gamma]
        Ys = [y * (1 - y) * sympy.Abs(y - theta) ** g for g in gamma]
        super().__init__(grid, model, phi_exprs=[1 / (1 + t)] * 5, phi_specs=[PhiSpec("inv1pt", (1.0,))] * 5,
                         Xs=Xs, Ys=Ys)


def make_MMSCaseNonFullySmoothPol_cls(gamma):
    class the_MMSCaseNonFullySmoothPol(MMSCaseNonFullySmoothPol):
        def __init__(self, grid, model):
            super().__init__(grid=grid, model=model, gamma=gamma)
    return the_MMSCaseNonFullySmoothPol


MMSCaseNonFullySmoothPol_cpcsH2_TclcdH3 = make_MMSCaseNonFullySmoothPol_cls(gamma=[2.1, 3.1])
MMSCaseNonFullySmoothPol_cpcsH1_TclcdH2 = make_MMSCaseNonFullySmoothPol_cls(gamma=[1.1, 2.1])
MMSCaseNonFullySmoothPol_cpcsH2_TclcdH2 = make_MMSCaseNonFullySmoothPol_cls(gamma=2.1)
MMSCaseNonFullySmoothPol_cpcsH3_TclcdH4 = make_MMSCaseNonFullySmoothPol_cls(gamma=[3.1, 4.1])
