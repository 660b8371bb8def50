"""B200-native drop-in for the time-stepping API of the reference's `prob1base.py`.

Same class names, argument meaning and error behaviour as the reference module
(cited per class as file:line relative to the reference root), but every
evaluation of the semidiscrete field and every time step runs in hand-written
sm_100a CUDA kernels behind the C ABI of libdd_b200.so (see `ddcore.Batch`).
NumPy is used only for what the reference API hands back to its callers as
ndarrays (grids, exact solutions, host-side source terms of user-supplied
forcing objects, the lazy convenience fields of `StateVars`).

There is no CPU fallback: stepping without the library or without a CUDA
device raises `DDLibraryError`.
"""

from __future__ import annotations

import numbers
from abc import ABC, abstractmethod
from typing import Callable, Dict, List, NamedTuple, Optional, Tuple

import numpy as np
import sympy

import ddcore
from _ddlib import MODE_ARRAYS, MODE_NONE, VARS
from ddcore import ExpSinSpec, PhiSpec, SeparableSpec


# ----------------------------------------------------------------------------
# model (reference src/prob1base.py:28-217)
# ----------------------------------------------------------------------------

class ModelConsts(NamedTuple):
    R0: float
    Ea: float
    K1: float
    K2: float
    K3: float
    K4: float
    DT: float
    Dl_max: float
    phi_l: float
    gamma_T: float
    Kd: float
    Sd: float
    Dd_max: float
    phi_d: float
    phi_T: float
    r_sp: float
    T_ref: float = 300


R0 = 8.3144621
Ea = 1.60217662e-19
default_model_consts = ModelConsts(
    R0=R0, Ea=Ea, K1=1e-2, K2=1e-2, K3=1e-2, K4=1e-2, DT=1e-3, Dl_max=8.01e-4, phi_l=1e-5, gamma_T=1e-9,
    Kd=1e-8, Sd=10, Dd_max=2.46e-6, phi_d=1e-5, phi_T=Ea / R0, r_sp=5e-2, T_ref=300)


class DefaultModel01:
    """Constants as mutable attributes + coefficient closures (reference :71-202)."""

    dd_kind = 1

    def __init__(self, mc: ModelConsts):
        for k, v in mc._asdict().items():
            setattr(self, k, v)

    def with_changes(self, **kwargs):
        # always a DefaultModel01, like the reference (:76-82): DefaultModel02(...).copy() drops the T_ref shift
        # of Dd, and tests/test_spatial_isolated_T_accuracy.py:407-409 of the reference relies on it
        out = DefaultModel01(ModelConsts(**{k: getattr(self, k) for k in ModelConsts._fields}))
        for k, v in kwargs.items():
            setattr(out, k, v)
        return out

    def copy(self):
        return self.with_changes()

    def Dl(self, cp, *, d=0):
        if isinstance(cp, sympy.Expr):
            return sympy.diff(self.Dl_max * sympy.exp(-self.phi_l * cp), cp, d)
        return ((-self.phi_l) ** d) * (self.Dl_max * np.exp(-self.phi_l * cp))

    def V1(self, T, *, d=0):
        if isinstance(T, sympy.Expr):
            return sympy.diff(self.gamma_T * T, T, d)
        if d == 0:
            return self.gamma_T * T
        if d == 1:
            return self.gamma_T * np.ones_like(T)
        return np.zeros_like(T)

    def V2(self, T, *, d=0):
        if isinstance(T, sympy.Expr):
            return sympy.S(0)
        return np.zeros_like(T)

    def _Dd_numeric(self, cp, Teff, d):
        cp = np.asarray(cp, dtype=np.float64)
        Teff = np.asarray(Teff, dtype=np.float64)
        assert cp.shape == Teff.shape
        nz = Teff != 0
        out = np.zeros_like(Teff, dtype=np.float64)
        out[nz] = self.Dd_max * np.exp(-self.phi_d * cp[nz]) * np.exp(-self.phi_T / Teff[nz])
        if d == (0, 0):
            return out
        if d == (1, 0):
            return -self.phi_d * out
        if d == (0, 1):
            out[nz] *= self.phi_T / (Teff[nz] ** 2)
            return out
        raise ValueError(f"unsupported derivative order {d}")

    def Dd(self, cp, T, *, d=(0, 0)):
        sym_cp, sym_T = isinstance(cp, sympy.Expr), isinstance(T, sympy.Expr)
        assert sym_cp == sym_T
        if sym_cp:
            e = self.Dd_max * sympy.exp(-self.phi_d * cp) * sympy.exp(-self.phi_T / T)
            return sympy.diff(sympy.diff(e, cp, d[0]), T, d[1])
        return self._Dd_numeric(cp, T, d)


class DefaultModel02(DefaultModel01):
    """Dd evaluated at T + T_ref (reference :205-217)."""

    dd_kind = 2

    def Dd(self, cp, T, *, d=(0, 0)):
        return super().Dd(cp, T + self.T_ref, d=d)


# ----------------------------------------------------------------------------
# grid (reference src/prob1base.py:220-490).  The O(NM) Python-loop asserts of the
# reference constructor are deliberately not reproduced; 2-D arrays are lazy.
# ----------------------------------------------------------------------------

class Grid:
    def __init__(self, x: np.ndarray, y: np.ndarray):
        x, y = np.asarray(x), np.asarray(y)
        assert len(x.shape) == len(y.shape)
        assert len(x.shape) in [1, 2], "Grid: x,y's shape must be 1D or 2D."
        if len(x.shape) == 2:
            assert x.shape == y.shape, "Grid: for meshgrid'ed x,y, same shape is required."
            x, y = x[:, 0], y[0, :]
        self.x, self.y = np.asarray(x, dtype=np.float64), np.asarray(y, dtype=np.float64)
        self.N, self.M = len(x) - 1, len(y) - 1
        self.h = np.concatenate([[np.inf], self.x[1:] - self.x[:-1]])
        self.k = np.concatenate([[np.inf], self.y[1:] - self.y[:-1]])
        self.h_phalf = np.concatenate([(self.h[:-1] + self.h[1:]) * 0.5, [np.inf]])
        self.k_phalf = np.concatenate([(self.k[:-1] + self.k[1:]) * 0.5, [np.inf]])
        self._lazy = {}
        self._dd_batch = None

    def _get(self, name, make):
        if name not in self._lazy:
            self._lazy[name] = make()
        return self._lazy[name]

    @property
    def xx(self):
        return self._get("xx", lambda: np.meshgrid(self.x, self.y, indexing="ij")[0])

    @property
    def yy(self):
        return self._get("yy", lambda: np.meshgrid(self.x, self.y, indexing="ij")[1])

    @property
    def hh(self):
        return self._get("hh", lambda: np.meshgrid(self.k, self.h)[1])

    @property
    def kk(self):
        return self._get("kk", lambda: np.meshgrid(self.k, self.h)[0])

    @property
    def hh_phalf(self):
        return self._get("hh_phalf", lambda: np.meshgrid(self.k_phalf, self.h_phalf)[1])

    @property
    def kk_phalf(self):
        return self._get("kk_phalf", lambda: np.meshgrid(self.k_phalf, self.h_phalf)[0])

    @property
    def xx_phalf(self):
        def make():
            a = np.zeros(self.full_shape)
            a[:-1, :] = 0.5 * (self.xx[:-1, :] + self.xx[1:, :])
            return a
        return self._get("xx_phalf", make)

    @property
    def yy_phalf(self):
        def make():
            a = np.zeros(self.full_shape)
            a[:, :-1] = 0.5 * (self.yy[:, :-1] + self.yy[:, 1:])
            return a
        return self._get("yy_phalf", make)

    @property
    def full_shape(self):
        return (self.N + 1, self.M + 1)

    @property
    def interior_shape(self):
        return (self.N - 1, self.M - 1)

    def make_full0(self):
        return np.zeros(self.full_shape)

    @property
    def null_bd_mask(self):
        return self._get("mask", lambda: self.const_with_nullbd(1))

    def const_with_nullbd(self, x):
        a = x * np.ones(self.full_shape)
        a[:, 0] = 0
        a[:, -1] = 0
        a[0, :] = 0
        a[-1, :] = 0
        return a

    # discrete inner products / norms (reference :387-433)
    def inner_product_H(self, u, v):
        return np.sum(u[1:-1, 1:-1] * np.conjugate(v[1:-1, 1:-1]) * self.h_phalf[1:-1, None] * self.k_phalf[None, 1:-1])

    def norm_H(self, u):
        return np.sqrt(self.inner_product_H(u, u))

    def inner_product_pk(self, u, v):
        return np.sum(u[1:, 1:-1] * np.conjugate(v[1:, 1:-1]) * self.h[1:, None] * self.k_phalf[None, 1:-1])

    def norm_pk(self, u):
        return np.sqrt(self.inner_product_pk(u, u))

    def inner_product_hp(self, u, v):
        return np.sum(u[1:-1, 1:] * np.conjugate(v[1:-1, 1:]) * self.h_phalf[1:-1, None] * self.k[None, 1:])

    def norm_hp(self, u):
        return np.sqrt(self.inner_product_hp(u, u))

    def inner_product_p(self, ux, uy, vx, vy):
        return self.inner_product_pk(ux, vx) + self.inner_product_hp(uy, vy)

    def norm_p(self, ux, uy):
        return np.sqrt(self.inner_product_p(ux, uy, ux, uy))

    def Dx_reg(self, u):
        return Dx_reg(u, self.h[:, None])

    def Dy_reg(self, u):
        return Dy_reg(u, self.k[None, :])

    def Dx_star(self, u):
        return Dx_star(u, self.h_phalf[:, None])

    def Dy_star(self, u):
        return Dy_star(u, self.k_phalf[None, :])

    def grad_H(self, u):
        return (self.Dx_reg(u), self.Dy_reg(u))

    # device batch holding one trajectory on this grid
    def device_batch(self) -> ddcore.Batch:
        if self._dd_batch is None:
            self._dd_batch = ddcore.Batch(self.x, self.y, 1, nslots=4)
        return self._dd_batch


def make_uniform_grid(N: int, M: int):
    return Grid(np.linspace(0, 1, N + 1), np.linspace(0, 1, M + 1))


def avg_int(f: Callable[[np.ndarray, np.ndarray], np.ndarray], grid: Grid):
    """3x3 Gauss cell average of f over the dual cells, interior only (reference :493-598).
    Host helper: used when a forcing object is evaluated by the host."""
    w = np.array([5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0])
    px = ddcore.quadrature_points(grid.x)[1:grid.N]
    py = ddcore.quadrature_points(grid.y)[1:grid.M]
    acc = np.zeros(grid.interior_shape)
    for a in range(3):
        for b in range(3):
            P, Q = np.meshgrid(px[:, a], py[:, b], indexing="ij")
            vals = f(P, Q)
            assert vals.shape == grid.interior_shape
            acc += w[a] * w[b] * vals
    out = grid.make_full0()
    out[1:-1, 1:-1] = 0.25 * acc
    return out


# stencil operators (reference :1499-1550)

def Mx_reg(u):
    o = np.zeros_like(u)
    o[1:, :] = 0.5 * (u[1:, :] + u[:-1, :])
    return o


def My_reg(u):
    o = np.zeros_like(u)
    o[:, 1:] = 0.5 * (u[:, 1:] + u[:, :-1])
    return o


def Dx_reg(u, hh):
    o = np.zeros_like(u)
    o[1:, :] = (u[1:, :] - u[:-1, :]) / hh[1:, :]
    return o


def Dy_reg(u, kk):
    o = np.zeros_like(u)
    o[:, 1:] = (u[:, 1:] - u[:, :-1]) / kk[:, 1:]
    return o


def Dx_star(u, hh_phalf):
    o = np.zeros_like(u)
    o[1:-1, :] = (u[2:, :] - u[1:-1, :]) / hh_phalf[1:-1, :]
    return o


def Dy_star(u, kk_phalf):
    o = np.zeros_like(u)
    o[:, 1:-1] = (u[:, 2:] - u[:, 1:-1]) / kk_phalf[:, 1:-1]
    return o


# ----------------------------------------------------------------------------
# MMS plumbing (reference :714-889, 1158-1487)
# ----------------------------------------------------------------------------

_MMS_METHODS = [p + v for v in VARS for p in ("", "dt_", "dx_", "dy_")] + ["lap_T", "lap_cl", "lap_cd"]


class MMSCaseBase(ABC):
    def __init__(self, grid: Grid, model: DefaultModel01):
        self._model = model
        self._grid = grid
        self._xx = grid.xx
        self._yy = grid.yy

    @property
    def grid(self):
        return self._grid

    @property
    def model(self):
        return self._model

    def device_spec(self):
        """A SeparableSpec / ExpSinSpec when the device can evaluate this solution itself, else None."""
        return None


class ForcingTermsBase(ABC):
    @abstractmethod
    def fcp(self, t, xx, yy) -> np.ndarray:
        pass

    @abstractmethod
    def fT(self, t, xx, yy) -> np.ndarray:
        pass

    @abstractmethod
    def fcl(self, t, xx, yy) -> np.ndarray:
        pass

    @abstractmethod
    def fcd(self, t, xx, yy) -> np.ndarray:
        pass

    @abstractmethod
    def fcs(self, t, xx, yy) -> np.ndarray:
        pass

    def asdict(self) -> Dict[str, Callable]:
        return {n: getattr(self, n) for n in ("fcp", "fT", "fcl", "fcd", "fcs")}


class NoForcingTerms(ForcingTermsBase):
    def __init__(self, grid: Grid):
        self._grid = grid

    def fcp(self, t, xx, yy):
        return self._grid.make_full0()

    def fT(self, t, xx, yy):
        return self._grid.make_full0()

    def fcl(self, t, xx, yy):
        return self._grid.make_full0()

    def fcd(self, t, xx, yy):
        return self._grid.make_full0()

    def fcs(self, t, xx, yy):
        return self._grid.make_full0()


class ForcingTermsFromDict(ForcingTermsBase):
    def __init__(self, forcing_terms_dict: Dict):
        self._d = forcing_terms_dict

    def fcp(self, t, xx, yy):
        return self._d["fcp"](t, xx, yy)

    def fT(self, t, xx, yy):
        return self._d["fT"](t, xx, yy)

    def fcl(self, t, xx, yy):
        return self._d["fcl"](t, xx, yy)

    def fcd(self, t, xx, yy):
        return self._d["fcd"](t, xx, yy)

    def fcs(self, t, xx, yy):
        return self._d["fcs"](t, xx, yy)


t_sym, x_sym, y_sym = sympy.symbols("t x y", negative=False, real=True)

_LAMBDIFY_MODULES = [{"DiracDelta": lambda arg: np.where(abs(arg) < 1e-13, 1.0, 0.0)}, "numpy"]


def _shape_adjusting(raw):
    def wrapped(t_num, x_num, y_num):
        assert isinstance(t_num, numbers.Number)
        shape = np.shape(x_num)
        assert shape == np.shape(y_num)
        r = np.asarray(raw(t_num, x_num, y_num))
        if r.size == 1:
            return np.full(shape, r.reshape(-1)[0], dtype=np.float64)
        assert r.size == int(np.prod(shape))
        return np.reshape(r, shape).astype(np.float64)
    return wrapped


def pack_symbolic_txy_with_derivatives(*, base_expr, t_var=t_sym, x_var=x_sym, y_var=y_sym) -> Dict[str, Callable]:
    dt = sympy.diff(base_expr, t_var)
    dx = sympy.diff(base_expr, x_var)
    dy = sympy.diff(base_expr, y_var)
    dxx = sympy.diff(dx, x_var)
    dyy = sympy.diff(dy, y_var)
    table = {"base": base_expr, "dt": dt, "dtt": sympy.diff(dt, t_var), "dx": dx, "dy": dy, "dxx": dxx,
             "dyy": dyy, "lap": dxx + dyy}
    return {k: _shape_adjusting(sympy.lambdify([t_var, x_var, y_var], e, modules=_LAMBDIFY_MODULES))
            for k, e in table.items()}


def _lambdify_1d(expr, var):
    return sympy.lambdify([var], expr, modules=_LAMBDIFY_MODULES)


def _split_abs_products(e):
    """|a b|^g -> |a|^g |b|^g so that products inside Abs separate by variable."""
    e = e.replace(lambda q: isinstance(q, sympy.Abs) and q.args[0].is_Mul,
                  lambda q: sympy.Mul(*[sympy.Abs(f) for f in q.args[0].args]))
    return sympy.expand_power_base(e, force=True)


MAX_SEPARABLE_TERMS = 8


def separable_terms(expr, t_var=t_sym, x_var=x_sym, y_var=y_sym):
    """(f(t), [(X_r(x), Y_r(y)), ...]) with expr == f * sum_r X_r Y_r, or None when the expression
    does not have that form (or needs more than MAX_SEPARABLE_TERMS terms)."""
    expr = sympy.sympify(expr)
    if expr == 0:
        return sympy.S(0), [(sympy.S(1), sympy.S(1))]
    e = _split_abs_products(sympy.factor_terms(expr))
    ft, rest = e.as_independent(x_var, y_var, as_Add=False)
    if rest.has(t_var):
        return None
    terms = []
    for term in sympy.Add.make_args(sympy.expand(rest)):
        term = _split_abs_products(term)
        gx, hy = term.as_independent(y_var, as_Add=False)
        if hy.has(x_var) or gx.has(y_var):
            return None
        terms.append((gx, hy))
    if len(terms) > MAX_SEPARABLE_TERMS:
        return None
    return ft, terms


def separable_factors(expr, t_var=t_sym, x_var=x_sym, y_var=y_sym):
    """(f(t), X(x), Y(y)) with expr == f X Y, or None when it is not a single product."""
    e = _split_abs_products(sympy.factor_terms(sympy.sympify(expr)))
    if e != 0:
        # a product such as x (1 - x) y (1 - y) splits as it stands (expanding it would give four terms)
        ft, rest = e.as_independent(x_var, y_var, as_Add=False)
        if not rest.has(t_var):
            for cand in (rest, sympy.factor(rest)):
                gx, hy = cand.as_independent(y_var, as_Add=False)
                if not gx.has(y_var) and not hy.has(x_var):
                    return ft, gx, hy
    st = separable_terms(expr, t_var, x_var, y_var)
    if st is None or len(st[1]) != 1:
        return None
    return st[0], st[1][0][0], st[1][0][1]


class MMSCaseSymbolic(MMSCaseBase):
    """Exact solution from SymPy expressions (reference :1283-1487)."""

    def __init__(self, *, grid, model, cp_sym_expr, T_sym_expr, cl_sym_expr, cd_sym_expr, cs_sym_expr,
                 t_var=t_sym, x_var=x_sym, y_var=y_sym):
        super().__init__(grid, model)
        self._vars3 = (t_var, x_var, y_var)
        self._exprs = {"cp": sympy.sympify(cp_sym_expr), "T": sympy.sympify(T_sym_expr),
                       "cl": sympy.sympify(cl_sym_expr), "cd": sympy.sympify(cd_sym_expr),
                       "cs": sympy.sympify(cs_sym_expr)}
        pk = dict(t_var=t_var, x_var=x_var, y_var=y_var)
        self._packs = {v: pack_symbolic_txy_with_derivatives(base_expr=e, **pk) for v, e in self._exprs.items()}
        self._spec = False  # not yet derived

    cp_pack = property(lambda self: self._packs["cp"])
    T_pack = property(lambda self: self._packs["T"])
    cl_pack = property(lambda self: self._packs["cl"])
    cd_pack = property(lambda self: self._packs["cd"])
    cs_pack = property(lambda self: self._packs["cs"])

    def device_spec(self):
        """Try u_v = f_v(t) X_v(x) Y_v(y) for all five variables (time profile evaluated by the host)."""
        if self._spec is False:
            self._spec = self._derive_separable_spec()
        return self._spec

    def device_program(self):
        """Generated forcing program (ddprogram.ProgramSpec) for expressions that are not separable: the five
        expressions and their derivatives printed as CUDA C and compiled with NVRTC (the device counterpart of
        the reference's lambdify, :1226-1280).  None when the expressions cannot be printed."""
        if getattr(self, "_program", False) is False:
            import ddprogram
            t_var, x_var, y_var = self._vars3
            self._program = ddprogram.program_for(self._exprs, t_var, x_var, y_var)
        return self._program

    def _derive_separable_spec(self):
        t_var, x_var, y_var = self._vars3
        try:
            parts = []
            for v in VARS:
                st = separable_terms(self._exprs[v], t_var, x_var, y_var)
                if st is None:
                    return None
                parts.append(st)
            return build_separable_spec([PhiSpec("host", f=sympy.lambdify([t_var], ft, modules="numpy"),
                                                 df=sympy.lambdify([t_var], sympy.diff(ft, t_var), modules="numpy"))
                                         for ft, _ in parts],
                                        [terms for _, terms in parts], x_var, y_var)
        except Exception:
            return None


def build_separable_spec(phis, terms_per_var, x_var=x_sym, y_var=y_sym) -> SeparableSpec:
    """SeparableSpec from per-variable lists of (X_r(x), Y_r(y)) SymPy factors; shorter lists are padded
    with zero terms so that every variable has the same number of terms."""
    R = max(len(t) for t in terms_per_var)
    X, Y = [], []
    for terms in terms_per_var:
        terms = list(terms) + [(sympy.S(0), sympy.S(0))] * (R - len(terms))
        X.append([[_lambdify_1d(sympy.diff(sympy.sympify(gx), x_var, d), x_var) for d in range(3)]
                  for gx, _ in terms])
        Y.append([[_lambdify_1d(sympy.diff(sympy.sympify(hy), y_var, d), y_var) for d in range(3)]
                  for _, hy in terms])
    return SeparableSpec(phi=list(phis), X=X, Y=Y)


def _make_pack_method(var, key):
    def method(self, t, xx, yy):
        return self._packs[var][key](t, xx, yy)
    return method


for _v in VARS:
    setattr(MMSCaseSymbolic, _v, _make_pack_method(_v, "base"))
    for _k in ("dt", "dtt", "dx", "dy", "dxx", "dyy", "lap"):
        setattr(MMSCaseSymbolic, f"{_k}_{_v}", _make_pack_method(_v, _k))


# ----------------------------------------------------------------------------
# MMS case from plain callables (reference :895-1155): derivatives by second-order finite differences.
# Host-only by nature (the solution is an opaque Python function): its forcing goes through the ARRAYS mode.
# ----------------------------------------------------------------------------

# (offsets in units of eps, integer weights, divisor as a function of eps) of the second-order difference
# formulas; summed left to right so that the rounding is that of the reference's written-out expressions
_FD_FIRST = {"center": ((1, -1), (1, -1), lambda e: 2 * e),
             "forward": ((0, 1, 2), (-3, 4, -1), lambda e: 2 * e),
             "backward": ((0, -1, -2), (3, -4, 1), lambda e: 2 * e)}
_FD_SECOND = {"center": ((1, 0, -1), (1, -2, 1), lambda e: e * e),
              "forward": ((0, 1, 2, 3), (2, -5, 4, -1), lambda e: e * e),
              "backward": ((0, -1, -2, -3), (2, -5, 4, -1), lambda e: e * e)}


def pack_analytical_txy_with_o2fdm_derivatives(fn, *, default_eps: float = 1e-6, time_stepping: str = "center"):
    """`fn(t, x, y)` -> `g(t, x, y, *, d=(dt, dx, dy), op=None, small_eps=None)`: derivatives of combined order
    <= 2 by second-order finite differences of step eps (one-sided in t for "forward" / "backward"), and
    `op="laplacian"` (or "lap") for the five-point Laplacian.  Same interface and formulas as the reference's
    helper (:895-1031)."""
    if time_stepping not in _FD_FIRST:
        raise ValueError("Invalid time stepping strategy")

    def line(t, x, y, axis, table, eps):
        offsets, weights, divisor = table
        acc = None
        for off, w in zip(offsets, weights):
            s = off * eps
            term = fn(t + s if axis == 0 else t, x + s if axis == 1 else x, y + s if axis == 2 else y)
            if abs(w) != 1:
                term = abs(w) * term
            acc = (term if w > 0 else -term) if acc is None else (acc + term if w > 0 else acc - term)
        return acc / divisor(eps)

    def enhanced(t, x, y, *, d=(0, 0, 0), op=None, small_eps=None):
        eps = small_eps or default_eps
        if op is not None:
            if op.lower() not in ("laplacian", "lap"):
                raise ValueError(f"Unknown operator: {op}. Use 'laplacian'/'lap'")
            return (fn(t, x + eps, y) + fn(t, x - eps, y) + fn(t, x, y + eps) + fn(t, x, y - eps)
                    - 4 * fn(t, x, y)) / (eps * eps)
        nt, nx, ny = d
        if any(k not in (0, 1, 2) for k in (nt, nx, ny)):
            raise ValueError("Individual derivatives must be 0, 1, or 2")
        if nt + nx + ny > 2:
            raise ValueError("Combined derivative order must be 0, 1, or 2")
        if nt:  # (a time derivative takes precedence over spatial ones, as in the reference)
            return line(t, x, y, 0, (_FD_FIRST if nt == 1 else _FD_SECOND)[time_stepping], eps)
        if nx == 1 and ny == 1:
            return (fn(t, x + eps, y + eps) - fn(t, x + eps, y - eps) - fn(t, x - eps, y + eps)
                    + fn(t, x - eps, y - eps)) / (4 * eps * eps)
        if nx:
            return line(t, x, y, 1, (_FD_FIRST if nx == 1 else _FD_SECOND)["center"], eps)
        if ny:
            return line(t, x, y, 2, (_FD_FIRST if ny == 1 else _FD_SECOND)["center"], eps)
        return fn(t, x, y)

    return enhanced


class MMSCaseFromAnalytic(MMSCaseBase):
    """Exact solution given as five Python callables `f(t, xx, yy)` (reference :1034-1155); the derivatives the
    forcing needs come from `pack_analytical_txy_with_o2fdm_derivatives`."""

    def __init__(self, model, *, grid, cp_base, T_base, cl_base, cd_base, cs_base):
        super().__init__(grid, model)
        self.cp_ex, self.T_ex, self.cl_ex, self.cd_ex, self.cs_ex = (
            pack_analytical_txy_with_o2fdm_derivatives(f) for f in (cp_base, T_base, cl_base, cd_base, cs_base))


def _make_analytic_method(var, kw):
    def method(self, t, xx, yy):
        return getattr(self, var + "_ex")(t, xx, yy, **kw)
    return method


for _v in VARS:
    setattr(MMSCaseFromAnalytic, _v, _make_analytic_method(_v, {}))
    for _k, _d in (("dt", (1, 0, 0)), ("dx", (0, 1, 0)), ("dy", (0, 0, 1))):
        setattr(MMSCaseFromAnalytic, f"{_k}_{_v}", _make_analytic_method(_v, {"d": _d}))
for _v in ("T", "cl", "cd"):
    setattr(MMSCaseFromAnalytic, f"lap_{_v}", _make_analytic_method(_v, {"op": "lap"}))


# ----------------------------------------------------------------------------
# state (reference src/prob1base.py:1913-2085)
# ----------------------------------------------------------------------------

def _lazy(name, fn):
    cache = "_cache_" + name

    def getter(self):
        try:
            return object.__getattribute__(self, cache)
        except AttributeError:
            val = fn(self)
            object.__setattr__(self, cache, val)
            return val
    return property(getter, doc=f"lazily evaluated immutable property: {name}")


class StateVars:
    """Immutable bundle of the five fields with lazily cached derived fields.  The derived fields
    exist for API compatibility; on the device they are in-register temporaries of the kernels."""

    _COMPUTED = {
        "MxT": lambda s: Mx_reg(s.T), "MyT": lambda s: My_reg(s.T),
        "Mxcp": lambda s: Mx_reg(s.cp), "Mycp": lambda s: My_reg(s.cp),
        "DmxT": lambda s: Dx_reg(s.T, s.hh), "DmyT": lambda s: Dy_reg(s.T, s.kk),
        "Dmxcl": lambda s: Dx_reg(s.cl, s.hh), "Dmycl": lambda s: Dy_reg(s.cl, s.kk),
        "Dmxcd": lambda s: Dx_reg(s.cd, s.hh), "Dmycd": lambda s: Dy_reg(s.cd, s.kk),
        "Dl_Mxcp": lambda s: s._model.Dl(s.Mxcp), "Dl_Mycp": lambda s: s._model.Dl(s.Mycp),
        "dDl_Mxcp": lambda s: s._model.Dl(s.Mxcp, d=1), "dDl_Mycp": lambda s: s._model.Dl(s.Mycp, d=1),
        "V1T": lambda s: s._model.V1(s.T), "V2T": lambda s: s._model.V2(s.T),
        "dV1T": lambda s: s._model.V1(s.T, d=1), "dV2T": lambda s: s._model.V2(s.T, d=1),
        "Dd_MxcpT": lambda s: s._model.Dd(s.Mxcp, s.MxT), "Dd_MycpT": lambda s: s._model.Dd(s.Mycp, s.MyT),
        "delcp_Dd_MxcpT": lambda s: s._model.Dd(s.Mxcp, s.MxT, d=(1, 0)),
        "delcp_Dd_MycpT": lambda s: s._model.Dd(s.Mycp, s.MyT, d=(1, 0)),
        "delT_Dd_MxcpT": lambda s: s._model.Dd(s.Mxcp, s.MxT, d=(0, 1)),
        "delT_Dd_MycpT": lambda s: s._model.Dd(s.Mycp, s.MyT, d=(0, 1)),
    }
    _COMPUTED_PROPERTIES = _COMPUTED

    def __init__(self, cp, T, cl, cd, cs, *, model, hh, kk):
        for name, val in zip(VARS, (cp, T, cl, cd, cs)):
            object.__setattr__(self, f"_{name}_data", np.asarray(val))
        object.__setattr__(self, "_model", model)
        object.__setattr__(self, "_hh", hh)
        object.__setattr__(self, "_kk", kk)
        object.__setattr__(self, "_initialized", True)

    cp = property(lambda self: self._cp_data)
    T = property(lambda self: self._T_data)
    cl = property(lambda self: self._cl_data)
    cd = property(lambda self: self._cd_data)
    cs = property(lambda self: self._cs_data)
    model = property(lambda self: self._model)
    hh = property(lambda self: self._hh)
    kk = property(lambda self: self._kk)

    def __setattr__(self, name, value):
        if name.startswith("_cache_") or not getattr(self, "_initialized", False):
            super().__setattr__(name, value)
        else:
            raise AttributeError(f"Cannot set attribute '{name}'. '{type(self).__name__}' instance is immutable.")

    def __delattr__(self, name):
        if name.startswith("_cache_") or not getattr(self, "_initialized", False):
            super().__delattr__(name)
        else:
            raise AttributeError(f"Cannot delete attribute '{name}'. '{type(self).__name__}' instance is immutable.")

    def into_dict(self, recipient: dict, which: List[str] = None):
        if which is None:
            which = list(self._COMPUTED) + list(VARS)
        for n in which:
            recipient[n] = getattr(self, n)
        return recipient

    def with_changes(self, **kwargs):
        cur = {v: getattr(self, v) for v in VARS}
        for k, val in kwargs.items():
            if k not in cur:
                raise ValueError(f"{k}: Invalid change. Can only change: {list(VARS)}.")
            cur[k] = val
        return StateVars(cur["cp"], cur["T"], cur["cl"], cur["cd"], cur["cs"], model=self.model, hh=self.hh,
                         kk=self.kk)

    def copy(self):
        return self.with_changes()

    def fields(self) -> Dict[str, np.ndarray]:
        return {v: getattr(self, v) for v in VARS}


for _n, _f in StateVars._COMPUTED.items():
    setattr(StateVars, _n, _lazy(_n, _f))


def state_from_mms_when(*, mms_case, t, grid):
    """Exact state at time t (reference :3433-3449)."""
    xx, yy = grid.xx, grid.yy
    return StateVars(mms_case.cp(t, xx, yy), mms_case.T(t, xx, yy), mms_case.cl(t, xx, yy), mms_case.cd(t, xx, yy),
                     mms_case.cs(t, xx, yy), model=mms_case.model, hh=grid.hh, kk=grid.kk)


def heaviside_regularized(x, regularization_factor: float):
    """H_eta(x) = 1 / (1 + exp(-eta x))  (reference :3452-3466)."""
    return 1 / (1 + np.exp(-regularization_factor * x))


# ----------------------------------------------------------------------------
# forcing terms of the RegHCsTriple family (reference :2296-2378, 3468-3551).
# Host (NumPy) evaluation, used when a caller asks for the source arrays and when
# the device cannot evaluate the manufactured solution itself.
# ----------------------------------------------------------------------------

class ForcingTerms_CsTriple(ForcingTermsBase):
    """MMS sources of the field with [Cs-Cd-int] = Kd (Sd - cd)(1 + cl) cs (reference :2296-2425).  The other
    two variants only differ in F2(cs) (`_F2`)."""

    dd_reaction = "cs"

    def __init__(self, *, mms_case: MMSCaseBase, model: DefaultModel01):
        self._mms_case = mms_case
        self._model = model

    mms_case = property(lambda self: self._mms_case)
    model = property(lambda self: self._model)

    @property
    def grid(self):
        return self._mms_case.grid

    def _F2(self, cs):
        return cs

    def fcp_ptwise(self, t, xx, yy):
        c, m = self.mms_case, self.model
        cp = c.cp(t, xx, yy)
        return c.dt_cp(t, xx, yy) - (-cp * (m.K1 * (1 + c.cl(t, xx, yy)) + m.K2 * c.T(t, xx, yy)))

    def fcp(self, t, xx, yy):
        return avg_int(lambda p, q: self.fcp_ptwise(t, p, q), grid=Grid(xx, yy))

    def fT(self, t, xx, yy):
        c, m = self.mms_case, self.model
        return c.dt_T(t, xx, yy) - (m.DT * c.lap_T(t, xx, yy) - m.K3 * c.cp(t, xx, yy) * c.T(t, xx, yy))

    def fcl(self, t, xx, yy):
        c, m = self.mms_case, self.model
        cp, T, cl = c.cp(t, xx, yy), c.T(t, xx, yy), c.cl(t, xx, yy)
        dxcl, dycl = c.dx_cl(t, xx, yy), c.dy_cl(t, xx, yy)
        return c.dt_cl(t, xx, yy) - (
            m.Dl(cp, d=1) * (c.dx_cp(t, xx, yy) * dxcl + c.dy_cp(t, xx, yy) * dycl)
            + m.Dl(cp) * c.lap_cl(t, xx, yy)
            - m.V1(T) * dxcl - m.V2(T) * dycl
            - (cl + 1) * (m.V1(T, d=1) * c.dx_T(t, xx, yy) + m.V2(T, d=1) * c.dy_T(t, xx, yy))
            - m.K4 * cp * (cl + 1))

    def fcd(self, t, xx, yy):
        c, m = self.mms_case, self.model
        cp, T = c.cp(t, xx, yy), c.T(t, xx, yy)
        dC, dT = m.Dd(cp, T, d=(1, 0)), m.Dd(cp, T, d=(0, 1))
        H = self._F2(c.cs(t, xx, yy))
        return c.dt_cd(t, xx, yy) - (
            (dC * c.dx_cp(t, xx, yy) + dT * c.dx_T(t, xx, yy)) * c.dx_cd(t, xx, yy)
            + (dC * c.dy_cp(t, xx, yy) + dT * c.dy_T(t, xx, yy)) * c.dy_cd(t, xx, yy)
            + m.Dd(cp, T) * c.lap_cd(t, xx, yy)
            + m.Kd * (m.Sd - c.cd(t, xx, yy)) * (c.cl(t, xx, yy) + 1) * H)

    def fcs(self, t, xx, yy):
        c, m = self.mms_case, self.model
        H = self._F2(c.cs(t, xx, yy))
        return c.dt_cs(t, xx, yy) - (-m.Kd * H * (1 + c.cl(t, xx, yy)) * (m.Sd - c.cd(t, xx, yy)))


class ForcingTerms_HCsTriple(ForcingTerms_CsTriple):
    """F2(cs) = (cs > 0) (reference :3222-3300)."""

    dd_reaction = "h"

    def __init__(self, *, mms_case: MMSCaseBase, model: DefaultModel01):
        super().__init__(mms_case=mms_case, model=model)
        self.cs3_forcing_terms = ForcingTerms_CsTriple(mms_case=mms_case, model=model)

    def _F2(self, cs):
        return cs > 0


class ForcingTerms_RegHCsTriple(ForcingTerms_CsTriple):
    """F2(cs) = H_eta(cs) (reference :3473-3551)."""

    dd_reaction = "regh"

    def __init__(self, *, mms_case: MMSCaseBase, model: DefaultModel01, regularization_factor: float):
        super().__init__(mms_case=mms_case, model=model)
        self._regularization_factor = regularization_factor

    regularization_factor = property(lambda self: self._regularization_factor)

    def _F2(self, cs):
        return heaviside_regularized(cs, self.regularization_factor)

    def fcs(self, t, xx, yy):  # factor order of the reference's RegH class (:3538-3551)
        c, m = self.mms_case, self.model
        H = self._F2(c.cs(t, xx, yy))
        return c.dt_cs(t, xx, yy) - (-m.Kd * (1 + c.cl(t, xx, yy)) * (m.Sd - c.cd(t, xx, yy)) * H)


# ----------------------------------------------------------------------------
# device plumbing shared by the field and the integrators
# ----------------------------------------------------------------------------

_FORCING_NAMES = ("fcp", "fT", "fcl", "fcd", "fcs")


def _model_signature(model, eta, reaction):
    return tuple(float(getattr(model, n)) for n in ModelConsts._fields[2:]) + (float(eta), model.dd_kind, reaction)


class _DeviceBinding:
    """Keeps the grid's device batch configured for one field object: model constants and forcing
    are read at call time (the reference's tests mutate `model.Kd` and rebind `field.fcs`)."""

    def __init__(self, field):
        self.field = field
        self.batch = field.grid.device_batch()

    def _owner(self):
        fns = [getattr(self.field, n) for n in _FORCING_NAMES]
        owners = {id(getattr(f, "__self__", None)) for f in fns}
        if len(owners) != 1:
            return None
        o = getattr(fns[0], "__self__", None)
        if o is None:
            return None
        for n, f in zip(_FORCING_NAMES, fns):
            if getattr(f, "__func__", None) is not getattr(type(o), n, None):
                return None  # rebound / overridden on the instance
        return o

    def configure(self, t0, dt):
        """Upload model constants and select the forcing mode.  Returns True when the device evaluates
        the sources itself, False when the host has to (ARRAYS mode)."""
        b, fld = self.batch, self.field
        sig = _model_signature(fld.model, fld.regularization_factor, fld.dd_reaction)
        if getattr(b, "_model_sig", None) != sig:
            b.set_model(fld.model, fld.regularization_factor, fld.dd_reaction)
            b._model_sig = sig
        o = self._owner()
        if isinstance(o, NoForcingTerms):
            if b.mode != MODE_NONE:
                b.forcing_none()
                b._spec_owner = None
            return True
        if (type(o) in (ForcingTerms_RegHCsTriple, ForcingTerms_CsTriple, ForcingTerms_HCsTriple)
                and o.model is fld.model and o.dd_reaction == fld.dd_reaction
                and float(getattr(o, "regularization_factor", 0.0)) == float(fld.regularization_factor)):
            spec = o.mms_case.device_spec() if hasattr(o.mms_case, "device_spec") else None
            if spec is not None and o.mms_case.model is fld.model:
                if getattr(b, "_spec_owner", None) is not spec:
                    b.forcing_spec(spec, t0, dt)
                    b._spec_owner = spec
                return True
            prog = (o.mms_case.device_program() if spec is None and hasattr(o.mms_case, "device_program")
                    else None)
            if prog is not None and o.mms_case.model is fld.model:
                if getattr(b, "_spec_owner", None) is not prog:
                    b.forcing_program(prog)
                    b._spec_owner = prog
                return True
        # generic path: evaluate the caller's source callables on the host and upload them
        g = fld.grid
        f = {}
        for v, n in zip(VARS, _FORCING_NAMES):
            fn = getattr(fld, n)
            f[v] = (np.asarray(fn(t0, g.xx, g.yy), dtype=np.float64),
                    np.asarray(fn(t0 + dt, g.xx, g.yy), dtype=np.float64))
        b.forcing_arrays(f)
        b._spec_owner = None
        return False


# ----------------------------------------------------------------------------
# semidiscrete field (reference :2133-2293 base, 2429-2839 Field01, 3553-3593 RegH)
# ----------------------------------------------------------------------------

class SemiDiscreteFieldBase(ABC):
    """Name kept for callers that annotate with it (reference :2133; src/mms_trial_utils.py,
    src/cvg_studies_base.py evaluate `p1.SemiDiscreteFieldBase` at import time)."""


class SemiDiscreteField_RegHCsTriple(SemiDiscreteFieldBase):
    """[Cs-Cd-int] = Kd (Sd - cd)(1 + cl) H_eta(cs).  F*(state, t) run on the device."""

    dd_reaction = "regh"

    def __init__(self, *, grid: Grid, model: DefaultModel01, forcing_terms: ForcingTermsBase,
                 regularization_factor: float = 0.0):
        self._model = model
        self._grid = grid
        self._regularization_factor = regularization_factor
        for n in _FORCING_NAMES:
            setattr(self, n, getattr(forcing_terms, n))
        self._binding = None

    model = property(lambda self: self._model)
    grid = property(lambda self: self._grid)
    regularization_factor = property(lambda self: self._regularization_factor)

    def binding(self) -> _DeviceBinding:
        if self._binding is None:
            self._binding = _DeviceBinding(self)
        return self._binding

    # reaction-term helpers kept for API compatibility (:3580-3593)
    def cscd_reaction_T(self):
        return (0, 1)

    def cscd_reaction_cl(self):
        return (1, 1)

    def cscd_reaction_cd(self):
        return (-1, self.model.Sd)

    def cscd_reaction_cp(self, cp):
        return self.grid.const_with_nullbd(1)

    def cscd_reaction_cs(self, cs):
        return self.model.Kd * heaviside_regularized(cs, self.regularization_factor)

    def cscd_reaction_term(self, state: StateVars):
        return ((self.model.Sd - state.cd) * (state.cl + 1) * self.cscd_reaction_cs(state.cs)
                * self.grid.null_bd_mask)

    # derivative of the reaction term at (i, j) with respect to T, cl, cd at (i + a, j + b): a product of affine
    # factors, so only the node itself contributes (host helpers of the reference's API, :2511-2597; the device
    # kernels carry the same factors inside the cd Jacobian rows)
    def _del_reaction(self, state: StateVars, a, b, wrt: str):
        assert a in (-1, 0, 1) and b in (-1, 0, 1) and (a == 0 or b == 0)
        if a != 0 or b != 0:
            return self.grid.make_full0()
        coef = {"T": self.cscd_reaction_T(), "cl": self.cscd_reaction_cl(), "cd": self.cscd_reaction_cd()}
        if coef[wrt][0] == 0.0:
            return self.grid.make_full0()
        out = self.cscd_reaction_cp(state.cp) * self.cscd_reaction_cs(state.cs) * self.grid.null_bd_mask
        for v, (slope, shift) in coef.items():
            out = out * (slope if v == wrt else slope * getattr(state, v) + shift)
        return out

    def delT_ab_cscd_reaction_ij(self, state: StateVars, *, a, b):
        return self._del_reaction(state, a, b, "T")

    def delcl_ab_cscd_reaction_ij(self, state: StateVars, *, a, b):
        return self._del_reaction(state, a, b, "cl")

    def delcd_ab_cscd_reaction_ij(self, state: StateVars, *, a, b):
        return self._del_reaction(state, a, b, "cd")

    def _all_F(self, at_t: StateVars, t: float) -> Dict[str, np.ndarray]:
        bind = self.binding()
        bind.configure(t, 1.0)
        b = bind.batch
        b.upload(0, at_t.fields())
        b.eval_fields(0, 1, t)
        return b.download(1)

    def Fcp(self, at_t, t):
        return self._all_F(at_t, t)["cp"]

    def FT(self, at_t, t):
        return self._all_F(at_t, t)["T"]

    def Fcl(self, at_t, t):
        return self._all_F(at_t, t)["cl"]

    def Fcd(self, at_t, t):
        return self._all_F(at_t, t)["cd"]

    def Fcs(self, at_t, t):
        return self._all_F(at_t, t)["cs"]


class SemiDiscreteField_CsTriple(SemiDiscreteField_RegHCsTriple):
    """[Cs-Cd-int] = Kd (Sd - cd)(1 + cl) cs (reference :2842-2876); same kernels, F2(cs) = cs."""

    dd_reaction = "cs"

    def __init__(self, *, grid: Grid, model: DefaultModel01, forcing_terms: ForcingTermsBase):
        super().__init__(grid=grid, model=model, forcing_terms=forcing_terms, regularization_factor=0.0)

    def cscd_reaction_cs(self, cs):
        return self.model.Kd * cs


class SemiDiscreteField_HCsTriple(SemiDiscreteField_RegHCsTriple):
    """[Cs-Cd-int] = Kd (Sd - cd)(1 + cl) H(cs), H(cs) = (cs > 0) (reference :3303-3340)."""

    dd_reaction = "h"

    def __init__(self, *, grid: Grid, model: DefaultModel01, forcing_terms: ForcingTermsBase):
        super().__init__(grid=grid, model=model, forcing_terms=forcing_terms, regularization_factor=0.0)

    def cscd_reaction_cs(self, cs):
        return self.model.Kd * (cs > 0)


# ----------------------------------------------------------------------------
# time integrators
# ----------------------------------------------------------------------------

class TimeIntegratorBase(ABC):
    @abstractmethod
    def step(self, at_t0: StateVars, *, t0, dt):
        pass


class ForwardEulerIntegrator(TimeIntegratorBase):
    """u1 = u0 + dt F(u0, t0) on all nodes (reference :2885-2903)."""

    def __init__(self, semi_discrete_field, **kwargs):
        self.semi_discrete_field = semi_discrete_field

    def step(self, at_t0: StateVars, *, t0, dt):
        bind = self.semi_discrete_field.binding()
        bind.configure(t0, dt)
        b = bind.batch
        b.upload(0, at_t0.fields())
        b.step_feuler(0, 1, t0, dt)
        return at_t0.with_changes(**b.download(1))


class _LazyResiduals(dict):
    """`last_residual` of the reference (:2943, 3041-3043, 3076-3078, 3111-3113), computed on first access."""

    def __init__(self, compute):
        super().__init__()
        self._compute = compute

    def _fill(self):
        if self._compute is not None:
            c, self._compute = self._compute, None
            self.update(c())

    def __getitem__(self, k):
        self._fill()
        return super().__getitem__(k)

    def __contains__(self, k):
        self._fill()
        return super().__contains__(k)

    def keys(self):
        self._fill()
        return super().keys()

    def items(self):
        self._fill()
        return super().items()

    def __len__(self):
        self._fill()
        return super().__len__()


class P_ModifiedEuler_C_Trapezoidal_TimeIntegratorBase(ABC):
    """Name kept for callers that annotate with it (reference :2906)."""


class P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple(P_ModifiedEuler_C_Trapezoidal_TimeIntegratorBase,
                                                                TimeIntegratorBase):
    """Predictor-corrector / Newton step of the RegHCsTriple family (reference :2906-3149, 3596-3702)."""

    def __init__(self, semi_discrete_field, *, num_pc_steps=1, num_newton_steps=1, regularization_factor: float,
                 num_newton_iterations: int = 5, consec_xs_rtol: float = 1e-6):
        self.semi_discrete_field = semi_discrete_field
        self._model = semi_discrete_field.model
        self._grid = semi_discrete_field.grid
        self.num_pc_steps = num_pc_steps
        self.num_newton_steps = num_newton_steps
        self._regularization_factor = regularization_factor
        self._num_newton_iterations = num_newton_iterations
        self._consec_xs_rtol = consec_xs_rtol
        self.last_residual: Dict = {}
        self.last_stats: Dict = {}
        self.cd_band_swap = True  # reproduce reference :3094-3100

    def _opts(self, **over):
        kw = dict(num_pc_steps=self.num_pc_steps, num_newton_steps=self.num_newton_steps,
                  num_newton_iterations=self._num_newton_iterations, consec_xs_rtol=self._consec_xs_rtol,
                  cd_band_swap=self.cd_band_swap)
        kw.update(over)
        return ddcore.pc_options(**kw)

    def _bind(self, t0, dt):
        assert dt > 0
        bind = self.semi_discrete_field.binding()
        bind.configure(t0, dt)
        return bind.batch

    # -- the fused step ----------------------------------------------------------
    def step(self, at_t0: StateVars, *, t0, dt):
        b = self._bind(t0, dt)
        b.upload(0, at_t0.fields())
        self.last_stats = b.step_pc(0, 1, t0, dt, self._opts())
        out = at_t0.with_changes(**b.download(1))
        self.last_residual = _LazyResiduals(lambda: self._replay_residuals(at_t0, t0, dt))
        return out

    # -- pieces (called directly by the reference's tests) --------------------------
    _cs_pred_masked = True  # RegHCsTriple / HCsTriple multiply the predicted cs by the boundary mask, CsTriple does not

    def _user_F(self, name: str) -> bool:
        """True when `semi_discrete_field.<name>` is user code (a subclass override or a rebound attribute, as the
        reference's tests do with mock fields): an opaque Python callable cannot run on the device, the piece is
        then evaluated with it on the host exactly as the reference composes it."""
        f = self.semi_discrete_field
        return name in vars(f) or getattr(type(f), name, None) is not getattr(SemiDiscreteField_RegHCsTriple, name)

    def _heun(self, F, at_t, t, dt, var):
        F0 = F(at_t, t)
        star = at_t.with_changes(**{var: getattr(at_t, var) + dt * F0})
        return getattr(at_t, var) + (0.5 * dt) * (F0 + F(star, t + dt))

    def initial_cp_pred(self, at_t, t, *, dt):
        if self._user_F("Fcp"):
            return self._heun(self.semi_discrete_field.Fcp, at_t, t, dt, "cp")
        b = self._bind(t, dt)
        b.upload(0, at_t.fields())
        b.pc_predict(0, t, dt)
        return b.work_download("cp1p")

    def initial_cs_pred(self, at_t, t, *, dt):
        if self._user_F("Fcs"):
            cs1 = self._heun(self.semi_discrete_field.Fcs, at_t, t, dt, "cs")
            return cs1 * self._grid.null_bd_mask if self._cs_pred_masked else cs1
        b = self._bind(t, dt)
        b.upload(0, at_t.fields())
        b.pc_predict(0, t, dt)
        return b.work_download("cs1p")

    def _correct(self, T1, cl1, cd1, at_t0, t0, dt):
        b = self._bind(t0, dt)
        b.upload(0, at_t0.fields())
        b.upload(1, {"T": T1, "cl": cl1, "cd": cd1})
        iters = b.pc_correct(0, 1, t0, dt, self._opts())
        self.last_cs_newton_iterations = int(iters[0])
        return b.download(1, which=("cp", "cs"))

    def corrector_cp_step(self, T1, cl1, _cd1_ignroed, *, at_t0, t0, dt):  # (parameter name as in the reference, :2967)
        cd1 = _cd1_ignroed if _cd1_ignroed is not None else at_t0.cd
        return self._correct(T1, cl1, cd1, at_t0, t0, dt)["cp"]

    def corrector_cs_step(self, _T1_ignored, cl1, cd1, *, at_t0, t0, dt):
        T1 = _T1_ignored if _T1_ignored is not None else at_t0.T
        return self._correct(T1, cl1, cd1, at_t0, t0, dt)["cs"]

    def _newton(self, var, ustar: StateVars, new_fields: Dict[str, np.ndarray], t0, dt, Y):
        b = self._bind(t0, dt)
        b.upload(2, ustar.fields())
        if new_fields:
            b.upload(3, new_fields)
        b.work_upload({"T": "YT", "cl": "Ycl", "cd": "Ycd"}[var], Y)
        self.last_stats = b.pc_newton(var, 2, 3, t0, dt, self._opts())
        vnew = b.download(3, which=(var,))[var]
        # residual at the state with the variables updated so far
        b.upload(3, {"cp": ustar.cp, "cs": ustar.cs, **{v: getattr(ustar, v) for v in ("T", "cl", "cd")
                                                         if v not in new_fields and v != var}})
        self.last_residual[var] = b.pc_residual(var, 3, t0, dt)[0]
        return vnew

    def newton_step_T(self, at_t0: StateVars, *, t0, dt, YT0):
        if isinstance(self.last_residual, _LazyResiduals):
            self.last_residual = {}
        return self._newton("T", at_t0, {}, t0, dt, YT0)

    def newton_step_cl(self, at_t0, T1, *, t0, dt, Ycl0):
        if isinstance(self.last_residual, _LazyResiduals):
            self.last_residual = {}
        return self._newton("cl", at_t0, {"T": T1}, t0, dt, Ycl0)

    def newton_step_cd(self, at_t0, T1, cl1, *, t0, dt, Ycd0):
        if isinstance(self.last_residual, _LazyResiduals):
            self.last_residual = {}
        return self._newton("cd", at_t0, {"T": T1, "cl": cl1}, t0, dt, Ycd0)

    def _replay_residuals(self, at_t0, t0, dt):
        """Re-run the step piece by piece to obtain the Newton residuals of its last iteration."""
        saved = self.last_residual
        self.last_residual = {}
        b = self._bind(t0, dt)
        b.upload(0, at_t0.fields())
        b.pc_predict(0, t0, dt)
        Y = {v: b.work_download(n) for v, n in (("T", "YT"), ("cl", "Ycl"), ("cd", "Ycd"))}
        cp1, cs1 = b.work_download("cp1p"), b.work_download("cs1p")
        T1, cl1, cd1 = at_t0.T, at_t0.cl, at_t0.cd
        for _ in range(self.num_pc_steps):
            for _ in range(self.num_newton_steps):
                u = at_t0.with_changes(cp=cp1, T=T1, cl=cl1, cd=cd1, cs=cs1)
                T1 = self.newton_step_T(u, t0=t0, dt=dt, YT0=Y["T"])
                cl1 = self.newton_step_cl(u, T1, t0=t0, dt=dt, Ycl0=Y["cl"])
                cd1 = self.newton_step_cd(u, T1, cl1, t0=t0, dt=dt, Ycd0=Y["cd"])
            c = self._correct(T1, cl1, cd1, at_t0, t0, dt)
            cp1, cs1 = c["cp"], c["cs"]
        out = dict(self.last_residual)
        self.last_residual = saved
        return out


class P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_CsTriple(P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple):
    """The same predictor-corrector / Newton step for the CsTriple field (reference :3152-3219): Heun cs
    predictor without boundary mask, closed-form trapezoidal cs corrector (no Newton iterations)."""

    _cs_pred_masked = False

    def __init__(self, semi_discrete_field, *, num_pc_steps=1, num_newton_steps=1):
        super().__init__(semi_discrete_field, num_pc_steps=num_pc_steps, num_newton_steps=num_newton_steps,
                         regularization_factor=0.0, num_newton_iterations=0, consec_xs_rtol=0.0)


class P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_HCsTriple(P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple):
    """... for the HCsTriple field (reference :3343-3430): masked Heun cs predictor, piecewise closed-form cs
    corrector; raises ValueError when 2 - dt Kd (Sd - cd1)(1 + cl1) is not safely positive."""

    def __init__(self, semi_discrete_field, *, num_pc_steps=1, num_newton_steps=1):
        super().__init__(semi_discrete_field, num_pc_steps=num_pc_steps, num_newton_steps=num_newton_steps,
                         regularization_factor=0.0, num_newton_iterations=0, consec_xs_rtol=0.0)

    def corrector_cs_step(self, _T1_ignored, cl1, cd1, *, at_t0, t0, dt):
        if not self._user_F("Fcs"):
            return super().corrector_cs_step(_T1_ignored, cl1, cd1, at_t0=at_t0, t0=t0, dt=dt)
        # a user-supplied Fcs (reference tests mock it): the same piecewise closed form on the host (:3393-3430) --
        # Y0 = 2 cs0 + dt Fcs(u0, t0) + dt fcs(t1); cs1 = Y0 / (2 - dt R1) where Y0 > 0, Y0 / 2 where Y0 < 0
        field, g, m = self.semi_discrete_field, self._grid, self._model
        tol = np.finfo(float).eps * 100
        denom = 2 - dt * ((m.Sd - cd1) * (1 + cl1) * m.Kd)
        if np.any(denom < tol):
            raise ValueError("Denominator 2 - \u0394t Kd (Sd - Cd1) (1 + Cl1) below positiveness treshold.")
        Y0 = 2 * at_t0.cs + dt * field.Fcs(at_t0, t0) + dt * field.fcs(t0 + dt, g.xx, g.yy)
        cs1 = g.make_full0()
        pos, neg = Y0 > tol, Y0 < -tol
        cs1[pos] = Y0[pos] / denom[pos]
        cs1[neg] = Y0[neg] / 2.0
        return cs1 * g.null_bd_mask
