"""Host-side handle on a device batch (dd_batch) and the MMS table builders.

`Batch` is the NumPy-facing wrapper of the C ABI: it owns B trajectories on one
grid in HBM and exposes the hot path (forward Euler, the RegHCsTriple
predictor-corrector step and its pieces, error norms).  The reference-compatible
class API in `prob1base.py` and the trial harness in `mms_trial_utils.py` are
thin layers over it.
"""

from __future__ import annotations

import ctypes as C
from dataclasses import dataclass, field
from typing import Dict, List, Optional, Sequence

import numpy as np

from _ddlib import DDNotConverged  # noqa: F401
from _ddlib import (DD_OK, MODE_ARRAYS, MODE_EXPSIN, MODE_NONE, MODE_PROGRAM, MODE_SEPARABLE, PHI_CONST, PHI_EXP, PHI_HOST,
                    PHI_INV1PT, PHI_LINEAR, PHI_OSC, VARS, Context, as_f64, dd_model, dd_pc_options,
                    dd_step_stats, dptr, _dp, _vp)

GAUSS_NODES = np.array([-np.sqrt(3.0 / 5.0), 0, np.sqrt(3.0 / 5.0)])  # reference src/prob1base.py:509


# ----------------------------------------------------------------------------
# device-evaluable descriptions of a manufactured solution
# ----------------------------------------------------------------------------

@dataclass
class PhiSpec:
    """Time profile phi(t).  kind in {inv1pt, exp, linear, osc, const, host};
    `host` carries callables f(t), f'(t) evaluated by the host before each step."""
    kind: str
    p: Sequence[float] = (0.0, 0.0, 0.0, 0.0)
    f: Optional[callable] = None
    df: Optional[callable] = None

    _KINDS = {"inv1pt": PHI_INV1PT, "exp": PHI_EXP, "linear": PHI_LINEAR, "osc": PHI_OSC, "const": PHI_CONST,
              "host": PHI_HOST}

    def code(self):
        return self._KINDS[self.kind]

    def params(self, t0=None, dt=None):
        if self.kind == "host":
            t1 = t0 + dt
            return [float(self.f(t0)), float(self.df(t0)), float(self.f(t1)), float(self.df(t1))]
        p = list(self.p) + [0.0] * 4
        return [float(v) for v in p[:4]]

    def value(self, t):
        k, p = self.kind, list(self.p) + [0.0] * 4
        if k == "inv1pt":
            return p[0] / (1.0 + t)
        if k == "exp":
            return p[0] * np.exp(-p[1] * t)
        if k == "linear":
            return p[0] - p[1] * t
        if k == "osc":
            return p[0] * (1.0 + p[1] * np.sin(p[2] * t))
        if k == "const":
            return p[0]
        return float(self.f(t))


@dataclass
class SeparableSpec:
    """u_v(t,x,y) = phi_v(t) sum_r X_{v,r}(x) Y_{v,r}(y) for v in cp, T, cl, cd, cs.
    X[v][r] / Y[v][r] are triples of callables (value, first, second derivative) of a 1-D array;
    every variable carries the same number of terms (pad with zero terms)."""
    phi: List[PhiSpec]
    X: List[Sequence[Sequence[callable]]]
    Y: List[Sequence[Sequence[callable]]]
    mode: int = MODE_SEPARABLE


@dataclass
class ExpSinSpec:
    """MMSCaseExpSin (reference src/prob1_mms_cases.py:296-337): evaluated in closed form on the device
    from the member's model constants; no parameters of its own."""
    mode: int = MODE_EXPSIN


def quadrature_points(c: np.ndarray) -> np.ndarray:
    """Gauss abscissae of the dual cells [c_{i-1/2}, c_{i+1/2}], shape (n+1, 3); rows 0 and n unused.
    Same floating-point expression as the reference's avg_int (src/prob1base.py:538-570)."""
    n = len(c) - 1
    h = np.concatenate([[np.inf], c[1:] - c[:-1]])
    hp = np.concatenate([(h[:-1] + h[1:]) * 0.5, [np.inf]])
    ph = np.zeros_like(c)
    ph[:-1] = 0.5 * (c[:-1] + c[1:])
    q = np.zeros((n + 1, 3))
    for a, node in enumerate(GAUSS_NODES):
        q[1:n, a] = ph[0:n - 1] + (node + 1.0) * 0.5 * hp[1:n]
    return q


def _eval1d(f, c):
    r = np.asarray(f(c), dtype=np.float64)
    if r.size == 1:
        r = np.full(np.shape(c), float(r.reshape(-1)[0]))
    return np.ascontiguousarray(r.reshape(np.shape(c)), dtype=np.float64)


# ----------------------------------------------------------------------------
# options / stats
# ----------------------------------------------------------------------------

def pc_options(num_pc_steps=1, num_newton_steps=1, num_newton_iterations=5, consec_xs_rtol=1e-6,
               cd_band_swap=True, solve_tol=1e-14, max_sweeps=20000, fixed_sweeps=0,
               extrapolate_guess=False) -> dd_pc_options:
    o = dd_pc_options()
    o.num_pc_steps = int(num_pc_steps)
    o.num_newton_steps = int(num_newton_steps)
    o.num_newton_iterations = int(num_newton_iterations)
    o.cd_band_swap = 1 if cd_band_swap else 0
    o.consec_xs_rtol = float(consec_xs_rtol)
    o.solve_tol = float(solve_tol)
    o.max_sweeps = int(max_sweeps)
    o.fixed_sweeps = int(fixed_sweeps)
    o.extrapolate_guess = 1 if extrapolate_guess else 0
    return o


REACTIONS = {"regh": 0, "cs": 1, "h": 2}  # F2(cs) = H_eta(cs) | cs | (cs > 0): RegHCsTriple, CsTriple, HCsTriple


def model_struct(model, eta: float, reaction="regh") -> dd_model:
    """dd_model from any object with the ModelConsts attributes (reference src/prob1base.py:28-45).
    `kind` comes from `model.dd_kind` (2 for DefaultModel02, else 1); `reaction` selects the field variant."""
    m = dd_model()
    for n in ("K1", "K2", "K3", "K4", "DT", "Dl_max", "phi_l", "gamma_T", "Kd", "Sd", "Dd_max", "phi_d", "phi_T",
              "r_sp", "T_ref"):
        setattr(m, n, float(getattr(model, n)))
    m.eta = float(eta)
    m.kind = int(getattr(model, "dd_kind", 1))
    m.reaction = REACTIONS[reaction] if isinstance(reaction, str) else int(reaction)
    return m


# ----------------------------------------------------------------------------
# the batch
# ----------------------------------------------------------------------------

class Batch:
    """B trajectories on one (N+1) x (M+1) grid, resident in HBM.

    row0 / nrows / own select a row slab of the global grid (domain decomposition);
    by default the batch holds the whole grid.
    """

    def __init__(self, x, y, nmembers: int = 1, *, ctx: Optional[Context] = None, nslots: int = 3,
                 row0: int = 0, nrows: Optional[int] = None, own: Optional[Sequence[int]] = None):
        self.ctx = ctx or Context.default()
        self.lib = self.ctx.lib
        self.x, self.y = as_f64(x), as_f64(y)
        self.N, self.M = len(self.x) - 1, len(self.y) - 1
        self.B = int(nmembers)
        self.row0 = int(row0)
        self.nrows = int(nrows) if nrows is not None else self.N + 1 - self.row0
        self.own = (0, self.nrows) if own is None else (int(own[0]), int(own[1]))
        self.nslots = nslots
        self.shape = (self.nrows, self.M + 1)
        h = _vp()
        rc = self.lib.dd_batch_create(self.ctx.handle, self.N, self.M, dptr(self.x), dptr(self.y), self.B,
                                      self.row0, self.nrows, self.own[0], self.own[1], nslots, C.byref(h))
        self.ctx.check(rc, "dd_batch_create")
        self.handle = h
        self.mode = MODE_NONE
        self.spec = None
        self._keep = []

    def close(self):
        if getattr(self, "handle", None):
            self.lib.dd_batch_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass

    # -- parameters ----------------------------------------------------------
    def set_models(self, models: Sequence[dd_model], first: int = 0):
        arr = (dd_model * len(models))(*models)
        self.ctx.check(self.lib.dd_batch_set_models(self.handle, first, len(models), arr), "set_models")

    def set_model(self, model, eta: float, reaction="regh"):
        """same model for every member"""
        self.set_models([model_struct(model, eta, reaction)] * self.B)

    def set_active(self, active: Sequence[int], first: int = 0):
        a = (C.c_int * len(active))(*[int(v) for v in active])
        self.ctx.check(self.lib.dd_batch_set_active(self.handle, first, len(active), a), "set_active")

    # -- forcing -----------------------------------------------------------------
    def forcing_none(self):
        self.ctx.check(self.lib.dd_forcing_none(self.handle), "forcing_none")
        self.mode, self.spec = MODE_NONE, None

    def forcing_spec(self, spec, t0: float = 0.0, dt: float = 1.0):
        """Select fused MMS forcing from a SeparableSpec / ExpSinSpec."""
        if isinstance(spec, ExpSinSpec):
            px, py = quadrature_points(self.x), quadrature_points(self.y)
            tabs = [np.sin(np.pi * self.x), np.cos(np.pi * self.x), np.sin(np.pi * self.y), np.cos(np.pi * self.y),
                    np.sin(np.pi * px).reshape(-1), np.sin(np.pi * py).reshape(-1)]
            tabs = [as_f64(t) for t in tabs]
            self.ctx.check(self.lib.dd_forcing_expsin(self.handle, *[dptr(t) for t in tabs]), "forcing_expsin")
            self.mode, self.spec = MODE_EXPSIN, spec
            return
        assert isinstance(spec, SeparableSpec)
        px, py = quadrature_points(self.x), quadrature_points(self.y)
        R = len(spec.X[0])
        assert all(len(spec.X[v]) == R and len(spec.Y[v]) == R for v in range(5))
        keep = []
        Xp = ((_dp * 3) * 5)()
        Yp = ((_dp * 3) * 5)()
        for v in range(5):
            for d in range(3):
                ax = as_f64(np.stack([_eval1d(spec.X[v][r][d], self.x) for r in range(R)]))
                ay = as_f64(np.stack([_eval1d(spec.Y[v][r][d], self.y) for r in range(R)]))
                keep += [ax, ay]
                Xp[v][d] = dptr(ax)
                Yp[v][d] = dptr(ay)
        XQ = (_dp * 3)()
        YQ = (_dp * 3)()
        for q, v in enumerate((0, 1, 2)):  # cp, T, cl
            ax = as_f64(np.stack([_eval1d(spec.X[v][r][0], px.reshape(-1)) for r in range(R)]))
            ay = as_f64(np.stack([_eval1d(spec.Y[v][r][0], py.reshape(-1)) for r in range(R)]))
            keep += [ax, ay]
            XQ[q] = dptr(ax)
            YQ[q] = dptr(ay)
        kinds = (C.c_int * 5)(*[p.code() for p in spec.phi])
        pp = ((C.c_double * 4) * 5)()
        for v in range(5):
            vals = spec.phi[v].params(t0, dt)
            for k in range(4):
                pp[v][k] = vals[k]
        self.ctx.check(self.lib.dd_forcing_separable(self.handle, R, C.byref(Xp), C.byref(Yp), C.byref(XQ),
                                                     C.byref(YQ), C.byref(kinds), C.byref(pp)),
                       "forcing_separable")
        self.mode, self.spec = MODE_SEPARABLE, spec

    def forcing_program(self, program):
        """Select generated forcing: `program` is a ddprogram.ProgramSpec (NVRTC image of the case's expressions)."""
        xq = as_f64(quadrature_points(self.x).reshape(-1))
        yq = as_f64(quadrature_points(self.y).reshape(-1))
        self.ctx.check(self.lib.dd_forcing_program(self.handle, program.image, len(program.image), dptr(xq),
                                                   dptr(yq)), "forcing_program")
        self.mode, self.spec = MODE_PROGRAM, program

    def refresh_host_phi(self, t0: float, dt: float):
        """phi kinds evaluated on the host need their four values renewed for each (t0, dt)."""
        if self.mode != MODE_SEPARABLE or not any(p.kind == "host" for p in self.spec.phi):
            return
        kinds = (C.c_int * (5 * self.B))(*([p.code() for p in self.spec.phi] * self.B))
        vals = []
        for p in self.spec.phi:
            vals += p.params(t0, dt)
        arr = as_f64(np.tile(np.array(vals), self.B))
        self.ctx.check(self.lib.dd_forcing_set_phi(self.handle, 0, self.B, kinds, dptr(arr)), "set_phi")

    def set_phi(self, kinds: np.ndarray, params: np.ndarray, first: int = 0):
        """per-member time profiles: kinds (count, 5) int, params (count, 5, 4)"""
        kinds = np.ascontiguousarray(kinds, dtype=np.int32)
        params = as_f64(params)
        n = kinds.shape[0]
        self.ctx.check(self.lib.dd_forcing_set_phi(self.handle, first, n,
                                                   kinds.ctypes.data_as(C.POINTER(C.c_int)), dptr(params)),
                       "set_phi")

    def forcing_arrays(self, f: Dict[str, Sequence[Optional[np.ndarray]]], member: int = 0):
        """host-evaluated sources: f[var] = (at t0, at t1); missing / None = zero"""
        arr = ((_dp * 2) * 5)()
        keep = []
        for v, name in enumerate(VARS):
            pair = f.get(name, (None, None))
            for s in range(2):
                if pair[s] is not None:
                    a = as_f64(pair[s])
                    assert a.shape == self.shape, f"forcing {name}: shape {a.shape} != {self.shape}"
                    keep.append(a)
                    arr[v][s] = dptr(a)
        self.ctx.check(self.lib.dd_forcing_arrays(self.handle, member, C.byref(arr)), "forcing_arrays")
        self.mode, self.spec = MODE_ARRAYS, None

    # -- state transfer ----------------------------------------------------------
    def upload(self, slot: int, fields: Dict[str, np.ndarray], member: int = 0):
        arr = (_dp * 5)()
        keep = []
        for v, name in enumerate(VARS):
            if name in fields and fields[name] is not None:
                a = as_f64(fields[name])
                assert a.shape == self.shape, f"{name}: shape {a.shape} != {self.shape}"
                keep.append(a)
                arr[v] = dptr(a)
        self.ctx.check(self.lib.dd_state_upload(self.handle, slot, member, C.byref(arr)), "state_upload")

    def download(self, slot: int, member: int = 0, which: Sequence[str] = VARS) -> Dict[str, np.ndarray]:
        arr = (_dp * 5)()
        out = {}
        for v, name in enumerate(VARS):
            if name in which:
                out[name] = np.empty(self.shape, dtype=np.float64)
                arr[v] = dptr(out[name])
        self.ctx.check(self.lib.dd_state_download(self.handle, slot, member, C.byref(arr)), "state_download")
        return out

    def download_into(self, slot: int, out: Dict[str, np.ndarray], member: int = 0):
        """device -> caller-provided host arrays (e.g. pinned memory)"""
        arr = (_dp * 5)()
        for v, name in enumerate(VARS):
            if name in out:
                assert out[name].shape == self.shape and out[name].flags["C_CONTIGUOUS"]
                arr[v] = dptr(out[name])
        self.ctx.check(self.lib.dd_state_download(self.handle, slot, member, C.byref(arr)), "state_download")

    def work_upload(self, name: str, a: np.ndarray, member: int = 0):
        a = as_f64(a)
        assert a.shape == self.shape
        self.ctx.check(self.lib.dd_work_upload(self.handle, name.encode(), member, dptr(a)), "work_upload")

    def work_download(self, name: str, member: int = 0) -> np.ndarray:
        a = np.empty(self.shape, dtype=np.float64)
        self.ctx.check(self.lib.dd_work_download(self.handle, name.encode(), member, dptr(a)), "work_download")
        return a

    def fill_exact(self, slot: int, t):
        t = as_f64(np.atleast_1d(t))
        self.ctx.check(self.lib.dd_state_fill_exact(self.handle, slot, dptr(t), len(t)), "fill_exact")

    def dev_ptr(self, slot: int, var: str):
        p, ms, ld = _vp(), C.c_longlong(), C.c_int()
        self.ctx.check(self.lib.dd_state_dev_ptr(self.handle, slot, VARS.index(var), C.byref(p), C.byref(ms),
                                                 C.byref(ld)), "dev_ptr")
        return p.value, ms.value, ld.value

    def work_dev_ptr(self, name: str):
        p = _vp()
        self.ctx.check(self.lib.dd_work_dev_ptr(self.handle, name.encode(), C.byref(p)), "work_dev_ptr")
        return p.value

    # -- hot path -----------------------------------------------------------------
    def _times(self, t0, dt):
        t0 = as_f64(np.atleast_1d(t0))
        dt = as_f64(np.atleast_1d(dt))
        n = max(len(t0), len(dt))
        if len(t0) != n:
            t0 = as_f64(np.full(n, t0[0]))
        if len(dt) != n:
            dt = as_f64(np.full(n, dt[0]))
        if n == 1:
            self.refresh_host_phi(float(t0[0]), float(dt[0]))
        return t0, dt, n

    def step_feuler(self, slot_in: int, slot_out: int, t0, dt):
        t0, dt, n = self._times(t0, dt)
        self.ctx.check(self.lib.dd_step_feuler(self.handle, slot_in, slot_out, dptr(t0), dptr(dt), n), "step_feuler")

    def step_pc(self, slot_in: int, slot_out: int, t0, dt, opt: Optional[dd_pc_options] = None,
                defer: bool = False) -> Optional[dict]:
        """One PC step.  defer=True: the step is only enqueued and the statistics of the previous deferred step
        are returned (None for the first); rotate three slots and call flush() at the end (dd_b200.h)."""
        t0, dt, n = self._times(t0, dt)
        opt = opt or pc_options()
        st = dd_step_stats()
        if defer:
            have = C.c_int(0)
            self.ctx.check(self.lib.dd_step_pc_deferred(self.handle, slot_in, slot_out, dptr(t0), dptr(dt), n,
                                                        C.byref(opt), C.byref(st), C.byref(have)), "step_pc_deferred")
            return st.as_dict() if have.value else None
        self.ctx.check(self.lib.dd_step_pc(self.handle, slot_in, slot_out, dptr(t0), dptr(dt), n, C.byref(opt),
                                           C.byref(st)), "step_pc")
        return st.as_dict()

    def flush(self) -> Optional[dict]:
        """Settles a pending deferred step; returns its statistics (None when nothing was pending)."""
        st, have = dd_step_stats(), C.c_int(0)
        self.ctx.check(self.lib.dd_step_pc_flush(self.handle, C.byref(st), C.byref(have)), "step_pc_flush")
        return st.as_dict() if have.value else None

    def run_pc(self, slot_a: int, slot_b: int, t0, dt, nsteps: int, opt: Optional[dd_pc_options] = None,
               norms: bool = False):
        """nsteps steps with device-side time advance; returns (final slot, norms or None, stats)"""
        t0, dt, n = self._times(t0, dt)
        if self.mode == MODE_SEPARABLE and any(p.kind == "host" for p in self.spec.phi):
            raise ValueError("run_pc needs device-evaluable time profiles (phi kind != host)")
        opt = opt or pc_options()
        st = dd_step_stats()
        out = np.zeros((nsteps + 1, self.B, 8)) if norms else None
        self.ctx.check(self.lib.dd_run_pc(self.handle, slot_a, slot_b, dptr(t0), dptr(dt), n, nsteps, C.byref(opt),
                                          dptr(out) if norms else None, C.byref(st)), "run_pc")
        return (slot_a if nsteps % 2 == 0 else slot_b), out, st.as_dict()

    def run_errors(self, slot_a: int, slot_b: int, t0, dt, nsteps: int, opt: Optional[dd_pc_options] = None,
                   integrator: str = "pc"):
        """nsteps steps with the error norms of every step taken and combined on the device: returns
        dict(overall (B,), per_var (B, 5)) = the combined max-integral norms of the reference
        (src/mms_trial_utils.py:15-53, 150-190) and, for "pc", the step statistics."""
        t0, dt, n = self._times(t0, dt)
        if self.mode == MODE_SEPARABLE and any(p.kind == "host" for p in self.spec.phi):
            raise ValueError("run_errors needs device-evaluable time profiles (phi kind != host)")
        out = np.zeros((self.B, 6))
        st = None
        if integrator == "pc":
            opt = opt or pc_options()
            st = dd_step_stats()
            self.ctx.check(self.lib.dd_run_pc_errors(self.handle, slot_a, slot_b, dptr(t0), dptr(dt), n, nsteps,
                                                     C.byref(opt), dptr(out), C.byref(st)), "run_pc_errors")
            st = st.as_dict()
        else:
            self.ctx.check(self.lib.dd_run_feuler_errors(self.handle, slot_a, slot_b, dptr(t0), dptr(dt), n, nsteps,
                                                         dptr(out)), "run_feuler_errors")
        return dict(overall=out[:, 0].copy(), per_var=out[:, 1:].copy()), st

    def run_feuler(self, slot_a: int, slot_b: int, t0, dt, nsteps: int, norms: bool = False):
        t0, dt, n = self._times(t0, dt)
        if self.mode == MODE_SEPARABLE and any(p.kind == "host" for p in self.spec.phi):
            raise ValueError("run_feuler needs device-evaluable time profiles (phi kind != host)")
        out = np.zeros((nsteps + 1, self.B, 8)) if norms else None
        self.ctx.check(self.lib.dd_run_feuler(self.handle, slot_a, slot_b, dptr(t0), dptr(dt), n, nsteps,
                                              dptr(out) if norms else None), "run_feuler")
        return (slot_a if nsteps % 2 == 0 else slot_b), out

    # -- pieces ---------------------------------------------------------------------
    def eval_fields(self, slot_in: int, slot_out: int, t):
        t = as_f64(np.atleast_1d(t))
        if len(t) == 1:
            self.refresh_host_phi(float(t[0]), 1.0)
        self.ctx.check(self.lib.dd_eval_fields(self.handle, slot_in, slot_out, dptr(t), len(t)), "eval_fields")

    def pc_predict(self, slot_in: int, t0, dt):
        t0, dt, n = self._times(t0, dt)
        self.ctx.check(self.lib.dd_pc_predict(self.handle, slot_in, dptr(t0), dptr(dt), n), "pc_predict")

    def pc_newton(self, var: str, slot_star: int, slot_new: int, t0, dt, opt=None) -> dict:
        t0, dt, n = self._times(t0, dt)
        opt = opt or pc_options()
        st = dd_step_stats()
        self.ctx.check(self.lib.dd_pc_newton(self.handle, VARS.index(var), slot_star, slot_new, dptr(t0), dptr(dt), n,
                                             C.byref(opt), C.byref(st)), "pc_newton")
        return st.as_dict()

    def pc_correct(self, slot0: int, slot_new: int, t0, dt, opt=None) -> np.ndarray:
        t0, dt, n = self._times(t0, dt)
        opt = opt or pc_options()
        iters = (C.c_int * self.B)()
        self.ctx.check(self.lib.dd_pc_correct(self.handle, slot0, slot_new, dptr(t0), dptr(dt), n, C.byref(opt),
                                              iters), "pc_correct")
        return np.array(list(iters))

    def pc_residual(self, var: str, slot_state: int, t0, dt) -> np.ndarray:
        t0, dt, n = self._times(t0, dt)
        out = np.empty((self.B,) + self.shape, dtype=np.float64)
        self.ctx.check(self.lib.dd_pc_residual(self.handle, VARS.index(var), slot_state, dptr(t0), dptr(dt), n,
                                               dptr(out)), "pc_residual")
        return out

    def error_norms(self, slot: int, t=None, slot_exact: int = -1) -> np.ndarray:
        """(B, 8): H2[cp,T,cl,cd,cs], P2[T,cl,cd] of state - exact"""
        out = np.zeros((self.B, 8))
        if slot_exact >= 0:
            self.ctx.check(self.lib.dd_error_norms(self.handle, slot, slot_exact, None, 0, dptr(out)), "error_norms")
        else:
            t = as_f64(np.atleast_1d(t))
            if len(t) == 1:
                self.refresh_host_phi(float(t[0]), 1.0)
            self.ctx.check(self.lib.dd_error_norms(self.handle, slot, -1, dptr(t), len(t), dptr(out)), "error_norms")
        return out
