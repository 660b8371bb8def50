"""Trial harness with the reference's `mms_trial_utils` API (src/mms_trial_utils.py), written from
scratch on the device path.

`MMSTrial.run_for_errors` keeps the reference's signature and result type.  When the case is
device-evaluable (every case of `prob1_mms_cases`) the whole time loop, the exact solution and the
error norms run on the GPU and only 8 scalars per step come back; otherwise it falls back to
stepping through the class API (host-evaluated sources, host norms) exactly like the reference.
"""

from __future__ import annotations

import math
from typing import Dict, List, NamedTuple, Optional, Tuple, Type

import numpy as np

import ddcore
import prob1base as p1

VARS = ("cp", "T", "cl", "cd", "cs")


class ErrorTimeSeries(NamedTuple):
    t: float
    h_norm_sq_errors: Dict[str, float]
    grad_h_norm_p_sq_errors: Dict[str, float]


def calculate_combined_error_norm(time_series_data, dt: float, integral_vars: List[str],
                                  all_variables: Optional[List[str]] = None) -> float:
    """max_k [ sum_v |e_v(t_k)|_H^2 + int_0^{t_k} sum_{v in integral_vars} |grad e_v|_p^2 ] ^ 1/2 with the
    trapezoidal rule in time (reference src/mms_trial_utils.py:15-53).  A NaN never replaces the running
    maximum, as in the reference's `max(0.0, nan)`."""
    if all_variables is not None:
        assert all(v in all_variables for v in integral_vars), "integral_vars must be a subset of all_variables."
    integrand = [sum(s.grad_h_norm_p_sq_errors[v] for v in integral_vars) for s in time_series_data]
    best, run = 0.0, 0.0
    for k, s in enumerate(time_series_data):
        if all_variables is None:
            hsq = sum(s.h_norm_sq_errors.values())
        else:
            hsq = sum(s.h_norm_sq_errors[v] for v in all_variables)
        if k > 0:
            run += 0.5 * dt * (integrand[k - 1] + integrand[k])
        best = max(best, hsq + run)
    return np.sqrt(best)


def _series_from_norms(times, norms, variable_names, integral_vars) -> List[ErrorTimeSeries]:
    out = []
    for t, row in zip(times, norms):
        h = {v: float(row[VARS.index(v)]) for v in variable_names}
        g = {v: (float(row[5 + ("T", "cl", "cd").index(v)]) if v in integral_vars and v in ("T", "cl", "cd") else 0.0)
             for v in variable_names}
        out.append(ErrorTimeSeries(t=t, h_norm_sq_errors=h, grad_h_norm_p_sq_errors=g))
    return out


def run_simulation_collect_data(*, grid, integrator, exact_sol_pack, initial_state, Tf: float, dt: float,
                                t0: float = 0.0, variable_names: List[str], integral_vars: List[str]
                                ) -> Tuple[List[ErrorTimeSeries], float]:
    """Time loop + per-step error norms (reference src/mms_trial_utils.py:56-147)."""
    num_steps = math.ceil((Tf - t0) / dt)
    dt = (Tf - t0) / num_steps
    fast = _device_trial(grid, integrator, exact_sol_pack, initial_state, t0, dt, num_steps, variable_names,
                         integral_vars)
    if fast is not None:
        return fast, dt
    # generic path: class API per step, norms on the host (same arithmetic as the reference harness)
    xx, yy = grid.xx, grid.yy

    def collect(state, t):
        ex = p1.state_from_mms_when(mms_case=exact_sol_pack, t=t, grid=grid)
        h, g = {}, {}
        for v in variable_names:
            num, exa = getattr(state, v), getattr(ex, v)
            h[v] = grid.norm_H(num - exa) ** 2
            if v in integral_vars:
                gn, ge = grid.grad_H(num), grid.grad_H(exa)
                g[v] = grid.norm_p(gn[0] - ge[0], gn[1] - ge[1]) ** 2
            else:
                g[v] = 0.0
        return ErrorTimeSeries(t=t, h_norm_sq_errors=h, grad_h_norm_p_sq_errors=g)

    t, state = t0, initial_state
    series = [collect(state, t)]
    for _ in range(num_steps):
        state = integrator.step(state, t0=t, dt=dt)
        t += dt
        series.append(collect(state, t))
    assert np.isclose(t, Tf), f"Final time mismatch: current_t={t}, Tf={Tf}"
    return series, dt


def _device_trial(grid, integrator, case, initial_state, t0, dt, num_steps, variable_names, integral_vars):
    """Whole trial on the device when the integrator is one of ours and the case is device-evaluable."""
    field = getattr(integrator, "semi_discrete_field", None)
    if not isinstance(field, p1.SemiDiscreteField_RegHCsTriple) or field.grid is not grid:
        return None
    if not set(integral_vars) <= {"T", "cl", "cd"}:
        return None
    bind = field.binding()
    if not bind.configure(t0, dt):
        return None
    b = bind.batch
    if b.mode not in (ddcore.MODE_SEPARABLE, ddcore.MODE_EXPSIN, ddcore.MODE_PROGRAM):
        return None
    owner = bind._owner()
    if owner is None or getattr(owner, "mms_case", None) is not case:
        return None
    if b.mode == ddcore.MODE_SEPARABLE and any(p.kind == "host" for p in b.spec.phi):
        return None
    b.upload(0, initial_state.fields())
    if isinstance(integrator, p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple):
        _, norms, stats = b.run_pc(0, 1, t0, dt, num_steps, integrator._opts(), norms=True)
        integrator.last_stats = stats
    elif isinstance(integrator, p1.ForwardEulerIntegrator):
        _, norms = b.run_feuler(0, 1, t0, dt, num_steps, norms=True)
    else:
        return None
    times, t = [t0], t0
    for _ in range(num_steps):
        t += dt
        times.append(t)
    return _series_from_norms(times, norms[:, 0, :], variable_names, integral_vars)


class NumericalErrorSummary:
    """Summary error norms of a trial (reference src/mms_trial_utils.py:150-198)."""

    def __init__(self, dt_used: float, time_series_data: List[ErrorTimeSeries], variable_names: List[str],
                 integral_vars: List[str]):
        self.dt_used = dt_used
        self.variable_names = variable_names
        self.integral_vars = integral_vars
        if not time_series_data:
            raise ValueError("time_series_data cannot be empty.")
        self.overall_combined_error: float = calculate_combined_error_norm(time_series_data, dt_used, integral_vars)
        self.per_variable_sup_errors: Dict[str, float] = {}
        for v in variable_names:
            iv = [v] if v in integral_vars else []
            self.per_variable_sup_errors[v] = calculate_combined_error_norm(time_series_data, dt_used,
                                                                           integral_vars=iv, all_variables=[v])

    def __repr__(self):
        pv = {k: f"{v:.4e}" for k, v in self.per_variable_sup_errors.items()}
        return (f"NumericalErrorSummary(dt={self.dt_used:.2e}, OverallCombinedError="
                f"{self.overall_combined_error:.4e}, PerVariableSupErrors={pv})")


class MMSTrial:
    """Setup and execution of one MMS trial (reference src/mms_trial_utils.py:201-280)."""

    def __init__(self, grid, model, mms_case_cls: Type, field_cls: Type, forcing_terms_cls: Type,
                 integrator_cls: Type, mms_case_params: Optional[Dict] = {}, integrator_params: Optional[Dict] = {},
                 forcing_terms_params: Optional[Dict] = {}, field_params: Optional[Dict] = {},
                 variable_names: List[str] = None, integral_vars: List[str] = None):
        self.grid = grid
        self.model = model
        self.mms_case_cls = mms_case_cls
        self.field_cls = field_cls
        self.forcing_terms_cls = forcing_terms_cls
        self.integrator_cls = integrator_cls
        self.variable_names = variable_names or ["cp", "T", "cl", "cd", "cs"]
        self.integral_vars = integral_vars or ["T", "cl", "cd"]
        self.mms_case = mms_case_cls(grid=self.grid, model=self.model, **mms_case_params)
        self.forcing_terms = forcing_terms_cls(mms_case=self.mms_case, model=self.model, **forcing_terms_params)
        self.field = field_cls(grid=self.grid, model=self.model, forcing_terms=self.forcing_terms, **field_params)
        self.integrator = integrator_cls(semi_discrete_field=self.field, **integrator_params)
        self.initial_state = p1.state_from_mms_when(mms_case=self.mms_case, t=0.0, grid=self.grid)

    def run_for_errors(self, Tf: float, dt: float, t0: float = 0.0) -> NumericalErrorSummary:
        series, dt_used = run_simulation_collect_data(
            grid=self.grid, integrator=self.integrator, exact_sol_pack=self.mms_case,
            initial_state=self.initial_state, Tf=Tf, dt=dt, t0=t0, variable_names=self.variable_names,
            integral_vars=self.integral_vars)
        return NumericalErrorSummary(dt_used=dt_used, time_series_data=series, variable_names=self.variable_names,
                                     integral_vars=self.integral_vars)
