"""Convergence-study driver with the reference's `cvg_studies_base` API (src/cvg_studies_base.py),
written from scratch on the device path.

Same names, arguments, result layout and status strings: `calculate_observed_rates`,
`run_simulation_collect_data`, `calculate_combined_error_norm`, `_setup_simulation_instances`,
`run_convergence_studies`, `TimeStepData`, `RateStatus`.  The time loops go through
`mms_trial_utils.run_simulation_collect_data`, i.e. whole trials run on the GPU (time loop, exact
solution, error norms) whenever the integrator and the case are device-evaluable, and through the class
API step by step otherwise.

The reference driver builds `forcing_terms_cls(mms_case=, model=)`, `field_cls(grid=, model=,
forcing_terms=)` and `integrator_cls(semi_discrete_field=, num_pc_steps=, num_newton_steps=)`
(src/cvg_studies_base.py:276-299); classes that need more (the RegHCsTriple ones take a regularisation
factor) are passed as `functools.partial` objects, exactly as with the reference.
"""

from __future__ import annotations

import math
import time
from collections import namedtuple
from typing import Any, Dict, List, Literal, NamedTuple, Optional, Tuple

import numpy as np

import mms_trial_utils as _mtu
import prob1base as p1

VERBOSE = True


def _say(*a, **k):
    if VERBOSE:
        print(*a, **k)


class _RateStatus(NamedTuple):
    OK: str = "OK"
    INSUFFICIENT_DATA: str = "Insufficient Data"
    ZERO_DENOMINATOR_ZERO_NUMERATOR: str = "Differences near zero (converged/stalled?)"
    ZERO_DENOMINATOR_NONZERO_NUMERATOR: str = "Unstable rate (denominator near zero)"
    NON_POSITIVE_RATIO: str = "Non-positive ratio (convergence issue?)"
    ERROR_INCREASING: str = "Error increasing significantly"


RateStatus = _RateStatus()

TimeStepData = namedtuple("TimeStepData", ["t", "h_norm_sq_errors", "grad_h_norm_p_sq_errors"])


def _triplet_rate(coarse: float, medium: float, fine: float, log_r: float) -> Tuple[float, str]:
    """Observed order from three consecutive errors: log_r[(e_c - e_m) / (e_m - e_f)] (reference
    src/cvg_studies_base.py:60-103, including its classification of the degenerate cases)."""
    drop_cm, drop_mf = coarse - medium, medium - fine
    if drop_mf < 0:
        return math.nan, RateStatus.ERROR_INCREASING
    if drop_cm <= 0:
        return math.nan, RateStatus.NON_POSITIVE_RATIO
    ratio = drop_cm / drop_mf  # an exactly stalled pair (e_m == e_f) divides by zero, as in the reference
    status = RateStatus.OK
    tiny = np.finfo(float).eps
    if abs(drop_mf) < tiny:
        status = (RateStatus.ZERO_DENOMINATOR_ZERO_NUMERATOR if abs(drop_cm) < tiny
                  else RateStatus.ZERO_DENOMINATOR_NONZERO_NUMERATOR)
    assert ratio > 0
    return math.log(ratio) / log_r, status


def calculate_observed_rates(errors: List[float], refinement_factor: float = 2.0) -> List[Tuple[float, str]]:
    """[(rate, status)] for every consecutive triple of `errors` (coarsest first)."""
    assert len(errors) >= 3, "At least 3 error values are required for rate calculation."
    assert refinement_factor > 1.0, "Refinement factor must be > 1.0"
    assert all(e >= 0 for e in errors), "All error values must be positive for rate calculation."
    log_r = math.log(refinement_factor)
    return [_triplet_rate(errors[k], errors[k + 1], errors[k + 2], log_r) for k in range(len(errors) - 2)]


def run_simulation_collect_data(*, grid, integrator, exact_sol_pack, initial_state_gfp, Tf: float, dt: float,
                                variable_names: List[str], integral_vars: List[str]
                                ) -> Tuple[List[TimeStepData], float]:
    """t = 0 .. Tf in ceil(Tf/dt) equal steps; per-step squared H-norm errors of every variable and squared
    gradient-norm errors of the integral variables (reference src/cvg_studies_base.py:117-221)."""
    t_start = time.time()
    series, dt_used = _mtu.run_simulation_collect_data(
        grid=grid, integrator=integrator, exact_sol_pack=exact_sol_pack, initial_state=initial_state_gfp, Tf=Tf,
        dt=dt, t0=0.0, variable_names=variable_names, integral_vars=integral_vars)
    _say(f"    Simulation finished in {time.time() - t_start:.2f} seconds.")
    return [TimeStepData(t=s.t, h_norm_sq_errors=s.h_norm_sq_errors,
                         grad_h_norm_p_sq_errors=s.grad_h_norm_p_sq_errors) for s in series], dt_used


def calculate_combined_error_norm(time_series_data, dt: float, integral_vars: List[str]) -> float:
    """Combined max-integral error norm (reference src/cvg_studies_base.py:224-250)."""
    err = _mtu.calculate_combined_error_norm(time_series_data, dt, integral_vars)
    _say(f"    Combined Max-Integral Error Norm: {err:.4e}")
    return err


def _setup_simulation_instances(*, field_cls, forcing_terms_cls, mms_case_cls, integrator_cls, grid, model,
                                variable_names: List[str], num_pc_steps: int, num_newton_steps: int):
    """(mms_case, integrator, initial state at t = 0) for one grid (reference src/cvg_studies_base.py:253-299)."""
    mms_case = mms_case_cls(grid=grid, model=model)
    forcing_terms = forcing_terms_cls(mms_case=mms_case, model=model)
    field = field_cls(grid=grid, model=model, forcing_terms=forcing_terms)
    integrator = integrator_cls(semi_discrete_field=field, num_pc_steps=num_pc_steps,
                                num_newton_steps=num_newton_steps)
    at0 = {v: getattr(mms_case, v)(0.0, grid.xx, grid.yy) for v in variable_names}
    initial = p1.StateVars(**at0, model=model, hh=grid.hh, kk=grid.kk)
    return mms_case, integrator, initial


CvgType = Literal["temporal", "spatial"]
CvgReport = Dict[str, List[Optional[float]]]
FullCvgReport = Dict[CvgType, CvgReport]
StudyConfig = Tuple[Any, Any, Any, Any, str]


def _report(errors: List[float], factor: float) -> dict:
    rated = calculate_observed_rates(errors, factor)
    return {"errors": errors, "rates": [r for r, _ in rated], "statuses": [s for _, s in rated]}


def run_convergence_studies(study_configs: List[StudyConfig], study_params: Dict[str, Any]
                            ) -> Dict[str, FullCvgReport]:
    """Spatial study (N = N_base * 2^k at fixed dt) and temporal study (dt = dt_base / 2^k on a fixed grid)
    for every (field_cls, mms_case_cls, forcing_terms_cls, integrator_cls, label) of `study_configs`
    (reference src/cvg_studies_base.py:319-486).  Returns label -> {"spatial" | "temporal" ->
    {"errors", "rates", "statuses"}}."""
    names = study_params.get("variable_names", ["cp", "T", "cl", "cd", "cs"])
    integral = study_params.get("integral_vars", ["T", "cl", "cd"])
    Tf, model = study_params["Tf"], study_params["model"]
    build = dict(variable_names=names, num_pc_steps=study_params.get("num_pc_steps", 1),
                 num_newton_steps=study_params.get("num_newton_steps", 1))
    factor = 2

    def trial_error(grid, integrator, case, initial, dt):
        series, dt_used = run_simulation_collect_data(
            grid=grid, integrator=integrator, exact_sol_pack=case, initial_state_gfp=initial, Tf=Tf, dt=dt,
            variable_names=names, integral_vars=integral)
        return calculate_combined_error_norm(series, dt_used, integral)

    results: Dict[str, FullCvgReport] = {}
    for field_cls, mms_case_cls, forcing_terms_cls, integrator_cls, label in study_configs:
        classes = dict(field_cls=field_cls, forcing_terms_cls=forcing_terms_cls, mms_case_cls=mms_case_cls,
                       integrator_cls=integrator_cls, model=model, **build)
        _say(f"\n===== Running Studies for Case: {label} =====")

        _say("\n--- Starting Spatial Convergence Study ---")
        dt_fixed = study_params["dt_fixed_spatial"]
        spatial: List[Optional[float]] = []
        for k in range(study_params["num_spatial_refinements"]):
            t_level = time.time()
            N = study_params["N_base_spatial"] * factor ** k
            _say(f"\n  Spatial Level {k} (N=M={N}, dt={dt_fixed:.1e})")
            grid = p1.make_uniform_grid(N, N)
            case, integrator, initial = _setup_simulation_instances(grid=grid, **classes)
            spatial.append(trial_error(grid, integrator, case, initial, dt_fixed))
            _say(f"  Spatial Level {k} finished in {time.time() - t_level:.2f} seconds.")

        _say("\n--- Starting Temporal Convergence Study ---")
        Nt = study_params["N_fixed_temporal"]
        _say(f"  Setting up fixed grid (N=M={Nt})...")
        grid = p1.make_uniform_grid(Nt, Nt)
        case, prototype, initial = _setup_simulation_instances(grid=grid, **classes)
        integrator = integrator_cls(semi_discrete_field=prototype.semi_discrete_field,
                                    num_pc_steps=build["num_pc_steps"], num_newton_steps=build["num_newton_steps"])
        temporal: List[Optional[float]] = []
        for k in range(study_params["num_temporal_refinements"]):
            t_level = time.time()
            dt = study_params["dt_base_temporal"] / factor ** k
            _say(f"\n  Temporal Level {k} (dt={dt:.4e})")
            temporal.append(trial_error(grid, integrator, case, initial, dt))
            _say(f"  Temporal Level {k} finished in {time.time() - t_level:.2f} seconds.")

        results[label] = {"spatial": _report(spatial, factor), "temporal": _report(temporal, factor)}
        _say(f"\n===== Finished Studies for Case: {label} =====")
    return results
