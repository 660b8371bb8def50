#!/usr/bin/env python
"""Generate tests/golden/*.npz from the LIVE reference (build container only).

Usage:  python oracle/make_golden.py [--only NAME]

Imports the unmodified reference from /root/reference/src (read-only; nothing
is copied) and records, for a fixed list of scenarios, what the reference
itself computes on the hot path:

  * per-step fields of ForwardEulerIntegrator.step and
    P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple.step,
  * `last_residual` after the final step,
  * the number of cs-corrector Newton iterations per step (counted by wrapping
    `_predictor_equation`, which the reference calls once per iteration,
    src/prob1base.py:3654-3663),
  * semidiscrete fields Fcp..Fcs on random states,
  * MMSTrial.run_for_errors summaries (overall / per-variable error norms).

Each fixture stores its scenario descriptor as JSON, so tests rebuild the same
inputs for the oracle and for the CUDA path without this script or the
reference.  The fixtures are the pin that ties `oracle/` to the reference.
"""

from __future__ import annotations

import argparse
import json
import os
import sys

import numpy as np

REF_SRC = "/root/reference/src"
OUT_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")

VARS = ("cp", "T", "cl", "cd", "cs")

NOTEBOOK = {
    "expsin": dict(K1=1e-3, K2=1e-3, K3=1e-3, K4=1e-3, DT=1e-3, Dl_max=1e-5, phi_l=1e-5, gamma_T=1e-9,
                   Kd=1e-2, Sd=1.0, Dd_max=1e-6, phi_d=1e-5, r_sp=5e-2, T_ref=300.0),
    "pol": dict(K1=1e-3, K2=1e-3, K3=1e-3, K4=1e-3, DT=1e-3, Dl_max=8.01e-4, phi_l=1e-5, gamma_T=1e-9,
                Kd=1e-2, Sd=1.0, Dd_max=2.46e-6, phi_d=1e-5, r_sp=5e-2, T_ref=300.0),
}
# O(1) constants that make every term of the scheme matter (used with random states)
STRESS = dict(K1=0.7, K2=0.4, K3=0.9, K4=0.6, DT=0.05, Dl_max=0.08, phi_l=0.3, gamma_T=0.2,
              Kd=0.8, Sd=1.5, Dd_max=0.06, phi_d=0.25, phi_T=2.0, r_sp=5e-2, T_ref=3.0)


def _consts_for(case):
    # cell 3 of each notebook: ExpSin and the NonFullySmoothPol studies share one set,
    # Pol and SlowlyChangingPeaks_Fast1e1 the other (Dl_max=8.01e-4, Dd_max=2.46e-6)
    return NOTEBOOK["expsin"] if case in ("expsin", "nfsp_h1h2", "nfsp_h2h2", "nfsp_h2h3") else NOTEBOOK["pol"]


NONSEP_MODEL = dict(K1=1e-3, K2=2e-3, K3=1.5e-3, K4=1e-3, DT=1e-3, Dl_max=8.01e-4, phi_l=1e-5, gamma_T=1e-3, Kd=1e-2,
                    Sd=1.0, Dd_max=2.46e-6, phi_d=1e-5, r_sp=5e-2, T_ref=300.0, kind=2)


sys.path.insert(0, os.path.join(os.path.dirname(os.path.abspath(__file__)), ".."))
from oracle.mms import nonsep_exprs  # noqa: E402  (one definition for the oracle and for the reference run)


def nonsep_expr_strings():
    import sympy
    t, x, y = sympy.symbols("t x y")
    return {k: str(v) for k, v in nonsep_exprs(t, x, y).items()}


def scenarios():
    S = []
    # A. MMS per-step fields, uniform grids, notebook constants
    for case in ("pol", "expsin", "scp_fast1e1", "nfsp_h1h2"):
        for (N, M) in ((8, 8), (12, 9)):
            h = 1.0 / N
            for integ in ("pc", "fe"):
                S.append(dict(name=f"steps_{case}_{N}x{M}_{integ}", kind="steps", case=case,
                              model=dict(_consts_for(case), kind=2), grid=dict(N=N, M=M), eta=50.0,
                              dt=h ** 1.5 if integ == "pc" else 1e-3, t0=0.0, nsteps=3, integrator=integ,
                              init="exact", pc={}))
    S.append(dict(name="steps_pol_32x32_pc", kind="steps", case="pol", model=dict(NOTEBOOK["pol"], kind=2),
                  grid=dict(N=32, M=32), eta=50.0, dt=5e-4, t0=0.0, nsteps=2, integrator="pc", init="exact",
                  pc={}, keep="last"))
    S.append(dict(name="steps_expsin_16x16_pc_t0", kind="steps", case="expsin",
                  model=dict(NOTEBOOK["expsin"], kind=2), grid=dict(N=16, M=16), eta=100.0, dt=2e-3, t0=0.25,
                  nsteps=2, integrator="pc", init="exact", pc=dict(num_pc_steps=2, num_newton_steps=2)))
    # large dt: weak diagonal dominance of the Newton matrices
    S.append(dict(name="steps_scp_10x10_pc_bigdt", kind="steps", case="scp_fast1e1",
                  model=dict(NOTEBOOK["pol"], kind=2, DT=0.5, Dl_max=0.3, Dd_max=0.2), grid=dict(N=10, M=10),
                  eta=50.0, dt=0.25, t0=0.0, nsteps=2, integrator="pc", init="exact", pc={}))
    # B. random states, non-uniform grid, stress constants, no forcing / MMS forcing
    for kind in (1, 2):
        for pc in ({}, dict(num_pc_steps=2, num_newton_steps=2)):
            tag = "pc22" if pc else "pc11"
            S.append(dict(name=f"random_nonuniform_model{kind}_{tag}", kind="steps", case=None,
                          model=dict(STRESS, kind=kind), grid=dict(N=6, M=5, nonuniform=True), eta=7.0,
                          dt=0.03, t0=0.1, nsteps=2, integrator="pc", init="random", pc=pc))
    S.append(dict(name="random_nonuniform_model2_fe", kind="steps", case=None, model=dict(STRESS, kind=2),
                  grid=dict(N=6, M=5, nonuniform=True), eta=7.0, dt=0.01, t0=0.0, nsteps=2, integrator="fe",
                  init="random", pc={}))
    S.append(dict(name="random_uniform_model2_pol_forcing", kind="steps", case="pol",
                  model=dict(STRESS, kind=2), grid=dict(N=7, M=9), eta=20.0, dt=0.02, t0=0.0, nsteps=2,
                  integrator="pc", init="random", pc={}))
    # cs-Newton exit test that actually fires (cs != 0 on every node, boundary included)
    S.append(dict(name="random_csnewton_exit", kind="steps", case=None, model=dict(STRESS, kind=2),
                  grid=dict(N=6, M=6), eta=3.0, dt=0.01, t0=0.0, nsteps=3, integrator="pc", init="random_pos",
                  pc=dict(num_newton_iterations=50, consec_xs_rtol=1e-6)))
    # D. iteration-count stress (tests/test_reghcstriple_system.py:372-377 settings)
    S.append(dict(name="steps_nfsp_8x8_pc_newton1000", kind="steps", case="nfsp_h1h2",
                  model=dict(NOTEBOOK["expsin"], kind=2), grid=dict(N=8, M=8), eta=50.0, dt=5e-4, t0=0.0,
                  nsteps=2, integrator="pc", init="exact",
                  pc=dict(num_newton_iterations=1000, consec_xs_rtol=1e-9)))
    # C. convergence-study numbers (MMSTrial.run_for_errors)
    for case, Tf, levels in (("expsin", 0.01, (2, 4, 8, 16, 32, 64)), ("pol", 0.01, (2, 4, 8, 16, 32, 64)),
                             ("scp_fast1e1", 1.0, (2, 4, 8, 16)), ("nfsp_h1h2", 1.0, (2, 4, 8, 16))):
        S.append(dict(name=f"trial_spatial_{case}", kind="trial", case=case,
                      model=dict(_consts_for(case), kind=2), eta=50.0, Tf=Tf,
                      levels=[dict(N=n, M=n, dt=(1.0 / n) ** 1.5) for n in levels], integrator="pc", pc={}))
    S.append(dict(name="trial_eta_pol", kind="trial", case="pol", model=dict(NOTEBOOK["pol"], kind=2), eta=None,
                  Tf=0.01, levels=[dict(N=32, M=32, dt=5e-4, eta=e) for e in (10.0, 50.0, 100.0, 1000.0)],
                  integrator="pc", pc={}))
    S.append(dict(name="trial_temporal_expsin", kind="trial", case="expsin",
                  model=dict(NOTEBOOK["expsin"], kind=2), eta=50.0, Tf=0.01,
                  levels=[dict(N=32, M=32, dt=1e-2 / 2 ** k) for k in range(4)], integrator="pc", pc={}))
    S.append(dict(name="trial_fe_pol", kind="trial", case="pol", model=dict(NOTEBOOK["pol"], kind=2), eta=50.0,
                  Tf=0.01, levels=[dict(N=8, M=8, dt=1e-3), dict(N=8, M=8, dt=5e-4)], integrator="fe", pc={}))
    # F. the CsTriple and HCsTriple field variants (src/prob1base.py:2842-2876, 3152-3430): same stencils, other
    #    F2(cs) and closed-form cs correctors
    for variant in ("cs", "h"):
        for case in ("pol", "expsin"):
            for integ in ("pc", "fe"):
                S.append(dict(name=f"steps_{variant}triple_{case}_12x9_{integ}", kind="steps", case=case,
                              variant=variant, model=dict(_consts_for(case), kind=2), grid=dict(N=12, M=9),
                              eta=50.0, dt=(1.0 / 12) ** 1.5 if integ == "pc" else 1e-3, t0=0.0, nsteps=3,
                              integrator=integ, init="exact", pc={}))
        S.append(dict(name=f"random_nonuniform_{variant}triple_pc22", kind="steps", case=None, variant=variant,
                      model=dict(STRESS, kind=2), grid=dict(N=6, M=5, nonuniform=True), eta=7.0, dt=0.03, t0=0.1,
                      nsteps=2, integrator="pc", init="random", pc=dict(num_pc_steps=2, num_newton_steps=2)))
        S.append(dict(name=f"random_uniform_{variant}triple_pol_forcing", kind="steps", case="pol", variant=variant,
                      model=dict(STRESS, kind=2), grid=dict(N=7, M=9), eta=20.0, dt=0.02, t0=0.0, nsteps=2,
                      integrator="pc", init="random", pc={}))
    # E. the convergence-study driver (src/cvg_studies_base.py): observed rates with their status strings and one
    #    small spatial + temporal study (the driver's constructor convention has no regularisation factor, so the
    #    RegHCsTriple classes are passed as functools.partial objects)
    S.append(dict(name="cvg_study_pol", kind="cvg", case="pol", model=dict(NOTEBOOK["pol"], kind=2), eta=50.0,
                  params=dict(Tf=4e-3, N_base_spatial=4, num_spatial_refinements=4, dt_fixed_spatial=1e-3,
                              N_fixed_temporal=12, dt_base_temporal=2e-3, num_temporal_refinements=3),
                  rate_cases=[[1.0, 0.25, 0.0625, 0.015625], [4e-5, 1.6e-5, 4.3e-6, 1.1e-6, 2.8e-7],
                              [1.0, 2.0, 1.5], [1.0, 1.0, 0.5], [0.5, 0.25, 0.3], [3.0, 2.0, 2.0 - 1e-17, 1.0],
                              [0.0, 0.0, 0.0], [1.0, 0.1, 0.01, 0.001], [1.0, 0.5, 0.25]],
                  rate_factors=[2.0, 2.0, 2.0, 2.0, 2.0, 2.0, 2.0, 10.0, 3.0]))
    # G. a manufactured solution that is NOT a sum of separable terms (MMSCaseSymbolic on the expressions of
    #    tests/test_program_codegen.py:nonseparable_exprs): pins the generated forcing programs to the reference
    for integ in ("pc", "fe"):
        S.append(dict(name=f"steps_nonsep_14x11_{integ}", kind="steps", case="nonsep", model=dict(NONSEP_MODEL),
                      grid=dict(N=14, M=11, nonuniform=True), eta=50.0, dt=2e-3 if integ == "pc" else 1e-3, t0=0.05,
                      nsteps=3, integrator=integ, init="exact", pc={}, exprs=nonsep_expr_strings()))
    # the finite-difference helper behind MMSCaseFromAnalytic
    S.append(dict(name="analytic_fd", kind="analytic", nx=7, ny=5, t=0.3))
    return S


# ----------------------------------------------------------------------------

ANALYTIC_DERIVS = [(0, 0, 0), (1, 0, 0), (2, 0, 0), (0, 1, 0), (0, 2, 0), (0, 0, 1), (0, 0, 2), (0, 1, 1), (1, 1, 0)]


def analytic_fn(t, x, y):
    """the callable differentiated by the `analytic` fixture (the test restates it)"""
    return np.exp(-t) * np.sin(2 * x + y * t) + x * y ** 2


def run_analytic(d):
    """pack_analytical_txy_with_o2fdm_derivatives of the live reference (src/prob1base.py:895-1031): every
    derivative selector, the Laplacian and a caller-chosen step, for the three time-stepping strategies."""
    p1, _, _ = _ref()
    X, Y = np.meshgrid(np.linspace(0, 1, d["nx"]), np.linspace(0, 1, d["ny"]), indexing="ij")
    out = {"X": X, "Y": Y}
    for ts in ("center", "forward", "backward"):
        g = p1.pack_analytical_txy_with_o2fdm_derivatives(analytic_fn, time_stepping=ts)
        for dd in ANALYTIC_DERIVS:
            out[f"{ts}_d{dd[0]}{dd[1]}{dd[2]}"] = g(d["t"], X, Y, d=dd)
        out[f"{ts}_lap"] = g(d["t"], X, Y, op="lap")
        out[f"{ts}_d020_eps1e-4"] = g(d["t"], X, Y, d=(0, 2, 0), small_eps=1e-4)
    return out


def make_xy(gd):
    N, M = gd["N"], gd["M"]
    if gd.get("nonuniform"):
        rng = np.random.default_rng(7)
        x = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, N))])
        y = np.concatenate([[0.0], np.cumsum(rng.uniform(0.5, 1.5, M))])
        return x / x[-1], y / y[-1]
    return np.linspace(0, 1, N + 1), np.linspace(0, 1, M + 1)


def random_fields(shape, positive=False):
    # the RNG recipe of the reference's own state test (tests/test_statevars.py:17)
    rng = np.random.default_rng(20250503)
    lo = 0.5 if positive else -0.5
    # T stays positive: DefaultModel01 evaluates exp(-phi_T / T)
    return {v: rng.uniform(0.5 if v == "T" else lo, 1.5, shape) for v in VARS}


def _ref():
    if REF_SRC not in sys.path:
        sys.path.insert(0, REF_SRC)
    import prob1base as p1  # noqa
    import prob1_mms_cases as p1mc  # noqa
    import mms_trial_utils as mtu  # noqa
    return p1, p1mc, mtu


def ref_case_cls(p1mc, case):
    return {"pol": p1mc.MMSCasePol, "expsin": p1mc.MMSCaseExpSin,
            "scp_fast1e1": p1mc.MMSCaseSlowlyChangingPeaks_Fast1e1,
            "nfsp_h1h2": p1mc.MMSCaseNonFullySmoothPol_cpcsH1_TclcdH2,
            "nfsp_h2h2": p1mc.MMSCaseNonFullySmoothPol_cpcsH2_TclcdH2,
            "nfsp_h2h3": p1mc.MMSCaseNonFullySmoothPol_cpcsH2_TclcdH3}[case]


def ref_model(p1, md):
    md = dict(md)
    kind = md.pop("kind")
    md.setdefault("R0", p1.R0)
    md.setdefault("Ea", p1.Ea)
    md.setdefault("phi_T", p1.Ea / p1.R0)
    mc = p1.ModelConsts(**md)
    return (p1.DefaultModel02 if kind == 2 else p1.DefaultModel01)(mc)


def run_steps(d):
    p1, p1mc, _ = _ref()
    x, y = make_xy(d["grid"])
    grid = p1.Grid(x, y)
    model = ref_model(p1, d["model"])
    eta = d["eta"]
    variant = d.get("variant", "regh")
    forcing_cls, field_cls, integ_cls = {
        "regh": (p1.ForcingTerms_RegHCsTriple, p1.SemiDiscreteField_RegHCsTriple,
                 p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple),
        "cs": (p1.ForcingTerms_CsTriple, p1.SemiDiscreteField_CsTriple,
               p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_CsTriple),
        "h": (p1.ForcingTerms_HCsTriple, p1.SemiDiscreteField_HCsTriple,
              p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_HCsTriple)}[variant]
    extra = dict(regularization_factor=eta) if variant == "regh" else {}
    if d["case"] == "nonsep":
        ex = nonsep_exprs(p1.t_sym, p1.x_sym, p1.y_sym)
        case = p1.MMSCaseSymbolic(grid=grid, model=model, **{k + "_sym_expr": v for k, v in ex.items()})
        forcing = forcing_cls(mms_case=case, model=model, **extra)
    elif d["case"] is not None:
        case = ref_case_cls(p1mc, d["case"])(grid=grid, model=model)
        forcing = forcing_cls(mms_case=case, model=model, **extra)
    else:
        case, forcing = None, p1.NoForcingTerms(grid)
    field = field_cls(grid=grid, model=model, forcing_terms=forcing, **extra)
    if d["init"] == "exact":
        s = p1.state_from_mms_when(mms_case=case, t=d["t0"], grid=grid)
    else:
        f = random_fields(grid.full_shape, positive=(d["init"] == "random_pos"))
        s = p1.StateVars(f["cp"], f["T"], f["cl"], f["cd"], f["cs"], model=model, hh=grid.hh, kk=grid.kk)
    out = {"x": x, "y": y}
    for v in VARS:
        out["init_" + v] = np.array(getattr(s, v))
    t = d["t0"]
    for v, F in zip(VARS, (field.Fcp, field.FT, field.Fcl, field.Fcd, field.Fcs)):
        out["F0_" + v] = F(s, t)
    counts = []
    if d["integrator"] == "pc":
        integ = integ_cls(field, **extra, **d["pc"])
        calls = [0]
        if variant == "regh":
            orig = integ._predictor_equation

            def counting(*a, **k):
                calls[0] += 1
                return orig(*a, **k)

            integ._predictor_equation = counting
    else:
        integ = p1.ForwardEulerIntegrator(field)
    keep_all = d.get("keep", "all") == "all"
    for n in range(d["nsteps"]):
        if d["integrator"] == "pc":
            before = calls[0]
        s = integ.step(s, t0=t, dt=d["dt"])
        t += d["dt"]
        if d["integrator"] == "pc":
            # corrector_cs_step runs once per pc step -> list per call
            counts.append(calls[0] - before)
        if keep_all or n == d["nsteps"] - 1:
            for v in VARS:
                out[f"step{n + 1}_{v}"] = np.array(getattr(s, v))
    if d["integrator"] == "pc":
        out["cs_newton_calls_per_step"] = np.array(counts)
        for v in ("T", "cl", "cd"):
            out["resid_" + v] = integ.last_residual[v]
    return out


def run_trial(d):
    p1, p1mc, mtu = _ref()
    model = ref_model(p1, d["model"])
    out = {}
    for li, lv in enumerate(d["levels"]):
        eta = lv.get("eta", d["eta"])
        grid = p1.make_uniform_grid(lv["N"], lv["M"])
        if d["integrator"] == "pc":
            icls = p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple
            ipar = dict(regularization_factor=eta, **d["pc"])
        else:
            icls, ipar = p1.ForwardEulerIntegrator, {}
        trial = mtu.MMSTrial(grid=grid, model=model, mms_case_cls=ref_case_cls(p1mc, d["case"]),
                             field_cls=p1.SemiDiscreteField_RegHCsTriple,
                             forcing_terms_cls=p1.ForcingTerms_RegHCsTriple, integrator_cls=icls,
                             forcing_terms_params={"regularization_factor": eta},
                             field_params={"regularization_factor": eta}, integrator_params=ipar)
        summ = trial.run_for_errors(Tf=d["Tf"], dt=lv["dt"])
        out[f"L{li}_overall"] = np.array(summ.overall_combined_error)
        out[f"L{li}_per_var"] = np.array([summ.per_variable_sup_errors[v] for v in VARS])
        out[f"L{li}_dt_used"] = np.array(summ.dt_used)
        print(f"   level {li}: N={lv['N']} overall={summ.overall_combined_error:.12e}", flush=True)
    return out


def run_cvg(d):
    import contextlib
    import functools
    import io
    p1, p1mc, _ = _ref()
    import cvg_studies_base as cvg  # the reference's driver
    model = ref_model(p1, d["model"])
    eta = d["eta"]
    cfg = (functools.partial(p1.SemiDiscreteField_RegHCsTriple, regularization_factor=eta),
           ref_case_cls(p1mc, d["case"]),
           functools.partial(p1.ForcingTerms_RegHCsTriple, regularization_factor=eta),
           functools.partial(p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple,
                             regularization_factor=eta),
           "study")
    with contextlib.redirect_stdout(io.StringIO()):
        rep = cvg.run_convergence_studies([cfg], dict(d["params"], model=model))["study"]
    out = {}
    for kind in ("spatial", "temporal"):
        out[kind + "_errors"] = np.array(rep[kind]["errors"], dtype=np.float64)
        out[kind + "_rates"] = np.array(rep[kind]["rates"], dtype=np.float64)
        out[kind + "_statuses"] = np.array(json.dumps(rep[kind]["statuses"]))
        print("  ", kind, rep[kind]["errors"], rep[kind]["rates"], flush=True)
    for k, (errs, fac) in enumerate(zip(d["rate_cases"], d["rate_factors"])):
        try:
            res = cvg.calculate_observed_rates(errs, fac)
            out[f"rates{k}"] = np.array([r for r, _ in res], dtype=np.float64)
            out[f"rates{k}_status"] = np.array(json.dumps([st for _, st in res]))
        except Exception as e:  # the driver's own failure modes are part of its behaviour
            out[f"rates{k}_raises"] = np.array(type(e).__name__)
    return out


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    import scipy
    import sympy
    os.makedirs(OUT_DIR, exist_ok=True)
    meta = dict(numpy=np.__version__, scipy=scipy.__version__, sympy=sympy.__version__,
                reference=REF_SRC)
    for d in scenarios():
        if args.only and args.only not in d["name"]:
            continue
        print("scenario", d["name"], flush=True)
        arrays = {"steps": run_steps, "trial": run_trial, "cvg": run_cvg, "analytic": run_analytic}[d["kind"]](d)
        np.savez_compressed(os.path.join(OUT_DIR, d["name"] + ".npz"),
                            __desc__=np.array(json.dumps(d)), __meta__=np.array(json.dumps(meta)), **arrays)


if __name__ == "__main__":
    main()
