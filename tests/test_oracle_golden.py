"""Pins the oracle to the reference: oracle vs tests/golden/*.npz (CPU only).

The fixtures were produced by the live reference (oracle/make_golden.py).
Tolerance: 1e-12 norm-wise relative on fields (the north_star bound; observed
~1e-15), exact equality on cs-Newton iteration counts, 1e-9 relative on the
convergence-study error norms (these are differences of nearly equal numbers).
"""
import numpy as np
import pytest

from golden_util import VARS, fixture_names, load_fixture, oracle_model, rel_err
from oracle import (OForcing, OGrid, OState, PCStepper, feuler_step, fields_F, make_case, run_trial,
                    uniform_grid)

FIELD_TOL = 1e-12


def _setup(desc, z):
    g = OGrid(z["x"], z["y"])
    m = oracle_model(desc["model"]).with_changes(reaction=desc.get("variant", "regh"))
    if desc["case"] is not None:
        case = make_case(desc["case"], m)
        forcing = OForcing(case, m, desc["eta"], g)
    else:
        case, forcing = None, None
    s = OState(**{v: z["init_" + v] for v in VARS})
    return g, m, case, forcing, s


@pytest.mark.parametrize("name", fixture_names(kind="steps") + fixture_names(kind="steps", program=True))
def test_steps_match_reference(name):
    desc, z = load_fixture(name)
    g, m, case, forcing, s = _setup(desc, z)
    eta, dt, t = desc["eta"], desc["dt"], desc["t0"]
    if desc["init"] == "exact":
        for v in VARS:
            assert rel_err(getattr(case, v)(t, g.xx, g.yy), z["init_" + v]) <= 1e-14
    F = fields_F(s, t, g, m, eta, forcing)
    for v in VARS:
        assert rel_err(F[v], z["F0_" + v]) <= FIELD_TOL, f"F0_{v}"
    stepper = PCStepper(g, m, eta, forcing, **desc["pc"]) if desc["integrator"] == "pc" else None
    for n in range(desc["nsteps"]):
        s = stepper.step(s, t, dt) if stepper else feuler_step(s, t, dt, g, m, eta, forcing)
        t += dt
        if f"step{n + 1}_cp" in z:
            for v in VARS:
                assert rel_err(getattr(s, v), z[f"step{n + 1}_{v}"]) <= FIELD_TOL, f"step {n + 1} {v}"
    if stepper:
        per_step = desc["pc"].get("num_pc_steps", 1)
        got = np.array(stepper.cs_newton_iters).reshape(-1, per_step).sum(axis=1)
        assert np.array_equal(got, z["cs_newton_calls_per_step"])  # 0 for the closed-form cs correctors
        for v in ("T", "cl", "cd"):
            scale = max(np.max(np.abs(z[f"step{desc['nsteps']}_{v}"])), 1e-300)
            assert np.max(np.abs(stepper.last_residual[v] - z["resid_" + v])) <= 1e-11 * scale


@pytest.mark.parametrize("name", fixture_names(kind="trial"))
def test_trial_errors_match_reference(name):
    desc, z = load_fixture(name)
    m = oracle_model(desc["model"])
    case = make_case(desc["case"], m)
    for li, lv in enumerate(desc["levels"]):
        if lv["N"] > 32:
            continue  # keep the CPU suite short; covered by the GPU tests
        eta = lv.get("eta", desc["eta"])
        g = uniform_grid(lv["N"], lv["M"])
        forcing = OForcing(case, m, eta, g)
        r = run_trial(case, g, m, eta, forcing, Tf=desc["Tf"], dt=lv["dt"], integrator=desc["integrator"],
                      keep_residuals=False, **desc["pc"]) if desc["integrator"] == "pc" else \
            run_trial(case, g, m, eta, forcing, Tf=desc["Tf"], dt=lv["dt"], integrator="fe")
        assert r["dt"] == float(z[f"L{li}_dt_used"])
        assert abs(r["overall"] - float(z[f"L{li}_overall"])) <= 1e-9 * float(z[f"L{li}_overall"]) + 1e-13
        ref_pv = z[f"L{li}_per_var"]
        got_pv = np.array([r["per_var"][v] for v in VARS])
        # error norms are differences of O(1) fields: allow 1e-13 absolute (fields agree to 1e-12)
        assert np.all(np.abs(got_pv - ref_pv) <= 1e-9 * np.abs(ref_pv) + 1e-13)


def test_published_notebook_numbers():
    """12-digit errors printed in the notebooks."""
    from oracle import NOTEBOOK_CONSTS
    published = {
        "expsin": [1.942652829989e-05, 5.197056624911e-06, 1.322695968641e-06, 3.372248813359e-07],
        "pol": [4.93452e-05, 1.59616e-05, 4.28269e-06, 1.08800e-06],
    }
    for cname, errs in published.items():
        m = NOTEBOOK_CONSTS[cname]
        case = make_case(cname, m)
        for n, e in zip((2, 4, 8, 16), errs):
            g = uniform_grid(n, n)
            r = run_trial(case, g, m, 50.0, OForcing(case, m, 50.0, g), Tf=0.01, dt=(1.0 / n) ** 1.5,
                          keep_residuals=False)
            tol = 1e-9 if cname == "expsin" else 1e-5  # Pol is printed with 6 digits
            assert abs(r["overall"] - e) <= tol * e, (cname, n, r["overall"], e)
