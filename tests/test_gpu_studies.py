"""GPU tests (-m gpu) of the study layer on the device path: the convergence-study driver against the
reference driver's golden numbers, a refinement sweep against the reference's trial fixtures, and an
ensemble with per-member constants against the oracle.  Tolerance of the error norms: 1e-9 relative +
1e-13 absolute (differences of nearly equal fields); rates follow from the errors."""
import functools
import json

import os

import numpy as np
import pytest

from golden_util import VARS, load_fixture, rel_err

pytestmark = pytest.mark.gpu
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def close(a, b):
    a, b = np.asarray(a, dtype=np.float64), np.asarray(b, dtype=np.float64)
    return np.all(np.abs(a - b) <= 1e-9 * np.abs(b) + 1e-13)


@pytest.fixture(scope="module")
def mods():
    import cvg_studies_base as cvg
    import ddcore
    import ddensemble
    import prob1base as p1
    from test_hostsim import CASES, product_model
    return dict(cvg=cvg, ens=ddensemble, p1=p1, ddcore=ddcore, CASES=CASES, product_model=product_model)


def test_convergence_study_driver_matches_reference(mods):
    cvg, p1 = mods["cvg"], mods["p1"]
    desc, z = load_fixture("cvg_study_pol")
    model, eta = mods["product_model"](desc["model"]), desc["eta"]
    cfg = (functools.partial(p1.SemiDiscreteField_RegHCsTriple, regularization_factor=eta),
           mods["CASES"][desc["case"]],
           functools.partial(p1.ForcingTerms_RegHCsTriple, regularization_factor=eta),
           functools.partial(p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple,
                             regularization_factor=eta),
           "study")
    cvg.VERBOSE = False
    rep = cvg.run_convergence_studies([cfg], dict(desc["params"], model=model))["study"]
    for kind in ("spatial", "temporal"):
        assert close(rep[kind]["errors"], z[kind + "_errors"]), (kind, rep[kind]["errors"], z[kind + "_errors"])
        assert rep[kind]["statuses"] == json.loads(str(z[kind + "_statuses"]))
        got, want = np.array(rep[kind]["rates"]), z[kind + "_rates"]
        assert np.array_equal(np.isnan(got), np.isnan(want))
        ok = ~np.isnan(want)
        assert np.all(np.abs(got[ok] - want[ok]) <= 1e-6)


@pytest.mark.parametrize("name", ["trial_spatial_pol", "trial_eta_pol", "trial_temporal_expsin"])
def test_refinement_sweep_matches_trial_fixtures(mods, name):
    """All levels of a study launched together (shared batches, concurrent streams) give the reference's
    per-level numbers."""
    desc, z = load_fixture(name)
    model = mods["product_model"](desc["model"])
    trials = [dict(N=lv["N"], M=lv["M"], dt=lv["dt"], Tf=desc["Tf"], eta=lv.get("eta", desc["eta"]))
              for lv in desc["levels"]]
    sw = mods["ens"].RefinementSweep(mods["CASES"][desc["case"]], model, trials, pc=desc["pc"])
    res = sw.run_for_errors()
    assert res["owned"].all()
    for li in range(len(trials)):
        assert sw.trials[li]["dt_used"] == float(z[f"L{li}_dt_used"])
        assert close(res["overall"][li], z[f"L{li}_overall"]), (li, res["overall"][li], z[f"L{li}_overall"])
        assert close(res["per_var"][li], z[f"L{li}_per_var"])


def test_sweep_split_over_two_ranks_covers_every_trial(mods):
    desc, z = load_fixture("trial_spatial_pol")
    model = mods["product_model"](desc["model"])
    trials = [dict(N=lv["N"], M=lv["M"], dt=lv["dt"], Tf=desc["Tf"], eta=desc["eta"]) for lv in desc["levels"]]
    parts = [mods["ens"].RefinementSweep(mods["CASES"]["pol"], model, trials, world=2, rank=r).run_for_errors()
             for r in range(2)]
    assert np.array_equal(parts[0]["owned"] ^ parts[1]["owned"], np.ones(len(trials), dtype=bool))
    merged = np.where(parts[0]["owned"], parts[0]["overall"], parts[1]["overall"])
    assert close(merged, [float(z[f"L{li}_overall"]) for li in range(len(trials))])


def test_ensemble_with_member_constants_matches_oracle(mods):
    """SCP_Fast1e1 members with perturbed K1..K4, DT, Kd and log-uniform eta (the draw of SURVEY 8d config 3):
    batched device run, sharded 2 ways, against the oracle member by member."""
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, make_case, run_trial
    import dataclasses
    p1, ens = mods["p1"], mods["ens"]
    rng = np.random.default_rng(20250503)
    n, N = 6, 12
    base = NOTEBOOK_CONSTS["pol"]
    etas = 10.0 ** rng.uniform(1.0, 3.0, n)
    omodels, models = [], []
    for _ in range(n):
        f = rng.uniform(0.5, 1.5, 6)
        om = dataclasses.replace(base, K1=base.K1 * f[0], K2=base.K2 * f[1], K3=base.K3 * f[2], K4=base.K4 * f[3],
                                 DT=base.DT * f[4], Kd=base.Kd * f[5])
        omodels.append(om)
        models.append(mods["product_model"](dict(
            K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max, phi_l=om.phi_l,
            gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max, phi_d=om.phi_d, r_sp=om.r_sp,
            T_ref=om.T_ref, kind=2)))
    grid = p1.make_uniform_grid(N, N)
    og = OGrid(np.array(grid.x), np.array(grid.y))
    Tf, dt = 0.004, 5e-4
    want = []
    for om, eta in zip(omodels, etas):
        oc = make_case("scp_fast1e1", om)
        r = run_trial(oc, og, om, float(eta), OForcing(oc, om, float(eta), og), Tf=Tf, dt=dt)
        want.append([r["overall"]] + [r["per_var"][v] for v in VARS])
    want = np.array(want)
    got = np.zeros_like(want)
    for rank in range(2):
        e = ens.TrajectoryEnsemble(grid, mods["CASES"]["scp_fast1e1"], models, etas, world=2, rank=rank, chunk=2)
        res = e.run_for_errors(Tf, dt)
        assert res["nsteps"] == 8
        got[e.first:e.last, 0] = res["overall"]
        got[e.first:e.last, 1:] = res["per_var"]
    assert close(got, want), (got, want)
    # one shared model object + scalar eta is the degenerate ensemble
    e = ens.TrajectoryEnsemble(grid, mods["CASES"]["scp_fast1e1"], models[0], [etas[0]] * 3)
    res = e.run_for_errors(Tf, dt)
    assert close(res["overall"], [want[0, 0]] * 3)


def test_ensemble_config3_spot_check_64_members(mods):
    """SURVEY 8d config 3 as written: MMSCaseSlowlyChangingPeaks_Fast1e1, N = M = 32, dt = 5e-4, Tf = 0.01 (20 steps),
    member parameters of the 10^5-member draw (rng 20250503: eta ~ logU[10, 1000]; K1..K4, DT, Kd each base * U[0.5,
    1.5]) -- 64 of those members, picked at random, in one batched device run against the oracle member by member."""
    import dataclasses
    import sys
    sys.path.insert(0, ROOT)
    import bench
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, make_case, run_trial
    p1, ens = mods["p1"], mods["ens"]
    n_all, n = 100000, 64
    rng = np.random.default_rng(20250503)
    etas_all = 10.0 ** rng.uniform(1.0, 3.0, n_all)
    f_all = rng.uniform(0.5, 1.5, (n_all, 6))
    # the bench draws its members the same way
    bm, be = bench.ensemble_members(8)
    assert np.allclose(be, 10.0 ** np.random.default_rng(20250503).uniform(1.0, 3.0, 8))
    pick = np.sort(np.random.default_rng(64).choice(n_all, n, replace=False))
    base = NOTEBOOK_CONSTS["pol"]
    names = ("K1", "K2", "K3", "K4", "DT", "Kd")
    omodels = [dataclasses.replace(base, **{k: getattr(base, k) * f_all[m, q] for q, k in enumerate(names)}) for m in pick]
    etas = etas_all[pick]
    models = [mods["product_model"](dict(K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max,
                                         phi_l=om.phi_l, gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max,
                                         phi_d=om.phi_d, r_sp=om.r_sp, T_ref=om.T_ref, kind=2)) for om in omodels]
    grid = p1.make_uniform_grid(32, 32)
    og = OGrid(np.array(grid.x), np.array(grid.y))
    Tf, dt = 0.01, 5e-4
    e = ens.TrajectoryEnsemble(grid, mods["CASES"]["scp_fast1e1"], models, etas, chunk=n)
    res = e.run_for_errors(Tf, dt)
    assert res["nsteps"] == 20
    want = []
    for om, eta in zip(omodels, etas):
        oc = make_case("scp_fast1e1", om)
        r = run_trial(oc, og, om, float(eta), OForcing(oc, om, float(eta), og), Tf=Tf, dt=dt)
        want.append([r["overall"]] + [r["per_var"][v] for v in VARS])
    got = np.column_stack([res["overall"], res["per_var"]])
    assert close(got, np.array(want))


# Published in the reference's notebooks (cell 9 outputs): overall error per spatial level
# N = 2 .. 256 at dt = h^1.5, Tf = 0.01, eta = 50, and the observed rates printed beneath them.
NOTEBOOK_SPATIAL = {
    "expsin": ([1.942652829989e-05, 5.197056624911e-06, 1.322695968641e-06, 3.372248813359e-07,
                8.344194130557e-08, 2.052209700229e-08, 5.119616858484e-09, 1.278782670173e-09], 5e-11,
               [1.877, 1.975, 1.957, 2.012, 2.030, 2.004]),
    "pol": ([4.93452e-05, 1.59616e-05, 4.28269e-06, 1.08800e-06, 2.75006e-07, 6.96085e-08, 1.74802e-08,
             4.38284e-09], 5e-6, [None, None, None, None, None, 1.993]),
}
NOTEBOOK_MODEL = {
    "expsin": dict(K1=1e-3, K2=1e-3, K3=1e-3, K4=1e-3, DT=1e-3, Dl_max=1e-5, phi_l=1e-5, gamma_T=1e-9, Kd=1e-2,
                   Sd=1.0, Dd_max=1e-6, phi_d=1e-5, r_sp=5e-2, T_ref=300.0, kind=2),
    "pol": dict(K1=1e-3, K2=1e-3, K3=1e-3, K4=1e-3, DT=1e-3, Dl_max=8.01e-4, phi_l=1e-5, gamma_T=1e-9, Kd=1e-2,
                Sd=1.0, Dd_max=2.46e-6, phi_d=1e-5, r_sp=5e-2, T_ref=300.0, kind=2),
}


@pytest.mark.parametrize("case", ["expsin", "pol"])
def test_notebook_spatial_study_numbers_and_orders(mods, case):
    """The full spatial study of the notebooks (eight levels up to N = 256, launched as one sweep): the published
    error values and observed convergence orders are reproduced."""
    errors, rtol, rates = NOTEBOOK_SPATIAL[case]
    model = mods["product_model"](NOTEBOOK_MODEL[case])
    trials = [dict(N=n, dt=(1.0 / n) ** 1.5, Tf=0.01, eta=50.0) for n in (2, 4, 8, 16, 32, 64, 128, 256)]
    sw = mods["ens"].RefinementSweep(mods["CASES"][case], model, trials)
    got = sw.run_for_errors()["overall"]
    sw.close()
    # the error norm is a difference of nearly equal O(1) fields: some ulps of the fields (1e-14: it moves with
    # the number of solver sweeps) on top of the digits the notebook prints
    assert np.all(np.abs(got - np.array(errors)) <= rtol * np.array(errors) + 1e-14), (got, errors)
    mods["cvg"].VERBOSE = False
    got_rates = [r for r, status in mods["cvg"].calculate_observed_rates(list(got), 2.0)]
    for g, want in zip(got_rates, rates):
        if want is not None:
            assert abs(g - want) <= 6e-4, (got_rates, rates)   # printed with three decimals


def _library_cases():
    import inspect
    import prob1_mms_cases as cases
    import prob1base as p1
    return [n for n, cls in inspect.getmembers(cases, inspect.isclass)
            if issubclass(cls, p1.MMSCaseBase) and cls.__module__ == cases.__name__ and not n.startswith("_")
            and n != "MMSCaseNonFullySmoothPol"]


@pytest.mark.parametrize("name", _library_cases())
def test_every_library_case_device_forcing_equals_host_forcing(mods, name):
    """Each MMS case of the library, two PC steps: sources evaluated on the device from the case's tables /
    closed form vs the same step with the sources evaluated by the host callables (the reference's formulas)
    and uploaded.  Pins all device descriptions, not only the four cases that have reference fixtures."""
    import prob1_mms_cases as cases
    p1 = mods["p1"]
    model = mods["product_model"](NOTEBOOK_MODEL["pol"])
    grid = p1.make_uniform_grid(14, 11)
    eta, t0, dt = 50.0, 0.05, 2e-3
    case = getattr(cases, name)(grid=grid, model=model)
    out = []
    for host_sources in (False, True):
        forcing = p1.ForcingTerms_RegHCsTriple(mms_case=case, model=model, regularization_factor=eta)
        field = p1.SemiDiscreteField_RegHCsTriple(grid=grid, model=model, forcing_terms=forcing,
                                                  regularization_factor=eta)
        if host_sources:
            field.fcs = lambda t, xx, yy, f=forcing: f.fcs(t, xx, yy)   # a rebound callable: generic (array) path
        integ = p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple(field, regularization_factor=eta)
        s = p1.state_from_mms_when(mms_case=case, t=t0, grid=grid)
        t = t0
        for _ in range(2):
            s = integ.step(s, t0=t, dt=dt)
            t += dt
        out.append((s, field.binding().batch.mode))
    import ddcore
    assert out[0][1] in (ddcore.MODE_SEPARABLE, ddcore.MODE_EXPSIN) and out[1][1] == ddcore.MODE_ARRAYS, name
    everything = max(np.max(np.abs(getattr(out[1][0], v))) for v in VARS)
    for v in VARS:
        a, b = getattr(out[0][0], v), getattr(out[1][0], v)
        # (fields that are identically zero analytically hold rounding noise of the sources only)
        assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)) + 1e-16 * everything, (name, v)


def test_blown_up_trial_propagates_nan_like_the_reference(mods):
    """SlowlyChangingPeaks_Fast1e1 at dt = 1 (the reference's temporal study starts there and prints an overall
    error of 0.0): cs overflows to NaN in the first step, cd follows, cp / T / cl stay
    finite.  The reference's direct solves return NaN and `max(0.0, nan)` keeps 0.0; the device path must do the
    same - no NOT_CONVERGED for systems that are NaN on entry - and agree on the finite variables."""
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, make_case, run_trial
    p1 = mods["p1"]
    om = NOTEBOOK_CONSTS["pol"]
    N, Tf, dt, eta = 32, 10.0, 1.0, 50.0
    grid = p1.make_uniform_grid(N, N)
    og = OGrid(np.array(grid.x), np.array(grid.y))
    oc = make_case("scp_fast1e1", om)
    with np.errstate(all="ignore"):
        want = run_trial(oc, og, om, eta, OForcing(oc, om, eta, og), Tf=Tf, dt=dt)
    assert want["overall"] == 0.0 and want["per_var"]["cs"] == 0.0     # the fixture really blows up
    model = mods["product_model"](NOTEBOOK_MODEL["pol"])
    sw = mods["ens"].RefinementSweep(mods["CASES"]["scp_fast1e1"], model, [dict(N=N, dt=dt, Tf=Tf, eta=eta)])
    got = sw.run_for_errors()
    sw.close()
    assert got["overall"][0] == 0.0
    gv = got["per_var"][0]
    wv = np.array([want["per_var"][v] for v in VARS])
    assert gv[4] == 0.0
    assert np.all(np.abs(gv - wv) <= 1e-9 * np.abs(wv)), (gv, wv)


def test_device_combined_norms_equal_the_python_arithmetic_bit_for_bit(mods):
    """dd_run_pc_errors / dd_run_feuler_errors keep the per-step norms on the device and combine them there; the
    result must be what the reference's Python arithmetic (builtin sum(), trapezoid, max that skips NaN) gives on
    the same norms - compared exactly, on a healthy ensemble with per-member dt and on the blown-up trial."""
    import ddcore
    p1, ens = mods["p1"], mods["ens"]
    model = mods["product_model"](NOTEBOOK_MODEL["pol"])
    rng = np.random.default_rng(7)
    for N, B, Tf, dts, etas in ((12, 7, 0.004, 0.004 / 8 * rng.uniform(0.9, 1.0, 7), 10.0 ** rng.uniform(1, 3, 7)),
                                (32, 2, 10.0, np.array([1.0, 1.0]), np.array([50.0, 20.0]))):
        grid = p1.make_uniform_grid(N, N)
        nsteps = int(np.ceil(Tf / dts[0]))
        dt_used = np.full(B, Tf / nsteps) if N == 32 else dts
        b = ddcore.Batch(grid.x, grid.y, B)
        b.set_models([ddcore.model_struct(model, float(e)) for e in etas])
        b.forcing_spec(mods["CASES"]["scp_fast1e1"](grid=grid, model=model).device_spec())
        # a fixed sweep count makes two runs of the same trial produce the same fields bit for bit (the adaptive
        # plan learns between runs); the blown-up trial needs the adaptive plan and is compared to rounding level
        opt = ddcore.pc_options(fixed_sweeps=8) if N == 12 else ddcore.pc_options()
        for integrator in ("pc", "feuler"):
            b.fill_exact(0, 0.0)
            if integrator == "pc":
                _, norms, _ = b.run_pc(0, 1, 0.0, dt_used, nsteps, opt, norms=True)
            else:
                _, norms = b.run_feuler(0, 1, 0.0, dt_used, nsteps, norms=True)
            want = ens.combined_error_norms(norms, dt_used)
            b.fill_exact(0, 0.0)
            got, _ = b.run_errors(0, 1, 0.0, dt_used, nsteps, opt, integrator=integrator)
            if N == 12 or integrator == "feuler":
                assert np.array_equal(got["overall"], want["overall"]), (N, integrator, got["overall"], want["overall"])
                assert np.array_equal(got["per_var"], want["per_var"]), (N, integrator)
            else:
                assert np.array_equal(got["overall"] == 0, want["overall"] == 0)
                assert np.array_equal(got["per_var"] == 0, want["per_var"] == 0)
                assert np.allclose(got["per_var"], want["per_var"], rtol=1e-9, atol=0)
            if N == 32 and integrator == "pc":
                assert np.isnan(norms).any() and got["overall"][0] == 0.0        # the blow-up really is in this run
        b.close()


def _notebooks():
    import sys
    sys.path.insert(0, os.path.join(ROOT, "tools"))
    import notebook_studies
    return notebook_studies


@pytest.mark.parametrize("name", ["MMSCaseSlowlyChangingPeaks_Fast1e1", "MMSCaseNonFullySmoothPol_cpcsH1_TclcdH2",
                                  "MMSCaseNonFullySmoothPol_cpcsH2_TclcdH2", "MMSCaseNonFullySmoothPol_cpcsH2_TclcdH3"])
def test_notebook_tables_of_the_long_studies(mods, name):
    """The studies the notebooks run to Tf = 1 and Tf = 10 (up to 4096 steps at N = 256; 6 - 7 hours each on the
    author's workstation): published overall errors and final observed orders -- SlowlyChangingPeaks_Fast1e1
    2.092 (spatial) and 1.996 (temporal, including the two blown-up levels that print 0.0), NonFullySmoothPol
    H1/H2 1.054 (the expected order breakdown), H2/H2 4.482 and 2.065, H2/H3 1.961."""
    nbs = _notebooks()
    nb = nbs.NOTEBOOKS[name]
    res = nbs.run_notebook(name, which=("spatial", "temporal"))
    for key in ("spatial", "temporal"):
        r = res[key]
        pub, got = np.array(r["published"]), np.array(r["reproduced"])
        if name == "MMSCaseSlowlyChangingPeaks_Fast1e1" and key == "temporal":
            # Tf = 10 at dt = 1 .. 1/256 on N = 200: the coarse levels blow up or nearly do, and the numbers the
            # notebook prints for them are not what the reference's code computes in this environment (numpy 2.3 /
            # scipy 1.18 instead of the notebook's 2.2.6 / 1.15.2): the oracle, pinned to the live reference, gives
            # 5.589568384884182e-01 at dt = 1/4 and 1.558972019927606e-01 at dt = 1/8 (profiles/README.md) -- the
            # device reproduces those; the finer levels agree with the notebook to 0.1 - 1 % and the observed order
            # (published 1.996) is reproduced as 2.000.
            assert got[0] == 0.0
            assert abs(got[2] - 5.589568384884182e-01) <= 1e-9 and abs(got[3] - 1.558972019927606e-01) <= 1e-9, got
            assert np.all(np.abs(got[5:] - pub[5:]) <= 1e-2 * pub[5:]), (got, pub)
            assert abs(r["final_rate"] - r["published_final_rate"]) <= 1e-2, r["rates"]
            continue
        # error norms are differences of nearly equal O(1) fields: rounding of the fields (1e-13 after thousands
        # of steps) on top of the printed digits
        assert np.all(np.abs(got - pub) <= 2e-6 * pub + 2e-13), (key, got, pub)
        want = r["published_final_rate"]
        if want is None:
            assert not np.isfinite(r["final_rate"]), (key, r["rates"])
        else:
            assert abs(r["final_rate"] - want) <= 2e-3, (key, r["rates"], want)


@pytest.mark.parametrize("case,N,M", [("scp_fast1e1", 32, 32), ("expsin", 20, 27), ("nfsp_h1h2", 16, 16)])
def test_member_kernel_equals_batched_kernels(mods, case, N, M):
    """Small grids run whole trajectories in one CTA per member (csrc/dd_member.cu: the time loop on chip); the same
    members through the batched mesh kernels (the default) give the same final fields (1e-12), the same
    combined error norms and the same cs-Newton iteration counts."""
    p1, ddcore = mods["p1"], mods["ddcore"]
    model = mods["product_model"](NOTEBOOK_MODEL["expsin" if case == "expsin" else "pol"])
    grid = p1.Grid(np.linspace(0, 1, N + 1) ** 1.1, np.linspace(0, 1, M + 1))
    spec = mods["CASES"][case](grid=grid, model=model).device_spec()
    B, t0, dt, nsteps = 3, 0.05, 5e-4, 7
    out = {}
    for path in ("member", "batched"):
        if path == "member":
            os.environ["DD_MEMBER_KERNEL"] = "1"
        try:
            b = ddcore.Batch(grid.x, grid.y, B)
            b.set_models([ddcore.model_struct(model, eta) for eta in (10.0, 50.0, 400.0)])
            b.forcing_spec(spec)
            b.fill_exact(0, t0)
            res, st = b.run_errors(0, 1, t0, dt, nsteps)
            final = 0 if nsteps % 2 == 0 else 1
            out[path] = (res, st, [b.download(final, member=m) for m in range(B)])
            b.close()
        finally:
            os.environ.pop("DD_MEMBER_KERNEL", None)
    (ra, sa, fa), (rb, sb, fb) = out["member"], out["batched"]
    assert sa["cs_newton_iters"] == sb["cs_newton_iters"]
    assert max(sa["bound"]) <= 1e-13, sa
    for m in range(B):
        for v in VARS:
            assert rel_err(fa[m][v], fb[m][v]) <= 1e-12, (m, v)
    assert np.all(np.abs(ra["overall"] - rb["overall"]) <= 1e-9 * rb["overall"] + 1e-15)
    assert np.all(np.abs(ra["per_var"] - rb["per_var"]) <= 1e-9 * rb["per_var"] + 1e-15)


def test_member_kernel_cs_exit_test_fires(mods):
    """A manufactured cs that is non-zero on every node, boundary included: the reference's global exit test of the
    cs corrector (max |dx| < rtol |x| everywhere) can fire, and the one-CTA kernel stops at the same iteration as
    the batched kernels' decide / redo pair."""
    import prob1_mms_cases as p1mc
    p1, ddcore = mods["p1"], mods["ddcore"]
    t, x, y = p1.t_sym, p1.x_sym, p1.y_sym
    model = mods["product_model"](NOTEBOOK_MODEL["pol"])
    grid = p1.make_uniform_grid(14, 12)
    case = p1mc._SeparableCase(grid, model, phi_exprs=[1 / (1 + t)] * 5, phi_specs=[p1mc.PhiSpec("inv1pt", (1.0,))] * 5,
                               Xs=[x * (1 - x)] * 4 + [2 + x], Ys=[y * (1 - y)] * 4 + [2 + y])
    spec = case.device_spec()
    opt = ddcore.pc_options(num_newton_iterations=40, consec_xs_rtol=1e-6)
    out = {}
    for path in ("member", "batched"):
        if path == "member":
            os.environ["DD_MEMBER_KERNEL"] = "1"
        try:
            b = ddcore.Batch(grid.x, grid.y, 1)
            b.set_model(model, 5.0)
            b.forcing_spec(spec)
            b.fill_exact(0, 0.0)
            iters = []
            for k in range(3):   # one-step runs: the count of every step
                res, st = b.run_errors(k % 2, (k + 1) % 2, k * 1e-3, 1e-3, 1, opt)
                iters.append(st["cs_newton_iters"])
            out[path] = (iters, b.download(1))
            b.close()
        finally:
            os.environ.pop("DD_MEMBER_KERNEL", None)
    assert out["member"][0] == out["batched"][0], out
    assert min(out["member"][0]) < 40, out["member"][0]   # the test really fired
    for v in VARS:
        assert rel_err(out["member"][1][v], out["batched"][1][v]) <= 1e-12, v
