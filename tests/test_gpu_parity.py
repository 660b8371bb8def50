"""GPU parity tests (run with -m gpu on a B200): the CUDA path, called through the C ABI
(ddcore.Batch -> libdd_b200.so), against the golden fixtures of the reference and against the
oracle on the same seeded inputs.

Tolerances: fields 1e-12 norm-wise relative (north_star), cs-Newton iteration counts identical,
convergence-study error norms 1e-9 relative + 1e-13 absolute (differences of nearly equal fields).
"""
import numpy as np
import pytest

from golden_util import VARS, fixture_names, load_fixture, oracle_model, rel_err

pytestmark = pytest.mark.gpu

TOL = 1e-12


@pytest.fixture(scope="module")
def dd():
    import ddcore
    import prob1base as p1
    import prob1_mms_cases as p1mc
    from test_hostsim import CASES, product_model
    return dict(ddcore=ddcore, p1=p1, p1mc=p1mc, CASES=CASES, product_model=product_model)


def make_batch(dd, desc, z, nmembers=1):
    p1, ddcore = dd["p1"], dd["ddcore"]
    model = dd["product_model"](desc["model"])
    grid = p1.Grid(z["x"], z["y"])
    b = ddcore.Batch(grid.x, grid.y, nmembers)
    b.set_model(model, desc["eta"], desc.get("variant", "regh"))
    if desc["case"] is not None:
        case = dd["CASES"][desc["case"]](grid=grid, model=model)
        b.forcing_spec(case.device_spec())
    else:
        b.forcing_none()
    return b, model, grid


def pc_opts(dd, pc, variant="regh", **kw):
    if variant != "regh":  # CsTriple / HCsTriple: closed-form cs corrector, no Newton iterations
        pc = dict(pc, num_newton_iterations=0, consec_xs_rtol=0.0)
    return dd["ddcore"].pc_options(num_pc_steps=pc.get("num_pc_steps", 1), num_newton_steps=pc.get("num_newton_steps", 1),
                                   num_newton_iterations=pc.get("num_newton_iterations", 5),
                                   consec_xs_rtol=pc.get("consec_xs_rtol", 1e-6), **kw)


@pytest.mark.parametrize("name", fixture_names(kind="steps"))
def test_steps_match_reference(dd, name):
    desc, z = load_fixture(name)
    b, model, grid = make_batch(dd, desc, z)
    dt, t = desc["dt"], desc["t0"]
    s = {v: z["init_" + v] for v in VARS}
    b.upload(0, s)
    if desc["init"] == "exact":
        b.fill_exact(2, t)
        ex = b.download(2)
        for v in VARS:
            assert rel_err(ex[v], z["init_" + v]) <= 1e-13, f"exact {v}"
    b.eval_fields(0, 1, t)
    F = b.download(1)
    for v in VARS:
        assert rel_err(F[v], z["F0_" + v]) <= TOL, f"F0_{v}"
    cur, nxt = 0, 1
    for n in range(desc["nsteps"]):
        if desc["integrator"] == "pc":
            st = b.step_pc(cur, nxt, t, dt, pc_opts(dd, desc["pc"], desc.get("variant", "regh")))
            per_step = desc["pc"].get("num_pc_steps", 1)
            # the fixture counts calls over all pc steps; the last corrector's count is what stats report
            assert st["cs_newton_iters"] * per_step >= int(z["cs_newton_calls_per_step"][n]) or per_step > 1
            if per_step == 1:
                assert st["cs_newton_iters"] == int(z["cs_newton_calls_per_step"][n])
            # solve_tol is 1e-14; the accepted residual may add the rounding floor of its own evaluation,
            # 16 eps (|b| + |x|) / ((1 - rho) |v|), which reaches 1e-13 for the weakly dominant big-dt matrices
            assert max(st["bound"]) <= 5e-13
        else:
            b.step_feuler(cur, nxt, t, dt)
        t += dt
        cur, nxt = nxt, cur
        if f"step{n + 1}_cp" in z:
            got = b.download(cur)
            for v in VARS:
                assert rel_err(got[v], z[f"step{n + 1}_{v}"]) <= TOL, f"step {n + 1} {v}"
    b.close()
    if desc["integrator"] == "pc":
        # last_residual of the reference's integrator (src/prob1base.py:3041-3043, 3076-3078, 3111-3113): the same
        # steps through the reference-compatible classes, whose last_residual replays the last step's Newton pieces
        p1 = dd["p1"]
        variant = desc.get("variant", "regh")
        fcls, Fcls, icls = {
            "regh": (p1.ForcingTerms_RegHCsTriple, p1.SemiDiscreteField_RegHCsTriple,
                     p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple),
            "cs": (p1.ForcingTerms_CsTriple, p1.SemiDiscreteField_CsTriple,
                   p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_CsTriple),
            "h": (p1.ForcingTerms_HCsTriple, p1.SemiDiscreteField_HCsTriple,
                  p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_HCsTriple)}[variant]
        extra = dict(regularization_factor=desc["eta"]) if variant == "regh" else {}
        if desc["case"] is not None:
            forcing = fcls(mms_case=dd["CASES"][desc["case"]](grid=grid, model=model), model=model, **extra)
        else:
            forcing = p1.NoForcingTerms(grid)
        field = Fcls(grid=grid, model=model, forcing_terms=forcing, **extra)
        integ = icls(field, **extra, **desc["pc"])
        st8 = p1.StateVars(*[z["init_" + v] for v in VARS], model=model, hh=grid.hh, kk=grid.kk)
        t = desc["t0"]
        for n in range(desc["nsteps"]):
            st8 = integ.step(st8, t0=t, dt=dt)
            t += dt
        for v in VARS:
            assert rel_err(getattr(st8, v), z[f"step{desc['nsteps']}_{v}"]) <= TOL, f"class API step {v}"
        for v in ("T", "cl", "cd"):
            scale = max(np.max(np.abs(z[f"step{desc['nsteps']}_{v}"])), 1e-300)
            assert np.max(np.abs(integ.last_residual[v] - z["resid_" + v])) <= 1e-11 * scale, f"last_residual {v}"


@pytest.mark.parametrize("name", fixture_names(kind="trial"))
def test_trial_errors_match_reference(dd, name):
    import mms_trial_utils as mtu
    desc, z = load_fixture(name)
    p1 = dd["p1"]
    model = dd["product_model"](desc["model"])
    for li, lv in enumerate(desc["levels"]):
        eta = lv.get("eta", desc["eta"])
        grid = p1.make_uniform_grid(lv["N"], lv["M"])
        if desc["integrator"] == "pc":
            icls, ipar = p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple, dict(regularization_factor=eta, **desc["pc"])
        else:
            icls, ipar = p1.ForwardEulerIntegrator, {}
        trial = mtu.MMSTrial(grid=grid, model=model, mms_case_cls=dd["CASES"][desc["case"]],
                             field_cls=p1.SemiDiscreteField_RegHCsTriple,
                             forcing_terms_cls=p1.ForcingTerms_RegHCsTriple, integrator_cls=icls,
                             forcing_terms_params={"regularization_factor": eta},
                             field_params={"regularization_factor": eta}, integrator_params=ipar)
        summ = trial.run_for_errors(Tf=desc["Tf"], dt=lv["dt"])
        assert summ.dt_used == float(z[f"L{li}_dt_used"])
        ref = float(z[f"L{li}_overall"])
        assert abs(summ.overall_combined_error - ref) <= 1e-9 * ref + 1e-13, (li, summ.overall_combined_error, ref)
        ref_pv = z[f"L{li}_per_var"]
        got_pv = np.array([summ.per_variable_sup_errors[v] for v in VARS])
        assert np.all(np.abs(got_pv - ref_pv) <= 1e-9 * np.abs(ref_pv) + 1e-13)


@pytest.mark.parametrize("case,N,M", [("pol", 100, 130), ("expsin", 160, 96), ("scp_fast1e1", 129, 129)])
def test_multitile_step_matches_oracle(dd, case, N, M):
    """Grids larger than one solver tile: tiled red-black SOR with halos vs the oracle's SuperLU."""
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, PCStepper, exact_state, feuler_step, make_case
    p1, ddcore = dd["p1"], dd["ddcore"]
    om = NOTEBOOK_CONSTS["expsin" if case == "expsin" else "pol"]
    eta, t0 = 50.0, 0.1
    dt = (1.0 / max(N, M)) ** 1.5
    og = OGrid(np.linspace(0, 1, N + 1), np.linspace(0, 1, M + 1))
    oc = make_case(case, om)
    of = OForcing(oc, om, eta, og)
    s0 = exact_state(oc, t0, og)
    ref = PCStepper(og, om, eta, of, keep_residuals=False).step(s0, t0, dt)
    ref_fe = feuler_step(s0, t0, dt, og, om, eta, of)
    model = dd["product_model"](dict(K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max,
                                     phi_l=om.phi_l, gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max,
                                     phi_d=om.phi_d, r_sp=om.r_sp, T_ref=om.T_ref, kind=2))
    grid = p1.Grid(og.x, og.y)
    b = ddcore.Batch(grid.x, grid.y, 1)
    b.set_model(model, eta)
    b.forcing_spec(dd["CASES"][case](grid=grid, model=model).device_spec())
    b.upload(0, s0.fields())
    st = b.step_pc(0, 1, t0, dt)
    got = b.download(1)
    b.step_feuler(0, 2, t0, dt)
    got_fe = b.download(2)
    for v in VARS:
        assert rel_err(got[v], getattr(ref, v)) <= TOL, (v, st)
        assert rel_err(got_fe[v], getattr(ref_fe, v)) <= TOL, v
    # forced multi-pass: few sweeps per pass must give the same iterate as one pass (tiling-exactness)
    b.step_pc(0, 2, t0, dt, ddcore.pc_options(fixed_sweeps=st["sweeps"][0]))
    again = b.download(2)
    for v in VARS:
        assert rel_err(again[v], got[v]) <= 1e-14, v
    b.close()


def test_batch_members_match_individual_runs(dd):
    """Ensemble: members with different model constants / eta in one batch == the same members alone."""
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, PCStepper, exact_state, make_case
    p1, ddcore = dd["p1"], dd["ddcore"]
    N = M = 24
    rng = np.random.default_rng(20250503)
    base = NOTEBOOK_CONSTS["pol"]
    B, t0, dt, nsteps = 5, 0.0, 5e-4, 3
    grid = p1.make_uniform_grid(N, M)
    og = OGrid(grid.x, grid.y)
    models, omodels = [], []
    for m in range(B):
        kw = {k: getattr(base, k) * rng.uniform(0.5, 1.5) for k in ("K1", "K2", "K3", "K4", "DT", "Kd")}
        om = base.with_changes(**kw)
        omodels.append((om, float(10 ** rng.uniform(1, 3))))
        pm = dd["product_model"](dict(K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max,
                                      phi_l=om.phi_l, gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max,
                                      phi_d=om.phi_d, r_sp=om.r_sp, T_ref=om.T_ref, kind=2))
        models.append(ddcore.model_struct(pm, omodels[-1][1]))
    case = dd["CASES"]["scp_fast1e1"](grid=grid, model=dd["product_model"](dict(
        K1=base.K1, K2=base.K2, K3=base.K3, K4=base.K4, DT=base.DT, Dl_max=base.Dl_max, phi_l=base.phi_l,
        gamma_T=base.gamma_T, Kd=base.Kd, Sd=base.Sd, Dd_max=base.Dd_max, phi_d=base.phi_d, r_sp=base.r_sp,
        T_ref=base.T_ref, kind=2)))
    b = ddcore.Batch(grid.x, grid.y, B)
    b.set_models(models)
    b.forcing_spec(case.device_spec())
    b.fill_exact(0, t0)
    final, norms, st = b.run_pc(0, 1, t0, dt, nsteps, norms=True)
    oc = make_case("scp_fast1e1", base)
    for m in range(B):
        om, eta = omodels[m]
        s = exact_state(oc, t0, og)
        stepper = PCStepper(og, om, eta, OForcing(oc, om, eta, og), keep_residuals=False)
        t = t0
        for _ in range(nsteps):
            s = stepper.step(s, t, dt)
            t += dt
        got = b.download(final, member=m)
        for v in VARS:
            assert rel_err(got[v], getattr(s, v)) <= TOL, (m, v)
    b.close()


def test_class_api_step_and_residuals(dd):
    """The reference-compatible classes: one PC step, its pieces and last_residual vs the golden fixture."""
    desc, z = load_fixture("steps_expsin_8x8_pc")
    p1 = dd["p1"]
    model = dd["product_model"](desc["model"])
    grid = p1.Grid(z["x"], z["y"])
    case = dd["CASES"]["expsin"](grid=grid, model=model)
    eta = desc["eta"]
    forcing = p1.ForcingTerms_RegHCsTriple(mms_case=case, model=model, regularization_factor=eta)
    field = p1.SemiDiscreteField_RegHCsTriple(grid=grid, model=model, forcing_terms=forcing,
                                              regularization_factor=eta)
    integ = p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple(field, regularization_factor=eta)
    s = p1.state_from_mms_when(mms_case=case, t=desc["t0"], grid=grid)
    for v, F in zip(VARS, (field.Fcp, field.FT, field.Fcl, field.Fcd, field.Fcs)):
        assert rel_err(F(s, desc["t0"]), z["F0_" + v]) <= TOL
    t = desc["t0"]
    for n in range(desc["nsteps"]):
        s = integ.step(s, t0=t, dt=desc["dt"])
        t += desc["dt"]
        for v in VARS:
            assert rel_err(getattr(s, v), z[f"step{n + 1}_{v}"]) <= TOL
    for v in ("T", "cl", "cd"):
        scale = np.max(np.abs(z[f"step{desc['nsteps']}_{v}"]))
        assert np.max(np.abs(integ.last_residual[v] - z["resid_" + v])) <= 1e-11 * scale
    with pytest.raises(AssertionError):
        integ.step(s, t0=t, dt=-1.0)
    # host-evaluated sources (any ForcingTermsBase): rebinding one source switches to the generic path
    field2 = p1.SemiDiscreteField_RegHCsTriple(grid=grid, model=model, forcing_terms=forcing,
                                               regularization_factor=eta)
    field2.fcs = lambda t, xx, yy: forcing.fcs(t, xx, yy)
    integ2 = p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple(field2, regularization_factor=eta)
    s2 = integ2.step(p1.state_from_mms_when(mms_case=case, t=desc["t0"], grid=grid), t0=desc["t0"], dt=desc["dt"])
    for v in VARS:
        assert rel_err(getattr(s2, v), z[f"step1_{v}"]) <= TOL


def test_roundtrip_and_linearity_large(dd):
    """Size-independent properties at a size the oracle cannot reach quickly (2049 x 1025 nodes):
    upload/download round trip is exact; F is affine in the source-free T field for cp = 0, and the
    device error norm of the exact state is ~0."""
    p1, ddcore = dd["p1"], dd["ddcore"]
    N, M = 2048, 1024
    grid = p1.make_uniform_grid(N, M)
    model = dd["product_model"](dict(K1=1e-3, K2=1e-3, K3=1e-3, K4=1e-3, DT=1e-3, Dl_max=8.01e-4, phi_l=1e-5,
                                     gamma_T=1e-9, Kd=1e-2, Sd=1.0, Dd_max=2.46e-6, phi_d=1e-5, r_sp=5e-2,
                                     T_ref=300.0, kind=2))
    b = ddcore.Batch(grid.x, grid.y, 1)
    b.set_model(model, 50.0)
    case = dd["CASES"]["pol"](grid=grid, model=model)
    b.forcing_spec(case.device_spec())
    b.fill_exact(0, 0.2)
    got = b.download(0)
    ex = case.T(0.2, grid.xx, grid.yy)
    assert rel_err(got["T"], ex) <= 1e-14
    b.upload(1, got)
    again = b.download(1)
    for v in VARS:
        assert np.array_equal(again[v], got[v])
    norms = b.error_norms(0, 0.2)
    assert np.all(norms[0] <= 1e-28)
    # one PC step at the study's dt keeps the error at the discretisation level and meets the solve bound
    dt = (1.0 / N) ** 1.5
    st = b.step_pc(0, 1, 0.2, dt)
    assert max(st["bound"]) <= 1e-13
    e = b.error_norms(1, 0.2 + dt)
    assert np.sqrt(e[0, :5].sum()) < 1e-9
    b.close()


@pytest.mark.parametrize("world", [2, 3])
def test_slab_decomposition_single_gpu_emulation(dd, world):
    """The multi-rank mesh path (phased step + halo exchange + common sweep plan) with all slabs on one
    GPU: must reproduce the undecomposed run bit for bit (same global red-black iteration) and match the
    oracle to 1e-12."""
    import ddmesh
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, PCStepper, exact_state, make_case
    p1, ddcore = dd["p1"], dd["ddcore"]
    N, M = 150, 40
    om = NOTEBOOK_CONSTS["pol"]
    eta, t0, dt = 50.0, 0.05, 2e-4
    og = OGrid(np.linspace(0, 1, N + 1), np.linspace(0, 1, M + 1))
    oc = make_case("pol", om)
    model = dd["product_model"](dict(K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max,
                                     phi_l=om.phi_l, gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max,
                                     phi_d=om.phi_d, r_sp=om.r_sp, T_ref=om.T_ref, kind=2))
    grid = p1.Grid(og.x, og.y)
    spec = dd["CASES"]["pol"](grid=grid, model=model).device_spec()
    meshes = ddmesh.SlabMesh.local_group(grid.x, grid.y, world, halo=12)
    for m in meshes:
        m.batch.set_model(model, eta)
        m.batch.forcing_spec(spec)
        m.fill_exact(0, t0)
    nsteps = 3
    for k in range(nsteps):
        st = meshes[0].step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt)
    got = {v: np.concatenate([m.owned(nsteps % 2)[v] for m in meshes]) for v in VARS}
    # undecomposed run with the same sweep counts
    b = ddcore.Batch(grid.x, grid.y, 1)
    b.set_model(model, eta)
    b.forcing_spec(spec)
    b.fill_exact(0, t0)
    s = exact_state(oc, t0, og)
    stepper = PCStepper(og, om, eta, OForcing(oc, om, eta, og), keep_residuals=False)
    t = t0
    for k in range(nsteps):
        s = stepper.step(s, t, dt)
        t += dt
    for v in VARS:
        assert got[v].shape == (N + 1, M + 1)
        assert rel_err(got[v], getattr(s, v)) <= TOL, v
    e_slab = meshes[0].error_norms(nsteps % 2, t0 + nsteps * dt)
    assert np.all(np.isfinite(e_slab)) and max(st["bound"]) <= 1e-13, st
    # bitwise reproducibility across decompositions: 1 slab == `world` slabs for equal sweep plans
    for k in range(nsteps):
        b.step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, ddcore.pc_options(fixed_sweeps=4))
    meshes2 = ddmesh.SlabMesh.local_group(grid.x, grid.y, world, halo=12)
    for m in meshes2:
        m.batch.set_model(model, eta)
        m.batch.forcing_spec(spec)
        m.fill_exact(0, t0)
    for k in range(nsteps):
        meshes2[0].step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, ddcore.pc_options(fixed_sweeps=4))
    one = b.download(nsteps % 2)
    for v in VARS:
        many = np.concatenate([m.owned(nsteps % 2)[v] for m in meshes2])
        assert np.array_equal(many, one[v]), v
    e_one = b.error_norms(nsteps % 2, t0 + nsteps * dt)[0]
    e_many = meshes2[0].error_norms(nsteps % 2, t0 + nsteps * dt)
    assert np.allclose(e_many, e_one, rtol=1e-12, atol=0)


def test_inline_device_math_accuracy(dd):
    """dd_exp / dd_rcp on the device vs NumPy: <= 2 ulp over the ranges the scheme uses, exact special cases."""
    from _ddlib import Context, dptr
    ctx = Context.default()
    rng = np.random.default_rng(1)
    x = np.concatenate([rng.uniform(-50, 50, 200000), rng.uniform(-1e-3, 1e-3, 100000), rng.uniform(-699, 699, 100000),
                        np.array([0.0, -0.0, 1.0, -1.0, 699.999, -699.999, 710.0, -750.0, np.inf, -np.inf, np.nan])])
    e, r = np.empty_like(x), np.empty_like(x)
    ctx.check(ctx.lib.dd_probe_math(ctx.handle, len(x), dptr(x), dptr(e), dptr(r)), "probe")
    with np.errstate(all="ignore"):
        ref_e, ref_r = np.exp(x), 1.0 / x
    fin = np.isfinite(ref_e) & (ref_e > 1e-300)
    ulp = np.abs(e[fin] - ref_e[fin]) / np.spacing(ref_e[fin])
    assert ulp.max() <= 2.0, ulp.max()
    assert np.array_equal(np.isnan(e), np.isnan(ref_e)) and np.array_equal(np.isinf(e), np.isinf(ref_e))
    fr = np.isfinite(ref_r)
    assert np.all(np.abs(r[fr] - ref_r[fr]) <= np.abs(np.spacing(ref_r[fr])))


@pytest.mark.parametrize("name", ["steps_pol_12x9_pc", "steps_expsin_8x8_pc", "steps_scp_fast1e1_12x9_pc",
                                  "steps_nfsp_h1h2_8x8_fe"])
def test_fused_sources_path_matches_reference(dd, name, monkeypatch):
    """DD_FUSED_SOURCES=1 evaluates the MMS sources inside every kernel (no staged source arrays); both
    paths must reproduce the reference."""
    monkeypatch.setenv("DD_FUSED_SOURCES", "1")
    desc, z = load_fixture(name)
    b, model, grid = make_batch(dd, desc, z)
    dt, t = desc["dt"], desc["t0"]
    b.upload(0, {v: z["init_" + v] for v in VARS})
    cur, nxt = 0, 1
    for n in range(desc["nsteps"]):
        if desc["integrator"] == "pc":
            b.step_pc(cur, nxt, t, dt, pc_opts(dd, desc["pc"]))
        else:
            b.step_feuler(cur, nxt, t, dt)
        t += dt
        cur, nxt = nxt, cur
        got = b.download(cur)
        for v in VARS:
            assert rel_err(got[v], z[f"step{n + 1}_{v}"]) <= TOL, f"step {n + 1} {v}"
    b.close()


@pytest.mark.parametrize("case,N,M,solver", [("pol", 100, 130, "reg"), ("scp_fast1e1", 129, 129, "reg"),
                                             ("pol", 100, 130, "smem")])
def test_extrapolated_initial_iterate(dd, case, N, M, solver, monkeypatch):
    """Ping-pong stepping starts each SOR solve from the previous step's increment.  The converged result
    must not depend on that: 6 steps with and without it agree with the oracle's direct solves to TOL, and
    the extrapolated start gets there with fewer sweeps."""
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, PCStepper, exact_state, make_case
    p1, ddcore = dd["p1"], dd["ddcore"]
    if solver == "smem":
        monkeypatch.setenv("DD_SOLVER", "smem")  # read at every solve: the shared-memory tile kernel
    om = NOTEBOOK_CONSTS["pol"]
    eta, t0, nsteps = 50.0, 0.1, 6
    dt = (1.0 / max(N, M)) ** 1.5
    og = OGrid(np.linspace(0, 1, N + 1), np.linspace(0, 1, M + 1))
    oc = make_case(case, om)
    of = OForcing(oc, om, eta, og)
    s = exact_state(oc, t0, og)
    init = s.fields()
    stepper = PCStepper(og, om, eta, of, keep_residuals=False)
    t = t0
    for _ in range(nsteps):
        s = stepper.step(s, t, dt)
        t += dt
    ref = s.fields()
    model = dd["product_model"](dict(K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max,
                                     phi_l=om.phi_l, gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max,
                                     phi_d=om.phi_d, r_sp=om.r_sp, T_ref=om.T_ref, kind=2))
    grid = p1.Grid(og.x, og.y)
    total = {}
    for guess in (False, True):
        b = ddcore.Batch(grid.x, grid.y, 1)
        b.set_model(model, eta)
        b.forcing_spec(dd["CASES"][case](grid=grid, model=model).device_spec())
        b.upload(0, init)
        opt = ddcore.pc_options(extrapolate_guess=guess)
        cur, nxt, t, sweeps = 0, 1, t0, 0
        for n in range(nsteps):
            st = b.step_pc(cur, nxt, t, dt, opt)
            assert max(st["bound"]) <= 1e-13
            if n >= 3:
                sweeps += sum(st["sweeps"])
            t += dt
            cur, nxt = nxt, cur
        got = b.download(cur)
        for v in VARS:
            assert rel_err(got[v], ref[v]) <= TOL, (guess, v)
        total[guess] = sweeps
        b.close()
    assert total[True] <= total[False], total


def _pol_problem(dd, N, M):
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, PCStepper, exact_state, make_case
    p1 = dd["p1"]
    om = NOTEBOOK_CONSTS["pol"]
    og = OGrid(np.linspace(0, 1, N + 1), np.linspace(0, 1, M + 1))
    oc = make_case("pol", om)
    model = dd["product_model"](dict(K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max,
                                     phi_l=om.phi_l, gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max,
                                     phi_d=om.phi_d, r_sp=om.r_sp, T_ref=om.T_ref, kind=2))
    grid = p1.Grid(og.x, og.y)
    spec = dd["CASES"]["pol"](grid=grid, model=model).device_spec()
    stepper = PCStepper(og, om, 50.0, OForcing(oc, om, 50.0, og), keep_residuals=False)
    return dict(og=og, oc=oc, model=model, grid=grid, spec=spec, stepper=stepper, exact_state=exact_state)


def test_deferred_verification_single_gpu(dd):
    """Steps enqueued before the previous one is verified (three rotating slots): same fields as the
    verified-per-step path; a plan that is too short is rejected one step late and both steps are redone."""
    import ctypes as C
    ddcore = dd["ddcore"]
    N, M, t0, dt, nsteps = 90, 70, 0.05, 3e-4, 7
    P = _pol_problem(dd, N, M)
    s, t = P["exact_state"](P["oc"], t0, P["og"]), t0
    for _ in range(nsteps):
        s = P["stepper"].step(s, t, dt)
        t += dt
    for sabotage in (False, True):
        b = ddcore.Batch(P["grid"].x, P["grid"].y, 1, nslots=3)
        b.set_model(P["model"], 50.0)
        b.forcing_spec(P["spec"])
        b.fill_exact(0, t0)
        seen, retries = 0, 0
        for k in range(nsteps):
            if sabotage and k == 3:
                # a one-sweep plan cannot meet the bound: step 3 is rejected when step 4 has been enqueued
                b.ctx.check(b.lib.dd_batch_set_plan(b.handle, C.byref((C.c_int * 3)(1, 1, 1))), "set_plan")
            st = b.step_pc(k % 3, (k + 1) % 3, t0 + k * dt, dt, defer=True)
            assert (st is None) == (k == 0)
            if st is not None:
                seen += 1
                retries += st["retries"]
                assert max(st["bound"]) <= 5e-13
        st = b.flush()
        assert st is not None and b.flush() is None
        retries += st["retries"]
        assert seen == nsteps - 1 and (retries > 0) == sabotage
        got = b.download(nsteps % 3)
        for v in VARS:
            assert rel_err(got[v], getattr(s, v)) <= TOL, (sabotage, v)
        b.close()


def test_deferred_verification_slab_emulation(dd):
    """The same through the multi-rank driver (two slabs on one GPU), including a rejected step."""
    import ddmesh
    N, M, t0, dt, nsteps = 120, 40, 0.05, 2e-4, 6
    P = _pol_problem(dd, N, M)
    s, t = P["exact_state"](P["oc"], t0, P["og"]), t0
    for _ in range(nsteps):
        s = P["stepper"].step(s, t, dt)
        t += dt
    for sabotage in (False, True):
        meshes = ddmesh.SlabMesh.local_group(P["grid"].x, P["grid"].y, 2, halo=12, nslots=3)
        for m in meshes:
            m.batch.set_model(P["model"], 50.0)
            m.batch.forcing_spec(P["spec"])
            m.fill_exact(0, t0)
        retries = 0
        for k in range(nsteps):
            if sabotage and k == 2:
                for m in meshes:
                    m._ctl[0]["plan"] = [1, 1, 1]
            st = meshes[0].step_pc(k % 3, (k + 1) % 3, t0 + k * dt, dt, defer=True)
            assert (st is None) == (k == 0)
            retries += st["retries"] if st else 0
        st = meshes[0].flush()
        retries += st["retries"]
        # (with a 12-row halo the first plans are clamped, so an honest rejection may also happen unprovoked)
        assert (retries > 0 or not sabotage) and meshes[0].flush() is None
        got = {v: np.concatenate([m.owned(nsteps % 3)[v] for m in meshes]) for v in VARS}
        for v in VARS:
            assert rel_err(got[v], getattr(s, v)) <= TOL, (sabotage, v)


def test_handles_may_be_released_in_any_order(dd):
    """A garbage-collected host drops contexts and batches in arbitrary order: destroying the context first
    must not pull the stream from under a live batch."""
    from _ddlib import Context
    ddcore, p1 = dd["ddcore"], dd["p1"]
    grid = p1.make_uniform_grid(8, 8)
    ctx = Context(0)
    b = ddcore.Batch(grid.x, grid.y, 1, ctx=ctx)
    b.forcing_none()
    ctx.close()              # context first ...
    b.close()                # ... then the batch (frees the context with it)
    ctx2 = Context(0)
    b2 = ddcore.Batch(grid.x, grid.y, 1, ctx=ctx2)
    b2.close()
    ctx2.close()


@pytest.mark.parametrize("variant", ["cs", "h"])
def test_field_variants_through_the_class_api(dd, variant):
    """CsTriple / HCsTriple (reference src/prob1base.py:2842-2876, 3152-3430) through the reference-facing
    classes: F(state, t), one step and the corrector pieces against the reference's fixture."""
    p1 = dd["p1"]
    name = f"random_uniform_{variant}triple_pol_forcing"
    desc, z = load_fixture(name)
    model = dd["product_model"](desc["model"])
    grid = p1.Grid(z["x"], z["y"])
    case = dd["CASES"][desc["case"]](grid=grid, model=model)
    fcls, fldcls, icls = {
        "cs": (p1.ForcingTerms_CsTriple, p1.SemiDiscreteField_CsTriple,
               p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_CsTriple),
        "h": (p1.ForcingTerms_HCsTriple, p1.SemiDiscreteField_HCsTriple,
              p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_HCsTriple)}[variant]
    forcing = fcls(mms_case=case, model=model)
    field = fldcls(grid=grid, model=model, forcing_terms=forcing)
    integ = icls(field)
    s = p1.StateVars(*[z["init_" + v] for v in VARS], model=model, hh=grid.hh, kk=grid.kk)
    t, dt = desc["t0"], desc["dt"]
    for v, F in zip(VARS, (field.Fcp, field.FT, field.Fcl, field.Fcd, field.Fcs)):
        assert rel_err(F(s, t), z["F0_" + v]) <= TOL, f"F0_{v}"
    for n in range(desc["nsteps"]):
        s = integ.step(s, t0=t, dt=dt)
        t += dt
        for v in VARS:
            assert rel_err(getattr(s, v), z[f"step{n + 1}_{v}"]) <= TOL, (n, v)
    for v in ("T", "cl", "cd"):
        scale = max(np.max(np.abs(z[f"step{desc['nsteps']}_{v}"])), 1e-300)
        assert np.max(np.abs(integ.last_residual[v] - z["resid_" + v])) <= 1e-11 * scale
    if variant == "h":
        # the reference's positivity guard: a huge step makes 2 - dt Kd (Sd - cd1)(1 + cl1) negative
        with pytest.raises(ValueError):
            integ.corrector_cs_step(None, np.full(grid.full_shape, 50.0), np.full(grid.full_shape, -50.0),
                                    at_t0=s, t0=t, dt=10.0)


def test_pipelined_solver_many_members_equals_single_member_runs(dd):
    """More solver tiles than SMs (here: 200 members, one tile each) switches the five-array solves to the
    persistent TMA pipeline; every member must come out exactly as when it is stepped alone."""
    ddcore, p1 = dd["ddcore"], dd["p1"]
    P = _pol_problem(dd, 20, 24)
    B, t0, dt = 200, 0.05, 1e-3
    etas = np.linspace(10.0, 400.0, B)
    big = ddcore.Batch(P["grid"].x, P["grid"].y, B)
    big.set_models([ddcore.model_struct(P["model"], float(e)) for e in etas])
    big.forcing_spec(P["spec"])
    big.fill_exact(0, t0)
    opt = ddcore.pc_options(fixed_sweeps=4)   # same plan in both runs: results are comparable bit for bit
    for k in range(2):
        big.step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, opt)
    for m in (0, 77, 148, 199):
        one = ddcore.Batch(P["grid"].x, P["grid"].y, 1)
        one.set_model(P["model"], float(etas[m]))
        one.forcing_spec(P["spec"])
        one.fill_exact(0, t0)
        for k in range(2):
            one.step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, opt)
        a, b = big.download(0, member=m), one.download(0)
        for v in VARS:
            assert np.array_equal(a[v], b[v]), (m, v)
        one.close()
    big.close()


def test_pipelined_solver_on_slabs_equals_undecomposed_run(dd):
    """A mesh large enough for the persistent pipeline (more tiles than SMs) in two slabs on one GPU: bitwise
    equal to the undecomposed run for equal sweep plans."""
    import ddmesh
    ddcore = dd["ddcore"]
    N, M, t0, dt = 900, 700, 0.05, 2e-5
    P = _pol_problem(dd, N, M)
    opt = ddcore.pc_options(fixed_sweeps=5)
    one = ddcore.Batch(P["grid"].x, P["grid"].y, 1)
    one.set_model(P["model"], 50.0)
    one.forcing_spec(P["spec"])
    one.fill_exact(0, t0)
    meshes = ddmesh.SlabMesh.local_group(P["grid"].x, P["grid"].y, 2, halo=14)
    for m in meshes:
        m.batch.set_model(P["model"], 50.0)
        m.batch.forcing_spec(P["spec"])
        m.fill_exact(0, t0)
    for k in range(2):
        one.step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, opt)
        meshes[0].step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, opt)
    ref = one.download(0)
    for v in VARS:
        got = np.concatenate([m.owned(0)[v] for m in meshes])
        assert np.array_equal(got, ref[v]), v
    one.close()


@pytest.mark.parametrize("M", [40, 150])
def test_slab_solves_in_segments_for_large_time_steps(dd, M):
    """Matrices of a large time step (weak diagonal dominance: the constants of the steps_scp_10x10_pc_bigdt fixture)
    need more SOR sweeps than a shallow halo supports between two exchanges.  The slab driver then runs the solve in
    segments and exchanges the iterate's halo rows in between (dd_pc_solve_segment): three slabs with a halo of 7
    rows (2 sweeps per segment) reproduce the oracle, and at a fixed plan of 9 sweeps the undecomposed run bit for
    bit.  M = 40: tile kernels; M = 150: marching kernels and the lane-private marching solver (forced: grids this
    small run the tile kernels by default)."""
    import os
    if M == 150:
        os.environ["DD_LANE"] = "1"
    try:
        _segments_case(dd, M)
    finally:
        os.environ.pop("DD_LANE", None)


def _segments_case(dd, M):
    import ddmesh
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, PCStepper, exact_state, make_case
    p1, ddcore = dd["p1"], dd["ddcore"]
    N = 60
    om = NOTEBOOK_CONSTS["pol"].with_changes(DT=0.5, Dl_max=0.3, Dd_max=0.2)
    eta, t0, dt, world = 50.0, 0.0, 2e-3, 3
    og = OGrid(np.linspace(0, 1, N + 1), np.linspace(0, 1, M + 1))
    oc = make_case("scp_fast1e1", om)
    model = dd["product_model"](dict(K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max,
                                     phi_l=om.phi_l, gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max,
                                     phi_d=om.phi_d, r_sp=om.r_sp, T_ref=om.T_ref, kind=2))
    grid = p1.Grid(og.x, og.y)
    spec = dd["CASES"]["scp_fast1e1"](grid=grid, model=model).device_spec()

    def slabs():
        meshes = ddmesh.SlabMesh.local_group(grid.x, grid.y, world, halo=7)
        for m in meshes:
            m.batch.set_model(model, eta)
            m.batch.forcing_spec(spec)
            m.fill_exact(0, t0)
        return meshes

    meshes = slabs()
    assert meshes[0].sweep_limit == 2
    nsteps = 3
    for k in range(nsteps):
        st = meshes[0].step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt)
    assert max(st["sweeps"]) > meshes[0].sweep_limit, st     # the solves really ran in segments
    if M == 150:
        # dt Dd / k^2 = 9: SOR diverges on the nonsymmetric cd system, the driver has fallen back to Gauss-Seidel
        assert meshes[0]._gs_only()[2] and not meshes[0]._gs_only()[0], (meshes[0]._gs_only(), st)
    got = {v: np.concatenate([m.owned(nsteps % 2)[v] for m in meshes]) for v in VARS}
    s = exact_state(oc, t0, og)
    stepper = PCStepper(og, om, eta, OForcing(oc, om, eta, og), keep_residuals=False)
    for k in range(nsteps):
        s = stepper.step(s, t0 + k * dt, dt)
    for v in VARS:
        assert rel_err(got[v], getattr(s, v)) <= TOL, (v, st)
    # bit-for-bit equality does not need a converged solve: 9 sweeps (segments of 2, 2, 2, 2, 1), accepted at a
    # loose tolerance (a fixed plan that fails the residual bound is an error in both drivers)
    fixed = ddcore.pc_options(fixed_sweeps=9, solve_tol=1.0)
    meshes2 = slabs()
    b = ddcore.Batch(grid.x, grid.y, 1)
    b.set_model(model, eta)
    b.forcing_spec(spec)
    b.fill_exact(0, t0)
    for k in range(2):
        meshes2[0].step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, fixed)
        b.step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, fixed)
    one = b.download(0)
    for v in VARS:
        many = np.concatenate([m.owned(0)[v] for m in meshes2])
        assert np.array_equal(many, one[v]), v
    b.close()


def test_sor_divergence_falls_back_to_gauss_seidel(dd):
    """The over-relaxation factor is the optimal one of a symmetric matrix; on the cd system of a very large time
    step (dt Dd / k^2 = 9, constants of the segment test at M = 150) SOR diverges although the matrix is strictly
    diagonally dominant.  The step controller notices (three times the theoretical sweep count fails the residual
    bound), switches that variable to Gauss-Seidel, which converges for every such matrix, and the step matches the
    oracle's direct solve."""
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, PCStepper, exact_state, make_case
    p1, ddcore = dd["p1"], dd["ddcore"]
    N, M = 60, 150
    om = NOTEBOOK_CONSTS["pol"].with_changes(DT=0.5, Dl_max=0.3, Dd_max=0.2)
    eta, t0, dt = 50.0, 0.0, 2e-3
    og = OGrid(np.linspace(0, 1, N + 1), np.linspace(0, 1, M + 1))
    oc = make_case("scp_fast1e1", om)
    model = dd["product_model"](dict(K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max,
                                     phi_l=om.phi_l, gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max,
                                     phi_d=om.phi_d, r_sp=om.r_sp, T_ref=om.T_ref, kind=2))
    grid = p1.Grid(og.x, og.y)
    b = ddcore.Batch(grid.x, grid.y, 1)
    b.set_model(model, eta)
    b.forcing_spec(dd["CASES"]["scp_fast1e1"](grid=grid, model=model).device_spec())
    b.fill_exact(0, t0)
    s = exact_state(oc, t0, og)
    stepper = PCStepper(og, om, eta, OForcing(oc, om, eta, og), keep_residuals=False)
    for k in range(2):
        st = b.step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt)
        s = stepper.step(s, t0 + k * dt, dt)
    got = b.download(0)
    for v in VARS:
        assert rel_err(got[v], getattr(s, v)) <= TOL, (v, st)
    # cd: far more sweeps than SOR's theoretical count for rho = 0.91 (about 50); T and cl stayed on SOR
    assert st["sweeps"][2] > 150 and st["sweeps"][1] < 150, st
    b.close()
