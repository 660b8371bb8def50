"""GPU tests (-m gpu) of generated forcing (DD_MODE_PROGRAM): a manufactured solution that no table form covers
runs on the device from its NVRTC-compiled SymPy expressions and must agree with the same steps driven by the
host callables (the reference's lambdified expressions + forcing object, src/prob1base.py:1226-1280, 3503-3551).
Tolerance on the fields: 1e-12 relative to the field's maximum."""
import numpy as np
import pytest

from golden_util import VARS, fixture_names, load_fixture, rel_err
from test_program_codegen import MODEL, nonseparable_exprs

pytestmark = pytest.mark.gpu


@pytest.fixture(scope="module")
def env():
    import cvg_studies_base as cvg
    import ddcore
    import ddensemble
    import ddmesh
    import mms_trial_utils as mtu
    import prob1base as p1
    from test_hostsim import product_model

    class NonSeparableCase(p1.MMSCaseSymbolic):
        def __init__(self, *, grid, model):
            super().__init__(grid=grid, model=model, **nonseparable_exprs())

    return dict(p1=p1, ddcore=ddcore, ens=ddensemble, mesh=ddmesh, mtu=mtu, cvg=cvg, product_model=product_model,
                Case=NonSeparableCase)


def _integrator(env, grid, model, case, eta, host_sources, integrator="pc"):
    p1 = env["p1"]
    forcing = p1.ForcingTerms_RegHCsTriple(mms_case=case, model=model, regularization_factor=eta)
    field = p1.SemiDiscreteField_RegHCsTriple(grid=grid, model=model, forcing_terms=forcing,
                                              regularization_factor=eta)
    if host_sources:
        field.fcs = lambda t, xx, yy, f=forcing: f.fcs(t, xx, yy)   # a rebound callable: generic (array) path
    if integrator == "pc":
        return p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple(field, regularization_factor=eta), field
    return p1.ForwardEulerIntegrator(field), field


@pytest.mark.parametrize("integrator", ["pc", "feuler"])
def test_program_steps_equal_host_forcing_steps(env, integrator):
    p1, ddcore = env["p1"], env["ddcore"]
    model = env["product_model"](MODEL)
    grid = p1.Grid(np.linspace(0, 1, 15) ** 1.1, np.linspace(0, 1, 12))
    case = env["Case"](grid=grid, model=model)
    eta, t0, dt = 50.0, 0.05, 2e-3
    out = []
    for host_sources in (False, True):
        integ, field = _integrator(env, grid, model, case, eta, host_sources, integrator)
        s = p1.state_from_mms_when(mms_case=case, t=t0, grid=grid)
        t = t0
        for _ in range(3):
            s = integ.step(s, t0=t, dt=dt)
            t += dt
        out.append((s, field.binding().batch.mode))
    assert out[0][1] == ddcore.MODE_PROGRAM and out[1][1] == ddcore.MODE_ARRAYS
    for v in VARS:
        a, b = getattr(out[0][0], v), getattr(out[1][0], v)
        assert np.max(np.abs(a - b)) <= 1e-12 * np.max(np.abs(b)), v
    # F(state, t) of the field, and the exact state filled by the program
    integ, field = _integrator(env, grid, model, case, eta, False)
    s = p1.state_from_mms_when(mms_case=case, t=t0, grid=grid)
    integ_h, field_h = _integrator(env, grid, model, case, eta, True)
    for n in ("Fcp", "FT", "Fcl", "Fcd", "Fcs"):
        a, b = getattr(field, n)(s, t0), getattr(field_h, n)(s, t0)
        assert np.max(np.abs(a - b)) <= 1e-12 * max(np.max(np.abs(b)), 1e-3), n
    assert field.binding().configure(t0, dt)          # the grid's batch is shared: back from ARRAYS to the program
    b = field.binding().batch
    assert b.mode == ddcore.MODE_PROGRAM
    b.fill_exact(1, t0 + 0.125)
    ex = b.download(1)
    for v in VARS:
        assert np.max(np.abs(ex[v] - getattr(case, v)(t0 + 0.125, grid.xx, grid.yy))) <= 4e-16, v


def test_program_whole_trial_on_device_equals_host_driven_trial(env):
    """run_simulation_collect_data: time loop, exact solution and error norms on the device (program mode) vs
    the class API step by step with host sources and host norms."""
    p1, mtu, ddcore = env["p1"], env["mtu"], env["ddcore"]
    model = env["product_model"](MODEL)
    grid = p1.make_uniform_grid(24, 20)
    case = env["Case"](grid=grid, model=model)
    eta, Tf, dt = 50.0, 0.02, 2.5e-3
    series = []
    for host_sources in (False, True):
        integ, field = _integrator(env, grid, model, case, eta, host_sources)
        initial = p1.state_from_mms_when(mms_case=case, t=0.0, grid=grid)
        ser, dt_used = mtu.run_simulation_collect_data(
            grid=grid, integrator=integ, exact_sol_pack=case, initial_state=initial, Tf=Tf, dt=dt, t0=0.0,
            variable_names=list(VARS), integral_vars=["T", "cl", "cd"])
        assert field.binding().batch.mode == (ddcore.MODE_ARRAYS if host_sources else ddcore.MODE_PROGRAM)
        series.append((ser, dt_used))
    assert series[0][1] == series[1][1] and len(series[0][0]) == len(series[1][0]) >= 9
    e = [mtu.calculate_combined_error_norm(s, d, ["T", "cl", "cd"]) for s, d in series]
    assert e[0] > 0 and abs(e[0] - e[1]) <= 1e-9 * e[1] + 1e-13, e
    for a, b in zip(*[s for s, _ in series]):
        for v in VARS:
            assert abs(a.h_norm_sq_errors[v] - b.h_norm_sq_errors[v]) <= 1e-8 * b.h_norm_sq_errors[v] + 1e-24, v


def test_program_sweep_ensemble_and_slabs(env):
    p1, ens, mtu, ddmesh = env["p1"], env["ens"], env["mtu"], env["mesh"]
    model = env["product_model"](MODEL)
    # refinement sweep: every trial against the single-trial device path
    trials = [dict(N=8, dt=2e-3, Tf=0.01, eta=50.0), dict(N=16, dt=1e-3, Tf=0.01, eta=50.0),
              dict(N=16, dt=1e-3, Tf=0.01, eta=300.0)]
    sw = ens.RefinementSweep(env["Case"], model, trials)
    got = sw.run_for_errors()["overall"]
    sw.close()
    for k, tr in enumerate(trials):
        grid = p1.make_uniform_grid(tr["N"], tr["N"])
        case = env["Case"](grid=grid, model=model)
        integ, _ = _integrator(env, grid, model, case, tr["eta"], False)
        initial = p1.state_from_mms_when(mms_case=case, t=0.0, grid=grid)
        ser, dt_used = mtu.run_simulation_collect_data(
            grid=grid, integrator=integ, exact_sol_pack=case, initial_state=initial, Tf=tr["Tf"], dt=tr["dt"], t0=0.0,
            variable_names=list(VARS), integral_vars=["T", "cl", "cd"])
        want = mtu.calculate_combined_error_norm(ser, dt_used, ["T", "cl", "cd"])
        assert abs(got[k] - want) <= 1e-9 * want + 1e-13, (k, got[k], want)
    # ensemble with per-member constants: member 1 differs in Kd, DT and eta
    grid = p1.make_uniform_grid(12, 12)
    other = env["product_model"](dict(MODEL, Kd=3e-2, DT=2e-3))
    e = ens.TrajectoryEnsemble(grid, env["Case"], [model, other], [50.0, 20.0])
    res = e.run_for_errors(0.01, 1e-3)
    for k, (md, eta) in enumerate(((model, 50.0), (other, 20.0))):
        case = env["Case"](grid=grid, model=md)
        integ, _ = _integrator(env, grid, md, case, eta, False)
        initial = p1.state_from_mms_when(mms_case=case, t=0.0, grid=grid)
        ser, dt_used = mtu.run_simulation_collect_data(
            grid=grid, integrator=integ, exact_sol_pack=case, initial_state=initial, Tf=0.01, dt=1e-3, t0=0.0,
            variable_names=list(VARS), integral_vars=["T", "cl", "cd"])
        want = mtu.calculate_combined_error_norm(ser, dt_used, ["T", "cl", "cd"])
        assert abs(res["overall"][k] - want) <= 1e-9 * want + 1e-13, (k, res["overall"][k], want)
    # slab decomposition: the program evaluates each slab's own rows (row0 offset); 3 slabs == 1 batch bitwise
    N, M, t0, dt = 150, 40, 0.05, 2e-4
    grid = p1.make_uniform_grid(N, M)
    prog = env["Case"](grid=grid, model=model).device_program()
    opts = env["ddcore"].pc_options(fixed_sweeps=4)
    meshes = ddmesh.SlabMesh.local_group(grid.x, grid.y, 3, halo=12)
    for m in meshes:
        m.batch.set_model(model, 50.0)
        m.batch.forcing_program(prog)
        m.fill_exact(0, t0)
    one = env["ddcore"].Batch(grid.x, grid.y, 1)
    one.set_model(model, 50.0)
    one.forcing_program(prog)
    one.fill_exact(0, t0)
    for k in range(3):
        meshes[0].step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, opts)
        one.step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, opts)
    ref = one.download(1)
    for v in VARS:
        assert np.array_equal(np.concatenate([m.owned(1)[v] for m in meshes]), ref[v]), v
    assert np.allclose(meshes[0].error_norms(1, t0 + 3 * dt), one.error_norms(1, t0 + 3 * dt)[0], rtol=1e-12, atol=0)


def test_analytic_callable_case_steps_through_the_array_path(env):
    """MMSCaseFromAnalytic (opaque callables, finite-difference derivatives): its forcing is evaluated on the host and
    uploaded (ARRAYS mode); the steps must agree with the same solution given symbolically (generated program) up
    to the finite-difference error of the sources (eps = 1e-6: ~1e-3 in the Laplacians, times dt)."""
    import sympy
    p1, ddcore = env["p1"], env["ddcore"]
    t, x, y = p1.t_sym, p1.x_sym, p1.y_sym
    model = env["product_model"](MODEL)
    grid = p1.make_uniform_grid(12, 10)
    exprs = {k[:-9]: v for k, v in nonseparable_exprs().items()}
    exprs["cs"] = sympy.sin(sympy.pi * (x + y * t)) * sympy.exp(-t) / 4          # smooth (no kink for the differences)
    sym = p1.MMSCaseSymbolic(grid=grid, model=model, **{k + "_sym_expr": v for k, v in exprs.items()})
    ana = p1.MMSCaseFromAnalytic(model, grid=grid, **{k + "_base": sympy.lambdify([t, x, y], v, "numpy")
                                                      for k, v in exprs.items()})
    eta, t0, dt = 50.0, 0.1, 1e-3
    out = []
    for case, mode in ((sym, ddcore.MODE_PROGRAM), (ana, ddcore.MODE_ARRAYS)):
        integ, field = _integrator(env, grid, model, case, eta, False)
        s = p1.state_from_mms_when(mms_case=sym, t=t0, grid=grid)
        tt = t0
        for _ in range(3):
            s = integ.step(s, t0=tt, dt=dt)
            tt += dt
        assert field.binding().batch.mode == mode
        out.append(s)
    for v in VARS:
        a, b = getattr(out[1], v), getattr(out[0], v)
        assert np.max(np.abs(a - b)) <= 2e-5 * max(np.max(np.abs(b)), 1e-3), v
        assert np.max(np.abs(a - b)) > 0 or v == "cp"      # (the two runs really used different sources)


@pytest.mark.parametrize("name", fixture_names(kind="steps", program=True))
def test_program_steps_match_the_reference(env, name):
    """Pins the generated forcing programs to the REFERENCE (not to this package's own array mode): per-step fields
    of the reference's MMSCaseSymbolic on the non-separable expressions, written by oracle/make_golden.py from the
    live reference, against the class API here, which runs the case from an NVRTC-compiled program."""
    import sympy
    desc, z = load_fixture(name)
    p1, ddcore = env["p1"], env["ddcore"]
    # the fixture was generated from the same expressions the tests here use
    plain = {p1.t_sym: sympy.Symbol("t"), p1.x_sym: sympy.Symbol("x"), p1.y_sym: sympy.Symbol("y")}
    for k, v in nonseparable_exprs().items():
        stored = sympy.sympify(desc["exprs"][k[:-9]])
        assert sympy.simplify(v.subs(plain) - stored) == 0, k
    model = env["product_model"](desc["model"])
    grid = p1.Grid(z["x"], z["y"])
    case = env["Case"](grid=grid, model=model)
    integ, field = _integrator(env, grid, model, case, desc["eta"], False, "pc" if desc["integrator"] == "pc" else "fe")
    t, dt = desc["t0"], desc["dt"]
    s = p1.state_from_mms_when(mms_case=case, t=t, grid=grid)
    for v in VARS:
        assert rel_err(getattr(s, v), z["init_" + v]) <= 1e-13, v
    for v, F in zip(VARS, (field.Fcp, field.FT, field.Fcl, field.Fcd, field.Fcs)):
        assert rel_err(F(s, t), z["F0_" + v]) <= 1e-12, v
    for n in range(desc["nsteps"]):
        s = integ.step(s, t0=t, dt=dt)
        t += dt
        for v in VARS:
            assert rel_err(getattr(s, v), z[f"step{n + 1}_{v}"]) <= 1e-12, (n, v)
    assert field.binding().batch.mode == ddcore.MODE_PROGRAM
    if desc["integrator"] == "pc":
        assert integ.last_stats["cs_newton_iters"] == int(z["cs_newton_calls_per_step"][-1])
        for v in ("T", "cl", "cd"):
            scale = np.max(np.abs(z[f"step{desc['nsteps']}_{v}"]))
            assert np.max(np.abs(integ.last_residual[v] - z["resid_" + v])) <= 1e-11 * scale
