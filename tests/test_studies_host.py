"""CPU tests of the study / ensemble host layer: observed rates against the reference driver's golden
values, the vectorised combined error norm against the scalar one, cost balancing, and the world_size-2
gathers over gloo."""
import json
import multiprocessing as mp
import os
import socket

import numpy as np
import pytest

from golden_util import load_fixture


def test_observed_rates_match_reference_driver():
    import cvg_studies_base as cvg
    desc, z = load_fixture("cvg_study_pol")
    for k, (errs, fac) in enumerate(zip(desc["rate_cases"], desc["rate_factors"])):
        if f"rates{k}_raises" in z:
            with pytest.raises(ZeroDivisionError):
                cvg.calculate_observed_rates(errs, fac)
            assert str(z[f"rates{k}_raises"]) == "ZeroDivisionError"
            continue
        got = cvg.calculate_observed_rates(errs, fac)
        assert [s for _, s in got] == json.loads(str(z[f"rates{k}_status"]))
        np.testing.assert_array_equal(np.array([r for r, _ in got]), z[f"rates{k}"])  # NaN == NaN here
    with pytest.raises(AssertionError):
        cvg.calculate_observed_rates([1.0, 0.5])
    with pytest.raises(AssertionError):
        cvg.calculate_observed_rates([1.0, 0.5, 0.25], 1.0)
    with pytest.raises(AssertionError):
        cvg.calculate_observed_rates([1.0, -0.5, 0.25])
    assert cvg.RateStatus.OK == "OK" and cvg.TimeStepData._fields == ("t", "h_norm_sq_errors",
                                                                      "grad_h_norm_p_sq_errors")


def test_vectorised_combined_norm_equals_scalar_one():
    import ddensemble
    import mms_trial_utils as mtu
    rng = np.random.default_rng(20250503)
    K, B = 9, 7
    norms = rng.random((K, B, 8)) * 1e-6
    norms[3, 2, 1] = np.nan          # a NaN never replaces the running maximum
    norms[:, 5, :] = 0.0
    dt = rng.random(B) * 1e-2 + 1e-3
    res = ddensemble.combined_error_norms(norms, dt)
    names, integral = list(mtu.VARS), ["T", "cl", "cd"]
    for m in range(B):
        series = mtu._series_from_norms(np.arange(K) * dt[m], norms[:, m, :], names, integral)
        want = mtu.NumericalErrorSummary(dt[m], series, names, integral)
        assert res["overall"][m] == want.overall_combined_error or (
            np.isnan(res["overall"][m]) and np.isnan(want.overall_combined_error))
        for v, name in enumerate(names):
            a, b = res["per_var"][m, v], want.per_variable_sup_errors[name]
            assert a == b or (np.isnan(a) and np.isnan(b)), (m, name, a, b)


def test_device_combination_arithmetic_equals_python_on_the_host():
    """csrc/dd_combine.cuh (what k_combine_update / k_combine_final run after every step of dd_run_*_errors),
    compiled for the host: bit for bit the scalar Python arithmetic, on magnitudes spread over many decades (so
    that the Neumaier correction of sum() matters), with NaN, Inf and all-zero members."""
    import ctypes as C
    import ddensemble
    import hostsim_util as hu
    import mms_trial_utils as mtu
    rng = np.random.default_rng(11)
    K, B = 23, 40
    norms = rng.random((K, B, 8)) * 10.0 ** rng.integers(-18, 2, (K, B, 8))
    norms[7:, 3, 4] = np.nan
    norms[5, 4, 6] = np.nan
    norms[9:, 6, 2] = np.inf
    norms[:, 8, :] = 0.0
    norms[0, 9, :] = 0.0
    dt = rng.random(B) * 1e-2 + 1e-4
    out = np.zeros((B, 6))
    dp = lambda a: a.ctypes.data_as(C.POINTER(C.c_double))
    series = np.ascontiguousarray(norms)
    assert hu.lib().hs_combine(dp(series), K, B, dp(dt), B, dp(out)) == 0
    res = ddensemble.combined_error_norms(norms, dt)
    assert np.array_equal(out[:, 0], res["overall"], equal_nan=True)
    assert np.array_equal(out[:, 1:], res["per_var"], equal_nan=True)
    names, integral = list(mtu.VARS), ["T", "cl", "cd"]
    for m in (0, 3, 4, 6, 8, 9):
        ser = mtu._series_from_norms(np.arange(K) * dt[m], norms[:, m, :], names, integral)
        want = mtu.NumericalErrorSummary(dt[m], ser, names, integral)
        assert out[m, 0] == want.overall_combined_error or (np.isnan(out[m, 0]) and np.isnan(want.overall_combined_error))
    # one dt for all members
    one = np.array([dt[0]])
    out1 = np.zeros((B, 6))
    assert hu.lib().hs_combine(dp(series), K, B, dp(one), 1, dp(out1)) == 0
    res1 = ddensemble.combined_error_norms(norms, dt[0])
    assert np.array_equal(out1[:, 0], res1["overall"], equal_nan=True)


def test_steps_and_dt_follow_the_trial_rule():
    import ddensemble
    assert ddensemble.steps_and_dt(0.01, (1 / 256) ** 1.5) == (41, 0.01 / 41)
    assert ddensemble.steps_and_dt(0.01, 0.3536) == (1, 0.01)
    assert ddensemble.steps_and_dt(1.0, 0.25, 0.5) == (2, 0.25)


def test_sweep_groups_and_balance():
    import ddensemble
    import prob1_mms_cases as p1mc
    trials = [dict(N=n, dt=(1.0 / n) ** 1.5, Tf=0.01, eta=50.0) for n in (2, 4, 8, 16, 32, 64, 128, 256)]
    trials += [dict(N=256, dt=1e-2 / 2 ** k, Tf=0.01, eta=50.0) for k in range(4)]
    trials += [dict(N=32, dt=5e-4, Tf=0.01, eta=e) for e in (10, 50, 100, 200, 300, 500, 1000)]
    sw = ddensemble.RefinementSweep(p1mc.MMSCasePol, None, trials, world=4, rank=1)
    assert len(sw.trials) == 19
    assert sum(len(g["members"]) for g in sw.groups) == 19
    eta_group = [g for g in sw.groups if g["key"] == (32, 32, 20)]
    assert len(eta_group) == 1 and len(eta_group[0]["members"]) == 7   # the eta study shares one batch
    assert [t["nsteps"] for t in sw.trials[:8]] == [1, 1, 1, 1, 2, 6, 15, 41]
    # every rank computes the same assignment; the heaviest group sits alone on its rank
    costs = [g["cost"] for g in sw.groups]
    owner = ddensemble.RefinementSweep.balance(costs, 4)
    assert owner == sw.assignment
    heavy = int(np.argmax(costs))
    assert sum(1 for o in owner if o == owner[heavy]) == 1
    loads = [sum(c for c, o in zip(costs, owner) if o == r) for r in range(4)]
    assert max(loads) == costs[heavy]
    assert ddensemble.RefinementSweep.balance([5, 5, 5, 5], 2) == [0, 1, 0, 1]


def _free_port():
    with socket.socket() as s:
        s.bind(("127.0.0.1", 0))
        return s.getsockname()[1]


def _gather_worker(rank, world, port, out):
    import torch.distributed as dist
    import ddensemble
    import prob1_mms_cases as p1mc
    from ddmesh import shard_members
    os.environ["MASTER_ADDR"], os.environ["MASTER_PORT"] = "127.0.0.1", str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    try:
        n = 11
        a, b = shard_members(n, world, rank)
        local = np.stack([np.arange(a, b, dtype=np.float64), 10.0 * np.arange(a, b)], axis=1)
        full = ddensemble.gather_members(local, n, world, rank, dist)
        trials = [dict(N=4 * (k + 1), dt=1e-3, Tf=2e-3, eta=50.0) for k in range(5)]
        sw = ddensemble.RefinementSweep(p1mc.MMSCasePol, None, trials, world=world, rank=rank)
        owned = np.array([sw.assignment[[k in g["members"] for g in sw.groups].index(True)] == rank
                          for k in range(5)])
        res = dict(overall=np.where(owned, np.arange(5) + 1.0, np.nan),
                   per_var=np.where(owned[:, None], np.arange(25).reshape(5, 5) + 0.5, np.nan), owned=owned)
        merged = sw.merge(res, dist)
        out.put((rank, full, merged["overall"], merged["per_var"]))
    finally:
        dist.destroy_process_group()


def test_member_gather_and_sweep_merge_gloo():
    world, port = 2, _free_port()
    ctx = mp.get_context("spawn")
    out = ctx.Queue()
    procs = [ctx.Process(target=_gather_worker, args=(r, world, port, out)) for r in range(world)]
    for p in procs:
        p.start()
    got = [out.get(timeout=120) for _ in range(world)]
    for p in procs:
        p.join(timeout=60)
        assert p.exitcode == 0
    for rank, full, overall, per_var in got:
        assert full.shape == (11, 2)
        np.testing.assert_array_equal(full[:, 0], np.arange(11.0))
        np.testing.assert_array_equal(full[:, 1], 10.0 * np.arange(11.0))
        np.testing.assert_array_equal(overall, np.arange(5) + 1.0)
        np.testing.assert_array_equal(per_var, np.arange(25).reshape(5, 5) + 0.5)


REF_SRC = "/root/reference/src"


@pytest.mark.skipif(not os.path.isdir(REF_SRC), reason="the reference tree only exists in the build container")
def test_reference_study_modules_load_on_this_core():
    """north_star: the reference's own mms_trial_utils / cvg_studies_base / prob1_mms_cases run unchanged on top of
    this package's prob1base.  Here (no GPU): they import (their annotations are evaluated against our names),
    their pure-host parts agree with ours, and the reference's case classes get a device description from our
    MMSCaseSymbolic."""
    import importlib.util
    import prob1base as p1
    import prob1_mms_cases as ours
    import cvg_studies_base as our_cvg

    def load(name):
        spec = importlib.util.spec_from_file_location("ref_" + name, os.path.join(REF_SRC, name + ".py"))
        mod = importlib.util.module_from_spec(spec)
        spec.loader.exec_module(mod)   # `import prob1base as p1` inside resolves to this package's module
        return mod

    ref_cvg, ref_mtu, ref_cases = load("cvg_studies_base"), load("mms_trial_utils"), load("prob1_mms_cases")
    assert ref_cvg.p1 is p1 and ref_mtu.p1 is p1
    errs = [4e-5, 1.6e-5, 4.3e-6, 1.1e-6, 2.8e-7]
    assert ref_cvg.calculate_observed_rates(errs) == our_cvg.calculate_observed_rates(errs)
    grid = p1.make_uniform_grid(6, 5)
    model = p1.DefaultModel02(p1.default_model_consts)
    import inspect
    names = [n for n, cls in inspect.getmembers(ref_cases, inspect.isclass)
             if issubclass(cls, p1.MMSCaseBase) and cls.__module__ == ref_cases.__name__
             and n not in ("MMSCaseNonFullySmoothPol",)]          # needs a gamma argument; its subclasses are covered
    assert len(names) >= 20
    for name in names:
        theirs, mine = getattr(ref_cases, name)(grid=grid, model=model), getattr(ours, name)(grid=grid, model=model)
        for v in ("cp", "T", "cl", "cd", "cs"):
            for fn in (v, "dt_" + v, "dx_" + v, "dy_" + v, "lap_" + v):
                a, b = getattr(theirs, fn)(0.3, grid.xx, grid.yy), getattr(mine, fn)(0.3, grid.xx, grid.yy)
                np.testing.assert_allclose(a, b, rtol=1e-13, atol=1e-300, err_msg=f"{name}.{fn}")
        assert mine.device_spec() is not None, name              # every library case runs on the device
        if name != "MMSCaseExpSin":   # ExpSin's closed form is keyed on our own class; the others factorise
            assert theirs.device_spec() is not None, name
    trial = ref_mtu.MMSTrial  # constructing it needs a device (state upload); the class itself resolves
    assert callable(trial)


def test_analytic_case_finite_differences_and_errors():
    """MMSCaseFromAnalytic / pack_analytical_txy_with_o2fdm_derivatives (reference src/prob1base.py:895-1155):
    second-order differences against exact derivatives, the precedence and error rules of the interface, and a
    host forcing object built on such a case against the same solution given symbolically."""
    import sympy
    import prob1base as p1
    from test_hostsim import product_model
    from test_program_codegen import MODEL
    t, x, y = p1.t_sym, p1.x_sym, p1.y_sym
    expr = sympy.exp(-t) * sympy.sin(2 * x + y * t) + x * y ** 2
    f = sympy.lambdify([t, x, y], expr, "numpy")
    X, Y = np.meshgrid(np.linspace(0.1, 0.9, 6), np.linspace(0.2, 0.8, 5), indexing="ij")
    exact = lambda d: sympy.lambdify([t, x, y], sympy.diff(expr, *([t] * d[0] + [x] * d[1] + [y] * d[2])), "numpy")(0.3, X, Y)
    for ts in ("center", "forward", "backward"):
        g = p1.pack_analytical_txy_with_o2fdm_derivatives(f, default_eps=1e-4, time_stepping=ts)
        assert np.array_equal(g(0.3, X, Y), f(0.3, X, Y))
        for d in ((1, 0, 0), (2, 0, 0), (0, 1, 0), (0, 2, 0), (0, 0, 1), (0, 0, 2), (0, 1, 1)):
            assert np.max(np.abs(g(0.3, X, Y, d=d) - exact(d))) <= 2e-6, (ts, d)   # O(eps^2) + rounding / eps^2
        assert np.max(np.abs(g(0.3, X, Y, op="Laplacian") - exact((0, 2, 0)) - exact((0, 0, 2)))) <= 2e-6
        assert np.array_equal(g(0.3, X, Y, d=(1, 1, 0)), g(0.3, X, Y, d=(1, 0, 0)))   # a time derivative wins
        assert np.max(np.abs(g(0.3, X, Y, d=(0, 1, 0), small_eps=1e-3) - exact((0, 1, 0)))) <= 1e-5
    g = p1.pack_analytical_txy_with_o2fdm_derivatives(f)
    for bad in (dict(op="grad"), dict(d=(3, 0, 0)), dict(d=(0, 2, 1))):
        with pytest.raises(ValueError):
            g(0.3, X, Y, **bad)
    with pytest.raises(ValueError):
        p1.pack_analytical_txy_with_o2fdm_derivatives(f, time_stepping="sideways")
    # a case from callables next to the same case from expressions
    grid = p1.make_uniform_grid(6, 5)
    model = product_model(MODEL)
    exprs = dict(cp=sympy.exp(-t) * x * (1 - x) * y * (1 - y), T=1 + sympy.sin(x + y + t) / 10, cl=sympy.cos(x * y + t) / 3,
                 cd=sympy.exp(-t - x) / 2, cs=sympy.sin(sympy.pi * x) * sympy.cos(t) / 4)
    sym = p1.MMSCaseSymbolic(grid=grid, model=model, **{k + "_sym_expr": v for k, v in exprs.items()})
    ana = p1.MMSCaseFromAnalytic(model, grid=grid, **{k + "_base": sympy.lambdify([t, x, y], v, "numpy")
                                                      for k, v in exprs.items()})
    assert ana.device_spec() is None and ana.grid is grid and ana.model is model
    for name in ("cp", "dt_cs", "dx_T", "dy_cl", "lap_cd", "lap_T", "dx_cp"):
        a, b = getattr(ana, name)(0.2, grid.xx, grid.yy), getattr(sym, name)(0.2, grid.xx, grid.yy)
        assert np.max(np.abs(a - b)) <= 5e-3 * max(1.0, np.max(np.abs(b))), name   # eps = 1e-6: rounding / eps^2 ~ 1e-3
    fa = p1.ForcingTerms_RegHCsTriple(mms_case=ana, model=model, regularization_factor=50.0)
    fs = p1.ForcingTerms_RegHCsTriple(mms_case=sym, model=model, regularization_factor=50.0)
    for name in ("fcp", "fT", "fcl", "fcd", "fcs"):
        a, b = getattr(fa, name)(0.2, grid.xx, grid.yy), getattr(fs, name)(0.2, grid.xx, grid.yy)
        assert np.max(np.abs(a - b)) <= 1e-6 * max(1.0, np.max(np.abs(b))) + 1e-6, name


def test_analytic_helper_equals_the_reference_bit_for_bit():
    """tests/golden/analytic_fd.npz holds the outputs of the live reference's finite-difference helper
    (oracle/make_golden.py, `run_analytic`) for the callable restated below."""
    import prob1base as p1
    from golden_util import load_fixture
    desc, z = load_fixture("analytic_fd")
    fn = lambda t, x, y: np.exp(-t) * np.sin(2 * x + y * t) + x * y ** 2
    X, Y, t = z["X"], z["Y"], desc["t"]
    checked = 0
    for ts in ("center", "forward", "backward"):
        g = p1.pack_analytical_txy_with_o2fdm_derivatives(fn, time_stepping=ts)
        for key in [k for k in z.files if k.startswith(ts + "_")]:
            sel = key[len(ts) + 1:]
            if sel == "lap":
                got = g(t, X, Y, op="lap")
            elif sel.endswith("_eps1e-4"):
                got = g(t, X, Y, d=(0, 2, 0), small_eps=1e-4)
            else:
                got = g(t, X, Y, d=tuple(int(c) for c in sel[1:]))
            assert np.array_equal(got, z[key]), key
            checked += 1
    assert checked == 33
