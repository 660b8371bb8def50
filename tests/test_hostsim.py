"""CPU check of the device arithmetic: the node programs of the CUDA kernels, compiled for the
host (tests/hostsim), against the golden fixtures of the reference.

This validates csrc/dd_physics.cuh + dd_nodeprog.cuh (stencils, Jacobians, forcing closed forms,
correctors, quirks) and the package's MMS table builder without a GPU.  The kernels' tiling,
reductions and the C ABI are covered by the `-m gpu` tests.
"""
import numpy as np
import pytest

import hostsim_util as hs
from golden_util import VARS, fixture_names, load_fixture, rel_err

import prob1base as p1
import prob1_mms_cases as p1mc

TOL = 1e-12

CASES = {"pol": p1mc.MMSCasePol, "expsin": p1mc.MMSCaseExpSin,
         "scp_fast1e1": p1mc.MMSCaseSlowlyChangingPeaks_Fast1e1,
         "nfsp_h1h2": p1mc.MMSCaseNonFullySmoothPol_cpcsH1_TclcdH2}


def product_model(md):
    md = dict(md)
    kind = md.pop("kind")
    md.setdefault("R0", p1.R0)
    md.setdefault("Ea", p1.Ea)
    md.setdefault("phi_T", p1.Ea / p1.R0)
    mc = p1.ModelConsts(**md)
    return (p1.DefaultModel02 if kind == 2 else p1.DefaultModel01)(mc)


def make_problem(desc, z, t0, dt):
    model = product_model(desc["model"])
    grid = p1.Grid(z["x"], z["y"])
    spec = None
    if desc["case"] is not None:
        case = CASES[desc["case"]](grid=grid, model=model)
        spec = case.device_spec()
        assert spec is not None
    return hs.Problem(z["x"], z["y"], model, desc["eta"], spec, t0, dt,
                      reaction=desc.get("variant", "regh")), model, grid


@pytest.mark.parametrize("name", fixture_names(kind="steps"))
def test_hostsim_steps_match_reference(name):
    desc, z = load_fixture(name)
    dt, t = desc["dt"], desc["t0"]
    prob, model, grid = make_problem(desc, z, t, dt)
    s = {v: z["init_" + v] for v in VARS}
    if desc["init"] == "exact":
        ex = prob.exact(t)
        for v in VARS:
            assert rel_err(ex[v], z["init_" + v]) <= 1e-13, f"exact {v}"
    F = prob.fields(s, t)
    for v in VARS:
        assert rel_err(F[v], z["F0_" + v]) <= TOL, f"F0_{v}"
    pc = desc["pc"]
    for n in range(desc["nsteps"]):
        prob, _, _ = make_problem(desc, z, t, dt)
        if desc["integrator"] == "pc":
            s, info, iters = prob.pc_step(s, t, dt, num_pc_steps=pc.get("num_pc_steps", 1),
                                          num_newton_steps=pc.get("num_newton_steps", 1),
                                          num_newton_iterations=pc.get("num_newton_iterations", 5),
                                          consec_xs_rtol=pc.get("consec_xs_rtol", 1e-6), sweeps=400)
            assert max(info[3:]) <= 1e-13 * max(1.0, max(np.max(np.abs(s[v])) for v in ("T", "cl", "cd")))
            assert sum(iters) == int(z["cs_newton_calls_per_step"][n])
        else:
            s = prob.feuler(s, t, dt)
        t += dt
        if f"step{n + 1}_cp" in z:
            for v in VARS:
                assert rel_err(s[v], z[f"step{n + 1}_{v}"]) <= TOL, f"step {n + 1} {v}"


def test_separable_detection():
    grid = p1.make_uniform_grid(4, 4)
    model = product_model(dict(K1=1e-3, K2=1e-3, K3=1e-3, K4=1e-3, DT=1e-3, Dl_max=1e-5, phi_l=1e-5, gamma_T=1e-9,
                               Kd=1e-2, Sd=1.0, Dd_max=1e-6, phi_d=1e-5, r_sp=5e-2, T_ref=300.0, kind=2))
    x, y, t = p1.x_sym, p1.y_sym, p1.t_sym
    import sympy
    f = sympy.exp(-3 * t) * (x**2 + y**2) * sympy.sin(x) * sympy.cos(y)
    case = p1.MMSCaseSymbolic(grid=grid, model=model, cp_sym_expr=f, T_sym_expr=2 * f, cl_sym_expr=f, cd_sym_expr=f,
                              cs_sym_expr=x * y / (1 + t))
    spec = case.device_spec()
    assert spec is not None and len(spec.X[0]) == 2
    g = sympy.exp(-t * x) * y  # not separable
    case2 = p1.MMSCaseSymbolic(grid=grid, model=model, cp_sym_expr=g, T_sym_expr=f, cl_sym_expr=f, cd_sym_expr=f,
                               cs_sym_expr=f)
    assert case2.device_spec() is None
    prob = hs.Problem(grid.x, grid.y, model, 50.0, spec, 0.3, 0.1)
    ex = prob.exact(0.3)
    for v in VARS:
        assert rel_err(ex[v], getattr(case, v)(0.3, grid.xx, grid.yy)) <= 1e-14
