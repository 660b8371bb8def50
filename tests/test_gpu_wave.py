"""GPU tests of the wavefront solver and of the marching kernels at the sizes where they are used
(M + 1 >= 124): bitwise against the tile kernels, and against the oracle for the cases the reference's
studies use plus a non-uniform grid, DefaultModel01, repeated Newton / PC steps and a time step large enough
to need several solver passes."""
import os

import numpy as np
import pytest

from golden_util import VARS, rel_err

pytestmark = pytest.mark.gpu
TOL = 1e-12


@pytest.fixture(scope="module")
def dd():
    import ddcore
    import prob1base as p1
    from test_hostsim import CASES, product_model
    return dict(ddcore=ddcore, p1=p1, CASES=CASES, product_model=product_model)


def _model_dict(om):
    return dict(K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max, phi_l=om.phi_l,
                gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max, phi_d=om.phi_d, r_sp=om.r_sp,
                T_ref=om.T_ref, kind=om.kind)


def _batch(dd, case, om, x, y, eta):
    p1, ddcore = dd["p1"], dd["ddcore"]
    model = dd["product_model"](_model_dict(om))
    grid = p1.Grid(x, y)
    b = ddcore.Batch(grid.x, grid.y, 1)
    b.set_model(model, eta)
    b.forcing_spec(dd["CASES"][case](grid=grid, model=model).device_spec())
    return b


@pytest.mark.parametrize("N,M,sweeps", [(150, 260, 7), (97, 131, 3), (64, 520, 11)])
def test_wave_solver_equals_tile_solver_bitwise(dd, N, M, sweeps):
    """Same global red-black iteration, same arithmetic (dd_sor.cuh): the wavefront kernel and the register-tile
    kernel give identical fields for equal sweep counts."""
    from oracle import NOTEBOOK_CONSTS
    om = NOTEBOOK_CONSTS["pol"]
    x, y = np.linspace(0, 1, N + 1) ** 1.05, np.linspace(0, 1, M + 1)
    opts = dd["ddcore"].pc_options(fixed_sweeps=sweeps)
    out = {}
    for mode in ("wave", "lane", "tile"):
        os.environ["DD_WAVE"] = "1" if mode == "wave" else "0"
        os.environ["DD_LANE"] = "1" if mode == "lane" else "0"
        try:
            b = _batch(dd, "pol", om, x, y, 50.0)
            b.fill_exact(0, 0.1)
            st = b.step_pc(0, 1, 0.1, 3e-4, opts)
            out[mode] = (b.download(1), st, [b.lib.dd_solver_kernel_name(k).decode() for k in (1, 2, 3)])
            b.close()
        finally:
            os.environ.pop("DD_WAVE", None)
            os.environ.pop("DD_LANE", None)
    assert all(n.startswith("k_sor_wave") for n in out["wave"][2]), out["wave"][2]
    assert all(n.startswith("k_sor_lane") for n in out["lane"][2]), out["lane"][2]
    for mode in ("wave", "lane"):
        for v in VARS:
            assert np.array_equal(out[mode][0][v], out["tile"][0][v]), (mode, v)
        # the marching kernels keep their statistics as high words (residual rounded up by at most 2^-20 relative)
        for a, b in zip(out[mode][1]["resid"], out["tile"][1]["resid"]):
            assert b <= a <= b * (1 + 2e-6) + 1e-300, (mode, a, b)


MARCH_CASES = [
    # id, case, constants, N, M, grid power, model kind, pc steps, newton steps, dt (None: h^1.5)
    ("expsin", "expsin", "expsin", 130, 140, 1.0, 2, 1, 1, None),
    ("nfsp_h1h2", "nfsp_h1h2", "pol", 128, 130, 1.0, 2, 1, 1, None),
    ("pol_nonuniform", "pol", "pol", 140, 150, 1.1, 2, 1, 1, None),
    ("scp_nonuniform_pc22", "scp_fast1e1", "pol", 126, 135, 1.15, 2, 2, 2, None),
    ("pol_model01", "pol", "pol", 125, 125, 1.0, 1, 1, 1, None),
    ("pol_pc22", "pol", "pol", 128, 128, 1.0, 2, 2, 2, None),
    ("scp_bigdt", "scp_fast1e1", "pol", 128, 140, 1.0, 2, 1, 1, 2e-2),
]


@pytest.fixture(params=["tile", "wave", "lane"])
def solver(request):
    """The solvers of the wide-grid regime: register-tile kernels, wavefront kernel, lane-private marching kernel."""
    os.environ["DD_WAVE"] = "1" if request.param == "wave" else "0"
    os.environ["DD_LANE"] = "1" if request.param == "lane" else "0"
    yield request.param
    os.environ.pop("DD_WAVE", None)
    os.environ.pop("DD_LANE", None)


@pytest.mark.parametrize("cid,case,consts,N,M,power,kind,P,Q,dt", MARCH_CASES, ids=[c[0] for c in MARCH_CASES])
def test_marching_sizes_match_oracle(dd, solver, cid, case, consts, N, M, power, kind, P, Q, dt):
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, PCStepper, exact_state, feuler_step, make_case
    ddcore = dd["ddcore"]
    om = NOTEBOOK_CONSTS[consts].with_changes(kind=kind)
    eta, t0 = 50.0, 0.1
    dt = (1.0 / max(N, M)) ** 1.5 if dt is None else dt
    x, y = np.linspace(0, 1, N + 1) ** power, np.linspace(0, 1, M + 1) ** (2.0 - power if power != 1.0 else 1.0)
    og = OGrid(x, y)
    oc = make_case(case, om)
    of = OForcing(oc, om, eta, og)
    s0 = exact_state(oc, t0, og)
    stepper = PCStepper(og, om, eta, of, num_pc_steps=P, num_newton_steps=Q, keep_residuals=False)
    ref1 = stepper.step(s0, t0, dt)
    ref2 = stepper.step(ref1, t0 + dt, dt)
    ref_fe = feuler_step(s0, t0, dt, og, om, eta, of)
    b = _batch(dd, case, om, x, y, eta)
    b.upload(0, s0.fields())
    opts = ddcore.pc_options(num_pc_steps=P, num_newton_steps=Q)
    st1 = b.step_pc(0, 1, t0, dt, opts)
    got1 = b.download(1)
    st2 = b.step_pc(1, 2, t0 + dt, dt, opts)
    got2 = b.download(2)
    b.step_feuler(0, 2, t0, dt)
    got_fe = b.download(2)
    for v in VARS:
        assert rel_err(got1[v], getattr(ref1, v)) <= TOL, (v, st1)
        assert rel_err(got2[v], getattr(ref2, v)) <= TOL, (v, st2)
        assert rel_err(got_fe[v], getattr(ref_fe, v)) <= TOL, v
    assert st1["cs_newton_iters"] == stepper.cs_newton_iters[0]
    assert max(st2["bound"]) <= 5e-13
    if cid == "scp_bigdt":
        assert max(st2["passes"]) >= 1 and max(st2["sweeps"]) >= 8, st2
    b.close()


@pytest.fixture(scope="module")
def oracle512():
    """Two PC steps of MMSCasePol at N = M = 512 by the oracle (SuperLU on 261 121 unknowns: ~20 s per step)."""
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, PCStepper, exact_state, make_case
    N = M = 512
    om = NOTEBOOK_CONSTS["pol"]
    eta, t0, dt = 50.0, 0.0, (1.0 / N) ** 1.5
    x, y = np.linspace(0, 1, N + 1), np.linspace(0, 1, M + 1)
    og = OGrid(x, y)
    oc = make_case("pol", om)
    stepper = PCStepper(og, om, eta, OForcing(oc, om, eta, og), keep_residuals=False)
    ref = exact_state(oc, t0, og)
    for k in range(2):
        ref = stepper.step(ref, t0 + k * dt, dt)
    return dict(N=N, M=M, om=om, eta=eta, t0=t0, dt=dt, x=x, y=y, ref=ref)


def test_slab_mesh_512_world8_matches_oracle(dd, solver, oracle512):
    """SURVEY 8d config 5 at its parity size: N = M = 512, row slabs over 8 ranks (all on one GPU, halo exchange by
    device copies), marching kernels and a multi-tile / multi-strip solve on every slab.  Two PC steps against the
    oracle (1e-12) and, at a fixed sweep plan, bit for bit against the undecomposed mesh."""
    import ddmesh
    ddcore, p1 = dd["ddcore"], dd["p1"]
    o = oracle512
    N, M, om, eta, t0, dt, x, y, ref = (o[k] for k in ("N", "M", "om", "eta", "t0", "dt", "x", "y", "ref"))
    world = 8
    model = dd["product_model"](_model_dict(om))
    spec = dd["CASES"]["pol"](grid=p1.Grid(x, y), model=model).device_spec()

    def run(world, opts):
        meshes = ddmesh.SlabMesh.local_group(x, y, world) if world > 1 else [ddmesh.SlabMesh(x, y)]
        for m in meshes:
            m.batch.set_model(model, eta)
            m.batch.forcing_spec(spec)
            m.fill_exact(0, t0)
        for k in range(2):
            st = meshes[0].step_pc(k % 2, (k + 1) % 2, t0 + k * dt, dt, opts)
        out = {v: np.concatenate([m.owned(0)[v] for m in meshes]) for v in VARS}
        for m in meshes:
            m.batch.close()
        return out, st

    got, st = run(world, ddcore.pc_options())
    for v in VARS:
        assert got[v].shape == (N + 1, M + 1)
        assert rel_err(got[v], getattr(ref, v)) <= TOL, (v, st)
    assert max(st["bound"]) <= 1e-13, st
    fixed = ddcore.pc_options(fixed_sweeps=6)
    many, _ = run(world, fixed)
    one, _ = run(1, fixed)
    for v in VARS:
        assert np.array_equal(many[v], one[v]), v
