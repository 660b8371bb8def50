import os, sys
import numpy as np
sys.path.insert(0, os.path.join(os.getcwd(), "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200"))
sys.path.insert(0, os.path.join(os.getcwd(), "tests"))
sys.path.insert(0, os.getcwd())
import ddcore, ddmesh
import prob1base as p1
from test_hostsim import CASES, product_model
from oracle import NOTEBOOK_CONSTS
N, M = 60, 150
om = NOTEBOOK_CONSTS["pol"].with_changes(DT=0.5, Dl_max=0.3, Dd_max=0.2)
eta, t0, dt, world = 50.0, 0.0, 2e-3, 3
model = product_model(dict(K1=om.K1, K2=om.K2, K3=om.K3, K4=om.K4, DT=om.DT, Dl_max=om.Dl_max, phi_l=om.phi_l, gamma_T=om.gamma_T, Kd=om.Kd, Sd=om.Sd, Dd_max=om.Dd_max, phi_d=om.phi_d, r_sp=om.r_sp, T_ref=om.T_ref, kind=2))
grid = p1.Grid(np.linspace(0, 1, N + 1), np.linspace(0, 1, M + 1))
spec = CASES["scp_fast1e1"](grid=grid, model=model).device_spec()
for lane in ("1", "0"):
    os.environ["DD_LANE"] = lane
    for fs in (40, 120):
        fixed = ddcore.pc_options(fixed_sweeps=fs)
        b = ddcore.Batch(grid.x, grid.y, 1)
        b.set_model(model, eta); b.forcing_spec(spec); b.fill_exact(0, t0)
        try:
            st = b.step_pc(0, 1, t0, dt, fixed)
            print("whole", lane, fs, st["resid"], st["ratio"] if "ratio" in st else None)
        except Exception as e:
            print("whole", lane, fs, "EXC", str(e)[:300])
        b.close()
        meshes = ddmesh.SlabMesh.local_group(grid.x, grid.y, world, halo=7)
        for m in meshes:
            m.batch.set_model(model, eta); m.batch.forcing_spec(spec); m.fill_exact(0, t0)
        try:
            st = meshes[0].step_pc(0, 1, t0, dt, fixed)
            print("slabs", lane, fs, st["resid"], st["ratio"])
        except Exception as e:
            print("slabs", lane, fs, "EXC", str(e)[-420:])
        for m in meshes: m.batch.close()
