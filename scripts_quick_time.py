"""Quick device timing of the PC / FE step at several grid sizes (development aid)."""
import sys, os, time, json
ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path[:0] = [ROOT, os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200"), os.path.join(ROOT, "tests")]
import numpy as np
import ddcore, prob1base as p1, prob1_mms_cases as p1mc
from test_hostsim import product_model

model = product_model(dict(K1=1e-3, K2=1e-3, K3=1e-3, K4=1e-3, DT=1e-3, Dl_max=8.01e-4, phi_l=1e-5, gamma_T=1e-9,
                           Kd=1e-2, Sd=1.0, Dd_max=2.46e-6, phi_d=1e-5, r_sp=5e-2, T_ref=300.0, kind=2))
for (N, M, B) in [(32, 32, 1), (256, 256, 1), (1024, 1024, 1), (4096, 2048, 1), (8192, 1024, 1), (32, 32, 4096)]:
    grid = p1.make_uniform_grid(N, M)
    b = ddcore.Batch(grid.x, grid.y, B)
    b.set_model(model, 50.0)
    b.forcing_spec(p1mc.MMSCasePol(grid=grid, model=model).device_spec())
    b.fill_exact(0, 0.0)
    dt = (1.0 / max(N, M)) ** 1.5
    st = b.step_pc(0, 1, 0.0, dt)
    nsteps = 10
    b.ctx.synchronize()
    t = time.perf_counter()
    final, _, st = b.run_pc(0, 1, 0.0, dt, nsteps)
    b.ctx.synchronize()
    el = time.perf_counter() - t
    t = time.perf_counter()
    b.run_feuler(0, 1, 0.0, dt, nsteps)
    b.ctx.synchronize()
    elf = time.perf_counter() - t
    cs = N * M * B * nsteps
    print(json.dumps(dict(N=N, M=M, B=B, pc_ms_per_step=el / nsteps * 1e3, pc_cellsteps_per_s=cs / el,
                          pc_roofline_frac=cs / el * 280 / 6.4515e12, fe_ms_per_step=elf / nsteps * 1e3,
                          fe_cellsteps_per_s=cs / elf, stats=st)), flush=True)
    b.close()
