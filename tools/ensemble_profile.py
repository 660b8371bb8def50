"""development aid: per-kernel-class device time of the ensemble workload (configs[2]), plus host wall clock"""
import os, sys, time
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200")]
import numpy as np
import torch
import bench
import ddensemble
import prob1_mms_cases as p1mc
import prob1base as p1
from _ddlib import Context, load_library, profile_read

members = int(sys.argv[1]) if len(sys.argv) > 1 else 12500
nsteps = int(sys.argv[2]) if len(sys.argv) > 2 else 20
lib = load_library()
models, etas = bench.ensemble_members(members)
grid = p1.make_uniform_grid(32, 32)
ens = ddensemble.TrajectoryEnsemble(grid, p1mc.MMSCaseSlowlyChangingPeaks_Fast1e1, models, etas, ctx=Context(0),
                                    chunk=members)
dt = 5e-4
ens.run_for_errors(3 * dt, dt)
torch.cuda.synchronize()
t0 = time.perf_counter()
ens.run_for_errors(nsteps * dt, dt)
torch.cuda.synchronize()
wall = time.perf_counter() - t0
print("wall ms/step %.3f  cell-steps/s %.3e" % (wall * 1e3 / nsteps, members * 1024 * nsteps / wall))
lib.dd_profile_enable(1)
ens.run_for_errors(nsteps * dt, dt)
torch.cuda.synchronize()
prof = profile_read()
lib.dd_profile_enable(0)
tot = sum(v[0] for v in prof.values())
for k, (ms, n) in sorted(prof.items(), key=lambda kv: -kv[1][0]):
    print("%-28s %8.3f ms/step  %5.1f %%  launches/step %.1f" % (k, ms / nsteps, 100 * ms / tot, n / nsteps))
print("device total ms/step %.3f" % (tot / nsteps), ens.last_stats[-1])

# where the wall clock goes: pieces of run_for_errors timed separately on the cached batch
import ctypes as C
from _ddlib import dptr
import ddcore
batch = ens._batch(0, members)
def timed(label, fn, reps=2):
    best = 1e9
    for _ in range(reps):
        torch.cuda.synchronize(); t0 = time.perf_counter(); r = fn(); torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    print("%-34s %8.3f ms  (%.3f ms/step)" % (label, best * 1e3, best * 1e3 / nsteps))
    return r
dts = np.full(1, dt)
timed("fill_exact", lambda: batch.fill_exact(0, 0.0))
timed("run_pc norms=False", lambda: batch.run_pc(0, 1, 0.0, dts, nsteps, ens.opt, norms=False))
_, norms, _ = timed("run_pc norms=True (pageable)", lambda: batch.run_pc(0, 1, 0.0, dts, nsteps, ens.opt, norms=True))
pinned = torch.empty((nsteps + 1, members, 8), dtype=torch.float64, pin_memory=True).numpy()
t0a, dta, n = batch._times(0.0, dts)
st = ddcore.dd_step_stats()
timed("run_pc norms=True (pinned)", lambda: batch.ctx.check(lib.dd_run_pc(batch.handle, 0, 1, dptr(t0a), dptr(dta), n, nsteps,
      C.byref(ens.opt), dptr(pinned), C.byref(st)), "run_pc"))
timed("combined_error_norms (host)", lambda: ddensemble.combined_error_norms(norms, np.full(members, dt)))
