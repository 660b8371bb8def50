#!/usr/bin/env python
"""bench.py -- fp64 cell-steps/s of the RegHCsTriple predictor-corrector step on B200.

    python bench.py --gpus N --steps K --warmup W            (N > 1: launched by torchrun, one rank per GPU)
    python bench.py --impl reference --gpus N --steps K --warmup W

A "step" is one PC step (num_pc_steps = num_newton_steps = 1, 5 cs-Newton iterations: the reference
defaults) of the whole workload.  Default workload `mesh`: MMSCasePol on 1024 rows per GPU of the
N = M = 8192 unit-square mesh (h = k = 1/8192, dt = h^1.5 -- BASELINE.json configs[4]; at 8 GPUs it is
exactly that config), slab-decomposed along i with halo exchange over NCCL.  Other workloads:
`sweep` (configs[1], the 19-trajectory Pol refinement sweep) and `ensemble` (configs[2]).

One JSON line on rank 0.  `value` = cell-steps/s with state resident in HBM, timed with CUDA events on
the launching stream (max over ranks); `e2e` = the same step through Batch.upload / step_pc / download
with pinned HOST buffers; `roofline` = dominant kernel vs the measured HBM peak; `cpu_baseline` = the
unmodified reference (baseline/_ref, staged by tools/stage_reference.sh; else the oracle port) on the host cores;
`check` = error norms of the timed state and, on several GPUs, slab-vs-whole-mesh digests.
"""
from __future__ import annotations

import argparse
import ctypes as C
import json
import os
import statistics
import subprocess
import sys
import tempfile
import time

ROOT = os.path.dirname(os.path.abspath(__file__))
PKG = os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

import numpy as np  # noqa: E402

METRIC = "fp64 cell-steps/sec (Newton+tridiag)"
UNIT = "cell-steps/s"
BYTES_PER_CELL_STEP = 280.0  # PC-RegH step, p = q = 1
POL = dict(K1=1e-3, K2=1e-3, K3=1e-3, K4=1e-3, DT=1e-3, Dl_max=8.01e-4, phi_l=1e-5, gamma_T=1e-9, Kd=1e-2, Sd=1.0,
           Dd_max=2.46e-6, phi_d=1e-5, r_sp=5e-2, T_ref=300.0)
ETA = 50.0
# algorithmic bytes per node and launch of each kernel class ("kernels").  The classes are those of the
# library's event profile (dd_profile_read): "k_rbsor_tile<v>" = all solver passes of variable v (kernel
# k_rbsor_reg<...> unless DD_SOLVER=smem), "k_predict" = k_predict_march or k_predict, "k_assemble<v>" likewise.
KERNEL_BYTES = {"k_predict": 88, "k_assemble<T>": 40, "k_assemble<cl>": 80, "k_assemble<cd>": 104,
                "k_rbsor_tile<T>": 32, "k_rbsor_tile<cl>": 56, "k_rbsor_tile<cd>": 56, "k_correct": 80,
                "k_feuler": 80, "k_eval_sources": 40}
# CUDA kernels behind the profile classes on the mesh workload (wide grid, staged sources); the solver's variant is
# asked from the library after the run (dd_solver_kernel_name)
KERNEL_NAMES = {"k_predict": "k_predict_march<true>", "k_assemble<cl>": "k_assemble_cl_march",
                "k_assemble<cd>": "k_assemble_cd_march", "k_correct": "k_correct<ARRAYS, false>",
                "k_eval_sources": "k_eval_sources<SEPARABLE>", "k_feuler": "k_feuler_march",
                "k_cs_decide+k_cs_redo": "k_cs_decide, k_cs_redo<ARRAYS>", "k_time_coefs": "k_time_coefs",
                "k_summarise": "k_summarise"}
PROFILE_TAG = "r02"  # profiles/<tag>_traffic.json, <tag>_fp64.json: the ncu captures the static figures come from


def captured_traffic(kernel):
    """DRAM bytes per launch of `kernel` from the committed ncu --set full capture (profiles/), or None."""
    try:
        with open(os.path.join(ROOT, "profiles", PROFILE_TAG + "_traffic.json")) as f:
            t = json.load(f)
        k = t["kernels"][kernel]
        return float(k["dram_bytes_per_launch"]), t["source"] + "; kernel " + k["ncu_kernel"]
    except Exception:
        return None, None


def captured_fp64():
    """fp64 flops per step (2 DFMA + DADD + DMUL thread instructions, predicated on) of the whole PC step from the
    committed ncu capture, or None."""
    try:
        with open(os.path.join(ROOT, "profiles", PROFILE_TAG + "_fp64.json")) as f:
            t = json.load(f)
        return float(t["flops_per_step"]), float(t["dfma_per_step"]), t["source"]
    except Exception:
        return None, None, None


def step_bytes_of(args):
    return 80.0 if args.integrator == "feuler" else BYTES_PER_CELL_STEP


def measured_peak():
    try:
        with open(os.path.join(ROOT, "MEASURED_PEAKS.json")) as f:
            return float(json.load(f)["hbm_gbs"]), "measured"
    except Exception:
        return 6650.0, "fallback"


def product_model():
    import prob1base as p1
    return p1.DefaultModel02(p1.ModelConsts(R0=p1.R0, Ea=p1.Ea, phi_T=p1.Ea / p1.R0, **POL))


# ----------------------------------------------------------------------------
# clocks
# ----------------------------------------------------------------------------
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region.  A helper process polls NVML every 2 ms
    (the default run's timed region lasts tens of milliseconds -- too short for nvidia-smi's polling, and a
    thread in this process would wait for the interpreter lock); the samples between begin() and stop() count."""
    REASONS = {0x8: "hw_slowdown", 0x40: "hw_thermal_slowdown", 0x20: "sw_thermal_slowdown", 0x4: "sw_power_cap"}
    CODE = (
        "import sys, time\n"
        "import pynvml as nv\n"
        "nv.nvmlInit()\n"
        "h = nv.nvmlDeviceGetHandleByIndex(int(sys.argv[1]))\n"
        "rs = getattr(nv, 'nvmlDeviceGetCurrentClocksEventReasons', None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons\n"
        "mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)\n"
        "while True:\n"
        "    print(time.time(), nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM), mx, int(rs(h)), flush=True)\n"
        "    time.sleep(0.002)\n")

    def __init__(self, device):
        vis = os.environ.get("CUDA_VISIBLE_DEVICES", "")
        ids = [v for v in vis.split(",") if v.strip().isdigit()]
        self.index = int(ids[device]) if device < len(ids) else device
        self.p = self.f = None
        self.t0 = self.t1 = None

    def start(self):
        """spawn the poller and wait until it delivers (call before the warm-up)"""
        try:
            self.f = tempfile.NamedTemporaryFile("w+", suffix=".clk", delete=False)
            self.p = subprocess.Popen([sys.executable, "-c", self.CODE, str(self.index)], stdout=self.f,
                                      stderr=subprocess.DEVNULL)
            for _ in range(400):  # wait for the first sample (interpreter + NVML start-up), at most 20 s
                if os.path.getsize(self.f.name) > 0 or self.p.poll() is not None:
                    break
                time.sleep(0.05)
        except Exception:
            self.p = None

    def begin(self):
        self.t0 = time.time()

    def stop(self):
        self.t1 = time.time()
        if self.p is None:
            return {"sm_mhz": None, "sm_max_mhz": None, "reasons": ["clock poller unavailable"], "samples": 0}
        if self.t0 is None:
            self.t0 = 0.0
        time.sleep(0.01)
        self.p.terminate()
        try:
            self.p.wait(timeout=5)
        except Exception:
            self.p.kill()
        self.f.flush()
        self.f.seek(0)
        sm, mx, reasons, total = [], [], set(), 0
        for line in self.f.read().splitlines():
            c = line.split()
            if len(c) != 4:
                continue
            total += 1
            try:
                ts, clk, cmax, mask = float(c[0]), float(c[1]), float(c[2]), int(c[3])
            except ValueError:
                continue
            mx.append(cmax)
            if ts < self.t0 or ts > self.t1:
                continue
            sm.append(clk)
            for bit, name in self.REASONS.items():
                if mask & bit:
                    reasons.add(name)
        try:
            os.unlink(self.f.name)
        except OSError:
            pass
        return {"sm_mhz": statistics.median(sm) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": sorted(reasons), "samples": len(sm), "samples_total": total,
                "source": "NVML polled every 2 ms by a helper process; samples inside the timed region"}


# ----------------------------------------------------------------------------
# workloads on the device
# ----------------------------------------------------------------------------
MESH_COLS = 8192
MESH_ROWS_PER_GPU = 1024


def mesh_setup(world, rank, ctx):
    """rows i of the N = M = 8192 unit-square mesh: 1024 per GPU (+ halo), dt = h^1.5."""
    import ddmesh
    import prob1_mms_cases as p1mc
    import prob1base as p1
    h = 1.0 / MESH_COLS
    N = MESH_ROWS_PER_GPU * world
    x = np.arange(N + 1) * h
    y = np.linspace(0.0, 1.0, MESH_COLS + 1)
    model = product_model()
    mesh = ddmesh.SlabMesh(x, y, world=world, rank=rank, ctx=ctx, nslots=3)
    mesh.batch.set_model(model, ETA)
    mesh.batch.forcing_spec(mesh_case_spec(float(x[-1])))
    mesh.fill_exact(0, 0.0)
    if os.environ.get("DD_BENCH_NOFORCING"):  # development probe: cost of the fused sources
        mesh.batch.forcing_none()
    return mesh, h ** 1.5, N * MESH_COLS


def ensemble_members(n, seed=20250503):
    """Member parameters of configs[2]: eta ~ logU[10, 1000]; K1..K4, DT, Kd each base * U[0.5, 1.5]."""
    import prob1base as p1
    rng = np.random.default_rng(seed)
    etas = 10.0 ** rng.uniform(1.0, 3.0, n)
    f = rng.uniform(0.5, 1.5, (n, 6))
    models = []
    for k in range(n):
        c = dict(POL)
        for q, name in enumerate(("K1", "K2", "K3", "K4", "DT", "Kd")):
            c[name] = POL[name] * f[k, q]
        models.append(p1.DefaultModel02(p1.ModelConsts(R0=p1.R0, Ea=p1.Ea, phi_T=p1.Ea / p1.R0, **c)))
    return models, etas


def sweep_trials():
    """configs[1]: Pol spatial N = 2..256 at dt = h^1.5, temporal N = 256, eta sweep at N = 32."""
    tr = [dict(N=n, dt=(1.0 / n) ** 1.5, Tf=0.01, eta=ETA) for n in (2, 4, 8, 16, 32, 64, 128, 256)]
    tr += [dict(N=256, dt=d, Tf=0.01, eta=ETA) for d in (1e-2, 5e-3, 2.5e-3, 1.25e-3)]
    tr += [dict(N=32, dt=5e-4, Tf=0.01, eta=float(e)) for e in (10, 50, 100, 200, 300, 500, 1000)]
    return tr


def run_studies(args):
    """`--workload ensemble | sweep`: independent trajectories split over the ranks, no communication while
    stepping, one gather of the error scalars at the end.  Wall-clock around whole trials
    (time loop + per-step error norms on the device), device idle on both sides."""
    import torch
    import ddensemble
    import prob1_mms_cases as p1mc
    import prob1base as p1
    from _ddlib import Context, load_library
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    torch.cuda.set_device(local)
    dist = None
    if world > 1:
        import torch.distributed as dist
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    lib = load_library()

    def barrier():
        torch.cuda.synchronize()
        if world > 1:
            dist.barrier()

    if args.workload == "ensemble":
        n = args.members * world
        models, etas = ensemble_members(n)
        grid = p1.make_uniform_grid(32, 32)
        ens = ddensemble.TrajectoryEnsemble(grid, p1mc.MMSCaseSlowlyChangingPeaks_Fast1e1, models, etas, world=world,
                                            rank=rank, ctx=Context(local), chunk=args.members)
        dt = 5e-4

        def job(nsteps):
            return ens.run_for_errors(nsteps * dt, dt)
        clocks = ClockSampler(local)
        clocks.start()
        job(max(1, args.warmup))
        barrier()
        launches0 = lib.dd_launch_count()
        clocks.begin()
        t0 = time.perf_counter()
        res = job(args.steps)
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        all_err = ens.gather(res["overall"], dist, "cuda")
        cell_steps = n * 32 * 32 * args.steps
        workload = (f"scp_fast1e1_ensemble: {args.members} members/GPU of MMSCaseSlowlyChangingPeaks_Fast1e1, N=M=32, "
                    f"dt=5e-4, per-member eta and K1..K4, DT, Kd (rng 20250503), error norms every step")
        extra = {"members": n, "finite_errors": int(np.isfinite(all_err).sum()), "solver": ens.last_stats[-1]}
    else:
        trials = sweep_trials()
        sw = ddensemble.RefinementSweep(p1mc.MMSCasePol, product_model(), trials, world=world, rank=rank,
                                        device=local)
        clocks = ClockSampler(local)
        clocks.start()
        sw.run_for_errors()
        barrier()
        launches0 = lib.dd_launch_count()
        clocks.begin()
        t0 = time.perf_counter()
        res = sw.run_for_errors()
        torch.cuda.synchronize()
        el = time.perf_counter() - t0
        merged = sw.merge(res, dist, "cuda")
        cell_steps = sum(t["N"] * t["M"] * t["nsteps"] for t in sw.trials)
        workload = ("pol_sweep: MMSCasePol spatial N=2..256 (dt=h^1.5) + temporal N=256 (4 dt) + eta sweep N=32 "
                    "(7 eta), Tf=0.01: 19 trials, batches by (grid, steps) on concurrent streams")
        extra = {"trials": len(trials), "overall_errors": [float(x) for x in merged["overall"]]}
    clk = clocks.stop()
    launches = lib.dd_launch_count() - launches0
    if world > 1:
        t = torch.tensor([el], device="cuda", dtype=torch.float64)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        el = float(t.item())
    if rank != 0:
        return
    peak, peak_kind = measured_peak()
    value = cell_steps / el
    line = {"metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps,
            "warmup": args.warmup, "ms_per_step": el * 1e3 / max(1, args.steps), "higher_is_better": True,
            "scaling": "weak" if args.workload == "ensemble" else "strong", "vs_baseline": None, "dtype": "f64",
            "data": "synthetic", "config": dict({"workload": workload, "timing": "host wall clock around whole "
                                                 "trials, device synchronised on both sides"}, **extra),
            "clocks": clk, "e2e": None, "gpu_launches": int(launches),
            "roofline": {"bound": "hbm", "kernel": "whole step (SMEM/L2-resident grids: HBM is not the binding "
                         "roof for these sizes, SURVEY.md 8d)", "achieved": value / world * BYTES_PER_CELL_STEP / 1e9,
                         "peak": peak, "peak_kind": peak_kind, "unit": "GB/s",
                         "frac": value / world * BYTES_PER_CELL_STEP / 1e9 / peak, "traffic": None}}
    _emit(line)


def field_digest(torch, mesh, slot, r0, r1):
    """Order-independent 64-bit digests (sum of the bit patterns, mod 2^64) of local rows [r0, r1) of the five fields."""
    out = []
    for v in ("cp", "T", "cl", "cd", "cs"):
        t = mesh.field(slot, v)[r0:r1]
        out.append(t.contiguous().view(torch.int64).sum())
    return torch.stack(out)


def mesh_check(args, torch, dist, world, rank, local, mesh, dt, slot, t_now):
    """(a) H- and p-norms of (state - exact) after the timed steps, all-reduced over the ranks;
    (b) world > 1: two PC steps from the exact state at a fixed sweep plan on the slabs and, on rank 0, on the
    undecomposed mesh: the slabs' owned rows must be bit-identical to the corresponding rows of the whole mesh."""
    import ddcore
    import ddmesh
    from _ddlib import Context
    e = mesh.error_norms(slot, t_now)
    out = {"err_H": float(np.sqrt(np.sum(e[:5]))), "err_P": float(np.sqrt(np.sum(e[5:]))),
           "what": "discrete H and p norms of (state - exact MMS state) after the timed steps, all ranks"}
    if world == 1 or args.integrator != "pc":
        return out
    fixed = ddcore.pc_options(fixed_sweeps=8)
    mesh.fill_exact(0, 0.0)
    for k in range(2):
        mesh.step_pc(k % 3, (k + 1) % 3, k * dt, dt, fixed)
    p = mesh.part
    mine = field_digest(torch, mesh, 2, p["own0"], p["own1"])
    gathered = [torch.zeros_like(mine) for _ in range(world)]
    dist.all_gather(gathered, mine)
    if rank == 0:
        whole = ddmesh.SlabMesh(mesh.x, mesh.y, world=1, rank=0, ctx=Context(local), nslots=3)
        whole.torch = torch
        whole.batch.set_model(product_model(), ETA)
        whole.batch.forcing_spec(mesh_case_spec(float(mesh.x[-1])))
        whole.fill_exact(0, 0.0)
        for k in range(2):
            whole.step_pc(k % 3, (k + 1) % 3, k * dt, dt, fixed)
        whole.batch.ctx.synchronize()
        equal = True
        for r in range(world):
            a, b = ddmesh.shard_members(len(mesh.x), world, r)
            ref = field_digest(torch, whole, 2, a, b)
            equal = equal and bool(torch.equal(ref, gathered[r]))
        out["slabs_equal_whole_mesh"] = equal
        out["digest"] = ("2 PC steps from the exact state, 8 sweeps per solve: 64-bit sums of the bit patterns of "
                         "each rank's owned rows (5 fields) == those of the same rows of an undecomposed run on rank 0")
        whole.batch.close()
    return out


def mesh_case_spec(L=1.0):
    """MMSCasePol on [0, L] x [0, 1]: u = (x/L)(1 - x/L) y (1 - y) / (1 + t) for all five variables.  L = 1 (8 GPUs:
    the N = M = 8192 unit square of configs[4]) is MMSCasePol itself; with fewer GPUs the weak-scaling mesh keeps
    h = k = 1/8192 and covers x in [0, L = rows / 8192] only, and the manufactured solution is scaled so that it
    still vanishes on the whole boundary (the scheme imposes T = 0 there) -- same tables, same work per cell."""
    import sympy
    import prob1_mms_cases as p1mc
    import prob1base as p1
    grid = p1.Grid(np.array([0.0, 0.5, 1.0]), np.array([0.0, 0.5, 1.0]))
    if L == 1.0:
        return p1mc.MMSCasePol(grid=grid, model=product_model()).device_spec()
    t, x, y = p1.t_sym, p1.x_sym, p1.y_sym
    Ls = sympy.Float(L)
    case = p1mc._SeparableCase(grid, product_model(), phi_exprs=[1 / (1 + t)] * 5,
                               phi_specs=[p1mc.PhiSpec("inv1pt", (1.0,))] * 5,
                               Xs=[(x / Ls) * (1 - x / Ls)] * 5, Ys=[y * (1 - y)] * 5)
    return case.device_spec()


def run_b200(args):
    if args.workload != "mesh":
        return run_studies(args)
    import torch
    import ddcore
    from _ddlib import Context, load_library, profile_read
    world = int(os.environ.get("WORLD_SIZE", "1"))
    rank = int(os.environ.get("RANK", "0"))
    local = int(os.environ.get("LOCAL_RANK", "0"))
    if world != args.gpus:
        if world == 1 and args.gpus > 1:
            raise SystemExit("bench.py --gpus N > 1 must be launched with torch.distributed.run (one rank per GPU)")
    torch.cuda.set_device(local)
    if world > 1:
        import torch.distributed as dist
        # (NCCL's INFO log goes to this process's C-level stdout, which main() has pointed at stderr: the one JSON
        # line is written to the saved original descriptor)
        dist.init_process_group("nccl", device_id=torch.device("cuda", local))
    stream = torch.cuda.Stream(device=local)
    lib = load_library()
    with torch.cuda.stream(stream):
        ctx = Context(local, stream.cuda_stream)
        if args.workload == "mesh":
            mesh, dt, cells = mesh_setup(world, rank, ctx)
            opt = ddcore.pc_options()
            if os.environ.get("DD_BENCH_NOTRACK"):  # development probe: cost of the cs exit-test statistics
                opt = ddcore.pc_options(consec_xs_rtol=0.0)

            clock = {"t": 0.0}

            def step(k):
                # times accumulate as in the reference's loop (current_t += dt, src/mms_trial_utils.py:128):
                # the t1 sources of a step are then bitwise the t0 sources of the next one
                # verification is deferred by one step (three rotating slots): the host enqueues step k + 1
                # before it reads the convergence summaries of step k; flush() settles the last one
                if args.integrator == "feuler":
                    mesh.step_feuler(k % 3, (k + 1) % 3, clock["t"], dt)  # ForwardEulerIntegrator.step (80 B/cell-step)
                else:
                    mesh.step_pc(k % 3, (k + 1) % 3, clock["t"], dt, opt, defer=not args.sync_steps)
                clock["t"] += dt
            workload = (f"pol_mesh: MMSCasePol, {MESH_ROWS_PER_GPU} rows/GPU x {MESH_COLS} cols of the N=M=8192 "
                        f"unit-square mesh (h=k=1/8192, dt=h^1.5), slab decomposition along i"
                        + ("" if world == 8 else f"; {world} GPU(s) cover x in [0, {world}/8] and the solution is "
                           f"scaled to vanish on that boundary"))
            extra_cfg = {"rows_per_gpu": MESH_ROWS_PER_GPU, "cols": MESH_COLS, "halo_rows": mesh.G,
                         "l2": "state + work arrays (>2 GB/GPU) exceed the 126 MB L2"}
        else:
            raise SystemExit(f"unknown workload {args.workload}")

        def barrier():
            torch.cuda.synchronize()
            if world > 1:
                dist.barrier()
            torch.cuda.synchronize()

        clocks = ClockSampler(local)
        if rank == 0:
            clocks.start()
        for k in range(args.warmup):
            step(k)
        mesh.flush()
        barrier()
        launches0 = lib.dd_launch_count()
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        barrier()
        clocks.begin()
        e0.record(stream)
        for k in range(args.steps):
            step(args.warmup + k)
        mesh.flush()  # the last step is verified inside the timed region
        e1.record(stream)
        barrier()
        ms = e0.elapsed_time(e1)
        clk = clocks.stop() if rank == 0 else None
        launches = lib.dd_launch_count() - launches0
        if world > 1:
            t = torch.tensor([ms], device="cuda", dtype=torch.float64)
            dist.all_reduce(t, op=dist.ReduceOp.MAX)
            ms = float(t.item())
        stats = mesh.last_stats

        # per-kernel device time with CUDA events on the same stream (second timed loop, same steps)
        lib.dd_profile_enable(1)
        barrier()
        for k in range(args.steps):
            step(args.warmup + args.steps + k)
        mesh.flush()
        barrier()
        prof = profile_read(True)
        lib.dd_profile_enable(0)
        solver_kernels = {v: lib.dd_solver_kernel_name(k).decode() for k, v in ((1, "T"), (2, "cl"), (3, "cd"))}

        # correctness of what was just timed: error norms of the state against the exact
        # manufactured solution, and -- on more than one GPU -- a digest of the slabs against an undecomposed run
        check = None
        if args.check:
            check = mesh_check(args, torch, dist if world > 1 else None, world, rank, local, mesh, dt,
                               (args.warmup + 2 * args.steps) % 3, clock["t"])
            barrier()
        fp64_peak = None
        if rank == 0:
            tf = C.c_double(0.0)
            ctx.check(lib.dd_probe_fp64(ctx.handle, 30.0, C.byref(tf)), "probe_fp64")
            fp64_peak = float(tf.value)

        # end to end through the host-buffer API (rank-local slab): H2D 5 fields, step, D2H 5 fields, every step
        e2e = None
        if args.e2e:
            shape = mesh.batch.shape
            got = mesh.batch.download((args.warmup + 2 * args.steps) % 3)
            nb = 5 * shape[0] * shape[1] * 8
            ne = max(2, min(args.steps, 5))

            def pinned_pair():
                pin = {v: torch.empty(shape, dtype=torch.float64).pin_memory().numpy() for v in ddcore.VARS}
                for v in ddcore.VARS:
                    pin[v][...] = got[v]
                return pin, {v: torch.empty(shape, dtype=torch.float64).pin_memory().numpy() for v in ddcore.VARS}

            def pipeline(m, pin, pout, nsteps):
                for k in range(nsteps):
                    m.batch.upload(0, pin)
                    if args.integrator == "feuler":
                        m.step_feuler(0, 1, k * dt, dt)
                    else:
                        m.step_pc(0, 1, k * dt, dt, opt)
                    m.batch.download_into(1, pout)

            if world == 1 and not args.e2e_single:
                # two host-resident trajectories, one host thread + context (stream) each: the D2H of one
                # overlaps the H2D and the kernels of the other (PCIe is full duplex); every call is synchronous,
                # so the host clock around the two threads brackets all copies and kernels
                import threading
                mesh_b, _, _ = mesh_setup(world, rank, Context(local))  # own context = own stream
                pipes = [(mesh,) + pinned_pair(), (mesh_b,) + pinned_pair()]
                for m, pin, pout in pipes:
                    pipeline(m, pin, pout, 1)
                barrier()
                th = [threading.Thread(target=pipeline, args=(m, pin, pout, ne)) for m, pin, pout in pipes]
                t_start = time.perf_counter()
                for t_ in th:
                    t_.start()
                for t_ in th:
                    t_.join()
                torch.cuda.synchronize()
                ems = (time.perf_counter() - t_start) * 1e3
                e2e = {"value": 2 * cells * ne / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nb,
                       "d2h_bytes_per_step": nb, "steps": 2 * ne,
                       "mode": "2 independent host-resident trajectories on 2 host threads / streams (duplex PCIe); "
                               "host clock around synchronous upload -> step -> download calls"}
            else:
                pin, pout = pinned_pair()
                pipeline(mesh, pin, pout, 1)
                barrier()
                e0.record(stream)
                pipeline(mesh, pin, pout, ne)
                e1.record(stream)
                barrier()
                ems = e0.elapsed_time(e1)
                if world > 1:
                    t = torch.tensor([ems], device="cuda", dtype=torch.float64)
                    dist.all_reduce(t, op=dist.ReduceOp.MAX)
                    ems = float(t.item())
                e2e = {"value": cells * ne / (ems * 1e-3), "unit": UNIT, "h2d_bytes_per_step": nb * world,
                       "d2h_bytes_per_step": nb * world, "steps": ne,
                       "mode": "one trajectory per GPU: upload -> step -> download on the launching stream"}
    if rank != 0:
        return
    peak, peak_kind = measured_peak()
    value = cells * args.steps / (ms * 1e-3)
    # dominant kernel
    nodes_rank = (MESH_ROWS_PER_GPU + 1) * (MESH_COLS + 1)
    top = max(((k, v) for k, v in prof.items() if k in KERNEL_BYTES), key=lambda kv: kv[1][0])
    tname, (tms, tcount) = top
    # a solve that is split into several passes spreads its algorithmic bytes over its launches
    lps = max(1.0, tcount / args.steps) if tname.startswith("k_rbsor") else 1.0
    alg_per_launch = KERNEL_BYTES[tname] * nodes_rank / lps
    achieved = alg_per_launch / (tms / tcount * 1e-3) / 1e9
    total_prof = sum(v[0] for v in prof.values())
    traffic, traffic_src = captured_traffic(tname)
    step_bytes = 80.0 if args.integrator == "feuler" else BYTES_PER_CELL_STEP
    knames = dict(KERNEL_NAMES)
    for v in ("T", "cl", "cd"):
        knames["k_rbsor_tile<%s>" % v] = solver_kernels.get(v, "")
    if "k_predict" in prof and "k_assemble<T>" in prof:
        knames["k_predict"] = "k_predict_march<false>"
    flops_step, dfma_step, fp64_src = captured_fp64()
    fp64 = None
    if fp64_peak and flops_step and args.integrator == "pc":
        # the capture counts thread instructions of one rank's step (1025 x 8193 nodes); time = this run's step
        ach = flops_step / (ms / args.steps * 1e-3) / 1e12
        fp64 = {"achieved": ach, "peak": fp64_peak, "unit": "TFLOP/s", "frac": ach / fp64_peak,
                "flops_per_step": flops_step, "dfma_per_step": dfma_step, "source": fp64_src,
                "peak_source": "dd_probe_fp64: 8 independent FMA chains per thread, 2 x 1024 threads per SM, "
                               "CUDA events, best of 3, this run"}
    hbm_frac = value / world * step_bytes_of(args) / 1e9 / peak
    bound = "fp64" if (fp64 and fp64["frac"] > hbm_frac) else "hbm"
    roof = {"bound": bound, "kernel": knames.get(tname, tname), "kernel_class": tname, "achieved": achieved,
            "peak": peak, "peak_kind": peak_kind, "fp64": fp64, "kernel_names": knames,
            "unit": "GB/s", "frac": achieved / peak, "traffic": traffic, "traffic_source": traffic_src,
            "algorithmic_bytes_per_launch": alg_per_launch, "launches_per_step": lps,
            "launch_ms": tms / tcount, "share_of_step": tms / total_prof,
            "step": {"bytes_per_cell_step": step_bytes,
                     "achieved": value / world * step_bytes / 1e9,
                     "frac": value / world * step_bytes / 1e9 / peak},
            "kernels": {k: {"ms_per_step": v[0] / args.steps, "launches_per_step": v[1] / args.steps}
                        for k, v in sorted(prof.items(), key=lambda kv: -kv[1][0])}}
    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": args.steps, "warmup": args.warmup,
        "ms_per_step": ms / args.steps, "higher_is_better": True, "scaling": "weak", "vs_baseline": None,
        "dtype": "f64", "data": "synthetic",
        "config": dict({"workload": workload, "integrator": ("forward Euler (all five fields, all nodes)" if args.integrator == "feuler"
                                       else "PC RegHCsTriple p=q=1, 5 cs-Newton iterations"),
                        "eta": ETA, "forcing": "fused MMS (separable tables)", "solver": stats}, **extra_cfg),
        "clocks": clk, "e2e": e2e, "gpu_launches": int(launches), "roofline": roof, "check": check,
    }
    if world == 1 and args.cpu_baseline:
        line["cpu_baseline"] = cpu_baseline(threads=1, steps=2, warmup=1)
    _emit(line)


# ----------------------------------------------------------------------------
# CPU side.  The UNMODIFIED reference (staged by tools/stage_reference.sh into the git-ignored baseline/_ref/, which
# travels to the GPU box) is timed when it is there: kind "reference".  Otherwise the oracle port (NumPy + SciPy
# SuperLU, test infrastructure, used here only as the timed baseline): kind "port".
# ----------------------------------------------------------------------------
CPU_SAMPLE_N = 256   # oracle port
REF_SAMPLE_N = 128   # live reference: 2 s per step and core at this size (8 s at 256)
REF_DIR = os.path.join(ROOT, "baseline", "_ref", "src")

REF_WORKER = r"""
import sys, time, json
import numpy as np
import prob1base as p1
import prob1_mms_cases as cases
N, steps, warmup, eta = int(sys.argv[1]), int(sys.argv[2]), int(sys.argv[3]), float(sys.argv[4])
consts = json.loads(sys.argv[5])
assert "_ref" in p1.__file__, p1.__file__
mc = p1.ModelConsts(R0=p1.R0, Ea=p1.Ea, phi_T=p1.Ea / p1.R0, **consts)
model = p1.DefaultModel02(mc)
grid = p1.make_uniform_grid(N, N)
case = cases.MMSCasePol(grid=grid, model=model)
forcing = p1.ForcingTerms_RegHCsTriple(mms_case=case, model=model, regularization_factor=eta)
field = p1.SemiDiscreteField_RegHCsTriple(grid=grid, model=model, forcing_terms=forcing, regularization_factor=eta)
integ = p1.P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple(field, regularization_factor=eta)
state = p1.state_from_mms_when(mms_case=case, t=0.0, grid=grid)
dt = (1.0 / N) ** 1.5
t = 0.0
for _ in range(warmup):
    state = integ.step(state, t0=t, dt=dt); t += dt
t0 = time.perf_counter()
for _ in range(steps):
    state = integ.step(state, t0=t, dt=dt); t += dt
print(json.dumps({"seconds": time.perf_counter() - t0, "finite": bool(np.isfinite(state.T).all())}))
"""


def _cpu_worker(job):
    steps, warmup = job
    from oracle import NOTEBOOK_CONSTS, OForcing, OGrid, PCStepper, exact_state, make_case
    N = CPU_SAMPLE_N
    om = NOTEBOOK_CONSTS["pol"]
    g = OGrid(np.linspace(0, 1, N + 1), np.linspace(0, 1, N + 1))
    case = make_case("pol", om)
    forcing = OForcing(case, om, ETA, g)
    s = exact_state(case, 0.0, g)
    stepper = PCStepper(g, om, ETA, forcing, keep_residuals=False)
    dt = (1.0 / N) ** 1.5
    t = 0.0
    for _ in range(warmup):
        s = stepper.step(s, t, dt)
        t += dt
    t0 = time.perf_counter()
    for _ in range(steps):
        s = stepper.step(s, t, dt)
        t += dt
    return time.perf_counter() - t0


def reference_baseline(threads: int, steps: int, warmup: int):
    """`threads` independent trajectories of the live reference, one fresh interpreter per core (the reference is
    single-threaded: NumPy element-wise work + serial SuperLU), module search path = baseline/_ref/src only."""
    env = dict(os.environ, PYTHONPATH=REF_DIR, OMP_NUM_THREADS="1", OPENBLAS_NUM_THREADS="1", MKL_NUM_THREADS="1")
    cmd = [sys.executable, "-c", REF_WORKER, str(REF_SAMPLE_N), str(steps), str(warmup), str(ETA), json.dumps(POL)]
    procs = [subprocess.Popen(cmd, env=env, stdout=subprocess.PIPE, stderr=subprocess.PIPE, text=True, cwd=REF_DIR)
             for _ in range(threads)]
    el = []
    for pr in procs:
        out, err = pr.communicate(timeout=1200)
        if pr.returncode != 0:
            raise RuntimeError("reference worker failed: " + err[-500:])
        el.append(json.loads(out.strip().splitlines()[-1])["seconds"])
    cells = REF_SAMPLE_N * REF_SAMPLE_N * steps
    value = sum(cells / e for e in el)
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "reference",
            "sample": (f"the unmodified reference (baseline/_ref/src, staged from /root/reference by "
                       f"tools/stage_reference.sh): P_ModifiedEuler_C_Trapezoidal_TimeIntegrator_RegHCsTriple.step, "
                       f"MMSCasePol N=M={REF_SAMPLE_N}, dt=h^1.5, {steps} steps per worker after {warmup} warm-up, "
                       f"{threads} independent trajectories (one interpreter per core; the reference is single-threaded)"),
            "host_cpus": os.cpu_count(), "s_per_step": max(el) / steps}


def cpu_baseline(threads: int, steps: int, warmup: int):
    if os.path.isdir(REF_DIR):
        try:
            return reference_baseline(threads, steps, warmup)
        except Exception as e:  # a broken staging must not take the bench line down: say so and time the port
            sys.stderr.write(f"reference baseline failed ({e}); timing the oracle port instead\n")
    os.environ.setdefault("OMP_NUM_THREADS", "1")
    os.environ.setdefault("OPENBLAS_NUM_THREADS", "1")
    if threads <= 1:
        el = [_cpu_worker((steps, warmup))]
    else:
        import multiprocessing as mp
        with mp.get_context("fork").Pool(threads) as pool:
            el = pool.map(_cpu_worker, [(steps, warmup)] * threads)
    cells = CPU_SAMPLE_N * CPU_SAMPLE_N * steps
    value = sum(cells / e for e in el) if threads > 1 else cells / el[0]
    return {"value": value, "unit": UNIT, "cores": threads, "kind": "port",
            "sample": (f"oracle port (NumPy + SciPy SuperLU), MMSCasePol N=M={CPU_SAMPLE_N}, dt=h^1.5, "
                       f"{steps} PC steps per worker after {warmup} warm-up, {threads} independent "
                       f"trajectories (one per core)"),
            "host_cpus": os.cpu_count(), "s_per_step": max(el) / steps}


def run_reference(args):
    if int(os.environ.get("RANK", "0")) != 0:
        return
    threads = os.cpu_count() or 1
    steps = max(1, min(args.steps, 4))
    base = cpu_baseline(threads=threads, steps=steps, warmup=min(args.warmup, 1))
    n = REF_SAMPLE_N if base["kind"] == "reference" else CPU_SAMPLE_N
    line = {"impl": "reference", "metric": METRIC, "value": base["value"], "unit": UNIT, "n_gpus": args.gpus,
            "steps": steps, "warmup": min(args.warmup, 1), "ms_per_step": base["s_per_step"] * 1e3,
            "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
            "config": {"workload": f"pol_mesh (bounded sample: N=M={n} per core, same case, constants, eta, "
                                   "dt rule and integrator settings as the CUDA arm)"},
            "cpu_baseline": base,
            "e2e": {"value": base["value"], "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0}}
    _emit(line)


def _emit(line: dict):
    """The one JSON line goes to the process's original stdout; everything else any library prints (NCCL's
    version banner included) was redirected to stderr in main()."""
    data = (json.dumps(line) + "\n").encode()
    fd = _REAL_STDOUT if _REAL_STDOUT is not None else 1
    while data:
        data = data[os.write(fd, data):]


_REAL_STDOUT = None


def main():
    global _REAL_STDOUT
    sys.stdout.flush()
    _REAL_STDOUT = os.dup(1)
    os.dup2(2, 1)  # C-level stdout of this process -> stderr
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=10)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--workload", default="mesh", choices=["mesh", "ensemble", "sweep"])
    ap.add_argument("--members", type=int, default=12500, help="ensemble workload: members per GPU")
    ap.add_argument("--integrator", default="pc", choices=["pc", "feuler"],
                    help="mesh workload: the predictor-corrector step (headline) or the forward-Euler step")
    ap.add_argument("--sync-steps", action="store_true", help="verify every step before enqueuing the next one")
    ap.add_argument("--no-e2e", dest="e2e", action="store_false")
    ap.add_argument("--e2e-single", action="store_true", help="e2e with one trajectory (no duplex overlap)")
    ap.add_argument("--no-cpu-baseline", dest="cpu_baseline", action="store_false")
    ap.add_argument("--no-check", dest="check", action="store_false",
                    help="skip the correctness check of the timed state (error norms; slab digests on > 1 GPU)")
    args = ap.parse_args()
    if args.impl == "reference":
        run_reference(args)
    else:
        run_b200(args)


if __name__ == "__main__":
    main()
