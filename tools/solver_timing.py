"""development aid: per-phase cycles of the tile solver (library built with -DDD_SOLVER_TIMING)"""
import ctypes as C, os, sys
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path[:0] = [ROOT, os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200")]
os.environ["DD_LIB"] = os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200", "libdd_b200_timing.so")
import torch
import bench
from _ddlib import Context, load_library
lib = load_library()
stream = torch.cuda.Stream()
with torch.cuda.stream(stream):
    ctx = Context(0, stream.cuda_stream)
    mesh, dt, cells = bench.mesh_setup(1, 0, ctx)
    t = 0.0
    out = (C.c_ulonglong * 4)()
    for cfg in sys.argv[1:] or [""]:
        for kv in cfg.split(","):
            if kv:
                k_, v_ = kv.split("=")
                os.environ[k_] = v_
        for k in range(10):
            mesh.step_pc(k % 2, (k + 1) % 2, t, dt); t += dt
        torch.cuda.synchronize()
        lib.dd_solver_timing_read(out, 1)
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(stream)
        for k in range(10, 14):
            mesh.step_pc(k % 2, (k + 1) % 2, t, dt); t += dt
        e1.record(stream)
        torch.cuda.synchronize()
        lib.dd_solver_timing_read(out, 1)
        n = out[3]
        print("%-40s ms/step %.3f CTAs/step %d cycles/CTA: staging %.0f sweeps %.0f epilogue %.0f" % (
            cfg, e0.elapsed_time(e1) / 4, n // 4, out[0] / n, out[1] / n, out[2] / n),
            mesh.last_stats["sweeps"], mesh.last_stats["passes"], flush=True)
