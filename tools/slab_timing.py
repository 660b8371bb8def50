"""Development probe: where a slab step spends its time at N ranks (sync after every part)."""
import os, sys, time, collections
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200"))
import numpy as np, torch, torch.distributed as dist
import bench, ddcore, ddmesh
from _ddlib import Context
world, rank, local = int(os.environ["WORLD_SIZE"]), int(os.environ["RANK"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
stream = torch.cuda.Stream(device=local)
acc = collections.defaultdict(float)
with torch.cuda.stream(stream):
    ctx = Context(local, stream.cuda_stream)
    mesh, dt, cells = bench.mesh_setup(world, rank, ctx)
    opt = ddcore.pc_options()
    comm = mesh._comm()
    for name in ("exchange", "allreduce"):
        orig = getattr(comm, name)
        def wrap(*a, _o=orig, _n=name, **k):
            torch.cuda.synchronize(); t = time.perf_counter()
            r = _o(*a, **k)
            torch.cuda.synchronize(); acc[_n] += time.perf_counter() - t
            return r
        setattr(comm, name, wrap)
    t = 0.0
    for k in range(5):
        mesh.step_pc(k % 2, (k + 1) % 2, t, dt, opt); t += dt
    acc.clear()
    torch.cuda.synchronize(); dist.barrier(); t0 = time.perf_counter()
    n = 10
    for k in range(5, 5 + n):
        mesh.step_pc(k % 2, (k + 1) % 2, t, dt, opt); t += dt
    torch.cuda.synchronize(); el = time.perf_counter() - t0
    if rank == 0:
        print("ms/step with syncs", el / n * 1e3, {k: v / n * 1e3 for k, v in acc.items()})
dist.destroy_process_group()
