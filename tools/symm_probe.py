import os, time, torch, torch.distributed as dist
import torch.distributed._symmetric_memory as symm
rank, world, local = int(os.environ["RANK"]), int(os.environ["WORLD_SIZE"]), int(os.environ["LOCAL_RANK"])
torch.cuda.set_device(local)
dist.init_process_group("nccl", device_id=torch.device("cuda", local))
G, ld = 24, 8193
buf = symm.empty((2, 2, 5, G * ld), dtype=torch.float64, device="cuda")
hdl = symm.rendezvous(buf, dist.group.WORLD)
print(rank, "rendezvous ok", hdl.world_size, hdl.buffer_size, hdl.signal_pad_size, flush=True)
up, dn = rank - 1, rank + 1
mine = torch.full((G * ld,), float(rank + 1), device="cuda", dtype=torch.float64)
halo_up = torch.zeros_like(mine); halo_dn = torch.zeros_like(mine)
def peer(r, par, side, v):
    off = ((par * 2 + side) * 5 + v) * G * ld
    return hdl.get_buffer(r, (G * ld,), torch.float64, off)
views = {}
def exchange(k, v=0):
    par = k % 2
    if up >= 0:
        views.setdefault((up, par, 1, v), peer(up, par, 1, v)).copy_(mine)
    if dn < world:
        views.setdefault((dn, par, 0, v), peer(dn, par, 0, v)).copy_(mine)
    if up >= 0: hdl.put_signal(up, 0)
    if dn < world: hdl.put_signal(dn, 0)
    if up >= 0: hdl.wait_signal(up, 0)
    if dn < world: hdl.wait_signal(dn, 0)
    if up >= 0: halo_up.copy_(buf[par, 0, v])
    if dn < world: halo_dn.copy_(buf[par, 1, v])
for k in range(4):
    mine.fill_(float(100 * k + rank + 1))
    exchange(k)
torch.cuda.synchronize()
exp_up = 300 + rank if up >= 0 else 0.0
exp_dn = 300 + rank + 2 if dn < world else 0.0
print(rank, "halo_up", halo_up[0].item(), halo_up[-1].item(), "expect", exp_up, "halo_dn", halo_dn[0].item(), "expect", exp_dn, flush=True)
dist.barrier(); torch.cuda.synchronize()
t0 = time.perf_counter()
n = 200
for k in range(n):
    exchange(k)
torch.cuda.synchronize()
el = (time.perf_counter() - t0) / n * 1e6
# NCCL reference: batch_isend_irecv of the same size
def nccl_ex():
    ops = []
    if up >= 0:
        ops += [dist.P2POp(dist.isend, mine, up), dist.P2POp(dist.irecv, halo_up, up)]
    if dn < world:
        ops += [dist.P2POp(dist.isend, mine, dn), dist.P2POp(dist.irecv, halo_dn, dn)]
    for r in dist.batch_isend_irecv(ops): r.wait()
for _ in range(5): nccl_ex()
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n): nccl_ex()
torch.cuda.synchronize(); el2 = (time.perf_counter() - t0) / n * 1e6
x = torch.zeros(3, device="cuda", dtype=torch.float64)
for _ in range(5): dist.all_reduce(x, op=dist.ReduceOp.MAX)
torch.cuda.synchronize(); t0 = time.perf_counter()
for _ in range(n): dist.all_reduce(x, op=dist.ReduceOp.MAX)
torch.cuda.synchronize(); el3 = (time.perf_counter() - t0) / n * 1e6
print(rank, f"symm exchange {el:.1f} us, nccl p2p exchange {el2:.1f} us, nccl allreduce(3) {el3:.1f} us", flush=True)
dist.destroy_process_group()
