"""Host-side logic of the multi-GPU paths on CPU: member sharding, slab row partition, and the halo
exchange pattern with world_size 2 and 3 over gloo (the same `exchange_halos` the NCCL path uses)."""
import os
import socket

import pytest
import torch
import torch.distributed as dist
import torch.multiprocessing as mp

import ddmesh


def test_shard_members_covers_everything():
    for n in (1, 7, 19, 100000):
        for w in (1, 2, 3, 8):
            spans = [ddmesh.shard_members(n, w, r) for r in range(w)]
            assert spans[0][0] == 0 and spans[-1][1] == n
            assert all(a[1] == b[0] for a, b in zip(spans, spans[1:]))
            sizes = [b - a for a, b in spans]
            assert max(sizes) - min(sizes) <= 1


def test_slab_rows():
    for nrows, w, G in ((8193, 8, 24), (2049, 2, 16), (33, 3, 4)):
        parts = [ddmesh.slab_rows(nrows, w, r, G) for r in range(w)]
        assert parts[0]["lo"] == 0 and parts[-1]["hi"] == 0
        assert sum(p["b"] - p["a"] for p in parts) == nrows
        for p in parts:
            assert p["row0"] >= 0 and p["row0"] + p["nrows"] <= nrows
            assert p["own1"] - p["own0"] == p["b"] - p["a"]


def _free_port():
    s = socket.socket()
    s.bind(("127.0.0.1", 0))
    port = s.getsockname()[1]
    s.close()
    return port


def _worker(rank, world, port, nrows, ncols, G, out):
    os.environ["MASTER_ADDR"] = "127.0.0.1"
    os.environ["MASTER_PORT"] = str(port)
    dist.init_process_group("gloo", rank=rank, world_size=world)
    part = ddmesh.slab_rows(nrows, world, rank, G)
    # global field value = 1000 * row + col; local storage holds owned rows, halos start as -1
    fields = []
    for f in range(2):
        t = torch.full((part["nrows"], ncols), -1.0, dtype=torch.float64)
        rows = torch.arange(part["a"], part["b"], dtype=torch.float64)[:, None]
        t[part["own0"]:part["own1"]] = (f + 1) * (1000.0 * rows + torch.arange(ncols, dtype=torch.float64)[None, :])
        fields.append(t)
    ddmesh.exchange_halos(fields, part, rank, world, G, dist)
    ok = True
    for f, t in enumerate(fields):
        rows = torch.arange(part["row0"], part["row0"] + part["nrows"], dtype=torch.float64)[:, None]
        want = (f + 1) * (1000.0 * rows + torch.arange(ncols, dtype=torch.float64)[None, :])
        ok = ok and bool(torch.equal(t, want))
    # all-reduce of per-rank partial error norms (sum)
    e = torch.tensor([float(rank + 1)] * 8, dtype=torch.float64)
    dist.all_reduce(e, op=dist.ReduceOp.SUM)
    ok = ok and float(e[0]) == world * (world + 1) / 2
    out[rank] = ok
    dist.destroy_process_group()


@pytest.mark.parametrize("world", [2, 3])
def test_halo_exchange_gloo(world):
    port = _free_port()
    mgr = mp.Manager()
    out = mgr.dict()
    mp.spawn(_worker, args=(world, port, 41, 7, 4, out), nprocs=world, join=True)
    assert all(out[r] for r in range(world))


def test_slab_thinner_than_halo_is_refused():
    import numpy as np
    import pytest
    """65 rows over 8 ranks with a halo of 24: rank 1 would expect 24 rows from rank 2, which owns 8.  The mesh
    refuses such a partition up front (before any device work: the check precedes the batch creation)."""
    import ddmesh
    x = np.linspace(0.0, 1.0, 65)
    with pytest.raises(ValueError, match="fewer than the halo"):
        ddmesh.SlabMesh(x, x, world=8, rank=1, halo=24)


def test_gauss_seidel_fallback_sweep_count():
    """Sweeps the slab driver plans for a variable that has fallen back to Gauss-Seidel (ddmesh._on_reject): error
    factor <= rho per sweep down to 1e-17, monotone in rho, capped by max_sweeps, defined at the ends of the range."""
    import math
    gs = ddmesh.SlabMesh._gs_sweeps
    prev = 0
    for rho in (0.0, 1e-6, 0.1, 0.5, 0.9, 0.97):
        n = gs(rho, 20000)
        assert n >= max(2, prev), (rho, n)
        r = rho * 1.02 + 1e-12
        assert r ** n <= 1e-17 * 1.0000001 or n == 2, (rho, n)
        if n > 2:
            assert r ** (n - 1) > 1e-17 * 0.999, (rho, n)
        prev = n
    assert gs(0.9126, 20000) == math.ceil(math.log(1e-17) / math.log(0.9126 * 1.02 + 1e-12))
    assert gs(0.999, 100) == 100          # capped
    assert gs(1.0, 500) == 500 and gs(float("nan"), 500) == 500 and gs(-0.1, 500) == 500  # not dominant: the cap
