"""Multi-GPU partitioning of the stepping path (one process per GPU, torch.distributed for plumbing).

Two ways the path shards:

* `shard_members`  -- ensembles of independent trajectories: contiguous blocks of members per rank,
  no communication while stepping; error norms are gathered / all-reduced at the end.
* `SlabMesh`       -- one fine mesh split into row slabs along i.  Every rank keeps G halo rows on
  each interior side, recomputes the cheap stencil work on them and exchanges G rows of a field
  after each kernel that produces it (state at step start; T1, cl1, cd1 after their solves).
  With a common sweep plan the tiled red-black SOR gives the same iterate as a single-GPU run.
"""

from __future__ import annotations

import ctypes as C
import math
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

import ddcore
from _ddlib import VARS, as_f64, dd_pc_options, dptr


def shard_members(nmembers: int, world: int, rank: int) -> Tuple[int, int]:
    """[first, last) of the members owned by `rank` (contiguous blocks, sizes differ by at most one)."""
    base, rem = divmod(nmembers, world)
    first = rank * base + min(rank, rem)
    return first, first + base + (1 if rank < rem else 0)


def slab_rows(nrows_global: int, world: int, rank: int, halo: int) -> Dict[str, int]:
    """Row partition of a grid with `nrows_global` node rows: owned global rows [a, b), local storage
    [a - lo, b + hi) with lo/hi = halo (0 at the physical boundary)."""
    a, b = shard_members(nrows_global, world, rank)
    lo = 0 if rank == 0 else min(halo, a)
    hi = 0 if rank == world - 1 else min(halo, nrows_global - b)
    return dict(a=a, b=b, lo=lo, hi=hi, row0=a - lo, nrows=(b - a) + lo + hi, own0=lo, own1=lo + (b - a))


class _DevArray:
    """__cuda_array_interface__ view of library-owned device memory (for torch.as_tensor)."""

    def __init__(self, ptr: int, shape, typestr: str = "<f8"):
        self.__cuda_array_interface__ = {"shape": tuple(shape), "typestr": typestr, "data": (int(ptr), False),
                                         "version": 2}


def halo_ops(t, part: Dict[str, int], rank: int, world: int, G: int, dist, tag_base: int = 0):
    """isend/irecv descriptors exchanging G rows of the 2-D tensor `t` (local rows x cols) with both
    neighbours.  Works for CPU (gloo) and CUDA (nccl) tensors alike."""
    ops = []
    o0, o1 = part["own0"], part["own1"]
    if rank > 0:
        g = min(G, part["lo"])
        ops.append(dist.P2POp(dist.isend, t[o0:o0 + g], rank - 1))
        ops.append(dist.P2POp(dist.irecv, t[o0 - g:o0], rank - 1))
    if rank < world - 1:
        g = min(G, part["hi"])
        ops.append(dist.P2POp(dist.isend, t[o1 - g:o1], rank + 1))
        ops.append(dist.P2POp(dist.irecv, t[o1:o1 + g], rank + 1))
    return ops


def exchange_halos(tensors: Sequence, part: Dict[str, int], rank: int, world: int, G: int, dist):
    if world == 1:
        return
    ops = []
    for t in tensors:
        ops += halo_ops(t, part, rank, world, G, dist)
    for r in dist.batch_isend_irecv(ops):
        r.wait()


def _shares_torch_stream(ctx, torch) -> bool:
    return ctx.stream is not None and ctx.stream == torch.cuda.current_stream().cuda_stream


class _Ordered:
    """Orders torch-side communication against the library's stream: when the context runs on torch's
    current stream nothing is needed, otherwise the two streams are joined by synchronising around the
    communication call."""

    def __init__(self, meshes, torch):
        self.meshes, self.torch = meshes, torch
        self.same = all(_shares_torch_stream(m.batch.ctx, torch) for m in meshes)

    def __enter__(self):
        if not self.same:
            for m in self.meshes:
                m.batch.ctx.synchronize()

    def __exit__(self, *a):
        if not self.same:
            self.torch.cuda.current_stream().synchronize()


class _PeerHalo:
    """Halo exchange of the state fields by direct NVLink stores (csrc/dd_halo.cu, include/dd_b200.h): every rank
    exports its field allocations and a flag block as CUDA IPC handles once, imports its neighbours', and an
    exchange is then ONE kernel on the library's stream (dd_halo_push) instead of NCCL send / recv pairs."""

    def __init__(self, mesh, dist):
        self.m, self.lib, self.ctx = mesh, mesh.batch.lib, mesh.batch.ctx
        lib, ctx, b = self.lib, self.ctx, mesh.batch
        self.flags = C.c_void_p()
        ctx.check(lib.dd_halo_flags_create(ctx.handle, C.byref(self.flags)), "halo_flags_create")
        self.seq = 0

        def export(ptr):
            h = C.create_string_buffer(64)
            ctx.check(lib.dd_ipc_export(ctx.handle, C.c_void_p(int(ptr)), h), "ipc_export")
            return h.raw

        self.nslots = b.nslots
        self.ptr = {(s, v): int(b.dev_ptr(s, v)[0]) for s in range(self.nslots) for v in VARS}
        mine = dict(rank=mesh.rank, part=dict(mesh.part), flags=export(self.flags.value),
                    fields={k: export(p) for k, p in self.ptr.items()})
        everyone = [None] * mesh.world
        dist.all_gather_object(everyone, mine)
        self.nb = {}
        self.opened = []
        for side, r in (("up", mesh.rank - 1), ("down", mesh.rank + 1)):
            if r < 0 or r >= mesh.world:
                continue
            info = everyone[r]

            def imp(handle):
                out = C.c_void_p()
                ctx.check(lib.dd_ipc_import(ctx.handle, handle, C.byref(out)), "ipc_import")
                self.opened.append(out.value)
                return out.value

            self.nb[side] = dict(part=info["part"], flags=imp(info["flags"]),
                                 fields={k: imp(h) for k, h in info["fields"].items()})
        dist.barrier()  # nobody pushes before every rank has zeroed its flags and imported the handles

    def usable(self) -> bool:
        p, G = self.m.part, self.m.G
        return all(min(G, p[k]) == G for k, side in (("lo", "up"), ("hi", "down")) if side in self.nb)

    def push(self, slot, var):
        m, p, G = self.m, self.m.part, self.m.G
        ld = m.batch.shape[1]
        base = self.ptr[(slot, var)]
        count = G * ld
        self.seq += 1
        src_top = dst_up = src_bot = dst_down = up_flags = down_flags = None
        if "up" in self.nb:
            nb = self.nb["up"]
            src_top = base + p["own0"] * ld * 8
            dst_up = nb["fields"][(slot, var)] + nb["part"]["own1"] * ld * 8
            up_flags = nb["flags"]
        if "down" in self.nb:
            nb = self.nb["down"]
            src_bot = base + (p["own1"] - G) * ld * 8
            dst_down = nb["fields"][(slot, var)] + (nb["part"]["own0"] - G) * ld * 8
            down_flags = nb["flags"]
        self.ctx.check(self.lib.dd_halo_push(self.ctx.handle, src_top, dst_up, src_bot, dst_down, count, self.flags,
                                             up_flags, down_flags, self.seq & 0xffffffff), "halo_push")

    def check(self):
        st = C.c_int(0)
        self.ctx.check(self.lib.dd_halo_status(self.ctx.handle, self.flags, C.byref(st)), "halo_status")
        if st.value:
            raise RuntimeError("halo exchange: a neighbour never arrived (handshake timed out)")

    def close(self):
        for ptr in self.opened:
            self.lib.dd_ipc_close(self.ctx.handle, C.c_void_p(ptr))
        self.opened = []
        if self.flags:
            self.lib.dd_halo_flags_destroy(self.ctx.handle, self.flags)
            self.flags = C.c_void_p()


class _DistComm:
    """Collectives of one rank per process (torch.distributed, NCCL on GPUs).  Halo rows of the state fields go by
    direct peer stores (_PeerHalo) when the ranks share a node with peer access (DD_HALO=nccl turns it off); the
    reductions and the odd exchange of a work array stay with NCCL."""

    def __init__(self):
        import os
        import torch
        import torch.distributed as dist
        self.torch, self.dist = torch, dist
        self.peer = None
        self.peer_tried = os.environ.get("DD_HALO", "peer") == "nccl" or dist.get_backend() != "nccl"

    def _peer(self, m):
        if not self.peer_tried:
            self.peer_tried = True
            ok = 1
            try:
                peer = _PeerHalo(m, self.dist)
                ok = 1 if peer.usable() else 0
            except Exception:  # no IPC / peer access between these devices: every rank must agree on the fallback
                peer, ok = None, 0
            flag = self.torch.tensor([ok], device="cuda", dtype=self.torch.int32)
            self.dist.all_reduce(flag, op=self.dist.ReduceOp.MIN)
            if int(flag.item()) == 1:
                self.peer = peer
            elif peer is not None:
                peer.close()
        return self.peer

    def exchange(self, meshes, slot, which):
        m = meshes[0]
        peer = self._peer(m)
        if peer is not None:
            for v in which:  # on the library's own stream: ordered with the kernels by construction
                peer.push(slot, v)
            return
        with _Ordered(meshes, self.torch):
            exchange_halos([m.field(slot, v) for v in which], m.part, m.rank, m.world, m.G, self.dist)

    def exchange_tensors(self, meshes, tensors):
        """Halo rows of one arbitrary (local rows x pitch) tensor per mesh, e.g. the iterate of a solve segment."""
        m = meshes[0]
        with _Ordered(meshes, self.torch):
            exchange_halos([tensors[0]], m.part, m.rank, m.world, m.G, self.dist)

    def allreduce(self, tensors, op, meshes=()):
        d = self.dist
        with _Ordered(meshes, self.torch):
            d.all_reduce(tensors[0], op={"max": d.ReduceOp.MAX, "min": d.ReduceOp.MIN, "sum": d.ReduceOp.SUM}[op])


class _LocalComm:
    """All slabs live in one process on one GPU (tests, single-GPU emulation of the multi-rank path):
    halo exchange = device-to-device copies between the slabs' tensors."""

    def __init__(self):
        import torch
        self.torch = torch

    def exchange(self, meshes, slot, which):
        with _Ordered(meshes, self.torch):
            self._exchange(meshes, slot, which)

    def _exchange(self, meshes, slot, which):
        for m in meshes:
            p = m.part
            for v in which:
                t = m.field(slot, v)
                if m.rank > 0:
                    up = meshes[m.rank - 1]
                    g = min(m.G, p["lo"])
                    t[p["own0"] - g:p["own0"]].copy_(up.field(slot, v)[up.part["own1"] - g:up.part["own1"]])
                if m.rank < m.world - 1:
                    dn = meshes[m.rank + 1]
                    g = min(m.G, p["hi"])
                    t[p["own1"]:p["own1"] + g].copy_(dn.field(slot, v)[dn.part["own0"]:dn.part["own0"] + g])

    def exchange_tensors(self, meshes, tensors):
        with _Ordered(meshes, self.torch):
            for m, t in zip(meshes, tensors):
                p = m.part
                if m.rank > 0:
                    up, tu = meshes[m.rank - 1], tensors[m.rank - 1]
                    g = min(m.G, p["lo"])
                    t[p["own0"] - g:p["own0"]].copy_(tu[up.part["own1"] - g:up.part["own1"]])
                if m.rank < m.world - 1:
                    dn, td = meshes[m.rank + 1], tensors[m.rank + 1]
                    g = min(m.G, p["hi"])
                    t[p["own1"]:p["own1"] + g].copy_(td[dn.part["own0"]:dn.part["own0"] + g])

    def allreduce(self, tensors, op, meshes=()):
        torch = self.torch
        with _Ordered(meshes, torch):
            stacked = torch.stack([t for t in tensors])
            red = {"max": stacked.max(dim=0).values, "min": stacked.min(dim=0).values,
                   "sum": stacked.sum(dim=0)}[op]
            for t in tensors:
                t.copy_(red)


class SlabMesh:
    """One trajectory on a mesh split into row slabs over `world` ranks.  `group` (optional) is the list
    of all slabs when they live in one process (single-GPU emulation); otherwise one rank per process."""

    def __init__(self, x, y, *, world: int = 1, rank: int = 0, ctx=None, halo: int = 24, nslots: int = 2):
        self.x, self.y = as_f64(x), as_f64(y)
        self.world, self.rank = world, rank
        self.G = halo if world > 1 else 0
        if world > 1:
            # every rank sends G of its OWN rows to each neighbour: a slab thinner than the halo would send stale
            # halo rows (and post a different size than its neighbour expects)
            thin = min(b - a for a, b in (shard_members(len(self.x), world, r) for r in range(world)))
            if thin < self.G:
                raise ValueError(f"SlabMesh: {len(self.x)} rows over {world} ranks leaves a rank with {thin} rows, "
                                 f"fewer than the halo of {self.G}; use fewer ranks or a shallower halo")
        self.part = slab_rows(len(self.x), world, rank, self.G)
        p = self.part
        self.batch = ddcore.Batch(self.x, self.y, 1, ctx=ctx, nslots=nslots, row0=p["row0"], nrows=p["nrows"],
                                  own=(p["own0"], p["own1"]))
        self.last_stats: dict = {}
        self._tensors: Dict = {}
        self.group: List["SlabMesh"] = [self]
        self.comm = None
        self._rho = None
        # sweep controller state, one set per kind of initial iterate (0: zero, 1: previous increment);
        # mirrors the library's own controller (dd_capi.cu), which the phased step bypasses
        self._ctl = [dict(extra=[0, 0, 0], floor=[1, 1, 1], plan=None) for _ in range(2)]
        self._prev = None  # (slot_in, slot_out, dt) of the last accepted step
        self._pending = None  # record of a deferred, not yet verified step
        if world > 1:
            import torch
            self.torch = torch

    @classmethod
    def local_group(cls, x, y, world: int, ctx=None, halo: int = 24, nslots: int = 2) -> List["SlabMesh"]:
        """`world` slabs in this process (one GPU) exchanging halos by device copies."""
        meshes = [cls(x, y, world=world, rank=r, ctx=ctx, halo=halo, nslots=nslots) for r in range(world)]
        comm = _LocalComm()
        for m in meshes:
            m.group, m.comm = meshes, comm
        return meshes

    def _comm(self):
        if self.comm is None:
            self.comm = _DistComm()
        return self.comm

    # -- torch views of device fields ---------------------------------------------
    def _tensor(self, key, ptr, shape=None, dtype="<f8"):
        # keyed on the address and shape as well: the library re-allocates some buffers (statistics, cs iteration
        # arrays) when they have to grow, and a view cached by name alone would then point at freed memory
        shape = tuple(shape or self.batch.shape)
        key = (key, int(ptr), shape, dtype)
        t = self._tensors.get(key)
        if t is None:
            t = self.torch.as_tensor(_DevArray(ptr, shape, dtype), device="cuda")
            self._tensors[key] = t
        return t

    def field(self, slot: int, var: str):
        return self._tensor(("s", slot, var), self.batch.dev_ptr(slot, var)[0])

    def exchange(self, slot: int, which: Sequence[str] = VARS):
        self._comm().exchange(self.group, slot, which)

    # -- state ------------------------------------------------------------------------
    def fill_exact(self, slot: int, t: float):
        self._prev = None
        self.batch.fill_exact(slot, t)

    def owned(self, slot: int) -> Dict[str, np.ndarray]:
        d = self.batch.download(slot)
        return {v: a[self.part["own0"]:self.part["own1"]] for v, a in d.items()}

    # -- stepping ---------------------------------------------------------------------
    def step_pc(self, slot_in: int, slot_out: int, t0: float, dt: float, opt: Optional[dd_pc_options] = None,
                defer: bool = False):
        """One PC step of the whole mesh.  With a local group, call it on any member: all slabs advance.

        defer=True: the step is only enqueued; the convergence summaries of the previous deferred step are read
        instead (its statistics are returned, None for the first), so the host never drains the stream between
        steps.  A rejected step is redone together with its successor, which needs its input: rotate three
        slots and call flush() at the end."""
        opt = opt or ddcore.pc_options()
        if self.world == 1:
            st = self.batch.step_pc(slot_in, slot_out, t0, dt, opt, defer=defer)
            if st is not None:
                self.last_stats = st
            return st
        if not defer:
            self.flush()
            return self._step_verified(slot_in, slot_out, t0, dt, opt)
        old = self._pending
        if old is not None and slot_in != old["slot_out"]:
            raise AssertionError("deferred step must continue from the output slot of the pending step")
        new = self._enqueue(slot_in, slot_out, t0, dt, opt, 0)
        stats = None
        if old is not None:
            ok, stats = self._finish(old)
            if not ok:
                if slot_out == old["slot_in"]:
                    raise ddcore.DDNotConverged("a deferred step was rejected after its input slot had been "
                                                "reused; rotate three slots")
                for m in self.group:
                    m.batch.ctx.synchronize()
                    m._pending = None
                self._on_reject(old, stats)
                stats = self._step_verified(old["slot_in"], old["slot_out"], old["t0"], old["dt"], old["opt"], 1)
                new = self._enqueue(slot_in, slot_out, t0, dt, opt, 0)
        for m in self.group:
            m._pending = new
        return stats

    def flush(self):
        """Settles a pending deferred step (redoing it with more sweeps when it was rejected)."""
        if self.world == 1:
            st = self.batch.flush()
            if st is not None:
                self.last_stats = st
            return st
        old = self._pending
        if old is None:
            return None
        for m in self.group:
            m._pending = None
        ok, stats = self._finish(old)
        peer = getattr(self.comm, "peer", None)
        if peer is not None:
            peer.check()
        if not ok:
            for m in self.group:
                m.batch.ctx.synchronize()
            self._on_reject(old, stats)
            stats = self._step_verified(old["slot_in"], old["slot_out"], old["t0"], old["dt"], old["opt"], 1)
        return stats

    def _step_verified(self, slot_in, slot_out, t0, dt, opt, first_attempt: int = 0):
        for attempt in range(first_attempt, 40):
            rec = self._enqueue(slot_in, slot_out, t0, dt, opt, attempt)
            ok, stats = self._finish(rec)
            if ok:
                return stats
            self._on_reject(rec, stats)
        raise ddcore.DDNotConverged("slab step: linear solve did not reach the residual bound")

    def _enqueue(self, slot_in, slot_out, t0, dt, opt, attempt):
        """All phases of one step, the reductions over ranks and the read-back of the reduced summaries into
        pinned host memory; nothing is waited for."""
        torch, comm, group = self.torch, self._comm(), self.group
        t0a, dta = as_f64([t0]), as_f64([dt])
        # same rule as the library (take_guess in dd_capi.cu): ping-pong steps start from the last increment
        mode = int(bool(opt.extrapolate_guess) and attempt == 0 and self._prev == (slot_out, slot_in, dt))
        ctl = self._ctl[mode]
        plan = self._common_plan(opt, ctl)

        def phase(k):
            for m in group:
                b = m.batch
                b.ctx.check(b.lib.dd_step_pc_phase(b.handle, k, slot_in, slot_out, dptr(t0a), dptr(dta), 1,
                                                   C.byref(opt), None, None), f"step_pc_phase {k}")
        track = opt.consec_xs_rtol > 0.0 and opt.num_newton_iterations > 0
        phase(0)
        # The SOR relaxation factor must be the same on every rank.  With a fixed sweep plan (bitwise
        # reproducibility across decompositions) and on the first steps it comes from the current Gershgorin
        # ratio, max-reduced before each solve; otherwise from the all-reduced ratios of the last verified step,
        # which the host already holds (they move by O(dt) per step and convergence is verified anyway) --
        # three collectives less on the critical path.
        lag = opt.fixed_sweeps <= 0 and self._rho is not None and all(0.0 <= float(r) < 1.0 for r in self._rho)
        rel = [float(r) for r in self._rho] if lag else [-1.0, -1.0, -1.0]
        # a variable whose SOR solves diverged runs plain Gauss-Seidel from then on (ratio 0: omega = 1; _on_reject)
        rel = [0.0 if g else r for r, g in zip(rel, self._gs_only())]
        arr = (C.c_double * 3)(*rel)
        for m in group:
            m.batch.ctx.check(m.batch.lib.dd_batch_set_relax_rho(m.batch.handle, C.byref(arr)), "set_relax_rho")
        limit = self.sweep_limit
        for k, var in ((1, "T"), (2, "cl"), (3, "cd")):
            phase(20 + k)  # assemble (T: done by the predictor on wide grids)
            if not lag:
                comm.allreduce([m._tensor("stats", m.batch.work_dev_ptr("solve_stats"), (3, 5))[k - 1, 0:1]
                                for m in group], "max", group)
            if plan[k - 1] <= limit:
                phase(30 + k)
            else:
                self._solve_in_segments(k, slot_in, slot_out, opt, plan[k - 1], limit)
            comm.exchange(group, slot_out, (var,))
        phase(4)
        if track:
            n = opt.num_newton_iterations
            comm.allreduce([m._tensor("itmax", m.batch.work_dev_ptr("cs_it_max"), (n,)) for m in group], "max",
                           group)
            comm.allreduce([m._tensor("itmin", m.batch.work_dev_ptr("cs_it_min"), (n,)) for m in group], "min",
                           group)
        phase(6)
        with _Ordered(group, torch):
            # rows 0..2: the three solves; row 3, column 0: the step's domain-error flag (HCsTriple corrector)
            summ = [m._tensor("summary", m.batch.work_dev_ptr("summary"), (4, 4)) for m in group]
            red = [torch.nan_to_num(t, nan=1e300) for t in summ]
        comm.allreduce(red, "max", group)
        with _Ordered(group, torch):
            host = self._host_buffers(attempt)
            host["summary"].copy_(red[0], non_blocking=True)
            if track:
                used = self._tensor("cs_used", self.batch.work_dev_ptr("cs_used"), (1,), dtype="<i4")
                host["used"].copy_(used, non_blocking=True)
            done = torch.cuda.Event()
            done.record()
        return dict(slot_in=slot_in, slot_out=slot_out, t0=t0, dt=dt, opt=opt, plan=plan, mode=mode, attempt=attempt,
                    host=host, done=done, track=track)

    @property
    def sweep_limit(self) -> int:
        """SOR sweeps a halo of G rows supports between two exchanges of the iterate (every sweep consumes two
        rows of valid data on each interior side, the assembled rows start one row in, the residual needs one more)."""
        return max(1, (self.G - 3) // 2)

    def _solve_in_segments(self, k, slot_in, slot_out, opt, sweeps, limit):
        """A solve that needs more sweeps than the halo supports in one go (weakly dominant matrices of large time
        steps): segments of at most `limit` sweeps, the iterate's halo rows exchanged in between.  Still the global
        red-black iteration, bit for bit."""
        comm, group = self._comm(), self.group
        chunks = [limit] * (sweeps // limit) + ([sweeps % limit] if sweeps % limit else [])
        for ci, c in enumerate(chunks):
            last = ci == len(chunks) - 1
            for m in group:
                b = m.batch
                b.ctx.check(b.lib.dd_pc_solve_segment(b.handle, k, slot_in, slot_out, C.byref(opt), c, int(ci == 0),
                                                      int(last)), "pc_solve_segment")
            if not last:
                ts = []
                for m in group:
                    ncols = m.batch.shape[1]
                    ts.append(m._tensor("x_cur", m.batch.work_dev_ptr("x_cur"), (m.batch.shape[0], ncols + (ncols & 1))))
                comm.exchange_tensors(group, ts)

    def _host_buffers(self, attempt):
        """Two alternating sets of pinned read-back buffers (one step may be pending while the next is enqueued)."""
        torch = self.torch
        if not hasattr(self, "_hostbuf"):
            self._hostbuf = [dict(summary=torch.zeros((4, 4), dtype=torch.float64).pin_memory(),
                                  used=torch.zeros((1,), dtype=torch.int32).pin_memory()) for _ in range(2)]
            self._hostsel = 0
        self._hostsel ^= 1
        return self._hostbuf[self._hostsel]

    def _finish(self, rec):
        """Waits for the reduced summaries of the step in `rec`; updates the sweep controller."""
        rec["done"].synchronize()
        opt, plan, mode = rec["opt"], rec["plan"], rec["mode"]
        ctl = self._ctl[mode]
        s4 = rec["host"]["summary"].numpy().copy()
        if s4[3, 0] > 0.0:  # same condition, same exception as the reference (src/prob1base.py:3417-3420)
            raise ValueError("Denominator 2 - dt Kd (Sd - Cd1) (1 + Cl1) below positiveness treshold.")
        s = s4[:3]
        iters = int(rec["host"]["used"][0]) if rec["track"] else int(opt.num_newton_iterations)
        stats = dict(sweeps=list(plan), guess=mode, rho=[float(x) for x in s[:, 0]],
                     resid=[float(x) for x in s[:, 2]], bound=[float(x) for x in s[:, 3]],
                     ratio=[float(x) for x in s[:, 1]], retries=rec["attempt"], cs_newton_iters=iters)
        ok = bool(np.all(s[:, 1] <= 1.0))
        for m in self.group:
            m._rho, m.last_stats = s[:, 0], stats
        if ok:
            # (a negative ratio marks a system that was NaN / Inf on entry: nothing to learn from it)
            nxt = [max(self.batch.lib.dd_next_plan(int(p), float(r), float(q), opt.max_sweeps), f)
                   if (q >= 0.0 and not g) else p
                   for p, r, q, f, g in zip(plan, s[:, 0], s[:, 1], ctl["floor"], self._gs_only())]
            for m in self.group:
                m._ctl[mode]["plan"] = nxt
                m._prev = (rec["slot_in"], rec["slot_out"], rec["dt"])
        return ok, stats

    def _on_reject(self, rec, stats):
        """Remembers the failing sweep counts (never plan below them again) and falls back to the theoretical
        plan; raises when the halo cannot support more sweeps."""
        opt, plan, mode = rec["opt"], rec["plan"], rec["mode"]
        ctl = self._ctl[mode]
        ratio, rho = stats["ratio"], stats["rho"]
        for m in self.group:
            m._ctl[mode]["plan"] = None
            m._prev = None
        if any(r > 1.0 and p >= opt.max_sweeps for p, r in zip(plan, ratio)) or opt.fixed_sweeps > 0:
            raise ddcore.DDNotConverged(f"slab step: {max(plan)} SOR sweeps do not reach the residual bound. "
                                        f"stats={stats}")
        lib = self.batch.lib
        # Same fallback as the library (gs_fallback in dd_capi.cu): SOR with the over-relaxation of a symmetric
        # matrix can diverge on the nonsymmetric cd system of very large time steps; once three times the
        # theoretical count has failed, the variable is solved by Gauss-Seidel (error factor <= rho per sweep).
        gs = self._gs_only()
        for k, (p, r, x) in enumerate(zip(plan, ratio, rho)):
            x = float(x)
            if r > 1.0 and not gs[k] and 0.0 <= x < 1.0 and \
                    p >= 3 * lib.dd_sweeps_for_rho(x * 1.02 + 1e-12, opt.max_sweeps) + 8:
                gs[k] = True
                for m in self.group:
                    for md in (0, 1):
                        m._ctl[md]["extra"][k], m._ctl[md]["floor"][k] = 0, 1
        floor = [max(f, p + 1) if r > 1.0 else f for f, p, r in zip(ctl["floor"], plan, ratio)]
        theory = [self._gs_sweeps(float(x), opt.max_sweeps) if g else
                  lib.dd_sweeps_for_rho(float(x) * 1.02 + 1e-12, opt.max_sweeps) for x, g in zip(rho, gs)]
        extra = [e + (p + 1) // 2 + 1 if (r > 1.0 and p >= th) else e
                 for e, p, r, th in zip(ctl["extra"], plan, ratio, theory)]
        for m in self.group:
            m._ctl[mode]["extra"], m._ctl[mode]["floor"] = extra, floor

    def _gs_only(self) -> List[bool]:
        """Per variable: its solves have fallen back to Gauss-Seidel (shared by the ranks of the group)."""
        lead = self.group[0]
        if not hasattr(lead, "_gs_flags"):
            lead._gs_flags = [False, False, False]
        return lead._gs_flags

    @staticmethod
    def _gs_sweeps(rho: float, max_sweeps: int) -> int:
        rho = rho * 1.02 + 1e-12
        if not (0.0 <= rho < 1.0):
            return max_sweeps
        if rho < 1e-300:
            return 2
        return int(min(max_sweeps, max(2, math.ceil(math.log(1e-17) / math.log(rho)))))

    def _common_plan(self, opt, ctl) -> List[int]:
        """Same number of SOR sweeps on every rank, planned from the all-reduced Gershgorin ratios of the
        previous step (first step: as many sweeps as the halo supports)."""
        lib = self.batch.lib
        limit = self.sweep_limit
        if opt.fixed_sweeps > 0:
            plan = [opt.fixed_sweeps] * 3
        elif self._rho is None:
            plan = [limit] * 3  # first step: what one halo exchange supports; the verified residual takes it from there
        elif ctl["plan"] is not None:
            plan = list(ctl["plan"])
        else:
            plan = [max((self._gs_sweeps(float(r), opt.max_sweeps) if g else
                         lib.dd_sweeps_for_rho(float(r) * 1.02 + 1e-12, opt.max_sweeps)) + e, f)
                    for r, e, f, g in zip(self._rho, ctl["extra"], ctl["floor"], self._gs_only())]
        # plans beyond `limit` sweeps are run in segments with the iterate's halo exchanged in between
        plan = [max(1, min(int(p), opt.max_sweeps)) for p in plan]
        arr = (C.c_int * 3)(*plan)
        for m in self.group:
            m.batch.ctx.check(lib.dd_batch_set_plan(m.batch.handle, C.byref(arr)), "set_plan")
        return plan

    def step_feuler(self, slot_in: int, slot_out: int, t0: float, dt: float):
        for m in self.group:
            m.batch.step_feuler(slot_in, slot_out, t0, dt)
        if self.world > 1:
            self._comm().exchange(self.group, slot_out, VARS)

    def error_norms(self, slot: int, t: float) -> np.ndarray:
        if self.world == 1:
            return self.batch.error_norms(slot, t)[0]
        ts = [self.torch.tensor(m.batch.error_norms(slot, t)[0], device="cuda") for m in self.group]
        self._comm().allreduce(ts, "sum")
        return ts[0].cpu().numpy()
