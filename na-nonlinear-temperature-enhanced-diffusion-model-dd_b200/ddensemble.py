"""Ensembles of independent trajectories and refinement sweeps on the device path (SURVEY.md 8d configs
2-4, 8e "ensembles / refinement sweeps").

Trajectories never talk to each other, so the members are cut into contiguous blocks, one block per rank
(`ddmesh.shard_members`), every rank steps its block with the batched kernels (one launch advances all
members of a block) and the only communication is one gather of the per-member error scalars at the end.

* `TrajectoryEnsemble`  -- many members on one grid: per-member model constants, eta, dt.
* `RefinementSweep`     -- trials on different grids / step counts (the spatial + temporal + eta studies of
  the notebooks launched together): trials that share grid and step count share a batch, the batches run
  concurrently on their own streams.

The error functional is the reference's combined max-integral norm (src/mms_trial_utils.py:15-53),
evaluated for all members at once from the per-step norms the device returns.
"""

from __future__ import annotations

import math
from concurrent.futures import ThreadPoolExecutor
from typing import Dict, List, Optional, Sequence, Tuple

import numpy as np

import ddcore
from _ddlib import Context
from ddmesh import shard_members

INTEGRAL_VARS = ("T", "cl", "cd")


def steps_and_dt(Tf: float, dt: float, t0: float = 0.0) -> Tuple[int, float]:
    """Number of steps and the adjusted step of a trial (reference src/mms_trial_utils.py:76-77)."""
    n = math.ceil((Tf - t0) / dt)
    return n, (Tf - t0) / n


def _builtin_sum(terms: Sequence[np.ndarray]) -> np.ndarray:
    """Elementwise replica of CPython's float sum() (Objects/bltinmodule.c: Neumaier compensation, the
    correction added once at the end when it is finite and non-zero)."""
    total = np.zeros_like(terms[0])
    comp = np.zeros_like(terms[0])
    with np.errstate(invalid="ignore"):
        for y in terms:
            t = total + y
            comp = comp + np.where(np.abs(total) >= np.abs(y), (total - t) + y, (y - t) + total)
            total = t
        return np.where((comp != 0.0) & np.isfinite(comp), total + comp, total)


def _device_description(case):
    """Table description of the case's exact solution, else its generated program (arbitrary SymPy cases)."""
    spec = case.device_spec()
    if spec is None and hasattr(case, "device_program"):
        spec = case.device_program()
    return spec


def _select_forcing(batch, spec):
    if hasattr(spec, "image"):
        batch.forcing_program(spec)
    else:
        batch.forcing_spec(spec)


def combined_error_norms(norms: np.ndarray, dt) -> Dict[str, np.ndarray]:
    """Combined max-integral error norms of B members at once.

    norms: (nsteps + 1, B, 8) = per step and member H2[cp,T,cl,cd,cs], P2[T,cl,cd] (Batch.run_pc(norms=True));
    dt: scalar or (B,).  Returns overall (B,) and per_var (B, 5), the quantities of
    NumericalErrorSummary (reference src/mms_trial_utils.py:150-190).  A NaN never replaces the running
    maximum (the reference's `max(0.0, nan)`)."""
    norms = np.asarray(norms, dtype=np.float64)
    K, B, _ = norms.shape
    dt = np.broadcast_to(np.asarray(dt, dtype=np.float64), (B,))
    planes = np.ascontiguousarray(np.moveaxis(norms, 2, 0))  # (8, K, B): one contiguous plane per quantity
    h2, p2 = planes[:5], planes[5:]

    def sup(hsq, integrand):
        best = np.zeros(B)
        run = np.zeros(B)
        for k in range(K):
            if k > 0:
                run = run + 0.5 * dt * (integrand[k - 1] + integrand[k])
            val = hsq[k] + run
            best = np.where(val > best, val, best)  # False for NaN: the maximum is kept
        return np.sqrt(best)

    # the reference adds the variables with the builtin sum(), which compensates (Neumaier) since Python 3.12
    hs = _builtin_sum([h2[v] for v in range(5)])
    ig = _builtin_sum([p2[v] for v in range(3)])
    overall = sup(hs, ig)
    per_var = np.zeros((B, 5))
    zero = np.zeros((K, B))
    for v in range(5):
        per_var[:, v] = sup(h2[v], p2[v - 1] if 1 <= v <= 3 else zero)
    return dict(overall=overall, per_var=per_var)


def gather_members(local: np.ndarray, nmembers: int, world: int, rank: int, dist=None, device=None) -> np.ndarray:
    """All ranks' member blocks (first axis) joined in member order.  One all_gather of equally padded
    blocks; `device` = "cuda" for NCCL, None for gloo."""
    local = np.ascontiguousarray(local, dtype=np.float64)
    if world == 1:
        return local
    import torch
    sizes = [b - a for a, b in (shard_members(nmembers, world, r) for r in range(world))]
    assert local.shape[0] == sizes[rank]
    pad = np.zeros((max(sizes),) + local.shape[1:])
    pad[:sizes[rank]] = local
    mine = torch.from_numpy(pad)
    if device is not None:
        mine = mine.to(device)
    parts = [torch.empty_like(mine) for _ in range(world)]
    dist.all_gather(parts, mine)
    return np.concatenate([p.cpu().numpy()[:n] for p, n in zip(parts, sizes)], axis=0)


class TrajectoryEnsemble:
    """`nmembers` independent trials of one MMS case on one grid; this rank owns a contiguous block.

    models: one model object (shared) or a sequence of `nmembers` objects with the ModelConsts attributes;
    etas: scalar or (nmembers,) regularisation factors.  Only the local block is ever materialised.
    """

    def __init__(self, grid, mms_case_cls, models, etas, *, world: int = 1, rank: int = 0,
                 ctx: Optional[Context] = None, chunk: int = 8192, integrator: str = "pc",
                 pc: Optional[dict] = None, mms_case_params: Optional[dict] = None):
        self.grid = grid
        single = not isinstance(models, (list, tuple))
        etas = np.atleast_1d(np.asarray(etas, dtype=np.float64))
        self.nmembers = len(etas) if single else len(models)
        if single:
            models = [models] * self.nmembers
        if len(etas) == 1 and self.nmembers > 1:
            etas = np.full(self.nmembers, etas[0])
        assert len(models) == len(etas) == self.nmembers
        self.world, self.rank = world, rank
        self.first, self.last = shard_members(self.nmembers, world, rank)
        self.models = list(models[self.first:self.last])
        self.etas = etas[self.first:self.last]
        self.ctx = ctx or Context.default()
        self.chunk = int(chunk)
        assert integrator in ("pc", "feuler")
        self.integrator = integrator
        self.opt = ddcore.pc_options(**(pc or {}))
        # the case supplies the device description of its exact solution; it only reads the grid and the
        # shared constants (per-member constants enter the sources on the device)
        self.case = mms_case_cls(grid=grid, model=self.models[0] if self.models else models[0],
                                 **(mms_case_params or {}))
        self.spec = _device_description(self.case)
        if self.spec is None:
            raise ValueError(f"{mms_case_cls.__name__} has no device description (device_spec() is None)")
        self.last_stats: List[dict] = []
        self._batches: Dict[Tuple[int, int], ddcore.Batch] = {}

    @property
    def nlocal(self) -> int:
        return self.last - self.first

    def _batch(self, a: int, b: int):
        """Device batch of the local members [a, b): created on first use, kept for later runs (allocating and
        freeing gigabytes per run costs more than the stepping)."""
        batch = self._batches.get((a, b))
        if batch is None:
            batch = ddcore.Batch(self.grid.x, self.grid.y, b - a, ctx=self.ctx, nslots=2)
            batch.set_models([ddcore.model_struct(m, e) for m, e in zip(self.models[a:b], self.etas[a:b])])
            _select_forcing(batch, self.spec)
            self._batches[(a, b)] = batch
        return batch

    def close(self):
        for batch in self._batches.values():
            batch.close()
        self._batches = {}

    def run_for_errors(self, Tf: float, dt, t0: float = 0.0) -> Dict[str, np.ndarray]:
        """Every local member from the exact state at t0 to Tf; dt scalar or (nmembers,) with a common number
        of steps.  Returns overall (nlocal,), per_var (nlocal, 5), dt_used (nlocal,), nsteps."""
        dt_all = np.broadcast_to(np.asarray(dt, dtype=np.float64), (self.nmembers,))
        sd = [steps_and_dt(Tf, float(d), t0) for d in np.unique(dt_all)]
        if len({n for n, _ in sd}) != 1:
            raise ValueError("members of one ensemble must take the same number of steps")
        nsteps = sd[0][0]
        dt_used = (Tf - t0) / nsteps if len(sd) == 1 else np.array([steps_and_dt(Tf, float(d), t0)[1] for d in dt_all])
        dt_used = np.broadcast_to(np.asarray(dt_used, dtype=np.float64), (self.nmembers,))[self.first:self.last]
        overall = np.zeros(self.nlocal)
        per_var = np.zeros((self.nlocal, 5))
        self.last_stats = []
        for a in range(0, self.nlocal, self.chunk):
            b = min(a + self.chunk, self.nlocal)
            batch = self._batch(a, b)
            batch.fill_exact(0, t0)
            dts = dt_used[a:b] if len(sd) > 1 else dt_used[a:a + 1]
            # time loop, per-step norms and their combination all on the device: 6 doubles per member come back
            res, st = batch.run_errors(0, 1, t0, dts, nsteps, self.opt, integrator=self.integrator)
            if st is not None:
                self.last_stats.append(st)
            overall[a:b], per_var[a:b] = res["overall"], res["per_var"]
        return dict(overall=overall, per_var=per_var, dt_used=np.array(dt_used), nsteps=nsteps)

    def gather(self, local: np.ndarray, dist=None, device=None) -> np.ndarray:
        return gather_members(local, self.nmembers, self.world, self.rank, dist, device)


class RefinementSweep:
    """A list of trials `dict(N=, M=, dt=, Tf=, eta=, t0=0.0)` of one case and model, e.g. the spatial,
    temporal and eta studies of a notebook, launched together.  Trials with the same grid and step count
    form one batch; the batches are split over ranks by cost (N*M*steps) and run concurrently, each on its
    own stream, from a small pool of host threads."""

    def __init__(self, mms_case_cls, model, trials: Sequence[dict], *, world: int = 1, rank: int = 0,
                 device: int = 0, make_grid=None, pc: Optional[dict] = None, streams: int = 8):
        import prob1base as p1
        self.p1 = p1
        self.case_cls, self.model = mms_case_cls, model
        self.trials = [dict(t) for t in trials]
        self.world, self.rank, self.device = world, rank, device
        self.make_grid = make_grid or p1.make_uniform_grid
        self.opt = ddcore.pc_options(**(pc or {}))
        self.streams = streams
        groups: Dict[Tuple[int, int, int], List[int]] = {}
        for k, t in enumerate(self.trials):
            t.setdefault("M", t["N"])
            t.setdefault("t0", 0.0)
            n, d = steps_and_dt(t["Tf"], t["dt"], t["t0"])
            t["nsteps"], t["dt_used"] = n, d
            groups.setdefault((t["N"], t["M"], n), []).append(k)
        self.groups = [dict(key=key, members=m, cost=key[0] * key[1] * key[2] * len(m)) for key, m in groups.items()]
        self.assignment = self.balance([g["cost"] for g in self.groups], world)

    @staticmethod
    def balance(costs: Sequence[float], world: int) -> List[int]:
        """Greedy longest-processing-time assignment of groups to ranks (balance
        heterogeneous levels by N*M*steps); deterministic, so every rank computes the same table."""
        load = [0.0] * world
        owner = [0] * len(costs)
        for g in sorted(range(len(costs)), key=lambda q: (-costs[q], q)):
            r = min(range(world), key=lambda q: (load[q], q))
            owner[g] = r
            load[r] += costs[g]
        return owner

    def _run_group(self, g) -> List[Tuple[int, float, np.ndarray]]:
        N, M, nsteps = g["key"]
        ts = [self.trials[k] for k in g["members"]]
        if "batch" not in g:  # context (own stream), batch and forcing tables are kept for later runs
            grid = self.make_grid(N, M)
            ctx = Context(self.device)
            batch = ddcore.Batch(grid.x, grid.y, len(ts), ctx=ctx, nslots=2)
            batch.set_models([ddcore.model_struct(self.model, t["eta"]) for t in ts])
            spec = _device_description(self.case_cls(grid=grid, model=self.model))
            if spec is None:
                batch.close()
                raise ValueError(f"{self.case_cls.__name__} has no device description")
            _select_forcing(batch, spec)
            g["ctx"], g["batch"] = ctx, batch
        batch = g["batch"]
        t0 = np.array([t["t0"] for t in ts])
        batch.fill_exact(0, t0)
        dts = np.array([t["dt_used"] for t in ts])
        res, _ = batch.run_errors(0, 1, t0, dts, nsteps, self.opt)
        return [(k, float(res["overall"][q]), res["per_var"][q]) for q, k in enumerate(g["members"])]

    def close(self):
        for g in self.groups:
            if "batch" in g:
                g.pop("batch").close()
                g.pop("ctx").close()

    def run_for_errors(self) -> Dict[str, np.ndarray]:
        """Runs this rank's groups; returns overall (ntrials,), per_var (ntrials, 5) with NaN for trials owned
        by other ranks, and `owned` (bool mask).  Use `merge` to combine ranks."""
        n = len(self.trials)
        overall = np.full(n, np.nan)
        per_var = np.full((n, 5), np.nan)
        owned = np.zeros(n, dtype=bool)
        mine = [g for g, r in zip(self.groups, self.assignment) if r == self.rank]
        mine.sort(key=lambda g: -g["cost"])
        if mine:
            with ThreadPoolExecutor(max_workers=max(1, min(self.streams, len(mine)))) as pool:
                for out in pool.map(self._run_group, mine):
                    for k, e, pv in out:
                        overall[k], per_var[k], owned[k] = e, pv, True
        return dict(overall=overall, per_var=per_var, owned=owned)

    def merge(self, res: Dict[str, np.ndarray], dist=None, device=None) -> Dict[str, np.ndarray]:
        """Every rank gets every trial's result (one all_reduce(sum) of the owner-masked table)."""
        if self.world == 1:
            return res
        import torch
        tab = np.concatenate([np.where(res["owned"], res["overall"], 0.0)[:, None],
                              np.where(res["owned"][:, None], res["per_var"], 0.0),
                              res["owned"][:, None].astype(np.float64)], axis=1)
        t = torch.from_numpy(tab)
        if device is not None:
            t = t.to(device)
        dist.all_reduce(t, op=dist.ReduceOp.SUM)
        tab = t.cpu().numpy()
        assert np.all(tab[:, 6] == 1.0), "every trial must be owned by exactly one rank"
        return dict(overall=tab[:, 0], per_var=tab[:, 1:6], owned=np.ones(len(tab), dtype=bool))
