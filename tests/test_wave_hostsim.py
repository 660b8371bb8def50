"""Host emulation of the wavefront SOR kernel (csrc/dd_wave.cuh compiled for the CPU, tests/hostsim/wavesim.cpp)
against the plain global red-black SOR with the same arithmetic: bit for bit, for every kernel variant, for
marches split over several CTAs, slab-local row ranges and passes that continue from a previous iterate."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200", "csrc")
BUILD = os.path.join(HERE, "hostsim", "_build")
LIB = os.path.join(BUILD, "libwavesim.so")
_dp = C.POINTER(C.c_double)


class WSProblem(C.Structure):
    _fields_ = [(n, C.c_int) for n in ("N", "M", "row0", "nrows", "ld", "ldR", "own0", "own1", "vr0", "vr1", "cb", "C",
                                       "nwarps", "nctas", "sweeps", "last_pass", "zero_boundary", "order")] + \
               [(n, C.c_double) for n in ("rho", "dt", "DT")] + \
               [(n, _dp) for n in ("rh", "rhp", "rk", "rkp", "bb", "aW", "aE", "aS", "aN", "xin", "vstar", "xout",
                                   "vnew")] + [("stats", C.c_double * 4), ("steps", C.c_longlong)]


def _lib():
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "hostsim", "wavesim.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in ("dd_wave.cuh", "dd_lane.cuh", "dd_sor.cuh", "dd_nodeprog.cuh", "dd_physics.cuh",
                                                    "dd_types.h")]
    if not (os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps)):
        subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I", CSRC,
                               "-o", LIB, src])
    lib = C.CDLL(LIB)
    lib.ws_wave.restype = C.c_int
    lib.ws_lane.restype = C.c_int
    lib.ws_reference.restype = C.c_int
    return lib


def _ptr(a):
    return a.ctypes.data_as(_dp) if a is not None else None


def _metrics(x):
    h = np.diff(x)
    n = len(x) - 1
    rh = np.zeros(n + 2)
    rh[1:n + 1] = 1.0 / h
    rhp = np.zeros(n + 1)
    rhp[1:n] = 1.0 / (0.5 * (h[:-1] + h[1:]))
    return rh, rhp


def _system(rng, N, M, cb, nonuniform):
    """Rows of a strictly diagonally dominant five-point system on the global grid (zero rows on the boundary)."""
    ldR = M + 1 + ((M + 1) & 1)
    x = np.linspace(0, 1, N + 1) ** (1.2 if nonuniform else 1.0)
    y = np.linspace(0, 1, M + 1) ** (1.1 if nonuniform else 1.0)
    rh, rhp = _metrics(x)
    rk, rkp = _metrics(y)
    inter = np.zeros((N + 1, M + 1), bool)
    inter[1:N, 1:M] = True
    arr = {k: np.zeros((N + 1, ldR)) for k in ("bb", "aW", "aE", "aS", "aN")}
    arr["bb"][:, :M + 1] = np.where(inter, rng.normal(size=(N + 1, M + 1)), 0.0)
    if cb:
        # dinv of 2 + dt (sum of couplings + K3 cp): the couplings are dt DT / (hhat h) ~ N^2 dt DT
        # strong coupling (Gershgorin ratio ~ 0.9): data one halo cell too far away must change the result visibly
        dt, DT = 4.0 / max(N, M) ** 2, 1.0
        fT = dt * DT
        rW = fT * rhp[:N + 1] * rh[:N + 1]
        rE = fT * rhp[:N + 1] * rh[1:N + 2]
        cS = fT * rkp[:M + 1] * rk[:M + 1]
        cN = fT * rkp[:M + 1] * rk[1:M + 2]
        d = 2.0 + (rW + rE)[:, None] + (cS + cN)[None, :] + 1e-3 * rng.uniform(size=(N + 1, M + 1))
        arr["aW"][:, :M + 1] = np.where(inter, 1.0 / d, 0.0)
        rho = float(np.max(((rW + rE)[:, None] + (cS + cN)[None, :]) / d))
    else:
        dt, DT = 1.0, 1.0
        for k in ("aW", "aE", "aS", "aN"):
            arr[k][:, :M + 1] = np.where(inter, rng.uniform(0.1, 0.24, size=(N + 1, M + 1)), 0.0)
        arr["aW"][1, :] = 0.0
        arr["aE"][N - 1, :] = 0.0
        arr["aS"][:, 1] = 0.0
        arr["aN"][:, M - 1] = 0.0
        rho = float(np.max(sum(np.abs(arr[k]) for k in ("aW", "aE", "aS", "aN"))))
    # garbage in the padding column must never be read into a result
    for k in arr:
        arr[k][:, M + 1:] = 1e300
    return arr, dict(rh=rh, rhp=rhp, rk=rk, rkp=rkp), rho, dt, DT, ldR


def _run(lib, N, M, arr, met, rho, dt, DT, ldR, *, cb, Cc, nwarps, nctas, sweeps, row0=0, nrows=None, own=None,
         vr=None, last=1, xin=None, order=0, zero_boundary=0, vstar=None, reference=False, kernel="wave"):
    nrows = N + 1 if nrows is None else nrows
    own = (0, nrows) if own is None else own
    vr = (0, nrows) if vr is None else vr
    loc = {k: np.ascontiguousarray(v[row0:row0 + nrows]) for k, v in arr.items()}
    ld = M + 1
    P = WSProblem()
    P.N, P.M, P.row0, P.nrows, P.ld, P.ldR = N, M, row0, nrows, ld, ldR
    P.own0, P.own1, P.vr0, P.vr1 = own[0], own[1], vr[0], vr[1]
    P.cb, P.C, P.nwarps, P.nctas, P.sweeps, P.last_pass, P.zero_boundary, P.order = cb, Cc, nwarps, nctas, sweeps, last, \
        zero_boundary, order
    P.rho, P.dt, P.DT = rho, dt, DT
    keep = [np.ascontiguousarray(met[k]) for k in ("rh", "rhp", "rk", "rkp")]
    P.rh, P.rhp, P.rk, P.rkp = [_ptr(a) for a in keep]
    for k in ("bb", "aW", "aE", "aS", "aN"):
        setattr(P, k, _ptr(loc[k]))
    xin_l = None if xin is None else np.ascontiguousarray(xin)
    P.xin = _ptr(xin_l)
    vs = np.ascontiguousarray(vstar[row0:row0 + nrows])
    P.vstar = _ptr(vs)
    xout = np.full((nrows, ldR), np.nan)
    vnew = np.full((nrows, ld), np.nan)
    P.xout, P.vnew = _ptr(xout), _ptr(vnew)
    if reference:
        x = np.zeros((nrows, ldR))
        assert lib.ws_reference(C.byref(P), _ptr(x)) == 0
        return x, vnew, list(P.stats), 0
    rc = (lib.ws_lane if kernel == "lane" else lib.ws_wave)(C.byref(P))
    assert rc == 0, rc
    return xout, vnew, list(P.stats), P.steps


CASES = [
    # cb, C, nwarps, sweeps, N, M, nctas
    (1, 4, 16, 7, 70, 300, 3),
    (1, 4, 16, 3, 45, 259, 1),
    (1, 3, 20, 9, 90, 200, 4),
    (1, 2, 24, 11, 60, 150, 2),
    (1, 2, 6, 1, 33, 140, 5),
    (0, 2, 12, 5, 80, 260, 3),
    (0, 2, 16, 7, 64, 131, 2),
    (0, 2, 8, 3, 50, 250, 7),
    (0, 1, 24, 10, 70, 70, 2),
    (0, 1, 4, 1, 20, 65, 1),
]


@pytest.mark.parametrize("cb,Cc,nwarps,sweeps,N,M,nctas", CASES)
@pytest.mark.parametrize("nonuniform", [False, True])
def test_wave_equals_global_sor(cb, Cc, nwarps, sweeps, N, M, nctas, nonuniform):
    lib = _lib()
    rng = np.random.default_rng(1000 * sweeps + N + M)
    arr, met, rho, dt, DT, ldR = _system(rng, N, M, cb, nonuniform)
    vstar = rng.normal(size=(N + 1, M + 1))
    kw = dict(cb=cb, Cc=Cc, nwarps=nwarps, nctas=nctas, sweeps=sweeps, vstar=vstar, zero_boundary=cb)
    xr, vr_, sr, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, reference=True, **kw)
    for order in (0, 1):
        _, vn, st, steps = _run(lib, N, M, arr, met, rho, dt, DT, ldR, order=order, **kw)
        assert np.array_equal(vn, vr_), (order, np.argwhere(vn != vr_)[:5])
        assert st == sr


@pytest.mark.parametrize("cb,Cc,nwarps,sweeps", [(1, 4, 16, 7), (0, 2, 12, 5), (0, 2, 8, 3)])
def test_wave_on_a_slab(cb, Cc, nwarps, sweeps):
    """Rows [row0, row0 + nrows) of a taller mesh, owned rows in the middle, halo rows recomputed: the owned rows
    equal the global iteration on the whole mesh."""
    lib = _lib()
    N, M = 140, 200
    rng = np.random.default_rng(7 + sweeps)
    arr, met, rho, dt, DT, ldR = _system(rng, N, M, cb, True)
    vstar = rng.normal(size=(N + 1, M + 1))
    kw = dict(cb=cb, Cc=Cc, nwarps=nwarps, sweeps=sweeps, vstar=vstar, zero_boundary=cb)
    _, vglob, _, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, nctas=1, reference=True, **kw)
    G = 2 * sweeps + 1
    for (a, b) in ((0, 51), (51, 97), (97, N + 1)):  # three slabs; odd first rows on purpose
        row0 = max(0, a - G - 2)
        row1 = min(N + 1, b + G + 2)
        own = (a - row0, b - row0)
        # assembled rows exist one row inside the local range unless it ends on the physical boundary
        vr = (0 if row0 == 0 else 1, row1 - row0 if row1 == N + 1 else row1 - row0 - 1)
        _, vn, _, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, nctas=3, row0=row0, nrows=row1 - row0, own=own, vr=vr,
                           **kw)
        assert np.array_equal(vn[own[0]:own[1]], vglob[a:b])


def test_wave_two_passes_continue_the_iteration():
    """5 + 4 sweeps in two passes (x through an array) = 9 sweeps in one."""
    lib = _lib()
    N, M, cb = 60, 180, 0
    rng = np.random.default_rng(99)
    arr, met, rho, dt, DT, ldR = _system(rng, N, M, cb, False)
    vstar = rng.normal(size=(N + 1, M + 1))
    kw = dict(cb=cb, vstar=vstar)
    _, vref, sref, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, Cc=1, nwarps=20, nctas=1, sweeps=9, reference=True,
                            **kw)
    x1, _, _, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, Cc=2, nwarps=12, nctas=2, sweeps=5, last=0, **kw)
    assert not np.isnan(x1[:, :M + 1]).any()
    x1[:, M + 1:] = 1e300
    _, v2, s2, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, Cc=2, nwarps=12, nctas=3, sweeps=4, xin=x1, **kw)
    assert np.array_equal(v2, vref)
    assert s2 == sref


# ---- the lane-private marching kernel (csrc/dd_lane.cuh) -------------------------------------------------------------
LANE_CASES = [
    # cb, sweeps, N, M, nctas (a 6th entry: rows per segment of the [segment][strip][row] work order)
    (1, 5, 70, 300, 3),
    (0, 3, 90, 200, 5, 31),
    (1, 2, 77, 131, 4, 20),
    (0, 5, 64, 260, 7, 33),
    (1, 4, 45, 259, 1),
    (1, 3, 90, 200, 4),
    (1, 2, 60, 150, 2),
    (1, 1, 33, 140, 5),
    (0, 5, 80, 260, 3),
    (0, 4, 64, 131, 2),
    (0, 3, 50, 250, 7),
    (0, 2, 70, 70, 2),
    (0, 1, 20, 65, 1),
]


@pytest.mark.parametrize("case", LANE_CASES, ids=[str(c) for c in LANE_CASES])
@pytest.mark.parametrize("nonuniform", [False, True])
def test_lane_equals_global_sor(case, nonuniform):
    cb, sweeps, N, M, nctas = case[:5]
    segrows = case[5] if len(case) > 5 else 1
    lib = _lib()
    rng = np.random.default_rng(2000 * sweeps + N + M)
    arr, met, rho, dt, DT, ldR = _system(rng, N, M, cb, nonuniform)
    vstar = rng.normal(size=(N + 1, M + 1))
    kw = dict(cb=cb, Cc=1, nwarps=segrows, nctas=nctas, sweeps=sweeps, vstar=vstar, zero_boundary=cb)
    xr, vr_, sr, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, reference=True, **kw)
    for order in (0, 1):
        _, vn, st, steps = _run(lib, N, M, arr, met, rho, dt, DT, ldR, order=order, kernel="lane", **kw)
        assert np.array_equal(vn, vr_), (order, np.argwhere(vn != vr_)[:5])
        assert st == sr


@pytest.mark.parametrize("cb,sweeps", [(1, 5), (0, 5), (0, 3), (1, 2)])
def test_lane_on_a_slab(cb, sweeps):
    lib = _lib()
    N, M = 140, 200
    rng = np.random.default_rng(17 + sweeps)
    arr, met, rho, dt, DT, ldR = _system(rng, N, M, cb, True)
    vstar = rng.normal(size=(N + 1, M + 1))
    kw = dict(cb=cb, Cc=1, nwarps=1, sweeps=sweeps, vstar=vstar, zero_boundary=cb)
    _, vglob, _, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, nctas=1, reference=True, **kw)
    G = 2 * sweeps + 1
    for (a, b) in ((0, 51), (51, 97), (97, N + 1)):
        row0 = max(0, a - G - 2)
        row1 = min(N + 1, b + G + 2)
        own = (a - row0, b - row0)
        vr = (0 if row0 == 0 else 1, row1 - row0 if row1 == N + 1 else row1 - row0 - 1)
        _, vn, _, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, nctas=3, row0=row0, nrows=row1 - row0, own=own, vr=vr,
                           kernel="lane", **kw)
        assert np.array_equal(vn[own[0]:own[1]], vglob[a:b])


def test_lane_two_passes_continue_the_iteration():
    lib = _lib()
    N, M, cb = 60, 180, 0
    rng = np.random.default_rng(199)
    arr, met, rho, dt, DT, ldR = _system(rng, N, M, cb, False)
    vstar = rng.normal(size=(N + 1, M + 1))
    kw = dict(cb=cb, vstar=vstar, Cc=1, nwarps=1)
    _, vref, sref, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, nctas=1, sweeps=9, reference=True, **kw)
    x1, _, _, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, nctas=2, sweeps=5, last=0, kernel="lane", **kw)
    assert not np.isnan(x1[:, :M + 1]).any()
    x1[:, M + 1:] = 1e300
    _, v2, s2, _ = _run(lib, N, M, arr, met, rho, dt, DT, ldR, nctas=3, sweeps=4, xin=x1, kernel="lane", **kw)
    assert np.array_equal(v2, vref)
    assert s2 == sref
