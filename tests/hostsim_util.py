"""Test-only driver for tests/hostsim (host compilation of the device node programs).

Builds the problem description exactly as the CUDA path receives it: model struct,
forcing mode, 1-D MMS tables produced by the package's own table builder
(`ddcore.Batch.forcing_spec` uses the same helper functions).
"""
import ctypes as C
import os
import subprocess

import numpy as np

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
CSRC = os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200", "csrc")
BUILD = os.path.join(HERE, "hostsim", "_build")
LIB = os.path.join(BUILD, "libhostsim.so")

_dp = C.POINTER(C.c_double)
VARS = ("cp", "T", "cl", "cd", "cs")


class HSProblem(C.Structure):
    _fields_ = [("N", C.c_int), ("M", C.c_int), ("x", _dp), ("y", _dp), ("model", C.c_double * 16),
                ("kind", C.c_int), ("mode", C.c_int), ("nterms", C.c_int), ("X", (_dp * 3) * 5),
                ("Y", (_dp * 3) * 5), ("XQ", _dp * 3), ("YQ", _dp * 3), ("phi_kind", C.c_int * 5),
                ("phi_p", (C.c_double * 4) * 5), ("farr", (_dp * 2) * 5)]


def build():
    os.makedirs(BUILD, exist_ok=True)
    src = os.path.join(HERE, "hostsim", "hostsim.cpp")
    deps = [src] + [os.path.join(CSRC, f) for f in ("dd_types.h", "dd_physics.cuh", "dd_nodeprog.cuh", "dd_combine.cuh")]
    if os.path.exists(LIB) and all(os.path.getmtime(LIB) >= os.path.getmtime(d) for d in deps):
        return LIB
    subprocess.check_call(["g++", "-O2", "-std=c++17", "-shared", "-fPIC", "-Wno-unknown-pragmas", "-I", CSRC,
                           "-o", LIB, src])
    return LIB


_lib = None


def lib():
    global _lib
    if _lib is None:
        _lib = C.CDLL(build())
        for n in ("hs_fields", "hs_feuler", "hs_exact", "hs_pc_step", "hs_combine"):
            getattr(_lib, n).restype = C.c_int
    return _lib


def f64(a):
    return np.ascontiguousarray(a, dtype=np.float64)


def ptr(a):
    return a.ctypes.data_as(_dp)


class Problem:
    """Holds an HSProblem and keeps the arrays it points to alive."""

    def __init__(self, x, y, model, eta, spec=None, t0=0.0, dt=1.0, arrays=None, reaction="regh"):
        import ddcore
        from _ddlib import MODE_ARRAYS, MODE_EXPSIN, MODE_NONE, MODE_SEPARABLE
        self.keep = []
        P = HSProblem()
        self.x, self.y = f64(x), f64(y)
        P.N, P.M = len(self.x) - 1, len(self.y) - 1
        P.x, P.y = ptr(self.x), ptr(self.y)
        names = ("K1", "K2", "K3", "K4", "DT", "Dl_max", "phi_l", "gamma_T", "Kd", "Sd", "Dd_max", "phi_d", "phi_T",
                 "r_sp", "T_ref")
        for k, n in enumerate(names):
            P.model[k] = float(getattr(model, n))
        P.model[15] = float(eta)
        P.kind = int(getattr(model, "dd_kind", getattr(model, "kind", 1))) + 10 * {"regh": 0, "cs": 1, "h": 2}[reaction]
        P.nterms = 1
        px, py = ddcore.quadrature_points(self.x), ddcore.quadrature_points(self.y)
        if arrays is not None:
            P.mode = MODE_ARRAYS
            for v, name in enumerate(VARS):
                for s in range(2):
                    a = arrays.get(name, (None, None))[s]
                    if a is not None:
                        a = f64(a)
                        self.keep.append(a)
                        P.farr[v][s] = ptr(a)
        elif spec is None:
            P.mode = MODE_NONE
        elif isinstance(spec, ddcore.ExpSinSpec):
            P.mode = MODE_EXPSIN
            tabs = [np.sin(np.pi * self.x), np.cos(np.pi * self.x), np.sin(np.pi * self.y), np.cos(np.pi * self.y),
                    np.sin(np.pi * px).reshape(-1), np.sin(np.pi * py).reshape(-1)]
            tabs = [f64(t) for t in tabs]
            self.keep += tabs
            P.X[0][0], P.X[0][1], P.Y[0][0], P.Y[0][1] = ptr(tabs[0]), ptr(tabs[1]), ptr(tabs[2]), ptr(tabs[3])
            P.XQ[0], P.YQ[0] = ptr(tabs[4]), ptr(tabs[5])
        else:
            P.mode = MODE_SEPARABLE
            R = len(spec.X[0])
            P.nterms = R
            for v in range(5):
                for d in range(3):
                    ax = f64(np.stack([ddcore._eval1d(spec.X[v][r][d], self.x) for r in range(R)]))
                    ay = f64(np.stack([ddcore._eval1d(spec.Y[v][r][d], self.y) for r in range(R)]))
                    self.keep += [ax, ay]
                    P.X[v][d], P.Y[v][d] = ptr(ax), ptr(ay)
                P.phi_kind[v] = spec.phi[v].code()
                for k, val in enumerate(spec.phi[v].params(t0, dt)):
                    P.phi_p[v][k] = val
            for q, v in enumerate((0, 1, 2)):
                ax = f64(np.stack([ddcore._eval1d(spec.X[v][r][0], px.reshape(-1)) for r in range(R)]))
                ay = f64(np.stack([ddcore._eval1d(spec.Y[v][r][0], py.reshape(-1)) for r in range(R)]))
                self.keep += [ax, ay]
                P.XQ[q], P.YQ[q] = ptr(ax), ptr(ay)
        self.P = P
        self.shape = (P.N + 1, P.M + 1)

    def _io(self, fields):
        ins = (_dp * 5)()
        arrs = [f64(fields[v]) for v in VARS]
        for k, a in enumerate(arrs):
            ins[k] = ptr(a)
        outs = (_dp * 5)()
        oarr = [np.zeros(self.shape) for _ in VARS]
        for k, a in enumerate(oarr):
            outs[k] = ptr(a)
        return ins, arrs, outs, oarr

    def fields(self, state, t):
        ins, a, outs, o = self._io(state)
        assert lib().hs_fields(C.byref(self.P), ins, outs, C.c_double(t)) == 0
        return dict(zip(VARS, o))

    def feuler(self, state, t0, dt):
        ins, a, outs, o = self._io(state)
        assert lib().hs_feuler(C.byref(self.P), ins, outs, C.c_double(t0), C.c_double(dt)) == 0
        return dict(zip(VARS, o))

    def exact(self, t):
        outs = (_dp * 5)()
        o = [np.zeros(self.shape) for _ in VARS]
        for k, a in enumerate(o):
            outs[k] = ptr(a)
        assert lib().hs_exact(C.byref(self.P), outs, C.c_double(t)) == 0
        return dict(zip(VARS, o))

    def pc_step(self, state, t0, dt, *, num_pc_steps=1, num_newton_steps=1, num_newton_iterations=5,
                consec_xs_rtol=1e-6, swap=1, sweeps=200):
        ins, a, outs, o = self._io(state)
        info = (C.c_double * 6)()
        iters = (C.c_int * max(num_pc_steps, 1))()
        rc = lib().hs_pc_step(C.byref(self.P), ins, outs, C.c_double(t0), C.c_double(dt), num_pc_steps,
                              num_newton_steps, num_newton_iterations, C.c_double(consec_xs_rtol), swap, sweeps,
                              info, iters)
        assert rc == 0
        return dict(zip(VARS, o)), list(info), list(iters)
