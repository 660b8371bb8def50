"""SymPy -> CUDA forcing programs (DD_MODE_PROGRAM, include/dd_b200_program.h).

The reference evaluates an arbitrary manufactured solution on the host: `MMSCaseSymbolic` lambdifies the five
expressions with their t / x / y derivatives (src/prob1base.py:1226-1280, 1283-1487) and the forcing object
composes the sources from those callables at every step (src/prob1base.py:2313-2378, 3503-3551; the cp source
is a 3x3 Gauss cell average, :493-598).  For cases that are not of the separable form the library evaluates
from tables, this module prints the same expressions as C (common subexpressions shared), wraps them in the
fixed composition code below, compiles the result with NVRTC for sm_100a and returns the image that
`dd_forcing_program` loads.  There is no host fallback inside: a case whose expressions cannot be printed
gives `None` and the caller keeps the array path (sources evaluated by the caller's own callables).

The generated text is also valid host C (the kernel wrapper is guarded by __CUDACC__): the CPU tests compile
it with gcc and compare it with the lambdified sources, so the generator is checked without a GPU.
"""

from __future__ import annotations

import hashlib
import os
import threading
from dataclasses import dataclass
from pathlib import Path
from typing import Dict, Optional

import sympy
from sympy.printing.c import C99CodePrinter

VARS = ("cp", "T", "cl", "cd", "cs")
ARCH = "sm_100a"
_HEADER = Path(__file__).resolve().parent.parent / "include" / "dd_b200_program.h"


class _Printer(C99CodePrinter):
    """C99 printer with the reference's lambdify rules: DiracDelta(a) -> |a| < 1e-13 (src/prob1base.py:1226-1247)."""

    def _print_DiracDelta(self, e):
        return "dd_dirac(%s)" % self._print(e.args[0])

    def _print_Heaviside(self, e):
        return "dd_step(%s)" % self._print(e.args[0])

    def _print_sign(self, e):
        return "dd_sign(%s)" % self._print(e.args[0])


_PRELUDE = r"""
#ifdef __CUDACC__
#define DD_FN __device__ __forceinline__
#else
#include <math.h>
#define DD_FN static inline
#endif
#ifndef M_PI
#define M_PI 3.14159265358979323846
#endif
#ifndef M_E
#define M_E 2.71828182845904523536
#endif

DD_FN double dd_dirac(double a) { return fabs(a) < 1e-13 ? 1.0 : 0.0; }
DD_FN double dd_step(double a) { return a > 0.0 ? 1.0 : (a < 0.0 ? 0.0 : 0.5); }
DD_FN double dd_sign(double a) { return (double)((a > 0.0) - (a < 0.0)); }

/* F2(cs) of the cs/cd interaction (src/prob1base.py:2842-2876, 3303-3340, 3452-3466) */
DD_FN double dd_F2(const dd_model* m, double cs) {
    if (m->reaction == 0) return 1.0 / (1.0 + exp(-m->eta * cs));
    return m->reaction == 1 ? cs : (cs > 0.0 ? 1.0 : 0.0);
}
"""

_COMPOSE = r"""
/* the five sources at one node: the PDE residual of the exact solution
 * (src/prob1base.py:2313-2378 fcp, fT, fcl; 3503-3551 fcd, fcs) */
DD_FN void dd_program_node(const dd_program_args* a, int m, int r, int j) {
    const dd_program_member* mb = &a->members[m];
    if (!mb->active) return;
    const dd_model* md = &mb->model;
    const int i = a->row0 + r;
    const long long o = (long long)m * a->mstride + (long long)r * a->ld + j;
    const double t = mb->t[a->tslot];
    const double x = a->x[i], y = a->y[j];
    if (a->what == 1) {
        double u[5];
        dd_user_exact(t, x, y, u);
        for (int v = 0; v < 5; ++v) a->out[v][o] = u[v];
        return;
    }
    double u[5], ut[5], ux[5], uy[5], lap[5];
    dd_user_point(t, x, y, u, ut, ux, uy, lap);
    const double cp = u[0], T = u[1], cl = u[2], cd = u[3], cs = u[4];
    /* T */
    a->out[1][o] = ut[1] - (md->DT * lap[1] - md->K3 * cp * T);
    /* cl: div(Dl(cp) grad cl) - d/dx(gamma_T T (cl + 1)) - K4 cp (cl + 1) */
    const double Dl = md->Dl_max * exp(-md->phi_l * cp);
    const double dDl = -md->phi_l * Dl;
    a->out[2][o] = ut[2] - (dDl * (ux[0] * ux[2] + uy[0] * uy[2]) + Dl * lap[2] - (md->gamma_T * T) * ux[2] -
                            (cl + 1.0) * (md->gamma_T * ux[1]) - md->K4 * cp * (cl + 1.0));
    /* cd: div(Dd(cp, T) grad cd) + Kd (Sd - cd)(1 + cl) F2(cs) */
    const double Te = T + md->T_ref;
    double Dd = 0.0, dDd_dT = 0.0;
    if (Te != 0.0) {
        Dd = md->Dd_max * exp(-md->phi_d * cp) * exp(-md->phi_T / Te);
        dDd_dT = Dd * (md->phi_T / (Te * Te));
    }
    const double dDd_dcp = -md->phi_d * Dd;
    const double F2 = dd_F2(md, cs);
    a->out[3][o] = ut[3] - ((dDd_dcp * ux[0] + dDd_dT * ux[1]) * ux[3] + (dDd_dcp * uy[0] + dDd_dT * uy[1]) * uy[3] +
                            Dd * lap[3] + md->Kd * (md->Sd - cd) * (cl + 1.0) * F2);
    /* cs */
    a->out[4][o] = ut[4] + md->Kd * (1.0 + cl) * (md->Sd - cd) * F2;
    /* cp: cell average of dt cp + cp (K1 (1 + cl) + K2 T) over the dual cell, interior nodes only */
    double fcp = 0.0;
    if (i > 0 && i < a->N && j > 0 && j < a->M) {
        const double w[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
        double acc = 0.0;
        for (int p = 0; p < 3; ++p)
            for (int q = 0; q < 3; ++q) {
                double g[4];
                dd_user_quad(t, a->xq[3 * i + p], a->yq[3 * j + q], g);
                acc += w[p] * w[q] * (g[1] + g[0] * (md->K1 * (1.0 + g[2]) + md->K2 * g[3]));
            }
        fcp = 0.25 * acc;
    }
    a->out[0][o] = fcp;
}

#ifdef __CUDACC__
extern "C" __global__ void __launch_bounds__(128) dd_program(dd_program_args a) {
    const int j = blockIdx.x * 128 + threadIdx.x;
    if (j > a.M) return;
    dd_program_node(&a, blockIdx.z, blockIdx.y, j);
}
#else
void dd_program_host(const dd_program_args* a) {
    for (int m = 0; m < a->nmembers; ++m)
        for (int r = 0; r < a->nrows; ++r)
            for (int j = 0; j <= a->M; ++j) dd_program_node(a, m, r, j);
}
#endif
"""


def _function(name: str, signature: str, outputs, printer: _Printer) -> str:
    """One C function assigning `outputs` = [(lvalue, expr)], common subexpressions hoisted."""
    repl, reduced = sympy.cse([e for _, e in outputs], symbols=sympy.numbered_symbols("w_"), optimizations="basic")
    lines = [f"DD_FN void {name}({signature}) {{"]
    for sym, e in repl:
        lines.append(f"    const double {printer.doprint(sym)} = {printer.doprint(e)};")
    for (lhs, _), e in zip(outputs, reduced):
        lines.append(f"    {lhs} = {printer.doprint(e)};")
    lines.append("}")
    return "\n".join(lines)


def generate_source(exprs: Dict[str, sympy.Expr], t_var, x_var, y_var) -> str:
    """CUDA / C text of the forcing program of the manufactured solution `exprs` (one expression per variable)."""
    t, x, y = sympy.symbols("t x y", real=True)
    ren = {t_var: t, x_var: x, y_var: y}
    ex = {v: sympy.sympify(exprs[v]).xreplace(ren) for v in VARS}
    free = set().union(*[e.free_symbols for e in ex.values()]) - {t, x, y}
    if free:
        raise ValueError(f"free symbols other than t, x, y: {sorted(map(str, free))}")
    pr = _Printer()
    point = []
    for k, v in enumerate(VARS):
        e = ex[v]
        dx, dy = sympy.diff(e, x), sympy.diff(e, y)
        point += [(f"u[{k}]", e), (f"ut[{k}]", sympy.diff(e, t)), (f"ux[{k}]", dx), (f"uy[{k}]", dy),
                  (f"lap[{k}]", sympy.diff(dx, x) + sympy.diff(dy, y))]
    quad = [("g[0]", ex["cp"]), ("g[1]", sympy.diff(ex["cp"], t)), ("g[2]", ex["cl"]), ("g[3]", ex["T"])]
    exact = [(f"u[{k}]", ex[v]) for k, v in enumerate(VARS)]
    args = "double t, double x, double y, "
    body = [
        _function("dd_user_point", args + "double* u, double* ut, double* ux, double* uy, double* lap", point, pr),
        _function("dd_user_quad", args + "double* g", quad, pr),
        _function("dd_user_exact", args + "double* u", exact, pr),
    ]
    text = "\n\n".join(body)
    if "Derivative" in text or "Subs(" in text or "// Not supported" in text:
        raise ValueError("expression with an unevaluated derivative / unsupported function")
    return '#include "dd_b200_program.h"\n' + _PRELUDE + "\n" + text + "\n" + _COMPOSE


def header_text() -> str:
    return _HEADER.read_text()


def _nvrtc():
    from cuda.bindings import nvrtc
    return nvrtc


def compile_program(source: str, arch: str = ARCH) -> bytes:
    """NVRTC: `source` -> cubin for `arch`.  Contraction into FMAs is off so that the sources round like the
    host evaluation of the same expressions (they run once per time level: not a hot spot)."""
    nvrtc = _nvrtc()

    def ok(res, what):
        err = res[0]
        if int(err) != 0:
            raise RuntimeError(f"NVRTC {what}: {nvrtc.nvrtcGetErrorString(err)[1].decode()}")
        return res[1:] if len(res) > 2 else (res[1] if len(res) == 2 else None)

    prog = ok(nvrtc.nvrtcCreateProgram(source.encode(), b"dd_program.cu", 1, [header_text().encode()],
                                       [b"dd_b200_program.h"]), "create")
    opts = [f"--gpu-architecture={arch}".encode(), b"--fmad=false", b"--std=c++17", b"-lineinfo"]
    res = nvrtc.nvrtcCompileProgram(prog, len(opts), opts)
    if int(res[0]) != 0:
        n = ok(nvrtc.nvrtcGetProgramLogSize(prog), "log size")
        log = b" " * n
        nvrtc.nvrtcGetProgramLog(prog, log)
        nvrtc.nvrtcDestroyProgram(prog)
        raise RuntimeError("NVRTC compile failed:\n" + log.decode(errors="replace"))
    n = ok(nvrtc.nvrtcGetCUBINSize(prog), "cubin size")
    image = b" " * n
    ok(nvrtc.nvrtcGetCUBIN(prog, image), "cubin")
    nvrtc.nvrtcDestroyProgram(prog)
    return bytes(image)


@dataclass(frozen=True)
class ProgramSpec:
    """A compiled forcing program: what `Batch.forcing_program` uploads."""
    source: str
    image: bytes
    key: str


_cache: Dict[str, ProgramSpec] = {}
_lock = threading.Lock()


def program_for(exprs: Dict[str, sympy.Expr], t_var, x_var, y_var) -> Optional[ProgramSpec]:
    """ProgramSpec of a manufactured solution, or None when its expressions cannot be printed as C.  Compiled
    once per distinct source text and process."""
    if os.environ.get("DD_NO_PROGRAM", "") == "1":
        return None
    try:
        src = generate_source(exprs, t_var, x_var, y_var)
    except Exception:
        return None
    key = hashlib.sha256(src.encode()).hexdigest()
    with _lock:
        spec = _cache.get(key)
        if spec is None:
            spec = _cache[key] = ProgramSpec(source=src, image=compile_program(src), key=key)
    return spec
