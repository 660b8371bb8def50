"""Generated forcing programs (ddprogram: SymPy -> CUDA C -> NVRTC) checked WITHOUT a GPU: the generated text
is also host C, so it is compiled with gcc and compared with the host forcing object built on the lambdified
expressions (the reference's formulas, src/prob1base.py:2313-2378, 3503-3551); NVRTC itself needs no device, so
the sm_100a image is built here as well."""
import ctypes as C
import os
import subprocess

import numpy as np
import pytest
import sympy

import _ddlib
import ddcore
import ddprogram
import prob1base as p1
from test_hostsim import product_model

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
MODEL = dict(K1=1e-3, K2=2e-3, K3=1.5e-3, K4=1e-3, DT=1e-3, Dl_max=8.01e-4, phi_l=1e-5, gamma_T=1e-3, Kd=1e-2,
             Sd=1.0, Dd_max=2.46e-6, phi_d=1e-5, r_sp=5e-2, T_ref=300.0, kind=2)
t, x, y = p1.t_sym, p1.x_sym, p1.y_sym


def nonseparable_exprs():
    """No variable is a product f(t) X(x) Y(y); cs changes sign and has a kink (|x - 1/2|^2.1)."""
    return dict(cp_sym_expr=sympy.exp(-t * x * y) / 2,
                T_sym_expr=1 + sympy.sin(sympy.pi * x * y + t) / 10,
                cl_sym_expr=sympy.cos(x + y * t) / 3,
                cd_sym_expr=sympy.exp(-(x - y) ** 2 - t) / 2,
                cs_sym_expr=(sympy.sin(sympy.pi * (x + y * t)) * sympy.exp(-t) - sympy.Rational(1, 5)
                             + sympy.Abs(x - sympy.Rational(1, 2)) ** sympy.Rational(21, 10)))


def make_case(grid, model):
    return p1.MMSCaseSymbolic(grid=grid, model=model, **nonseparable_exprs())


@pytest.fixture(scope="module")
def host_program(tmp_path_factory):
    d = tmp_path_factory.mktemp("prog")
    grid = p1.Grid(np.linspace(0, 1, 8) ** 1.2, np.linspace(0, 1, 6) ** 0.9)
    model = product_model(MODEL)
    case = make_case(grid, model)
    assert case.device_spec() is None                      # really outside the table form
    ex = case._exprs
    src = ddprogram.generate_source(ex, t, x, y)
    (d / "prog.c").write_text(src)
    so = d / "prog.so"
    subprocess.run(["gcc", "-std=c99", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                    "-o", str(so), str(d / "prog.c"), "-lm"], check=True)
    return grid, model, case, src, C.CDLL(str(so))


def run_host(lib, grid, members, what, tslot, row0=0, nrows=None):
    N, M = grid.N, grid.M
    nrows = N + 1 - row0 if nrows is None else nrows
    ld = M + 1 + 3                                          # a pitch wider than the row
    B = len(members)
    out = [np.full((B, nrows, ld), np.nan) for _ in range(5)]
    xq = np.ascontiguousarray(ddcore.quadrature_points(grid.x).reshape(-1))
    yq = np.ascontiguousarray(ddcore.quadrature_points(grid.y).reshape(-1))
    xs, ys = np.ascontiguousarray(grid.x), np.ascontiguousarray(grid.y)
    mem = (_ddlib.dd_program_member * B)(*members)
    a = _ddlib.dd_program_args()
    dp = lambda arr: arr.ctypes.data_as(C.POINTER(C.c_double))
    a.x, a.y, a.xq, a.yq, a.members = dp(xs), dp(ys), dp(xq), dp(yq), mem
    for v in range(5):
        a.out[v] = dp(out[v])
    a.mstride, a.N, a.M, a.row0, a.nrows, a.ld, a.nmembers, a.what, a.tslot = nrows * ld, N, M, row0, nrows, ld, B, what, tslot
    lib.dd_program_host(C.byref(a))
    assert all(np.all(np.isnan(o[:, :, M + 1:])) for o in out)      # nothing written into the padding
    return [o[:, :, :M + 1] for o in out]


def member(model, eta, reaction, t0, t1, active=1):
    m = _ddlib.dd_program_member()
    m.model = ddcore.model_struct(model, eta, reaction)
    m.t[0], m.t[1], m.active = t0, t1, active
    return m


def test_struct_mirrors_match_the_header(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include <stddef.h>\n#include "dd_b200_program.h"\nint main(void){printf("%zu %zu %zu %zu %zu\\n",'
                   'sizeof(dd_model), sizeof(dd_program_member), sizeof(dd_program_args), offsetof(dd_program_args, mstride),'
                   'offsetof(dd_program_args, what)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    A = _ddlib.dd_program_args
    assert got == [C.sizeof(_ddlib.dd_model), C.sizeof(_ddlib.dd_program_member), C.sizeof(A), A.mstride.offset,
                   A.what.offset]


@pytest.mark.parametrize("reaction,forcing_cls", [("regh", "ForcingTerms_RegHCsTriple"), ("cs", "ForcingTerms_CsTriple"),
                                                  ("h", "ForcingTerms_HCsTriple")])
def test_generated_sources_equal_host_forcing(host_program, reaction, forcing_cls):
    grid, model, case, _, lib = host_program
    eta, t0, t1 = 37.0, 0.3, 0.3125
    kw = dict(regularization_factor=eta) if reaction == "regh" else {}
    forcing = getattr(p1, forcing_cls)(mms_case=case, model=model, **kw)
    got = run_host(lib, grid, [member(model, eta, reaction, t0, t1)], 0, 1)
    want = [getattr(forcing, n)(t1, grid.xx, grid.yy) for n in ("fcp", "fT", "fcl", "fcd", "fcs")]
    for g, w, n in zip(got, want, "cp T cl cd cs".split()):
        assert np.max(np.abs(g[0] - w)) <= 2e-14 * max(1.0, np.max(np.abs(w))), n
    assert np.all(got[0][0][0, :] == 0) and np.all(got[0][0][:, -1] == 0)         # fcp: zero on the boundary
    assert np.max(np.abs(want[4])) > 1e-2 and np.min(case.cs(t1, grid.xx, grid.yy)) < 0 < np.max(case.cs(t1, grid.xx, grid.yy))


def test_generated_exact_members_slabs_and_inactive(host_program):
    grid, model, case, _, lib = host_program
    other = product_model(dict(MODEL, Kd=3e-2, kind=1))
    mem = [member(model, 50.0, "regh", 0.1, 0.2), member(other, 10.0, "regh", 0.5, 0.6), member(model, 50.0, "regh", 0.1, 0.2, 0)]
    ex = run_host(lib, grid, mem, 1, 0)
    for v, n in enumerate("cp T cl cd cs".split()):
        assert np.array_equal(ex[v][0], getattr(case, n)(0.1, grid.xx, grid.yy)) or \
            np.max(np.abs(ex[v][0] - getattr(case, n)(0.1, grid.xx, grid.yy))) <= 4e-16
        assert np.max(np.abs(ex[v][1] - getattr(case, n)(0.5, grid.xx, grid.yy))) <= 4e-16
        assert np.all(np.isnan(ex[v][2]))                                       # inactive member untouched
    # per-member constants and a row slab [2, 5)
    case1 = make_case(grid, other)
    f1 = p1.ForcingTerms_RegHCsTriple(mms_case=case1, model=other, regularization_factor=10.0)
    full = run_host(lib, grid, mem, 0, 1)
    slab = run_host(lib, grid, mem, 0, 1, row0=2, nrows=3)
    for v, n in enumerate(("fcp", "fT", "fcl", "fcd", "fcs")):
        w = getattr(f1, n)(0.6, grid.xx, grid.yy)
        assert np.max(np.abs(full[v][1] - w)) <= 2e-14 * max(1.0, np.max(np.abs(w))), n
        assert np.array_equal(slab[v][1], full[v][1][2:5])


def test_nvrtc_image_for_sm100a(host_program):
    grid, model, case, src, _ = host_program
    prog = case.device_program()
    assert prog is not None and prog.source == src and case.device_program() is prog
    assert prog.image[:4] == b"\x7fELF" and b"dd_program" in prog.image
    out = subprocess.run(["cuobjdump", "-elf", "/dev/stdin"], input=prog.image, capture_output=True)
    if out.returncode == 0:
        assert b"sm_100" in out.stdout


def test_separable_cases_keep_their_tables_and_unprintable_cases_have_no_program():
    import prob1_mms_cases as cases
    grid = p1.make_uniform_grid(4, 4)
    model = product_model(MODEL)
    assert cases.MMSCasePol(grid=grid, model=model).device_spec() is not None
    f = sympy.Function("mystery")
    bad = dict(nonseparable_exprs(), cs_sym_expr=f(x * y + t))
    assert ddprogram.program_for({k[:-9]: v for k, v in bad.items()}, t, x, y) is None


def test_dirac_and_sign_follow_the_reference_lambdify_rules(tmp_path):
    """|x - 1/2| with a grid node exactly on the kink: the Laplacian holds 2 DiracDelta(x - 1/2), which the
    reference's lambdify evaluates as (|arg| < 1e-13 ? 1 : 0) (src/prob1base.py:1226-1247), and the gradient holds
    sign(x - 1/2) = 0 there.  The generated program must give the same sources as the host callables."""
    grid = p1.make_uniform_grid(8, 6)
    model = product_model(MODEL)
    ex = dict(nonseparable_exprs(),
              T_sym_expr=1 + sympy.Abs(x - sympy.Rational(1, 2)) * (1 + y * t) / 5,
              cd_sym_expr=sympy.exp(-t) * sympy.Abs(y - sympy.Rational(1, 2)) * sympy.cos(x * y) / 2)
    case = p1.MMSCaseSymbolic(grid=grid, model=model, **ex)
    assert case.device_spec() is None
    src = ddprogram.generate_source(case._exprs, t, x, y)
    assert "dd_dirac(" in src and "dd_sign(" in src
    (tmp_path / "prog.c").write_text(src)
    subprocess.run(["gcc", "-std=c99", "-O1", "-ffp-contract=off", "-shared", "-fPIC", "-I", os.path.join(ROOT, "include"),
                    "-o", str(tmp_path / "prog.so"), str(tmp_path / "prog.c"), "-lm"], check=True)
    lib = C.CDLL(str(tmp_path / "prog.so"))
    forcing = p1.ForcingTerms_RegHCsTriple(mms_case=case, model=model, regularization_factor=50.0)
    got = run_host(lib, grid, [member(model, 50.0, "regh", 0.2, 0.25)], 0, 0)
    want = [getattr(forcing, n)(0.2, grid.xx, grid.yy) for n in ("fcp", "fT", "fcl", "fcd", "fcs")]
    for g, w, n in zip(got, want, "cp T cl cd cs".split()):
        assert np.max(np.abs(g[0] - w)) <= 2e-14 * max(1.0, np.max(np.abs(w))), n
    # the delta really contributes on the kink line (and only there)
    lapT = case.lap_T(0.2, grid.xx, grid.yy)
    assert np.all(lapT[4, :] != 0) and np.all(lapT[3, :] == 0)
