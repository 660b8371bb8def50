"""CPU checks of the drop-in boundary: libdd_b200.so loads without a GPU and exports every function that
include/dd_b200.h declares, the ctypes layer binds exactly that set, the structure mirrors have the sizes the C
compiler gives them, and the product path fails loudly (no CPU fallback) when there is no device.  No compute
call is made here."""
import ctypes as C
import os
import re
import subprocess

import pytest

import _ddlib

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
HEADER = os.path.join(ROOT, "include", "dd_b200.h")


def declared_functions():
    text = open(HEADER).read()
    text = re.sub(r"/\*.*?\*/", " ", text, flags=re.S)
    names = re.findall(r"^\s*(?:const\s+char\s*\*|int|void|long\s+long)\s+(dd_\w+)\s*\(", text, flags=re.M)
    assert len(names) == len(set(names)) and len(names) >= 40
    return names


def test_library_loads_and_exports_every_declared_symbol():
    lib = C.CDLL(_ddlib.LIB_PATH)          # plain dlopen: no GPU, no torch
    missing = [n for n in declared_functions() if not hasattr(lib, n)]
    assert not missing, missing


def test_ctypes_layer_binds_exactly_the_declared_functions():
    declared = set(declared_functions())
    bound = set(_ddlib.SIGNATURES)
    assert declared - bound == set(), sorted(declared - bound)
    assert bound - declared == set(), sorted(bound - declared)
    lib = _ddlib.load_library()
    assert lib.dd_version().decode().startswith("dd_b200")


def test_structure_mirrors_have_the_c_sizes(tmp_path):
    src = tmp_path / "sz.c"
    src.write_text('#include <stdio.h>\n#include "dd_b200.h"\nint main(void){printf("%zu %zu %zu\\n", sizeof(dd_model), '
                   'sizeof(dd_pc_options), sizeof(dd_step_stats)); return 0;}\n')
    exe = tmp_path / "sz"
    subprocess.run(["gcc", "-I", os.path.join(ROOT, "include"), "-o", str(exe), str(src)], check=True)
    got = [int(v) for v in subprocess.run([str(exe)], check=True, capture_output=True, text=True).stdout.split()]
    assert got == [C.sizeof(_ddlib.dd_model), C.sizeof(_ddlib.dd_pc_options), C.sizeof(_ddlib.dd_step_stats)]


def test_no_device_is_an_error_not_a_fallback():
    import torch
    if torch.cuda.is_available():
        pytest.skip("a CUDA device is present")
    with pytest.raises(_ddlib.DDLibraryError, match="no CUDA device"):
        _ddlib.Context(0)
    import prob1base as p1
    from test_hostsim import product_model
    from test_program_codegen import MODEL
    grid = p1.make_uniform_grid(4, 4)
    model = product_model(MODEL)
    field = p1.SemiDiscreteField_RegHCsTriple(grid=grid, model=model, forcing_terms=p1.NoForcingTerms(grid),
                                              regularization_factor=50.0)
    integ = p1.ForwardEulerIntegrator(field)
    state = p1.StateVars(**{v: grid.make_full0() for v in ("cp", "T", "cl", "cd", "cs")}, model=model, hh=grid.hh,
                         kk=grid.kk)
    with pytest.raises(_ddlib.DDLibraryError):
        integ.step(state, t0=0.0, dt=1e-3)
