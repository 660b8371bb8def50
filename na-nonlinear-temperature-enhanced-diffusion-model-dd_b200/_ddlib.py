"""ctypes binding of libdd_b200.so (the C ABI declared in include/dd_b200.h).

There is no CPU fallback: importing this module is cheap, but the first call
that needs the library raises `DDLibraryError` if the shared object is missing
or the machine has no CUDA device.
"""

from __future__ import annotations

import ctypes as C
import os

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.environ.get("DD_LIB", os.path.join(_HERE, "libdd_b200.so"))  # DD_LIB: development override

DD_OK, DD_ERR_INVALID, DD_ERR_CUDA, DD_ERR_NOT_CONVERGED, DD_ERR_NO_DEVICE, DD_ERR_DOMAIN = 0, -1, -2, -3, -4, -5
VAR_INDEX = {"cp": 0, "T": 1, "cl": 2, "cd": 3, "cs": 4}
VARS = ("cp", "T", "cl", "cd", "cs")
MODE_NONE, MODE_ARRAYS, MODE_SEPARABLE, MODE_EXPSIN, MODE_PROGRAM = 0, 1, 2, 3, 4
PHI_INV1PT, PHI_EXP, PHI_LINEAR, PHI_OSC, PHI_CONST, PHI_HOST = 0, 1, 2, 3, 4, 5


class DDLibraryError(RuntimeError):
    pass


class DDNotConverged(RuntimeError):
    pass


class dd_model(C.Structure):
    _fields_ = [(n, C.c_double) for n in
                ("K1", "K2", "K3", "K4", "DT", "Dl_max", "phi_l", "gamma_T", "Kd", "Sd", "Dd_max", "phi_d",
                 "phi_T", "r_sp", "T_ref", "eta")] + [("kind", C.c_int), ("reaction", C.c_int)]


class dd_program_member(C.Structure):
    """include/dd_b200_program.h"""
    _fields_ = [("model", dd_model), ("t", C.c_double * 2), ("active", C.c_int), ("_pad", C.c_int)]


class dd_program_args(C.Structure):
    """include/dd_b200_program.h"""
    _fields_ = [("x", C.POINTER(C.c_double)), ("y", C.POINTER(C.c_double)), ("xq", C.POINTER(C.c_double)),
                ("yq", C.POINTER(C.c_double)), ("members", C.POINTER(dd_program_member)),
                ("out", C.POINTER(C.c_double) * 5), ("mstride", C.c_longlong), ("N", C.c_int), ("M", C.c_int),
                ("row0", C.c_int), ("nrows", C.c_int), ("ld", C.c_int), ("nmembers", C.c_int), ("what", C.c_int),
                ("tslot", C.c_int)]


class dd_pc_options(C.Structure):
    _fields_ = [("num_pc_steps", C.c_int), ("num_newton_steps", C.c_int), ("num_newton_iterations", C.c_int),
                ("cd_band_swap", C.c_int), ("consec_xs_rtol", C.c_double), ("solve_tol", C.c_double),
                ("max_sweeps", C.c_int), ("fixed_sweeps", C.c_int),
                ("extrapolate_guess", C.c_int), ("_pad", C.c_int)]


class dd_step_stats(C.Structure):
    _fields_ = [("sweeps", C.c_int * 3), ("passes", C.c_int * 3), ("retries", C.c_int),
                ("cs_newton_iters", C.c_int), ("rho", C.c_double * 3), ("resid", C.c_double * 3),
                ("bound", C.c_double * 3)]

    def as_dict(self):
        return dict(sweeps=list(self.sweeps), passes=list(self.passes), retries=self.retries,
                    cs_newton_iters=self.cs_newton_iters, rho=list(self.rho), resid=list(self.resid),
                    bound=list(self.bound))


_P = C.POINTER
_dp = _P(C.c_double)
_vp = C.c_void_p

# every symbol include/dd_b200.h declares: name -> (restype, argtypes)
SIGNATURES = {
    "dd_ctx_create": (C.c_int, [C.c_int, _vp, _P(_vp)]),
    "dd_ctx_destroy": (C.c_int, [_vp]),
    "dd_ctx_synchronize": (C.c_int, [_vp]),
    "dd_last_error": (C.c_char_p, [_vp]),
    "dd_version": (C.c_char_p, []),
    "dd_batch_create": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, C.c_int, C.c_int, C.c_int,
                                  C.c_int, _P(_vp)]),
    "dd_batch_destroy": (C.c_int, [_vp]),
    "dd_batch_set_models": (C.c_int, [_vp, C.c_int, C.c_int, _P(dd_model)]),
    "dd_batch_set_active": (C.c_int, [_vp, C.c_int, C.c_int, _P(C.c_int)]),
    "dd_forcing_none": (C.c_int, [_vp]),
    "dd_forcing_separable": (C.c_int, [_vp, C.c_int, _P(_dp * 3 * 5), _P(_dp * 3 * 5), _P(_dp * 3), _P(_dp * 3),
                                       _P(C.c_int * 5), _P(C.c_double * 4 * 5)]),
    "dd_forcing_set_phi": (C.c_int, [_vp, C.c_int, C.c_int, _P(C.c_int), _dp]),
    "dd_forcing_expsin": (C.c_int, [_vp, _dp, _dp, _dp, _dp, _dp, _dp]),
    "dd_forcing_arrays": (C.c_int, [_vp, C.c_int, _P(_dp * 2 * 5)]),
    "dd_forcing_program": (C.c_int, [_vp, C.c_char_p, C.c_ulonglong, _dp, _dp]),
    "dd_state_upload": (C.c_int, [_vp, C.c_int, C.c_int, _P(_dp * 5)]),
    "dd_state_download": (C.c_int, [_vp, C.c_int, C.c_int, _P(_dp * 5)]),
    "dd_state_fill_exact": (C.c_int, [_vp, C.c_int, _dp, C.c_int]),
    "dd_state_dev_ptr": (C.c_int, [_vp, C.c_int, C.c_int, _P(_vp), _P(C.c_longlong), _P(C.c_int)]),
    "dd_work_dev_ptr": (C.c_int, [_vp, C.c_char_p, _P(_vp)]),
    "dd_work_upload": (C.c_int, [_vp, C.c_char_p, C.c_int, _dp]),
    "dd_work_download": (C.c_int, [_vp, C.c_char_p, C.c_int, _dp]),
    "dd_step_feuler": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int]),
    "dd_step_pc": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, _P(dd_pc_options), _P(dd_step_stats)]),
    "dd_step_pc_deferred": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, _P(dd_pc_options), _P(dd_step_stats),
                                      _P(C.c_int)]),
    "dd_step_pc_flush": (C.c_int, [_vp, _P(dd_step_stats), _P(C.c_int)]),
    "dd_run_pc": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _P(dd_pc_options), _dp,
                            _P(dd_step_stats)]),
    "dd_run_feuler": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp]),
    "dd_run_pc_errors": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _P(dd_pc_options), _dp,
                                   _P(dd_step_stats)]),
    "dd_run_feuler_errors": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, C.c_int, _dp]),
    "dd_pc_options_default": (None, [_P(dd_pc_options)]),
    "dd_eval_fields": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int]),
    "dd_pc_predict": (C.c_int, [_vp, C.c_int, _dp, _dp, C.c_int]),
    "dd_pc_newton": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, _P(dd_pc_options),
                               _P(dd_step_stats)]),
    "dd_pc_correct": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, _P(dd_pc_options), _P(C.c_int)]),
    "dd_pc_residual": (C.c_int, [_vp, C.c_int, C.c_int, _dp, _dp, C.c_int, _dp]),
    "dd_error_norms": (C.c_int, [_vp, C.c_int, C.c_int, _dp, C.c_int, _dp]),
    "dd_step_pc_phase": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _dp, _dp, C.c_int, _P(dd_pc_options), _dp,
                                   _P(C.c_int)]),
    "dd_batch_set_plan": (C.c_int, [_vp, _P(C.c_int * 3)]),
    "dd_batch_get_plan": (C.c_int, [_vp, _P(C.c_int * 3)]),
    "dd_batch_set_relax_rho": (C.c_int, [_vp, _P(C.c_double * 3)]),
    "dd_sweeps_for_rho": (C.c_int, [C.c_double, C.c_int]),
    "dd_next_plan": (C.c_int, [C.c_int, C.c_double, C.c_double, C.c_int]),
    "dd_pc_solve_segment": (C.c_int, [_vp, C.c_int, C.c_int, C.c_int, _P(dd_pc_options), C.c_int, C.c_int, C.c_int]),
    "dd_probe_math": (C.c_int, [_vp, C.c_int, _dp, _dp, _dp]),
    "dd_ipc_export": (C.c_int, [_vp, _vp, C.c_char_p]),
    "dd_ipc_import": (C.c_int, [_vp, C.c_char_p, C.POINTER(_vp)]),
    "dd_ipc_close": (C.c_int, [_vp, _vp]),
    "dd_halo_flags_create": (C.c_int, [_vp, C.POINTER(_vp)]),
    "dd_halo_flags_destroy": (C.c_int, [_vp, _vp]),
    "dd_halo_push": (C.c_int, [_vp, _vp, _vp, _vp, _vp, C.c_longlong, _vp, _vp, _vp, C.c_uint]),
    "dd_halo_status": (C.c_int, [_vp, _vp, C.POINTER(C.c_int)]),
    "dd_probe_fp64": (C.c_int, [_vp, C.c_double, _dp]),
    "dd_solver_kernel_name": (C.c_char_p, [C.c_int]),
    "dd_launch_count": (C.c_longlong, []),
    "dd_profile_enable": (C.c_int, [C.c_int]),
    "dd_profile_read": (C.c_int, [_P(C.c_char_p), _dp, _P(C.c_longlong), C.c_int]),
}


def profile_read(reset: bool = True):
    """{kernel class: (device ms, launches)} accumulated since the last reset (profiling must be enabled)."""
    lib = load_library()
    names = (C.c_char_p * 16)()
    ms = (C.c_double * 16)()
    cnt = (C.c_longlong * 16)()
    n = lib.dd_profile_read(names, ms, cnt, 1 if reset else 0)
    return {names[k].decode(): (ms[k], cnt[k]) for k in range(n) if cnt[k] > 0}

_lib = None


def load_library(path: str = LIB_PATH):
    """dlopen the library and attach signatures.  Needs no GPU (used by the CPU tests
    that check the exported symbols)."""
    global _lib
    if _lib is not None:
        return _lib
    if not os.path.exists(path):
        raise DDLibraryError(
            f"{path} not found: build it with `python __graft_entry__.py` (or ./build_lib.sh). "
            "There is no CPU fallback for the stepping path.")
    lib = C.CDLL(path)
    for name, (res, args) in SIGNATURES.items():
        fn = getattr(lib, name)  # AttributeError here = header/library mismatch
        fn.restype = res
        fn.argtypes = args
    _lib = lib
    return lib


def dptr(a: np.ndarray):
    return a.ctypes.data_as(_dp)


def as_f64(a) -> np.ndarray:
    return np.ascontiguousarray(a, dtype=np.float64)


class Context:
    """dd_ctx wrapper: one per (device, stream)."""

    _default = None

    def __init__(self, device: int = 0, stream: int | None = None):
        self.lib = load_library()
        h = _vp()
        rc = self.lib.dd_ctx_create(device, _vp(stream) if stream else None, C.byref(h))
        if rc == DD_ERR_NO_DEVICE:
            raise DDLibraryError("no CUDA device visible: the stepping path has no CPU fallback")
        if rc != DD_OK:
            raise DDLibraryError(f"dd_ctx_create failed with status {rc}")
        self.handle = h
        self.device = device
        self.stream = int(stream) if stream else None  # None: the library's own non-blocking stream

    @classmethod
    def default(cls) -> "Context":
        if cls._default is None:
            dev = int(os.environ.get("LOCAL_RANK", os.environ.get("DD_DEVICE", "0")))
            cls._default = Context(dev)
        return cls._default

    def check(self, rc: int, what: str = ""):
        if rc == DD_OK:
            return
        msg = self.lib.dd_last_error(self.handle)
        msg = msg.decode() if msg else ""
        if rc == DD_ERR_INVALID:
            # the reference signals bad arguments (dt <= 0, shape mismatch) with AssertionError
            raise AssertionError(f"{what}: {msg or 'invalid argument'}")
        if rc == DD_ERR_DOMAIN:
            raise ValueError(msg)  # the reference's HCsTriple corrector raises ValueError (src/prob1base.py:3410)
        if rc == DD_ERR_NOT_CONVERGED:
            raise DDNotConverged(f"{what}: {msg}")
        raise DDLibraryError(f"{what}: status {rc}: {msg}")

    def synchronize(self):
        self.check(self.lib.dd_ctx_synchronize(self.handle), "synchronize")

    def close(self):
        if getattr(self, "handle", None):
            self.lib.dd_ctx_destroy(self.handle)
            self.handle = None

    def __del__(self):
        try:
            self.close()
        except Exception:
            pass
