"""CPU oracle for the prob1base.py time-stepping hot path.

THIS PACKAGE IS TEST INFRASTRUCTURE, NOT PRODUCT CODE.

It is a from-scratch NumPy/SciPy restatement of the reference algorithm
(`/root/reference/src/prob1base.py`, cited per function as file:line) used
only as the *checker*:

  * `tests/` compare the CUDA path against it,
  * `__graft_entry__.smoke()` checks one small step against it,
  * `bench.py` times it as the `cpu_baseline` leg / the `--impl reference` arm.

Nothing under `na-nonlinear-temperature-enhanced-diffusion-model-dd_b200/`
imports it; the product path fails loudly when the CUDA library is missing.

Parity status: PINNED.  `oracle/make_golden.py` imports the *live* reference
from `/root/reference/src` (only possible in the build container) and writes
per-step fields, `last_residual`, cs-Newton iteration counts and convergence
study error norms to `tests/golden/*.npz`; `tests/test_oracle_golden.py` checks
this oracle against those fixtures (<= 1e-13 relative, counts identical) and,
when `/root/reference` is present, against the live reference as well.
The sparse direct solve of the reference is SciPy's bundled SuperLU
(`scipy.sparse.linalg.spsolve`, reference pin scipy=1.15.2, environment.yml:204;
call sites src/prob1base.py:2103,2130); the oracle calls the same routine.
"""

from .ddoracle import (  # noqa: F401
    OGrid,
    OModel,
    OState,
    uniform_grid,
    NOTEBOOK_CONSTS,
    heaviside_reg,
    fields_F,
    feuler_step,
    PCStepper,
    norm_H_sq,
    grad_norm_p_sq,
    run_trial,
    combined_error_norm,
    exact_state,
    error_norms,
)
from .mms import make_case, OForcing, CASE_NAMES  # noqa: F401
