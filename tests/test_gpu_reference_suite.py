"""The reference's OWN pytest files, unmodified, on top of this package's core (north_star:
"cvg_studies_base and mms_trial_utils run unchanged on top of it").

`tools/stage_reference.sh` copies /root/reference/{src,tests} into the git-ignored baseline/_ref/ (it travels to the
GPU box with the snapshot).  Each reference test file runs in a subprocess whose module search path puts this
package FIRST -- `import prob1base`, `prob1_mms_cases`, `mms_trial_utils`, `cvg_studies_base` resolve to the
CUDA-backed modules here -- and the reference's src/ LAST, so that only what this package does not provide
(`utils_for_testing`, pure host reporting) comes from the reference.  Skipped when baseline/_ref is absent."""
import os
import subprocess
import sys

import pytest

pytestmark = pytest.mark.gpu

HERE = os.path.dirname(os.path.abspath(__file__))
ROOT = os.path.dirname(HERE)
PKG = os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200")
REF = os.path.join(ROOT, "baseline", "_ref")

# (file, number of test items the reference collects)
FILES = [
    ("test_reghcstriple.py", 63),
    ("test_reghcstriple_system.py", 14),
    ("test_statevars.py", 10),
    ("test_newton_residuals.py", 44),
    ("test_semidiscrete_field_hcs_triple.py", 25),
    ("test_forcing_terms_hcs_triple.py", 12),
    ("test_time_integrator_hcs_triple.py", 15),
    ("test_time_integrator_hcs_triple_full_step.py", 5),
    ("test_isolated_correctors_cp_cs.py", 1),
    ("test_feuler_spatial_accuracy.py", 1),
    ("test_time_integration_fwd_euler_full_p1base.py", 24),
    ("test_spatial_isolated_T_accuracy.py", 5),
    ("test_spatial_h1_isolated_T_accuracy.py", 1),
    ("test_mms_trial_utils.py", 9),
]

CHECK = """
import sys, prob1base, mms_trial_utils, cvg_studies_base, prob1_mms_cases
pkg = sys.argv[1]
for m in (prob1base, mms_trial_utils, cvg_studies_base, prob1_mms_cases):
    assert m.__file__.startswith(pkg), m.__file__
import utils_for_testing
assert "_ref" in utils_for_testing.__file__
"""


def _env():
    env = dict(os.environ)
    env["PYTHONPATH"] = os.pathsep.join([PKG, os.path.join(REF, "src")])
    return env


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="baseline/_ref not staged")
def test_module_resolution():
    """The subprocess really imports this package's modules, and the reference's for what is out of scope."""
    subprocess.run([sys.executable, "-c", CHECK, PKG], env=_env(), check=True, cwd=REF)


@pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "tests")), reason="baseline/_ref not staged")
@pytest.mark.parametrize("fname,nitems", FILES, ids=[f[0] for f in FILES])
def test_reference_file_passes_on_this_core(fname, nitems):
    log = os.path.join(ROOT, "gpurun_out", "refsuite_" + fname.replace(".py", ".log"))
    os.makedirs(os.path.dirname(log), exist_ok=True)
    # the reference's own pytest configuration is `pythonpath = ["src"]` only; -c /dev/null keeps it (and this
    # repo's conftest) out, -p no:cacheprovider keeps the staged tree clean
    cmd = [sys.executable, "-m", "pytest", os.path.join("tests", fname), "-q", "-x", "-p", "no:cacheprovider", "-c",
           os.devnull, "--rootdir", REF]
    r = subprocess.run(cmd, env=_env(), cwd=REF, capture_output=True, text=True, timeout=1500)
    with open(log, "w") as f:
        f.write(r.stdout[-20000:] + "\n---- stderr ----\n" + r.stderr[-5000:])
    tail = "\n".join(r.stdout.strip().splitlines()[-15:])
    assert r.returncode == 0, tail
    assert f"{nitems} passed" in r.stdout, tail
