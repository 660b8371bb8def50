"""pytest configuration: markers and import paths.

`-m "not gpu"` runs on a CPU-only box; `-m gpu` tests need a B200 and call the
CUDA path through the C ABI.
"""
import os
import sys

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)


def pytest_configure(config):
    config.addinivalue_line("markers", "gpu: needs a CUDA device (B200); run with -m gpu")
