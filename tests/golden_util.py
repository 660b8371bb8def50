"""Helpers shared by the golden-fixture tests (oracle side).

Rebuilds a fixture's scenario for the oracle from the JSON descriptor stored in
the .npz by `oracle/make_golden.py`.
"""
import glob
import json
import os

import numpy as np

GOLDEN_DIR = os.path.join(os.path.dirname(os.path.abspath(__file__)), "golden")
VARS = ("cp", "T", "cl", "cd", "cs")


def fixture_names(kind=None, prefix=None, program=False):
    """Fixture names by kind / prefix.  The fixtures of the non-separable symbolic case ("nonsep": the device
    runs it from a generated forcing program) are listed only with program=True; they have tests of their own."""
    out = []
    for p in sorted(glob.glob(os.path.join(GOLDEN_DIR, "*.npz"))):
        name = os.path.basename(p)[:-4]
        if prefix and not name.startswith(prefix):
            continue
        with np.load(p) as z:
            desc = json.loads(str(z["__desc__"]))
        if kind and desc["kind"] != kind:
            continue
        if (desc.get("case") == "nonsep") != program:
            continue
        out.append(name)
    return out


def load_fixture(name):
    z = np.load(os.path.join(GOLDEN_DIR, name + ".npz"))
    desc = json.loads(str(z["__desc__"]))
    return desc, z


def oracle_model(md):
    from oracle import OModel
    md = dict(md)
    return OModel(**md)


def rel_err(a, b):
    """Norm-wise relative error max|a-b| / max|b|; 0/0 -> 0."""
    a = np.asarray(a, dtype=np.float64)
    b = np.asarray(b, dtype=np.float64)
    den = np.max(np.abs(b))
    num = np.max(np.abs(a - b)) if a.size else 0.0
    if den == 0.0:
        return float(num)
    return float(num / den)
