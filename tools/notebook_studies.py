#!/usr/bin/env python
"""The convergence studies of the reference's six notebooks, re-run on the device (GPU required).

For every notebook: the spatial study (N = 2 .. 256, dt = h^1.5), the temporal study and the regularisation-factor
study, with the notebook's constants (cell 3), case class (cell 5), final times and step sizes (cells 9 - 13), as
refinement sweeps of `ddensemble.RefinementSweep` (whole trials on the device: time loop, per-step error norms,
max-integral combination).  Prints and stores, side by side, the overall errors and the 3-point observed rates the
notebooks PUBLISH in their cell outputs and the reproduced ones.

    python tools/notebook_studies.py [--out gpurun_out/notebook_studies.json]

tests/test_gpu_studies.py::test_notebook_tables_of_all_six_notebooks asserts on the same comparison.
"""
import argparse
import json
import math
import os
import sys
import time

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
PKG = os.path.join(ROOT, "na-nonlinear-temperature-enhanced-diffusion-model-dd_b200")
for p in (ROOT, PKG):
    if p not in sys.path:
        sys.path.insert(0, p)

CONSTS_A = dict(K1=1e-3, K2=1e-3, K3=1e-3, K4=1e-3, DT=1e-3, Dl_max=1e-5, phi_l=1e-5, gamma_T=1e-9, Kd=1e-2, Sd=1.0,
                Dd_max=1e-6, phi_d=1e-5, r_sp=5e-2, T_ref=300.0)
CONSTS_B = dict(CONSTS_A, Dl_max=8.01e-4, Dd_max=2.46e-6)
NS = [2, 4, 8, 16, 32, 64, 128, 256]
ETAS = [10.0, 50.0, 100.0, 200.0, 300.0, 500.0, 1000.0]

# notebook -> (case class, constants, spatial Tf, published spatial errors, published final spatial rate,
#              temporal (N, Tf, dts), published temporal errors, published final temporal rate (None = NaN),
#              published eta-study errors)       [cell outputs of the notebooks in the reference's root directory]
NOTEBOOKS = {
    "MMSCaseExpSin": dict(
        cls="MMSCaseExpSin", consts=CONSTS_A, Tf=0.01,
        spatial=[1.942652829989e-05, 5.197056624911e-06, 1.322695968641e-06, 3.372248813359e-07, 8.344194130557e-08,
                 2.052209700229e-08, 5.119616858484e-09, 1.278782670173e-09], spatial_rate=2.004,
        temporal=dict(N=32, Tf=0.01, dts=[1e-2 / 2 ** k for k in range(6)]),
        temporal_err=[1.036215100290e-07, 8.344194130557e-08, 8.193792525959e-08, 8.181573405295e-08,
                      8.180115032463e-08, 8.179850160373e-08], temporal_rate=2.461,
        eta_err=[8.179982876369e-08, 8.179982920798e-08, 8.179982937825e-08, 8.179982942569e-08, 8.179982943273e-08,
                 8.179982943698e-08, 8.179982943940e-08]),
    "MMSCasePol": dict(
        cls="MMSCasePol", consts=CONSTS_B, Tf=0.01,
        spatial=[4.93452e-05, 1.59616e-05, 4.28269e-06, 1.08800e-06, 2.75006e-07, 6.96085e-08, 1.74802e-08,
                 4.38284e-09], spatial_rate=1.993,
        temporal=dict(N=256, Tf=0.01, dts=[1e-2 / 2 ** k for k in range(4)]),
        temporal_err=[3.60101e-08, 8.49854e-09, 4.01980e-09, 4.18199e-09], temporal_rate=None,
        eta_err=[2.78759e-07] * 7),
    "MMSCaseSlowlyChangingPeaks_Fast1e1": dict(
        cls="MMSCaseSlowlyChangingPeaks_Fast1e1", consts=CONSTS_B, Tf=1.0,
        spatial=[3.410697138975e-01, 2.998593199634e-01, 4.558178972447e-02, 6.673442252443e-03, 1.083722320571e-03,
                 2.251962441053e-04, 5.355729294822e-05, 1.329324479086e-05], spatial_rate=2.092,
        temporal=dict(N=200, Tf=10.0, dts=[1.0 / 2 ** k for k in range(9)]),
        temporal_err=[0.0, 0.0, 7.108884464820e-01, 9.864776561636e-01, 3.221709152999e-01, 8.453668420359e-02,
                      2.135342688180e-02, 5.351596147453e-03, 1.338721547608e-03], temporal_rate=1.996,
        eta_err=[7.498440503481e-05, 7.498440503313e-05, 7.498440503312e-05, 7.498440503312e-05, 7.498440503312e-05,
                 7.498440503312e-05, 7.498440503312e-05]),
    "MMSCaseNonFullySmoothPol_cpcsH1_TclcdH2": dict(
        cls="MMSCaseNonFullySmoothPol_cpcsH1_TclcdH2", consts=CONSTS_A, Tf=1.0,
        spatial=[1.387299517318e-05, 8.822763874973e-05, 3.383480896506e-05, 1.517524996184e-05, 3.747930839694e-06,
                 6.816794044645e-07, 2.799670822833e-07, 8.645407062247e-08], spatial_rate=1.054,
        temporal=dict(N=256, Tf=0.01, dts=[1e-2 / 2 ** k for k in range(4)]),
        temporal_err=[1.713006210334e-09, 1.353723332525e-09, 1.365070781516e-09, 1.374905111409e-09],
        temporal_rate=None,
        eta_err=[7.360771456680e-08, 7.360771456678e-08, 7.360771456676e-08, 7.360771456673e-08, 7.360771456671e-08,
                 7.360771456671e-08, 7.360771456676e-08]),
    "MMSCaseNonFullySmoothPol_cpcsH2_TclcdH2": dict(
        cls="MMSCaseNonFullySmoothPol_cpcsH2_TclcdH2", consts=CONSTS_A, Tf=1.0,
        spatial=[1.877869516145e-05, 2.037364736137e-05, 8.851905299491e-06, 3.085671522449e-06, 8.603180933041e-07,
                 2.660430202155e-07, 8.453466133015e-08, 7.641023132398e-08], spatial_rate=4.482,
        temporal=dict(N=256, Tf=0.01, dts=[1e-2 / 2 ** k for k in range(4)]),
        temporal_err=[1.336453753218e-09, 1.169004286281e-09, 1.132086126593e-09, 1.123260655389e-09],
        temporal_rate=2.065,
        eta_err=[1.335768487783e-08, 1.335768487784e-08, 1.335768487785e-08, 1.335768487788e-08, 1.335768487790e-08,
                 1.335768487794e-08, 1.335768487802e-08]),
    "MMSCaseNonFullySmoothPol_cpcsH2_TclcdH3": dict(
        cls="MMSCaseNonFullySmoothPol_cpcsH2_TclcdH3", consts=CONSTS_A, Tf=1.0,
        spatial=[1.706334182719e-05, 1.734312300666e-05, 8.519357549781e-06, 2.640813480048e-06, 7.168895437498e-07,
                 1.844956739082e-07, 4.710500390333e-08, 1.180466192215e-08], spatial_rate=1.961,
        temporal=dict(N=256, Tf=0.01, dts=[1e-2 / 2 ** k for k in range(4)]),
        temporal_err=[2.293853773997e-10, 1.984988392477e-10, 2.134913177290e-10, 2.184209586590e-10],
        temporal_rate=None,
        eta_err=[1.324294360382e-08, 1.324294360382e-08, 1.324294360383e-08, 1.324294360383e-08, 1.324294360384e-08,
                 1.324294360385e-08, 1.324294360388e-08]),
}


def rates_3point(errors):
    """The rate formula of the notebooks (reference src/utils_for_testing.py:98-140): log2 of the ratio of consecutive
    error differences; NaN when a difference is not positive."""
    out = []
    for k in range(len(errors) - 2):
        num, den = errors[k] - errors[k + 1], errors[k + 1] - errors[k + 2]
        out.append(math.log2(num / den) if (num > 1e-16 and den > 1e-16) else float("nan"))
    return out


def run_notebook(name, which=("spatial", "temporal", "eta")):
    import ddensemble
    import prob1_mms_cases as p1mc
    import prob1base as p1
    nb = NOTEBOOKS[name]
    model = p1.DefaultModel02(p1.ModelConsts(R0=p1.R0, Ea=p1.Ea, phi_T=p1.Ea / p1.R0, **nb["consts"]))
    cls = getattr(p1mc, nb["cls"])
    res = {}
    studies = {
        "spatial": [dict(N=n, dt=(1.0 / n) ** 1.5, Tf=nb["Tf"], eta=50.0) for n in NS],
        "temporal": [dict(N=nb["temporal"]["N"], dt=d, Tf=nb["temporal"]["Tf"], eta=50.0) for d in nb["temporal"]["dts"]],
        "eta": [dict(N=32, dt=5e-4, Tf=0.01, eta=e) for e in ETAS],
    }
    for key in which:
        t0 = time.perf_counter()
        sw = ddensemble.RefinementSweep(cls, model, studies[key])
        got = [float(v) for v in sw.run_for_errors()["overall"]]
        sw.close()
        published = nb[{"spatial": "spatial", "temporal": "temporal_err", "eta": "eta_err"}[key]]
        res[key] = dict(published=published, reproduced=got, seconds=time.perf_counter() - t0,
                        rel_diff=[abs(g - p) / p if p else abs(g) for g, p in zip(got, published)])
        if key != "eta":
            r = rates_3point(got)
            res[key]["rates"] = r
            res[key]["final_rate"] = r[-1]
            res[key]["published_final_rate"] = nb["spatial_rate" if key == "spatial" else "temporal_rate"]
    return res


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--out", default=os.path.join(ROOT, "gpurun_out", "notebook_studies.json"))
    ap.add_argument("--only", default=None)
    args = ap.parse_args()
    out = {}
    for name in NOTEBOOKS:
        if args.only and args.only not in name:
            continue
        out[name] = run_notebook(name)
        for key, r in out[name].items():
            print(f"== {name} / {key}  ({r['seconds']:.1f} s)")
            for p, g, d in zip(r["published"], r["reproduced"], r["rel_diff"]):
                print(f"   published {p:.12e}   reproduced {g:.12e}   rel. diff {d:.1e}")
            if "final_rate" in r:
                pub = r["published_final_rate"]
                print(f"   final observed rate: published {'NaN' if pub is None else '%.3f' % pub}, "
                      f"reproduced {r['final_rate']:.3f}")
    os.makedirs(os.path.dirname(args.out), exist_ok=True)
    with open(args.out, "w") as f:
        json.dump(out, f, indent=1)
    print("written", args.out)


if __name__ == "__main__":
    main()
