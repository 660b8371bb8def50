// dd_combine.cuh -- combined max-integral error norms, folded in step by step
// (calculate_combined_error_norm, reference src/mms_trial_utils.py:15-53, and the per-variable figures of
// NumericalErrorSummary, :150-190):
//   sup_k ( sum_v H2_v(t_k) + trapezoid_0^{t_k} sum_w P2_w ),  then the square root,
// from the per-step norms r[8] = H2[cp, T, cl, cd, cs], P2[T, cl, cd].  The arithmetic repeats the
// reference's Python operation by operation -- the builtin sum() with its Neumaier compensation,
// `0.5 * dt * (a + b)`, a maximum that a NaN never replaces (`max(0.0, nan)`) -- with explicitly rounded
// operations on the device (no contraction into FMAs), so the result is bit for bit what Python gives on the
// same norms.  Header-only so that tests/hostsim can compile it for the host.
#pragma once
#include <math.h>

#include "dd_types.h"

#ifdef __CUDA_ARCH__
#define DD_ADD_RN(a, b) __dadd_rn((a), (b))
#define DD_SUB_RN(a, b) __dsub_rn((a), (b))
#define DD_MUL_RN(a, b) __dmul_rn((a), (b))
#else
#define DD_ADD_RN(a, b) ((a) + (b))
#define DD_SUB_RN(a, b) ((a) - (b))
#define DD_MUL_RN(a, b) ((a) * (b))
#endif

// CPython's float sum() (Objects/bltinmodule.c): Neumaier compensation, the correction added once at the end
// when it is finite and non-zero
DD_HD double dd_py_sum(const double* y, int n) {
    double total = 0.0, comp = 0.0;
    for (int k = 0; k < n; ++k) {
        const double t = DD_ADD_RN(total, y[k]);
        if (fabs(total) >= fabs(y[k])) comp = DD_ADD_RN(comp, DD_ADD_RN(DD_SUB_RN(total, t), y[k]));
        else comp = DD_ADD_RN(comp, DD_ADD_RN(DD_SUB_RN(y[k], t), total));
        total = t;
    }
    if (comp != 0.0 && isfinite(comp)) total = DD_ADD_RN(total, comp);
    return total;
}

// One time level folded into the running state st[18] = best[6], run[6], last integrand[6] of
// (overall, cp, T, cl, cd, cs); first != 0 starts a series.
DD_HD void dd_combine_fold(const double* r, double dt, int first, double* st) {
    const double half_dt = DD_MUL_RN(0.5, dt);
    double H[5], P[3];
    for (int v = 0; v < 5; ++v) H[v] = r[v];
    for (int v = 0; v < 3; ++v) P[v] = r[5 + v];
    const double hsq[6] = {dd_py_sum(H, 5), H[0], H[1], H[2], H[3], H[4]};
    const double ig[6] = {dd_py_sum(P, 3), 0.0, P[0], P[1], P[2], 0.0};
    for (int q = 0; q < 6; ++q) {
        double best = first ? 0.0 : st[q], run = first ? 0.0 : st[6 + q];
        if (!first) run = DD_ADD_RN(run, DD_MUL_RN(half_dt, DD_ADD_RN(st[12 + q], ig[q])));
        const double val = DD_ADD_RN(hsq[q], run);
        if (val > best) best = val;
        st[q] = best;
        st[6 + q] = run;
        st[12 + q] = ig[q];
    }
}
