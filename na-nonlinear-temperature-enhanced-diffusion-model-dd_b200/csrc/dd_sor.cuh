// dd_sor.cuh -- the arithmetic of one red-black SOR relaxation, shared by every solver kernel (tile, register
// tile, wavefront) and by the test-only host build, with the fused multiply-adds written out so that all of them
// produce bit-identical iterates:
//   general rows      d = gs - x = (bb + aW xw + aS xs + aN xn - x) + aE xe
//   constant band (T) d = gs - x = (bb - x) + dinv (rW xw + cS xs + cN xn + rE xe)
//   relaxation        x <- x + omega d
// d is the cell's residual in the Jacobi-scaled system.  The term of the next row (xe) enters last and the old x is
// folded into the sum: the marching kernel (dd_lane.cuh) relaxes the next row a moment earlier in the same step,
// and only two (general) or three (constant band) dependent operations then follow it.
#pragma once

#include <math.h>

#include "dd_types.h"

DD_HD double dd_sor_d5(double bb, double aW, double aE, double aS, double aN, double xw, double xe, double xs,
                       double xn, double x) {
    return fma(aE, xe, fma(aN, xn, fma(aS, xs, fma(aW, xw, bb))) - x);
}

DD_HD double dd_sor_dT(double bb, double dinv, double rW, double rE, double cS, double cN, double xw, double xe,
                       double xs, double xn, double x) {
    return fma(dinv, fma(rE, xe, fma(cN, xn, fma(cS, xs, rW * xw))), bb - x);
}

DD_HD double dd_sor_relax(double x, double d, double omega) { return fma(omega, d, x); }

// max of two non-negative doubles by bit pattern: NaN (largest pattern) is sticky
DD_HD double dd_nn_max(double a, double b) {
    union { double d; unsigned long long u; } x, y;
    x.d = fabs(a);
    y.d = fabs(b);
    return x.u > y.u ? x.d : y.d;
}
