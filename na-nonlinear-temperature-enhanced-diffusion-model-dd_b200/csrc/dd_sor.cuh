// dd_sor.cuh -- the arithmetic of one red-black SOR relaxation, shared by every solver kernel (tile, register
// tile, wavefront) and by the test-only host build, with the fused multiply-adds written out so that all of them
// produce bit-identical iterates:
//   general rows      gs = bb + aW xw + aS xs + aN xn + aE xe
//   constant band (T) gs = bb + dinv (rW xw + cS xs + cN xn + rE xe)
//   relaxation        x <- x + omega (gs - x)
// The term of the next row (xe) enters last: the marching kernel (dd_lane.cuh) relaxes that row a moment earlier
// in the same step, and everything that does not depend on it is then off the critical path.
#pragma once

#include <math.h>

#include "dd_types.h"

DD_HD double dd_sor_gs5(double bb, double aW, double aE, double aS, double aN, double xw, double xe, double xs,
                        double xn) {
    return fma(aE, xe, fma(aN, xn, fma(aS, xs, fma(aW, xw, bb))));
}

DD_HD double dd_sor_gsT(double bb, double dinv, double rW, double rE, double cS, double cN, double xw, double xe,
                        double xs, double xn) {
    return fma(dinv, fma(rE, xe, fma(cN, xn, fma(cS, xs, rW * xw))), bb);
}

DD_HD double dd_sor_relax(double x, double gs, double omega) { return fma(omega, gs - x, x); }

// max of two non-negative doubles by bit pattern: NaN (largest pattern) is sticky
DD_HD double dd_nn_max(double a, double b) {
    union { double d; unsigned long long u; } x, y;
    x.d = fabs(a);
    y.d = fabs(b);
    return x.u > y.u ? x.d : y.d;
}
