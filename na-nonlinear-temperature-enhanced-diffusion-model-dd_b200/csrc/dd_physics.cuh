// dd_physics.cuh -- node-level arithmetic of the RegHCsTriple scheme.
//
// Everything here is a pure inline function of values handed in by the caller,
// so the CUDA kernels (dd_kernels.cu, dd_solver.cu) and the test-only host
// build (tests/hostsim) share one definition of the numerics.  Formulas follow
// the closed forms of SURVEY.md Appendix A; each block cites the reference
// lines (relative to /root/reference/) whose results it must reproduce.
#pragma once

#include <math.h>

#include "dd_types.h"

#define DD_PI 3.14159265358979323846

// ---------------------------------------------------------------------------
// exp / reciprocal with a short instruction sequence on the device.  The stencil kernels are
// instruction-issue bound (profiles/): the library exp() carries full-range special-case code, so the
// common range |x| < 700 is handled inline and everything else (overflow, underflow, NaN) still goes to exp().
//   x = (64 m + j) ln2/64 + r,  |r| <= ln2/128:   exp(x) = 2^m * T[j] * (1 + r + r^2 P(r))
// Cody-Waite reduction with a 34-bit ln2/64 (k * L_hi is exact for |k| < 2^17), T[j] = 2^(j/64) from a 64-entry
// table (512 B, L1-resident), degree-4 P (truncation 1e-19): 11 fp64 operations and 8 immediates, against 17 and
// 15 of a table-free degree-13 polynomial on |r| <= ln2/2 -- every 64-bit immediate costs two uniform moves, which
// were 17 % of the marching predictor's instructions.  Relative error < 0.9 eps (checked over 2e7 arguments against
// long double on the host, and on the device against numpy by tests/test_gpu_parity.py).  1/x uses the
// IEEE-rounded reciprocal instruction sequence.  Host builds (tests/hostsim) use libm.
// ---------------------------------------------------------------------------
#ifdef __CUDACC__
static __device__ const double dd_exp_tab[64] = {
    1.00000000000000000e+00, 1.01088928605170048e+00, 1.02189714865411663e+00, 1.03302487902122841e+00,
    1.04427378242741375e+00, 1.05564517836055716e+00, 1.06714040067682370e+00, 1.07876079775711986e+00,
    1.09050773266525769e+00, 1.10238258330784089e+00, 1.11438674259589243e+00, 1.12652161860824185e+00,
    1.13878863475669156e+00, 1.15118922995298267e+00, 1.16372485877757748e+00, 1.17639699165028122e+00,
    1.18920711500272103e+00, 1.20215673145270308e+00, 1.21524735998046896e+00, 1.22848053610687002e+00,
    1.24185781207348400e+00, 1.25538075702469110e+00, 1.26905095719173322e+00, 1.28287001607877826e+00,
    1.29683955465100964e+00, 1.31096121152476441e+00, 1.32523664315974132e+00, 1.33966752405330292e+00,
    1.35425554693689265e+00, 1.36900242297459052e+00, 1.38390988196383202e+00, 1.39897967253831124e+00,
    1.41421356237309515e+00, 1.42961333839197002e+00, 1.44518080697704665e+00, 1.46091779418064704e+00,
    1.47682614593949935e+00, 1.49290772829126484e+00, 1.50916442759342284e+00, 1.52559815074453842e+00,
    1.54221082540794074e+00, 1.55900440023783693e+00, 1.57598084510788650e+00, 1.59314215134226700e+00,
    1.61049033194925428e+00, 1.62802742185734783e+00, 1.64575547815396495e+00, 1.66367658032673638e+00,
    1.68179283050742900e+00, 1.70010635371852348e+00, 1.71861929812247793e+00, 1.73733383527370622e+00,
    1.75625216037329945e+00, 1.77537649252652119e+00, 1.79470907500310717e+00, 1.81425217550039886e+00,
    1.83400808640934243e+00, 1.85397912508338547e+00, 1.87416763411029996e+00, 1.89457598158696561e+00,
    1.91520656139714740e+00, 1.93606179349229435e+00, 1.95714412417540018e+00, 1.97845602638795093e+00};
#endif

// Branch-free variants (DD_BRANCHFREE_EXP / DD_BRANCHFREE_RCP = 1; off): a kernel evaluates six or more independent
// exponentials and reciprocals per node, and the slow-path branch inside each of them (the range check here,
// __drcp_rn's special-case call) ends a basic block, so their dependency chains cannot be interleaved.  Removing
// the branches -- clamping plus scaling by two exact powers of two for exp, hardware seed + two Newton steps + a
// select for 1/x; both pass the accuracy test -- was measured on B200 and does NOT pay: the extra instructions cost
// more than the interleaving gains (exp: predictor 0.307 -> 0.322 ms, corrector 0.257 -> 0.281; reciprocal:
// 0.304 / 0.262, sources 0.120 -> 0.128).
#ifndef DD_BRANCHFREE_EXP
#define DD_BRANCHFREE_EXP 0
#endif
#ifndef DD_BRANCHFREE_RCP
#define DD_BRANCHFREE_RCP 0
#endif

DD_HD double dd_exp(double x) {
#ifdef __CUDA_ARCH__
#if DD_BRANCHFREE_EXP
    const double xc = fmin(fmax(x, -746.0), 710.0);  // (a NaN becomes -746 here and is restored below)
    const double t = fma(xc, 9.23324826168936567683e+01, 6755399441055744.0);  // round(x * 64 / ln2) in the low word
#else
    if (!(fabs(x) < 700.0)) return exp(x);
    const double t = fma(x, 9.23324826168936567683e+01, 6755399441055744.0);  // round(x * 64 / ln2) in the low word
    const double xc = x;
#endif
    const int k = __double2loint(t);
    const double kf = t - 6755399441055744.0;
    double r = fma(-kf, 1.08304246959960437380e-02, xc);  // ln2/64, high 34 bits
    r = fma(-kf, 2.53101721666508769488e-13, r);          // ... and the rest
    double p = 1.0 / 720.0;
    p = fma(p, r, 1.0 / 120.0);
    p = fma(p, r, 1.0 / 24.0);
    p = fma(p, r, 1.0 / 6.0);
    p = fma(p, r, 0.5);
    const double q = fma(r * r, p, r);                   // exp(r) - 1
    const double T = __ldg(&dd_exp_tab[k & 63]);
    const double e = fma(T, q, T);                       // in [0.99, 2.01)
#if DD_BRANCHFREE_EXP
    // e * 2^(k >> 6) as two multiplications by exact powers of two, each a normal number (k >> 6 in [-1077, 1025])
    const int kq = k >> 6, k1 = kq >> 1, k2 = kq - k1;
    const double s1 = __hiloint2double((k1 + 1023) << 20, 0), s2 = __hiloint2double((k2 + 1023) << 20, 0);
    const double res = (e * s1) * s2;
    return x != x ? x : res;
#else
    return __hiloint2double(__double2hiint(e) + ((k >> 6) << 20), __double2loint(e));
#endif
#else
    return exp(x);
#endif
}

DD_HD double dd_rcp(double x) {
#ifdef __CUDA_ARCH__
#if DD_BRANCHFREE_RCP
    // hardware seed (about 20 bits), two Newton steps (error below one unit in the last place); the seed of 0, of
    // an infinity and of a NaN is already the result
    double y0;
    asm("rcp.approx.ftz.f64 %0, %1;" : "=d"(y0) : "d"(x));
    const double e0 = fma(-x, y0, 1.0);
    const double y1 = fma(y0, e0, y0);
    const double e1 = fma(-x, y1, 1.0);
    const double y2 = fma(y1, e1, y1);
    const double ay = fabs(y0);
    return (ay > 0.0 && ay < __longlong_as_double(0x7ff0000000000000LL)) ? y2 : y0;
#else
    return __drcp_rn(x);
#endif
#else
    return 1.0 / x;
#endif
}

// ---------------------------------------------------------------------------
// model coefficient functions (reference src/prob1base.py:96-217, 3452-3466)
// ---------------------------------------------------------------------------

DD_HD double dd_Dl(const DDModel& m, double cp) { return m.Dl_max * dd_exp(-m.phi_l * cp); }

// Dd(cp, T) = Dd_max e^{-phi_d cp} e^{-phi_T/(T+T_shift)}, 0 where T+T_shift == 0
DD_HD double dd_Dd(const DDModel& m, double cp, double T) {
    const double Te = T + m.T_shift;
    if (Te == 0.0) return 0.0;
    // one exponential: e^{-phi_d cp} e^{-phi_T/Te} = e^{-(phi_d cp + phi_T/Te)} (differs from the product by rounding only)
    return m.Dd_max * dd_exp(-(m.phi_d * cp + m.phi_T * dd_rcp(Te)));
}

// returns Dd and writes dDd/dT = Dd * phi_T / Te^2
DD_HD double dd_Dd_dT(const DDModel& m, double cp, double T, double* dT) {
    const double Te = T + m.T_shift;
    if (Te == 0.0) {
        *dT = 0.0;
        return 0.0;
    }
    const double iT = dd_rcp(Te);
    const double d = m.Dd_max * dd_exp(-(m.phi_d * cp + m.phi_T * iT));
    *dT = d * (m.phi_T * iT * iT);
    return d;
}

DD_HD double dd_H(double x, double eta) { return dd_rcp(1.0 + dd_exp(-eta * x)); }

// F2(cs) of the cs/cd interaction (cscd_reaction_cs / Kd of the reference's three field classes)
DD_HD double dd_F2(const DDModel& m, double cs) {
    if (m.react == DD_REACT_REGH) return dd_H(cs, m.eta);  // the common case first
    return (m.react == DD_REACT_CS) ? cs : (cs > 0.0 ? 1.0 : 0.0);
}

// ---------------------------------------------------------------------------
// stencil helpers
// ---------------------------------------------------------------------------

struct DDSten {
    double c, w, e, s, n;  // (i,j), (i-1,j), (i+1,j), (i,j-1), (i,j+1)
};

DD_HD DDSten dd_load_sten(const double* u, long long o, int ld) {
    DDSten r;
    r.c = u[o];
    r.w = u[o - ld];
    r.e = u[o + ld];
    r.s = u[o - 1];
    r.n = u[o + 1];
    return r;
}

// Geometry factors of an interior node (reference src/prob1base.py:287-304,
// 1517-1550):  D*x(a D-x u)_i = rhp_i [ a_{i+1} (u_{i+1}-u_i) rhE - a_i (u_i-u_{i-1}) rhW ].
struct DDNodeGeo {
    double rhp, rkp, rhW, rhE, rkS, rkN;
    double cW, cE, cS, cN;  // rhp*rhW, rhp*rhE, rkp*rkS, rkp*rkN
};

DD_HD DDNodeGeo dd_node_geo(const DDGeom& g, int i, int j) {
    DDNodeGeo q;
    q.rhp = g.rhp[i];
    q.rkp = g.rkp[j];
    q.rhW = g.rh[i];
    q.rhE = g.rh[i + 1];
    q.rkS = g.rk[j];
    q.rkN = g.rk[j + 1];
    q.cW = q.rhp * q.rhW;
    q.cE = q.rhp * q.rhE;
    q.cS = q.rkp * q.rkS;
    q.cN = q.rkp * q.rkN;
    return q;
}

struct DDFaces {
    double w, e, s, n;  // coefficient on faces (i-1/2,j), (i+1/2,j), (i,j-1/2), (i,j+1/2)
};

// Dl at face averages of cp (reference StateVars.Dl_Mxcp / Dl_Mycp, src/prob1base.py:1941-1942)
DD_HD DDFaces dd_faces_Dl(const DDModel& m, const DDSten& cp) {
    DDFaces f;
    f.w = dd_Dl(m, 0.5 * (cp.c + cp.w));
    f.e = dd_Dl(m, 0.5 * (cp.e + cp.c));
    f.s = dd_Dl(m, 0.5 * (cp.c + cp.s));
    f.n = dd_Dl(m, 0.5 * (cp.n + cp.c));
    return f;
}

// Dd at face averages of (cp, T) (src/prob1base.py:1951-1952)
DD_HD DDFaces dd_faces_Dd(const DDModel& m, const DDSten& cp, const DDSten& T) {
    DDFaces f;
    f.w = dd_Dd(m, 0.5 * (cp.c + cp.w), 0.5 * (T.c + T.w));
    f.e = dd_Dd(m, 0.5 * (cp.e + cp.c), 0.5 * (T.e + T.c));
    f.s = dd_Dd(m, 0.5 * (cp.c + cp.s), 0.5 * (T.c + T.s));
    f.n = dd_Dd(m, 0.5 * (cp.n + cp.c), 0.5 * (T.n + T.c));
    return f;
}

DD_HD DDFaces dd_faces_Dd_dT(const DDModel& m, const DDSten& cp, const DDSten& T, DDFaces* dT) {
    DDFaces f;
    f.w = dd_Dd_dT(m, 0.5 * (cp.c + cp.w), 0.5 * (T.c + T.w), &dT->w);
    f.e = dd_Dd_dT(m, 0.5 * (cp.e + cp.c), 0.5 * (T.e + T.c), &dT->e);
    f.s = dd_Dd_dT(m, 0.5 * (cp.c + cp.s), 0.5 * (T.c + T.s), &dT->s);
    f.n = dd_Dd_dT(m, 0.5 * (cp.n + cp.c), 0.5 * (T.n + T.c), &dT->n);
    return f;
}

// div(a grad u) with face coefficients a
DD_HD double dd_div_flux(const DDNodeGeo& q, const DDFaces& a, const DDSten& u) {
    return q.rhp * (a.e * ((u.e - u.c) * q.rhE) - a.w * ((u.c - u.w) * q.rhW)) +
           q.rkp * (a.n * ((u.n - u.c) * q.rkN) - a.s * ((u.c - u.s) * q.rkS));
}

// ---------------------------------------------------------------------------
// semidiscrete field without the MMS source (interior nodes)
// reference src/prob1base.py:2599-2672 (Fcp..Fcs), 2489-2509 + 3580-3593 (reaction)
// ---------------------------------------------------------------------------

DD_HD double dd_reaction(const DDModel& m, double cl, double cd, double cs) {
    return (m.Sd - cd) * (cl + 1.0) * (m.Kd * dd_F2(m, cs));
}

DD_HD double dd_Fcp_int(const DDModel& m, double cp, double T, double cl) {
    return -m.K1 * (cl + 1.0) * cp - m.K2 * T * cp;
}

DD_HD double dd_FT_int(const DDModel& m, const DDNodeGeo& q, const DDSten& T, double cp) {
    const double lap = q.rhp * ((T.e - T.c) * q.rhE - (T.c - T.w) * q.rhW) +
                       q.rkp * ((T.n - T.c) * q.rkN - (T.c - T.s) * q.rkS);
    return m.DT * lap - m.K3 * cp * T.c;
}

DD_HD double dd_Fcl_int(const DDModel& m, const DDNodeGeo& q, const DDFaces& Dl, const DDSten& T,
                        const DDSten& cl, double cp) {
    // advective flux Mx(V1(T)(cl+1)), V1 = gamma_T T, V2 == 0
    const double ac = m.gamma_T * T.c * (cl.c + 1.0);
    const double aw = m.gamma_T * T.w * (cl.w + 1.0);
    const double ae = m.gamma_T * T.e * (cl.e + 1.0);
    const double adv = q.rhp * (0.5 * (ae + ac) - 0.5 * (ac + aw));
    return dd_div_flux(q, Dl, cl) - adv - m.K4 * cp * (cl.c + 1.0);
}

DD_HD double dd_Fcd_int(const DDModel& m, const DDNodeGeo& q, const DDFaces& Dd, const DDSten& cd,
                        double cl, double cs) {
    return dd_div_flux(q, Dd, cd) + dd_reaction(m, cl, cd.c, cs);
}

// ---------------------------------------------------------------------------
// manufactured solutions: exact fields and derivatives at a point
// ---------------------------------------------------------------------------

struct DDExact {
    double u[DD_NVAR], ut[DD_NVAR], ux[DD_NVAR], uy[DD_NVAR], lap[DD_NVAR];
};

DD_HD void dd_phi_eval(int kind, const double* p, double t, int slot, double* phi, double* dphi) {
    switch (kind) {
        case DD_PHI_HOST:
            *phi = p[2 * slot];
            *dphi = p[2 * slot + 1];
            break;
        case DD_PHI_INV1PT: {
            const double r = 1.0 / (1.0 + t);
            *phi = p[0] * r;
            *dphi = -p[0] * r * r;
        } break;
        case DD_PHI_EXP: {
            const double e = p[0] * exp(-p[1] * t);
            *phi = e;
            *dphi = -p[1] * e;
        } break;
        case DD_PHI_LINEAR:
            *phi = p[0] - p[1] * t;
            *dphi = -p[1];
            break;
        case DD_PHI_OSC:
            *phi = p[0] * (1.0 + p[1] * sin(p[2] * t));
            *dphi = p[0] * p[1] * p[2] * cos(p[2] * t);
            break;
        default:
            *phi = p[0];
            *dphi = 0.0;
    }
}

// Time scalars of one member at time t.  EXPSIN layout:
//  c[0]=tau c[1]=tau' c[2]=e^{-t} | cp = W exp(A + B W): c[3]=A c[4]=B c[5]=A' c[6]=B'
//  cs = r_sp W exp(As + Bs W + Cs W^2): c[7..9]=As,Bs,Cs  c[10..12]=As',Bs',Cs'
// (closed forms of the integrals in reference src/prob1_mms_cases.py:319-325)
DD_HD void dd_time_coefs(int mode, const DDMember& mb, double t, int slot, DDTimeCoef* tc) {
    for (int k = 0; k < 16; ++k) tc->c[k] = 0.0;
    if (mode == DD_FORCING_SEPARABLE) {
        for (int v = 0; v < DD_NVAR; ++v) dd_phi_eval(mb.phi_kind[v], mb.phi_p[v], t, slot, &tc->c[v], &tc->c[5 + v]);
    } else if (mode == DD_FORCING_EXPSIN) {
        const DDModel& m = mb.m;
        const double cc = 2.0 * DD_PI * DD_PI * m.DT;
        const double tau = exp(-cc * t), em = exp(-t), em2 = exp(-2.0 * t);
        tc->c[0] = tau;
        tc->c[1] = -cc * tau;
        tc->c[2] = em;
        tc->c[3] = -m.K1 * t;
        tc->c[4] = m.K1 * (1.0 - em) - m.K2 * (1.0 - tau) / cc;
        tc->c[5] = -m.K1;
        tc->c[6] = m.K1 * em - m.K2 * tau;
        tc->c[7] = -m.Kd * m.Sd * t;
        tc->c[8] = m.Kd * (m.Sd + 1.0) * (1.0 - em);
        tc->c[9] = -0.5 * m.Kd * (1.0 - em2);
        tc->c[10] = -m.Kd * m.Sd;
        tc->c[11] = m.Kd * (m.Sd + 1.0) * em;
        tc->c[12] = -m.Kd * em2;
    }
}

// time-independent part of a separable manufactured solution at node (i, j)
struct DDSpatial {
    double S[DD_NVAR], Sx[DD_NVAR], Sy[DD_NVAR], Sl[DD_NVAR];  // sum_r X Y, X' Y, X Y', X'' Y + X Y''
    double Q1, Q2, Q3;                                          // cell-average sums of fcp (see DDTables)
};

DD_HD void dd_spatial_separable(const DDTables& tb, int i, int j, bool want_q, DDSpatial* sp) {
    if (tb.nprof == 1) {
        // one spatial profile shared by all five variables (MMSCasePol, SlowlyChangingPeaks): no selection logic
        double sXY = 0.0, sX1Y = 0.0, sXY1 = 0.0, sLap = 0.0;
        for (int r = 0; r < tb.nterms; ++r) {
            const int oi = r * tb.nx + i, oj = r * tb.ny + j;
            const double X0 = tb.X[0][0][oi], X1 = tb.X[0][1][oi], X2 = tb.X[0][2][oi];
            const double Y0 = tb.Y[0][0][oj], Y1 = tb.Y[0][1][oj], Y2 = tb.Y[0][2][oj];
            sXY += X0 * Y0;
            sX1Y += X1 * Y0;
            sXY1 += X0 * Y1;
            sLap += X2 * Y0 + X0 * Y2;
        }
#pragma unroll
        for (int v = 0; v < DD_NVAR; ++v) {
            sp->S[v] = sXY; sp->Sx[v] = sX1Y; sp->Sy[v] = sXY1; sp->Sl[v] = sLap;
        }
    } else {
    double pS[DD_NVAR], pSx[DD_NVAR], pSy[DD_NVAR], pSl[DD_NVAR];
#pragma unroll
    for (int p = 0; p < DD_NVAR; ++p) {
        double sXY = 0.0, sX1Y = 0.0, sXY1 = 0.0, sLap = 0.0;
        if (p < tb.nprof) {
            for (int r = 0; r < tb.nterms; ++r) {
                const int oi = r * tb.nx + i, oj = r * tb.ny + j;
                const double X0 = tb.X[p][0][oi], X1 = tb.X[p][1][oi], X2 = tb.X[p][2][oi];
                const double Y0 = tb.Y[p][0][oj], Y1 = tb.Y[p][1][oj], Y2 = tb.Y[p][2][oj];
                sXY += X0 * Y0;
                sX1Y += X1 * Y0;
                sXY1 += X0 * Y1;
                sLap += X2 * Y0 + X0 * Y2;
            }
        }
        pS[p] = sXY; pSx[p] = sX1Y; pSy[p] = sXY1; pSl[p] = sLap;
    }
#pragma unroll
    for (int v = 0; v < DD_NVAR; ++v) {
        const int q = tb.var_prof[v];
        double a = pS[0], bx = pSx[0], by = pSy[0], l = pSl[0];
#pragma unroll
        for (int p = 1; p < DD_NVAR; ++p)
            if (q == p) { a = pS[p]; bx = pSx[p]; by = pSy[p]; l = pSl[p]; }
        sp->S[v] = a; sp->Sx[v] = bx; sp->Sy[v] = by; sp->Sl[v] = l;
    }
    }
    sp->Q1 = sp->Q2 = sp->Q3 = 0.0;
    if (want_q) {
        const int R = tb.nterms;
        double q1 = 0.0, q2 = 0.0, q3 = 0.0;
        for (int r = 0; r < R; ++r) q1 += tb.QX1[r * tb.nx + i] * tb.QY1[r * tb.ny + j];
        for (int rs = 0; rs < R * R; ++rs) {
            q2 += tb.QX2[rs * tb.nx + i] * tb.QY2[rs * tb.ny + j];
            q3 += tb.QX3[rs * tb.nx + i] * tb.QY3[rs * tb.ny + j];
        }
        sp->Q1 = q1; sp->Q2 = q2; sp->Q3 = q3;
    }
}

DD_HD void dd_exact_separable(const DDSpatial& sp, const DDTimeCoef& tc, DDExact* e) {
#pragma unroll
    for (int v = 0; v < DD_NVAR; ++v) {
        const double phi = tc.c[v], dphi = tc.c[5 + v];
        e->u[v] = phi * sp.S[v];
        e->ut[v] = dphi * sp.S[v];
        e->ux[v] = phi * sp.Sx[v];
        e->uy[v] = phi * sp.Sy[v];
        e->lap[v] = phi * sp.Sl[v];
    }
}

DD_HD void dd_exact_expsin(const DDModel& m, const DDTables& tb, const DDTimeCoef& tc, int i, int j, DDExact* e) {
    const double sx = tb.X[0][0][i], cx = tb.X[0][1][i];
    const double sy = tb.Y[0][0][j], cy = tb.Y[0][1][j];
    const double W = sx * sy, Wx = DD_PI * cx * sy, Wy = DD_PI * sx * cy;
    const double lapW = -2.0 * DD_PI * DD_PI * W;
    const double tau = tc.c[0], dtau = tc.c[1], em = tc.c[2];
    e->u[DD_T] = tau * W;   e->ut[DD_T] = dtau * W;  e->ux[DD_T] = tau * Wx;  e->uy[DD_T] = tau * Wy;
    e->lap[DD_T] = tau * lapW;
    e->u[DD_CL] = -em * W;  e->ut[DD_CL] = em * W;   e->ux[DD_CL] = -em * Wx; e->uy[DD_CL] = -em * Wy;
    e->lap[DD_CL] = -em * lapW;
    e->u[DD_CD] = em * W;   e->ut[DD_CD] = -em * W;  e->ux[DD_CD] = em * Wx;  e->uy[DD_CD] = em * Wy;
    e->lap[DD_CD] = em * lapW;
    const double A = tc.c[3], B = tc.c[4], dA = tc.c[5], dB = tc.c[6];
    const double E = exp(A + B * W);
    const double g1 = E * (1.0 + B * W);
    e->u[DD_CP] = W * E;
    e->ut[DD_CP] = (W * E) * (dA + dB * W);
    e->ux[DD_CP] = g1 * Wx;
    e->uy[DD_CP] = g1 * Wy;
    e->lap[DD_CP] = 0.0;  // not used by any forcing term
    const double As = tc.c[7], Bs = tc.c[8], Cs = tc.c[9];
    const double Es = exp(As + W * (Bs + Cs * W));
    const double cs = m.r_sp * W * Es;
    e->u[DD_CS] = cs;
    e->ut[DD_CS] = cs * (tc.c[10] + W * (tc.c[11] + tc.c[12] * W));
    e->ux[DD_CS] = 0.0; e->uy[DD_CS] = 0.0; e->lap[DD_CS] = 0.0;  // not used
}

// MMS sources at a node from the exact point data
// (reference src/prob1base.py:2331-2378 fT, fcl; 3503-3551 fcd, fcs).
DD_HD double dd_src_fT(const DDModel& m, const DDExact& e) {
    return e.ut[DD_T] - (m.DT * e.lap[DD_T] - m.K3 * e.u[DD_CP] * e.u[DD_T]);
}

DD_HD double dd_src_fcl(const DDModel& m, const DDExact& e) {
    const double Dl = dd_Dl(m, e.u[DD_CP]);
    const double dDl = -m.phi_l * Dl;
    const double cl1 = e.u[DD_CL] + 1.0;
    return e.ut[DD_CL] - (dDl * (e.ux[DD_CP] * e.ux[DD_CL] + e.uy[DD_CP] * e.uy[DD_CL]) + Dl * e.lap[DD_CL] -
                          (m.gamma_T * e.u[DD_T]) * e.ux[DD_CL] - cl1 * (m.gamma_T * e.ux[DD_T]) -
                          m.K4 * e.u[DD_CP] * cl1);
}

// (Hcs = H_eta(cs): evaluated once by the caller for fcd and fcs)
DD_HD double dd_src_fcd(const DDModel& m, const DDExact& e, double Hcs) {
    double dT;
    const double Dd = dd_Dd_dT(m, e.u[DD_CP], e.u[DD_T], &dT);
    const double dC = -m.phi_d * Dd;
    return e.ut[DD_CD] - ((dC * e.ux[DD_CP] + dT * e.ux[DD_T]) * e.ux[DD_CD] +
                          (dC * e.uy[DD_CP] + dT * e.uy[DD_T]) * e.uy[DD_CD] + Dd * e.lap[DD_CD] +
                          m.Kd * (m.Sd - e.u[DD_CD]) * (e.u[DD_CL] + 1.0) * Hcs);
}

DD_HD double dd_src_fcs(const DDModel& m, const DDExact& e, double Hcs) {
    return e.ut[DD_CS] + m.Kd * (1.0 + e.u[DD_CL]) * (m.Sd - e.u[DD_CD]) * Hcs;
}

// 3x3 Gauss-Legendre cell average of fcp_ptwise = dt cp + cp (K1 (1+cl) + K2 T)
// over [x_{i-1/2}, x_{i+1/2}] x [y_{j-1/2}, y_{j+1/2}], interior nodes only
// (reference src/prob1base.py:493-598, 2313-2328); separable form, see DDTables.
DD_HD double dd_fcp_avg_separable(const DDModel& m, const DDSpatial& sp, const DDTimeCoef& tc) {
    return 0.25 * (tc.c[5 + DD_CP] * sp.Q1 +
                   tc.c[DD_CP] * (m.K1 * sp.Q1 + m.K1 * tc.c[DD_CL] * sp.Q2 + m.K2 * tc.c[DD_T] * sp.Q3));
}

DD_HD double dd_fcp_avg_expsin(const DDModel& m, const DDTables& tb, const DDTimeCoef& tc, int i, int j) {
    const double wq[3] = {5.0 / 9.0, 8.0 / 9.0, 5.0 / 9.0};
    const double tau = tc.c[0], em = tc.c[2], A = tc.c[3], B = tc.c[4], dA = tc.c[5], dB = tc.c[6];
    double acc = 0.0;
#pragma unroll
    for (int a = 0; a < 3; ++a) {
        const double sx = tb.XQ0[i * 3 + a];
#pragma unroll
        for (int b = 0; b < 3; ++b) {
            const double W = sx * tb.YQ0[j * 3 + b];
            const double cp = W * exp(A + B * W);
            const double f = cp * (dA + dB * W) + cp * (m.K1 * (1.0 - em * W) + m.K2 * (tau * W));
            acc += wq[a] * wq[b] * f;
        }
    }
    return 0.25 * acc;
}

// ---------------------------------------------------------------------------
// forcing front-end: one struct per kernel launch, evaluated per node
// ---------------------------------------------------------------------------

struct DDForcing {
    DDTables tab;
    DDForcingArrays arr;
};

struct DDSrc {
    double fcp, fT, fcl, fcd, fcs;
};

// Time-independent preparation of the sources at node (i, j) (global indices): for SEPARABLE
// solutions the spatial sums are evaluated once and reused for both time slots.
template <int MODE>
DD_HD void dd_src_prepare(const DDForcing& F, int i, int j, bool want_q, DDSpatial* sp) {
    if (MODE == DD_FORCING_SEPARABLE) dd_spatial_separable(F.tab, i, j, want_q, sp);
}

// All five sources at the node for time slot `slot`.  `interior` selects whether the cell-averaged
// fcp is evaluated (it is zero on the boundary); `off` is the node's element offset (ARRAYS mode).
template <int MODE>
DD_HD DDSrc dd_sources(const DDForcing& F, const DDMember& mb, const DDSpatial& sp, int slot, int i, int j,
                       long long off, bool interior, bool want_cp) {
    DDSrc s;
    s.fcp = s.fT = s.fcl = s.fcd = s.fcs = 0.0;
    if (MODE == DD_FORCING_ARRAYS) {
        s.fcp = F.arr.f[DD_CP][slot] ? F.arr.f[DD_CP][slot][off] : 0.0;
        s.fT = F.arr.f[DD_T][slot] ? F.arr.f[DD_T][slot][off] : 0.0;
        s.fcl = F.arr.f[DD_CL][slot] ? F.arr.f[DD_CL][slot][off] : 0.0;
        s.fcd = F.arr.f[DD_CD][slot] ? F.arr.f[DD_CD][slot][off] : 0.0;
        s.fcs = F.arr.f[DD_CS][slot] ? F.arr.f[DD_CS][slot][off] : 0.0;
    } else if (MODE == DD_FORCING_SEPARABLE || MODE == DD_FORCING_EXPSIN) {
        DDExact e;
        const DDTimeCoef& tc = mb.tc[slot];
        if (MODE == DD_FORCING_SEPARABLE)
            dd_exact_separable(sp, tc, &e);
        else
            dd_exact_expsin(mb.m, F.tab, tc, i, j, &e);
        s.fT = dd_src_fT(mb.m, e);
        s.fcl = dd_src_fcl(mb.m, e);
        const double Hcs = dd_F2(mb.m, e.u[DD_CS]);
        s.fcd = dd_src_fcd(mb.m, e, Hcs);
        s.fcs = dd_src_fcs(mb.m, e, Hcs);
        if (interior && want_cp) {
            s.fcp = (MODE == DD_FORCING_SEPARABLE) ? dd_fcp_avg_separable(mb.m, sp, tc)
                                                   : dd_fcp_avg_expsin(mb.m, F.tab, tc, i, j);
        }
    }
    return s;
}

// exact solution values only (error norms, initial states)
template <int MODE>
DD_HD void dd_exact_values(const DDForcing& F, const DDMember& mb, int slot, int i, int j, double* u) {
    DDExact e;
    if (MODE == DD_FORCING_SEPARABLE) {
        DDSpatial sp;
        dd_spatial_separable(F.tab, i, j, false, &sp);
        dd_exact_separable(sp, mb.tc[slot], &e);
    } else {
        dd_exact_expsin(mb.m, F.tab, mb.tc[slot], i, j, &e);
    }
    for (int v = 0; v < DD_NVAR; ++v) u[v] = e.u[v];
}

// ---------------------------------------------------------------------------
// Newton systems in Jacobi-scaled form
//   x_c = bb + aW x_w + aE x_e + aS x_s + aN x_n ,   x = v_new - v*
// (A = 2I - dt J with couplings to boundary nodes dropped,
//  reference src/prob1base.py:680-682, 2998-3115)
// ---------------------------------------------------------------------------

struct DDRow {
    double bb, aW, aE, aS, aN;
};

DD_HD DDRow dd_make_row(double d, double oW, double oE, double oS, double oN, double rhs, int i, int j, int N,
                        int M) {
    // d: diagonal of A;  oX: -A(i,j; neighbour) (so that x_c = (rhs + sum oX x_X)/d)
    DDRow r;
    const double inv = dd_rcp(d);
    r.bb = rhs * inv;
    r.aW = (i > 1) ? oW * inv : 0.0;
    r.aE = (i < N - 1) ? oE * inv : 0.0;
    r.aS = (j > 1) ? oS * inv : 0.0;
    r.aN = (j < M - 1) ? oN * inv : 0.0;
    return r;
}

// T system (reference newton_step_T 2998-3045, delT_ab_FT_ij 2674-2684)
DD_HD DDRow dd_row_T(const DDModel& m, const DDNodeGeo& q, double dt, const DDSten& Ts, double cps, double YT,
                     double fT1, int i, int j, int N, int M) {
    const double W = m.DT * q.cW, E = m.DT * q.cE, S = m.DT * q.cS, Nn = m.DT * q.cN;
    const double C = -(W + E + S + Nn) - m.K3 * cps;
    const double G0 = 2.0 * Ts.c - dt * (fT1 + dd_FT_int(m, q, Ts, cps));
    return dd_make_row(2.0 - dt * C, dt * W, dt * E, dt * S, dt * Nn, YT - G0, i, j, N, M);
}

// cl system (newton_step_cl 3047-3080, delcl_ab_Fcl_ij 2716-2750, delT_ab_Fcl_ij 2686-2714
// applied to interior w only, 2234-2255).  wW / wE = (T1 - T*) at (i-1,j) / (i+1,j).
DD_HD DDRow dd_row_cl(const DDModel& m, const DDNodeGeo& q, double dt, const DDSten& cps, const DDSten& Ts,
                      const DDSten& cls, double wW, double wE, double Ycl, double fcl1, int i, int j, int N,
                      int M) {
    const DDFaces Dl = dd_faces_Dl(m, cps);
    const double dW = Dl.w * q.cW, dE = Dl.e * q.cE, S = Dl.s * q.cS, Nn = Dl.n * q.cN;
    const double W = dW + (m.gamma_T * Ts.w) * (0.5 * q.rhp);
    const double E = dE - (m.gamma_T * Ts.e) * (0.5 * q.rhp);
    const double C = -(dW + dE + S + Nn) - m.K4 * cps.c;
    const double Fcl = fcl1 + dd_Fcl_int(m, q, Dl, Ts, cls, cps.c);
    const double jw = (i > 1) ? m.gamma_T * (1.0 + cls.w) * wW : 0.0;
    const double je = (i < N - 1) ? m.gamma_T * (1.0 + cls.e) * wE : 0.0;
    const double JT = (jw - je) * (0.5 * q.rhp);
    const double rhs = Ycl - 2.0 * cls.c + dt * Fcl + dt * JT;
    return dd_make_row(2.0 - dt * C, dt * W, dt * E, dt * S, dt * Nn, rhs, i, j, N, M);
}

// cd system (newton_step_cd 3082-3115, delcd_ab_Fcd_ij 2811-2839 with the band
// exchange of 3094-3100 when `swap`; delT_ab_Fcd_ij 2752-2800 and
// delcl_ab_Fcd_ij 2802-2809 applied to full-grid w, 2257-2293).
// w* = (T1 - T*) stencil, dcl = cl1 - cl* at the node.
DD_HD DDRow dd_row_cd(const DDModel& m, const DDNodeGeo& q, double dt, const DDSten& cps, const DDSten& Ts,
                      double cls, const DDSten& cds, double css, const DDSten& w, double dcl, double Ycd,
                      double fcd1, int swap, int i, int j, int N, int M) {
    DDFaces dT;
    const DDFaces Dd = dd_faces_Dd_dT(m, cps, Ts, &dT);
    const double W = Dd.w * q.cW, E = Dd.e * q.cE, S = Dd.s * q.cS, Nn = Dd.n * q.cN;
    const double KH = m.Kd * dd_F2(m, css);
    const double C = -(W + E + S + Nn) - KH * (cls + 1.0);
    const double Fcd = fcd1 + dd_div_flux(q, Dd, cds) + (m.Sd - cds.c) * (cls + 1.0) * KH;
    const double gxw = ((cds.c - cds.w) * q.rhW) * dT.w, gxe = ((cds.e - cds.c) * q.rhE) * dT.e;
    const double gys = ((cds.c - cds.s) * q.rkS) * dT.s, gyn = ((cds.n - cds.c) * q.rkN) * dT.n;
    const double JT = q.rhp * (-gxw * (0.5 * (w.c + w.w)) + gxe * (0.5 * (w.e + w.c))) +
                      q.rkp * (-gys * (0.5 * (w.c + w.s)) + gyn * (0.5 * (w.n + w.c)));
    const double Jcl = KH * (m.Sd - cds.c) * dcl;
    const double rhs = Ycd - 2.0 * cds.c + dt * Fcd + dt * JT + dt * Jcl;
    const double oW = swap ? S : W, oS = swap ? W : S;
    return dd_make_row(2.0 - dt * C, dt * oW, dt * E, dt * oS, dt * Nn, rhs, i, j, N, M);
}

// ---------------------------------------------------------------------------
// cp / cs predictors and correctors
// ---------------------------------------------------------------------------

// Heun predictor for cp (reference initial_cp_pred 2953-2965), interior node
DD_HD double dd_predict_cp(const DDModel& m, double dt, double cp0, double T0, double cl0, double fcp0, double fcp1) {
    const double F0 = fcp0 + dd_Fcp_int(m, cp0, T0, cl0);
    const double star = cp0 + dt * F0;
    const double Fs = fcp1 + dd_Fcp_int(m, star, T0, cl0);
    return cp0 + 0.5 * dt * (F0 + Fs);
}

// Heun predictor for cs (reference initial_cs_pred 3631-3645), interior node
// (react0 = dd_reaction(m, cl0, cd0, cs0), which the caller needs for Fcd as well)
DD_HD double dd_predict_cs_r(const DDModel& m, double dt, double cs0, double cl0, double cd0, double fcs0, double fcs1,
                             double react0) {
    const double F0 = fcs0 - react0;
    const double star = cs0 + dt * F0;
    const double Fs = fcs1 - dd_reaction(m, cl0, cd0, star);
    return cs0 + (0.5 * dt) * (F0 + Fs);
}

DD_HD double dd_predict_cs(const DDModel& m, double dt, double cs0, double cl0, double cd0, double fcs0, double fcs1) {
    const double F0 = fcs0 - dd_reaction(m, cl0, cd0, cs0);
    const double star = cs0 + dt * F0;
    const double Fs = fcs1 - dd_reaction(m, cl0, cd0, star);
    return cs0 + (0.5 * dt) * (F0 + Fs);
}

// trapezoidal corrector for cp (reference corrector_cp_step 2967-2996), interior node
DD_HD double dd_correct_cp(const DDModel& m, double dt, double cp0, double T0, double cl0, double T1, double cl1,
                           double fcp0, double fcp1) {
    const double a0 = -m.K2 * T0 - m.K1 * (cl0 + 1.0);
    const double a1 = -m.K2 * T1 - m.K1 * (cl1 + 1.0);
    const double num = (1.0 + (dt / 2.0) * a0) * cp0 + (dt / 2.0) * (fcp0 + fcp1);
    return num * dd_rcp(1.0 - (dt / 2.0) * a1);
}

// implicit cs corrector (reference corrector_cs_step 3665-3702): y, a of
//   2x + (2x - y) e^{-eta x} = y - a
DD_HD void dd_cs_ya(const DDModel& m, double dt, double cs0, double cl0, double cd0, double cl1, double cd1,
                    double fcs0, double fcs1, double* y, double* a) {
    *y = 2.0 * cs0 - dt * m.Kd * (m.Sd - cd0) * (cl0 + 1.0) * dd_H(cs0, m.eta) + dt * (fcs0 + fcs1);
    *a = dt * m.Kd * (m.Sd - cd1) * (cl1 + 1.0);
}

// closed-form cs correctors of the other two field variants (the regularised one iterates, see below)
// CsTriple (reference corrector_cs_step 3191-3219): trapezoidal rule solved for cs1, interior node
DD_HD double dd_correct_cs_cstriple(const DDModel& m, double dt, double cs0, double cl0, double cd0, double cl1,
                                    double cd1, double fcs0, double fcs1) {
    const double a0 = -m.Kd * (m.Sd - cd0) * (1.0 + cl0);
    const double a1 = -m.Kd * (m.Sd - cd1) * (1.0 + cl1);
    const double num = (1.0 + (dt / 2.0) * a0) * cs0 + (dt / 2.0) * (fcs0 + fcs1);
    return num * dd_rcp(1.0 - (dt / 2.0) * a1);
}

// HCsTriple (reference corrector_cs_step 3393-3430), any node: Y0 = 2 cs0 + dt Fcs(u0, t0) + dt fcs(t1);
// cs1 = Y0 / (2 - dt R1) where Y0 > tol, Y0 / 2 where Y0 < -tol, else 0; times the boundary mask.
// *bad is set when 2 - dt Kd (Sd - cd1)(1 + cl1) < tol (the reference raises ValueError if that holds anywhere).
DD_HD double dd_correct_cs_hcstriple(const DDModel& m, double dt, double cs0, double cl0, double cd0, double cl1,
                                     double cd1, double fcs0, double fcs1, bool interior, int* bad) {
    const double tol = 2.220446049250313e-16 * 100.0;
    const double dY1 = 2.0 - dt * ((m.Sd - cd1) * (1.0 + cl1) * m.Kd);
    if (dY1 < tol) *bad = 1;
    const double Fcs0 = interior ? fcs0 - dd_reaction(m, cl0, cd0, cs0) : (fcs0 - 0.0) * 0.0;
    const double Y0 = 2.0 * cs0 + dt * Fcs0 + dt * fcs1;
    double r = 0.0;
    if (Y0 > tol) r = Y0 / dY1;
    else if (Y0 < -tol) r = Y0 / 2.0;
    return r * (interior ? 1.0 : 0.0);
}

// one Newton update (reference _newton_iterations 3654-3663); returns dx
DD_HD double dd_cs_newton_dx(double x, double y, double a, double eta) {
    const double ex = dd_exp(-eta * x);
    const double f = 2.0 * x + (2.0 * x - y) * ex - y + a;
    const double J = 2.0 + 2.0 * ex - eta * (2.0 * x - y) * ex;
    return -f * dd_rcp(J);
}
