// dd_types.h -- plain-old-data shared by host code, CUDA kernels and the C ABI.
//
// Layout of one field in HBM: [member][row][col], float64, row pitch `ld`
// (= M+1, columns j contiguous), `nrows` local rows per member.  Local row r
// is global grid row i = row0 + r (row0 > 0 only for slab-decomposed meshes),
// which mirrors the reference's C-order (N+1, M+1) arrays
// (reference src/prob1base.py:242-248).
#pragma once

#include <stdint.h>

#ifdef __CUDACC__
#define DD_HD __host__ __device__ __forceinline__
#else
#define DD_HD inline
#endif

enum DDForcingMode {
    DD_FORCING_NONE = 0,       // NoForcingTerms (reference src/prob1base.py:852-869)
    DD_FORCING_ARRAYS = 1,     // host-evaluated source fields uploaded per step
    DD_FORCING_SEPARABLE = 2,  // u_v = phi_v(t) X_v(x) Y_v(y): 1-D tables + time profile
    DD_FORCING_EXPSIN = 3,     // MMSCaseExpSin closed form (reference src/prob1_mms_cases.py:296-337)
    DD_FORCING_PROGRAM = 4     // generated kernel (include/dd_b200_program.h); host-level mode only: the step kernels
                               // see its output as ARRAYS
};

enum DDPhiKind {
    DD_PHI_INV1PT = 0,  // p0 / (1 + t)
    DD_PHI_EXP = 1,     // p0 * exp(-p1 t)
    DD_PHI_LINEAR = 2,  // p0 - p1 t
    DD_PHI_OSC = 3,     // p0 * (1 + p1 sin(p2 t))
    DD_PHI_CONST = 4,   // p0
    DD_PHI_HOST = 5     // host supplies phi(t0), phi'(t0), phi(t1), phi'(t1) before each step
};

enum { DD_CP = 0, DD_T = 1, DD_CL = 2, DD_CD = 3, DD_CS = 4, DD_NVAR = 5 };

// Per-member physical parameters (reference ModelConsts, src/prob1base.py:28-45,
// plus the regularisation factor eta of the RegHCsTriple family, 3452-3466).
struct DDModel {
    double K1, K2, K3, K4, DT, Dl_max, phi_l, gamma_T, Kd, Sd, Dd_max, phi_d, phi_T, r_sp;
    double T_shift;  // T_ref for DefaultModel02 (src/prob1base.py:205-217), 0 for DefaultModel01
    double eta;
    int react, _pad;  // DD_REACT_*: which F2(cs) the cs/cd interaction uses
};

// [Cs-Cd-int] = Kd (Sd - cd)(1 + cl) F2(cs): the three field variants of the reference
enum {
    DD_REACT_REGH = 0,  // F2 = H_eta(cs)   RegHCsTriple, src/prob1base.py:3553-3593
    DD_REACT_CS = 1,    // F2 = cs          CsTriple,     src/prob1base.py:2842-2876
    DD_REACT_H = 2      // F2 = (cs > 0)    HCsTriple,    src/prob1base.py:3303-3340
};

// Per-member, per-time-slot scalars of the manufactured solution
// (slot 0 = t0, slot 1 = t1 = t0 + dt).
struct DDTimeCoef {
    double c[16];
    // SEPARABLE: c[v] = phi_v(t), c[5+v] = phi_v'(t)
    // EXPSIN   : see dd_physics.cuh (expsin_time_coefs)
};

struct DDMember {
    DDModel m;
    double t0, dt;
    int phi_kind[DD_NVAR];
    int active;  // 0 -> member is skipped (already finished its trajectory)
    double phi_p[DD_NVAR][4];
    DDTimeCoef tc[2];
};

// Grid geometry, global 1-D arrays on the device (length N+1 / M+1).
// rh[i] = 1/h_i (rh[0] = 0), rhp[i] = 1/hhat_i (0 at i = 0, N); same in y.
struct DDGeom {
    int N, M;        // global grid: nodes 0..N x 0..M
    int row0, nrows; // this batch holds global rows [row0, row0 + nrows)
    int ld;          // row pitch in doubles
    long long mstride;  // member stride in doubles (= nrows * ld)
    const double *x, *y, *h, *k, *hp, *kp, *rh, *rk, *rhp, *rkp;
};

// 1-D tables for the fused MMS forcing (device pointers).
// SEPARABLE: u_v = phi_v(t) sum_{r < nterms} X_{p,r}(x) Y_{p,r}(y), p = var_prof[v] (variables with
//   identical spatial tables share one "profile": MMSCasePol has one, NonFullySmoothPol two).
//   X[p][d][r*nx + i], d = 0,1,2 -> X_{p,r}, X', X'' at node i; Y likewise (stride ny).
//   The 3x3 Gauss cell average of fcp_ptwise = dt cp + cp (K1 (1 + cl) + K2 T) separates as well:
//     avg = 1/4 [ phi_cp' Q1 + phi_cp (K1 Q1 + K1 phi_cl Q2 + K2 phi_T Q3) ],
//     Q1 = sum_r QX1[r][i] QY1[r][j],  Q2 = sum_{r,s} QX2[r*R+s][i] QY2[..][j] (cp x cl),  Q3 (cp x T),
//   with QX1[r][i] = sum_a w_a X_cp,r(p_a(i)) etc. built on the host from the quadrature samples.
// EXPSIN: X[0][0] = sin(pi x), X[0][1] = cos(pi x); XQ0 = sin(pi x) at the 3 abscissae of each cell.
struct DDTables {
    int nterms, nx, ny, nprof;  // nx = N+1, ny = M+1
    int var_prof[DD_NVAR];
    const double* X[DD_NVAR][3];
    const double* Y[DD_NVAR][3];
    const double *QX1, *QY1, *QX2, *QY2, *QX3, *QY3;
    const double *XQ0, *YQ0;
};

// Host-evaluated forcing arrays: f[v][slot], same layout as the fields
// (fcp is the cell-averaged source, zero on the boundary).
struct DDForcingArrays {
    const double* f[DD_NVAR][2];
};

// Five state fields (device pointers, member 0 / local row 0).
struct DDState {
    double* v[DD_NVAR];
};
struct DDStateC {
    const double* v[DD_NVAR];
};

// Per-member statistics of one linear solve, reduced on the device.
struct DDSolveStats {
    double rho;        // max_i sum_j |a_ij| / |a_ii|  (Gershgorin / Jacobi norm)
    double resid;      // max |bb - (I - G) x| of the Jacobi-scaled system
    double xmax;       // max |x|
    double vmax;       // max |v_new|
    double bmax;       // max |bb|
};
