// dd_nodeprog.cuh -- what each kernel does at ONE grid node.
//
// A "node program" reads its 5-point neighbourhood straight from the field
// arrays (global memory on the GPU; plain host arrays in tests/hostsim) and
// writes the node's outputs.  The CUDA kernels in dd_kernels.cu only map
// threads to (member, row, col) and call these.
#pragma once

#include "dd_physics.cuh"

struct DDRows {
    double *bb, *aW, *aE, *aS, *aN;  // Jacobi-scaled five-point rows (see dd_physics.cuh)
    int ld;                          // row pitch of these arrays: even, so that the solver can fetch the two
    long long mstride;               // colours of a packed column with one aligned 16-byte load
};

DD_HD bool dd_is_interior(const DDGeom& g, int i, int j) { return i > 0 && i < g.N && j > 0 && j < g.M; }

// ---------------------------------------------------------------------------
// semidiscrete field F(state, t) at one node -- reference Fcp/FT/Fcl/Fcd/Fcs
// (src/prob1base.py:2599-2672): source on every node, operator on interior.
// ---------------------------------------------------------------------------
template <int MODE>
DD_HD void dd_node_F(const DDGeom& g, const DDMember& mb, const DDForcing& F, const DDStateC& s, long long mo,
                     int r, int j, int slot, double* Fout /*[5]*/) {
    const int i = g.row0 + r;
    const long long o = mo + (long long)r * g.ld + j;
    const bool inter = dd_is_interior(g, i, j);
    DDSpatial sp;
    dd_src_prepare<MODE>(F, i, j, inter, &sp);
    const DDSrc src = dd_sources<MODE>(F, mb, sp, slot, i, j, o, inter, true);
    if (!inter) {
        Fout[DD_CP] = src.fcp;  // cell average is zero-padded on the boundary
        Fout[DD_T] = src.fT;
        Fout[DD_CL] = src.fcl;
        Fout[DD_CD] = src.fcd;
        Fout[DD_CS] = (src.fcs - 0.0) * 0.0;  // (fcs - R) * mask
        if (MODE != DD_FORCING_ARRAYS) Fout[DD_CP] = 0.0;
        return;
    }
    const DDModel& m = mb.m;
    const DDNodeGeo q = dd_node_geo(g, i, j);
    const DDSten cp = dd_load_sten(s.v[DD_CP], o, g.ld);
    const DDSten T = dd_load_sten(s.v[DD_T], o, g.ld);
    const DDSten cl = dd_load_sten(s.v[DD_CL], o, g.ld);
    const DDSten cd = dd_load_sten(s.v[DD_CD], o, g.ld);
    const double cs = s.v[DD_CS][o];
    Fout[DD_CP] = src.fcp + dd_Fcp_int(m, cp.c, T.c, cl.c);
    Fout[DD_T] = src.fT + dd_FT_int(m, q, T, cp.c);
    Fout[DD_CL] = src.fcl + dd_Fcl_int(m, q, dd_faces_Dl(m, cp), T, cl, cp.c);
    const double react = dd_reaction(m, cl.c, cd.c, cs);  // once for Fcd and Fcs
    Fout[DD_CD] = src.fcd + (dd_div_flux(q, dd_faces_Dd(m, cp, T), cd) + react);
    Fout[DD_CS] = src.fcs - react;
}

// forward Euler (reference ForwardEulerIntegrator.step, src/prob1base.py:2889-2903):
// every node, boundary included.
template <int MODE>
DD_HD void dd_node_feuler(const DDGeom& g, const DDMember& mb, const DDForcing& F, const DDStateC& in,
                          const DDState& out, long long mo, int r, int j) {
    double Fv[DD_NVAR];
    dd_node_F<MODE>(g, mb, F, in, mo, r, j, 0, Fv);
    const long long o = mo + (long long)r * g.ld + j;
    for (int v = 0; v < DD_NVAR; ++v) out.v[v][o] = in.v[v][o] + mb.dt * Fv[v];
}

// ---------------------------------------------------------------------------
// PC step, phase 1: Y_T, Y_cl, Y_cd and the Heun predictors of cp, cs
// (reference step 3122-3130, initial_cp_pred 2953-2965, initial_cs_pred 3631-3645)
// ---------------------------------------------------------------------------
struct DDPredictOut {
    double *cp1p, *cs1p, *YT, *Ycl, *Ycd;
};

template <int MODE>
DD_HD void dd_node_predict(const DDGeom& g, const DDMember& mb, const DDForcing& F, const DDStateC& s,
                           const DDPredictOut& out, long long mo, int r, int j) {
    const int i = g.row0 + r;
    const long long o = mo + (long long)r * g.ld + j;
    const bool inter = dd_is_interior(g, i, j);
    const double dt = mb.dt;
    DDSpatial sp;
    dd_src_prepare<MODE>(F, i, j, inter, &sp);
    const DDSrc s0 = dd_sources<MODE>(F, mb, sp, 0, i, j, o, inter, true);
    if (!inter) {
        out.YT[o] = dt * s0.fT + 2.0 * s.v[DD_T][o];
        out.Ycl[o] = dt * s0.fcl + 2.0 * s.v[DD_CL][o];
        out.Ycd[o] = dt * s0.fcd + 2.0 * s.v[DD_CD][o];
        // cp: F = zero-padded cell average -> predictor keeps cp0 (ARRAYS mode: whatever the
        // caller's boundary values are); cs: (...) * mask = 0
        double cpb = s.v[DD_CP][o];
        if (MODE == DD_FORCING_ARRAYS) {
            const DDSrc s1b = dd_sources<MODE>(F, mb, sp, 1, i, j, o, inter, true);
            cpb = cpb + 0.5 * dt * (s0.fcp + s1b.fcp);
        }
        out.cp1p[o] = cpb;
        // cs: Fcs carries the mask, so the Heun predictor keeps cs0 on the boundary; the RegH / H integrators
        // then multiply by the mask (3644, 3391), the CsTriple one does not (3175-3189)
        out.cs1p[o] = (mb.m.react == DD_REACT_CS) ? s.v[DD_CS][o] : s.v[DD_CS][o] * 0.0;
        return;
    }
    const DDSrc s1 = dd_sources<MODE>(F, mb, sp, 1, i, j, o, inter, true);
    const DDModel& m = mb.m;
    const DDNodeGeo q = dd_node_geo(g, i, j);
    const DDSten cp = dd_load_sten(s.v[DD_CP], o, g.ld);
    const DDSten T = dd_load_sten(s.v[DD_T], o, g.ld);
    const DDSten cl = dd_load_sten(s.v[DD_CL], o, g.ld);
    const DDSten cd = dd_load_sten(s.v[DD_CD], o, g.ld);
    const double cs = s.v[DD_CS][o];
    out.YT[o] = dt * (s0.fT + dd_FT_int(m, q, T, cp.c)) + 2.0 * T.c;
    out.Ycl[o] = dt * (s0.fcl + dd_Fcl_int(m, q, dd_faces_Dl(m, cp), T, cl, cp.c)) + 2.0 * cl.c;
    const double react0 = dd_reaction(m, cl.c, cd.c, cs);
    out.Ycd[o] = dt * (s0.fcd + (dd_div_flux(q, dd_faces_Dd(m, cp, T), cd) + react0)) + 2.0 * cd.c;
    out.cp1p[o] = dd_predict_cp(m, dt, cp.c, T.c, cl.c, s0.fcp, s1.fcp);
    out.cs1p[o] = dd_predict_cs_r(m, dt, cs, cl.c, cd.c, s0.fcs, s1.fcs, react0);
}

// ---------------------------------------------------------------------------
// PC step, phases 2-4: assemble the Jacobi-scaled Newton rows for T, cl, cd.
// `u` is the linearisation state (cp1p, T*, cl*, cd*, cs1p).  Returns the
// node's Gershgorin ratio sum|a| (0 on boundary nodes).
// ---------------------------------------------------------------------------
DD_HD void dd_store_row(const DDRows& R, long long o, const DDRow& row) {
    R.bb[o] = row.bb;
    R.aW[o] = row.aW;
    R.aE[o] = row.aE;
    R.aS[o] = row.aS;
    R.aN[o] = row.aN;
}

DD_HD double dd_row_rho(const DDRow& r) { return fabs(r.aW) + fabs(r.aE) + fabs(r.aS) + fabs(r.aN); }

template <int MODE>
DD_HD double dd_node_asm_T(const DDGeom& g, const DDMember& mb, const DDForcing& F, const DDStateC& u,
                           const double* YT, const DDRows& R, long long mo, long long moR, int r, int j) {
    const int i = g.row0 + r;
    const long long o = mo + (long long)r * g.ld + j;
    const long long oR = moR + (long long)r * R.ld + j;
    DDRow row = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (dd_is_interior(g, i, j)) {
        DDSpatial sp;
        dd_src_prepare<MODE>(F, i, j, false, &sp);
        const DDSrc s1 = dd_sources<MODE>(F, mb, sp, 1, i, j, o, true, false);
        const DDNodeGeo q = dd_node_geo(g, i, j);
        const DDSten T = dd_load_sten(u.v[DD_T], o, g.ld);
        row = dd_row_T(mb.m, q, mb.dt, T, u.v[DD_CP][o], YT[o], s1.fT, i, j, g.N, g.M);
    }
    dd_store_row(R, oR, row);
    return dd_row_rho(row);
}

// T system in "constant band" form: the off-diagonals of A_T are dt DT / (hhat_i h_i) etc. -- pure grid
// geometry -- so only bb = rhs / d and dinv = 1 / d are stored (16 B instead of 40 B per node); the tile
// solver rebuilds a_W = dinv * dt DT cW_i from 1-D arrays.  Couplings to boundary nodes need no masking:
// x is identically 0 on boundary nodes (bb = dinv = 0 there).
template <int MODE>
DD_HD double dd_node_asm_T_const(const DDGeom& g, const DDMember& mb, const DDForcing& F, const DDStateC& u,
                                 const double* YT, const DDRows& R, long long mo, long long moR, int r, int j) {
    const int i = g.row0 + r;
    const long long o = mo + (long long)r * g.ld + j;
    const long long oR = moR + (long long)r * R.ld + j;
    double vb = 0.0, vd = 0.0, rho = 0.0;
    if (dd_is_interior(g, i, j)) {
        DDSpatial sp;
        dd_src_prepare<MODE>(F, i, j, false, &sp);
        const DDSrc s1 = dd_sources<MODE>(F, mb, sp, 1, i, j, o, true, false);
        const DDNodeGeo q = dd_node_geo(g, i, j);
        const DDSten T = dd_load_sten(u.v[DD_T], o, g.ld);
        const DDModel& m = mb.m;
        const double cps = u.v[DD_CP][o], dt = mb.dt;
        const double sumc = m.DT * (q.cW + q.cE + q.cS + q.cN);
        const double d = 2.0 + dt * (sumc + m.K3 * cps);
        const double G0 = 2.0 * T.c - dt * (s1.fT + dd_FT_int(m, q, T, cps));
        vd = dd_rcp(d);
        vb = (YT[o] - G0) * vd;
        rho = dt * sumc * vd;
    }
    R.bb[oR] = vb;
    R.aW[oR] = vd;  // aW holds dinv in the constant-band form
    return rho;
}

template <int MODE>
DD_HD double dd_node_asm_cl(const DDGeom& g, const DDMember& mb, const DDForcing& F, const DDStateC& u,
                            const double* T1, const double* Ycl, const DDRows& R, long long mo, long long moR, int r,
                            int j) {
    const int i = g.row0 + r;
    const long long o = mo + (long long)r * g.ld + j;
    const long long oR = moR + (long long)r * R.ld + j;
    DDRow row = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (dd_is_interior(g, i, j)) {
        DDSpatial sp;
        dd_src_prepare<MODE>(F, i, j, false, &sp);
        const DDSrc s1 = dd_sources<MODE>(F, mb, sp, 1, i, j, o, true, false);
        const DDNodeGeo q = dd_node_geo(g, i, j);
        const DDSten cp = dd_load_sten(u.v[DD_CP], o, g.ld);
        const DDSten T = dd_load_sten(u.v[DD_T], o, g.ld);
        const DDSten cl = dd_load_sten(u.v[DD_CL], o, g.ld);
        const double wW = T1[o - g.ld] - T.w, wE = T1[o + g.ld] - T.e;
        row = dd_row_cl(mb.m, q, mb.dt, cp, T, cl, wW, wE, Ycl[o], s1.fcl, i, j, g.N, g.M);
    }
    dd_store_row(R, oR, row);
    return dd_row_rho(row);
}

template <int MODE>
DD_HD double dd_node_asm_cd(const DDGeom& g, const DDMember& mb, const DDForcing& F, const DDStateC& u,
                            const double* T1, const double* cl1, const double* Ycd, int swap, const DDRows& R,
                            long long mo, long long moR, int r, int j) {
    const int i = g.row0 + r;
    const long long o = mo + (long long)r * g.ld + j;
    const long long oR = moR + (long long)r * R.ld + j;
    DDRow row = {0.0, 0.0, 0.0, 0.0, 0.0};
    if (dd_is_interior(g, i, j)) {
        DDSpatial sp;
        dd_src_prepare<MODE>(F, i, j, false, &sp);
        const DDSrc s1 = dd_sources<MODE>(F, mb, sp, 1, i, j, o, true, false);
        const DDNodeGeo q = dd_node_geo(g, i, j);
        const DDSten cp = dd_load_sten(u.v[DD_CP], o, g.ld);
        const DDSten T = dd_load_sten(u.v[DD_T], o, g.ld);
        const DDSten cd = dd_load_sten(u.v[DD_CD], o, g.ld);
        const DDSten t1 = dd_load_sten(T1, o, g.ld);
        DDSten w;
        w.c = t1.c - T.c; w.w = t1.w - T.w; w.e = t1.e - T.e; w.s = t1.s - T.s; w.n = t1.n - T.n;
        const double clc = u.v[DD_CL][o];
        row = dd_row_cd(mb.m, q, mb.dt, cp, T, clc, cd, u.v[DD_CS][o], w, cl1[o] - clc, Ycd[o], s1.fcd, swap, i, j,
                        g.N, g.M);
    }
    dd_store_row(R, oR, row);
    return dd_row_rho(row);
}

// Newton update after the solve: v_new = v* + x on the interior; on the boundary
// T := 0 (reference 3038-3039) while cl, cd keep v* (2102-2106).
DD_HD double dd_newton_update(bool interior, double vstar, double x, int zero_boundary) {
    if (interior) return vstar + x;
    return zero_boundary ? 0.0 : vstar;
}

// ---------------------------------------------------------------------------
// PC step, phase 5: correctors.  cp: trapezoidal closed form (interior, 0 on the
// boundary).  cs: y and a of the implicit equation on ALL nodes; the Newton
// iterations themselves run in the kernel because of the global exit test.
// ---------------------------------------------------------------------------
template <int MODE>
DD_HD void dd_node_correct_prepare(const DDGeom& g, const DDMember& mb, const DDForcing& F, const DDStateC& s0,
                                   const double* T1, const double* cl1, const double* cd1, long long mo, int r,
                                   int j, double* cp1, double* y, double* a, double* fcs0 = nullptr,
                                   double* fcs1 = nullptr) {
    const int i = g.row0 + r;
    const long long o = mo + (long long)r * g.ld + j;
    const bool inter = dd_is_interior(g, i, j);
    DDSpatial sp;
    dd_src_prepare<MODE>(F, i, j, inter, &sp);
    const DDSrc q0 = dd_sources<MODE>(F, mb, sp, 0, i, j, o, inter, true);
    const DDSrc q1 = dd_sources<MODE>(F, mb, sp, 1, i, j, o, inter, true);
    const DDModel& m = mb.m;
    *cp1 = inter ? dd_correct_cp(m, mb.dt, s0.v[DD_CP][o], s0.v[DD_T][o], s0.v[DD_CL][o], T1[o], cl1[o], q0.fcp,
                                 q1.fcp)
                 : 0.0;
    dd_cs_ya(m, mb.dt, s0.v[DD_CS][o], s0.v[DD_CL][o], s0.v[DD_CD][o], cl1[o], cd1[o], q0.fcs, q1.fcs, y, a);
    if (fcs0) *fcs0 = q0.fcs;
    if (fcs1) *fcs1 = q1.fcs;
}

// cs corrector of the closed-form variants (CsTriple, HCsTriple) at a node; see dd_physics.cuh
DD_HD double dd_node_correct_cs_closed(const DDGeom& g, const DDMember& mb, const DDStateC& s0, const double* cl1,
                                       const double* cd1, long long mo, int r, int j, double fcs0, double fcs1,
                                       int* bad) {
    const long long o = mo + (long long)r * g.ld + j;
    const bool inter = dd_is_interior(g, g.row0 + r, j);
    const DDModel& m = mb.m;
    if (m.react == DD_REACT_CS)
        return inter ? dd_correct_cs_cstriple(m, mb.dt, s0.v[DD_CS][o], s0.v[DD_CL][o], s0.v[DD_CD][o], cl1[o],
                                              cd1[o], fcs0, fcs1)
                     : 0.0;
    return dd_correct_cs_hcstriple(m, mb.dt, s0.v[DD_CS][o], s0.v[DD_CL][o], s0.v[DD_CD][o], cl1[o], cd1[o], fcs0,
                                   fcs1, inter, bad);
}
