// hostsim.cpp -- TEST-ONLY host compilation of the device node programs.
//
// Compiles na-.../csrc/dd_nodeprog.cuh with g++ (DD_HD expands to `inline`) and
// drives it with plain loops, so that the arithmetic of the CUDA kernels can be
// checked against the oracle on a machine without a GPU.  It is built by
// tests/test_hostsim.py into tests/hostsim/_build/ and is never loaded by the
// package: the product path has no CPU fallback.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "dd_nodeprog.cuh"
#include "dd_combine.cuh"
#include "dd_tables_host.h"

struct HSProblem {
    int N, M;
    const double *x, *y;
    double model[16];  // K1..T_ref (15) + eta
    int kind;
    int mode;
    int nterms;
    const double* X[5][3];
    const double* Y[5][3];
    const double* XQ[3];
    const double* YQ[3];
    int phi_kind[5];
    double phi_p[5][4];
    const double* farr[5][2];
};

struct HSCtx {
    DDGeom g;
    DDMember mb;
    DDForcing F;
    std::vector<double> h, k, hp, kp, rh, rk, rhp, rkp;
    DDHostTables ht;
};

static void setup(const HSProblem& P, HSCtx& c, double t0, double dt) {
    const int N = P.N, M = P.M;
    c.h.assign(N + 1, 0); c.k.assign(M + 1, 0); c.hp.assign(N + 1, 0); c.kp.assign(M + 1, 0);
    c.rh.assign(N + 1, 0); c.rk.assign(M + 1, 0); c.rhp.assign(N + 1, 0); c.rkp.assign(M + 1, 0);
    c.h[0] = INFINITY; c.k[0] = INFINITY;
    for (int i = 1; i <= N; ++i) c.h[i] = P.x[i] - P.x[i - 1];
    for (int j = 1; j <= M; ++j) c.k[j] = P.y[j] - P.y[j - 1];
    for (int i = 0; i < N; ++i) c.hp[i] = (c.h[i] + c.h[i + 1]) * 0.5;
    c.hp[N] = INFINITY;
    for (int j = 0; j < M; ++j) c.kp[j] = (c.k[j] + c.k[j + 1]) * 0.5;
    c.kp[M] = INFINITY;
    for (int i = 0; i <= N; ++i) { c.rh[i] = i ? 1.0 / c.h[i] : 0.0; c.rhp[i] = (i == 0 || i == N) ? 0.0 : 1.0 / c.hp[i]; }
    for (int j = 0; j <= M; ++j) { c.rk[j] = j ? 1.0 / c.k[j] : 0.0; c.rkp[j] = (j == 0 || j == M) ? 0.0 : 1.0 / c.kp[j]; }
    DDGeom& g = c.g;
    g.N = N; g.M = M; g.row0 = 0; g.nrows = N + 1; g.ld = M + 1; g.mstride = (long long)(N + 1) * (M + 1);
    g.x = P.x; g.y = P.y; g.h = c.h.data(); g.k = c.k.data(); g.hp = c.hp.data(); g.kp = c.kp.data();
    g.rh = c.rh.data(); g.rk = c.rk.data(); g.rhp = c.rhp.data(); g.rkp = c.rkp.data();
    memset(&c.mb, 0, sizeof(c.mb));
    DDModel& m = c.mb.m;
    const double* q = P.model;
    m.K1 = q[0]; m.K2 = q[1]; m.K3 = q[2]; m.K4 = q[3]; m.DT = q[4]; m.Dl_max = q[5]; m.phi_l = q[6];
    m.gamma_T = q[7]; m.Kd = q[8]; m.Sd = q[9]; m.Dd_max = q[10]; m.phi_d = q[11]; m.phi_T = q[12]; m.r_sp = q[13];
    m.T_shift = (P.kind % 10 == 2) ? q[14] : 0.0;  // kind = model kind + 10 * reaction kind
    m.eta = q[15];
    m.react = P.kind / 10;
    c.mb.active = 1;
    c.mb.t0 = t0; c.mb.dt = dt;
    for (int v = 0; v < 5; ++v) { c.mb.phi_kind[v] = P.phi_kind[v]; for (int s = 0; s < 4; ++s) c.mb.phi_p[v][s] = P.phi_p[v][s]; }
    memset(&c.F, 0, sizeof(c.F));
    DDTables& tb = c.F.tab;
    tb.nterms = P.nterms; tb.nx = N + 1; tb.ny = M + 1; tb.nprof = 1;
    if (P.mode == DD_FORCING_SEPARABLE) {
        dd_prepare_tables(P.nterms, N, M, P.X, P.Y, P.XQ, P.YQ, &c.ht);
        tb.nprof = c.ht.nprof;
        for (int v = 0; v < 5; ++v) tb.var_prof[v] = c.ht.var_prof[v];
        for (int p = 0; p < c.ht.nprof; ++p)
            for (int d = 0; d < 3; ++d) { tb.X[p][d] = c.ht.X[p][d].data(); tb.Y[p][d] = c.ht.Y[p][d].data(); }
        tb.QX1 = c.ht.QX1.data(); tb.QY1 = c.ht.QY1.data(); tb.QX2 = c.ht.QX2.data(); tb.QY2 = c.ht.QY2.data();
        tb.QX3 = c.ht.QX3.data(); tb.QY3 = c.ht.QY3.data();
    } else if (P.mode == DD_FORCING_EXPSIN) {
        tb.X[0][0] = P.X[0][0]; tb.X[0][1] = P.X[0][1]; tb.Y[0][0] = P.Y[0][0]; tb.Y[0][1] = P.Y[0][1];
        tb.XQ0 = P.XQ[0]; tb.YQ0 = P.YQ[0];
    }
    for (int v = 0; v < 5; ++v) for (int s = 0; s < 2; ++s) c.F.arr.f[v][s] = P.farr[v][s];
    dd_time_coefs(P.mode, c.mb, t0, 0, &c.mb.tc[0]);
    dd_time_coefs(P.mode, c.mb, t0 + dt, 1, &c.mb.tc[1]);
}

#define HS_MODE(mode, CALL)                                                            \
    switch (mode) {                                                                    \
        case DD_FORCING_NONE: { constexpr int MODE = DD_FORCING_NONE; CALL; } break;   \
        case DD_FORCING_ARRAYS: { constexpr int MODE = DD_FORCING_ARRAYS; CALL; } break; \
        case DD_FORCING_SEPARABLE: { constexpr int MODE = DD_FORCING_SEPARABLE; CALL; } break; \
        default: { constexpr int MODE = DD_FORCING_EXPSIN; CALL; } break;              \
    }

// global red-black SOR on the Jacobi-scaled rows; returns the max residual
static double rbsor(const DDGeom& g, const DDRows& R, std::vector<double>& x, double rho, int sweeps) {
    const int ld = g.ld;
    double omega = 1.0;
    if (rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - rho * rho));
    std::fill(x.begin(), x.end(), 0.0);
    auto nb = [&](int i, int j) { return (i < 0 || i > g.N || j < 0 || j > g.M) ? 0.0 : x[(size_t)i * ld + j]; };
    for (int s = 0; s < sweeps; ++s)
        for (int colour = 0; colour < 2; ++colour)
            for (int i = 0; i <= g.N; ++i)
                for (int j = 0; j <= g.M; ++j) {
                    if (((i + j) & 1) != colour) continue;
                    const size_t p = (size_t)i * ld + j;
                    const double gs = R.bb[p] + R.aW[p] * nb(i - 1, j) + R.aE[p] * nb(i + 1, j) + R.aS[p] * nb(i, j - 1) +
                                      R.aN[p] * nb(i, j + 1);
                    x[p] = x[p] + omega * (gs - x[p]);
                }
    double res = 0.0;
    for (int i = 0; i <= g.N; ++i)
        for (int j = 0; j <= g.M; ++j) {
            const size_t p = (size_t)i * ld + j;
            const double r = R.bb[p] + R.aW[p] * nb(i - 1, j) + R.aE[p] * nb(i + 1, j) + R.aS[p] * nb(i, j - 1) +
                             R.aN[p] * nb(i, j + 1) - x[p];
            res = fmax(res, fabs(r));
        }
    return res;
}

extern "C" int hs_fields(const HSProblem* P, const double* const in[5], double* const out[5], double t) {
    HSCtx c;
    setup(*P, c, t, 1.0);
    DDStateC s;
    for (int v = 0; v < 5; ++v) s.v[v] = in[v];
    for (int r = 0; r <= P->N; ++r)
        for (int j = 0; j <= P->M; ++j) {
            double Fv[5];
            HS_MODE(P->mode, (dd_node_F<MODE>(c.g, c.mb, c.F, s, 0, r, j, 0, Fv)));
            for (int v = 0; v < 5; ++v) out[v][(size_t)r * c.g.ld + j] = Fv[v];
        }
    return 0;
}

extern "C" int hs_feuler(const HSProblem* P, const double* const in[5], double* const out[5], double t0, double dt) {
    HSCtx c;
    setup(*P, c, t0, dt);
    DDStateC s;
    DDState o;
    for (int v = 0; v < 5; ++v) { s.v[v] = in[v]; o.v[v] = out[v]; }
    for (int r = 0; r <= P->N; ++r)
        for (int j = 0; j <= P->M; ++j) HS_MODE(P->mode, (dd_node_feuler<MODE>(c.g, c.mb, c.F, s, o, 0, r, j)));
    return 0;
}

extern "C" int hs_exact(const HSProblem* P, double* const out[5], double t) {
    HSCtx c;
    setup(*P, c, t, 1.0);
    for (int r = 0; r <= P->N; ++r)
        for (int j = 0; j <= P->M; ++j) {
            double u[5];
            if (P->mode == DD_FORCING_SEPARABLE) dd_exact_values<DD_FORCING_SEPARABLE>(c.F, c.mb, 0, r, j, u);
            else dd_exact_values<DD_FORCING_EXPSIN>(c.F, c.mb, 0, r, j, u);
            for (int v = 0; v < 5; ++v) out[v][(size_t)r * c.g.ld + j] = u[v];
        }
    return 0;
}

// one predictor-corrector step; info[0..2] = rho of the last T/cl/cd systems, info[3..5] = residuals,
// cs_iters[pc] = Newton iterations used by each cs corrector call
extern "C" int hs_pc_step(const HSProblem* P, const double* const in[5], double* const out[5], double t0, double dt,
                          int npc, int nnewton, int cap, double rtol, int swap, int sweeps, double* info,
                          int* cs_iters) {
    HSCtx c;
    setup(*P, c, t0, dt);
    const DDGeom& g = c.g;
    const size_t n = (size_t)(P->N + 1) * (P->M + 1);
    std::vector<double> cp1p(n), cs1p(n), YT(n), Ycl(n), Ycd(n), bb(n), aW(n), aE(n), aS(n), aN(n), x(n);
    std::vector<double> Tn[2] = {std::vector<double>(n), std::vector<double>(n)};
    std::vector<double> cln[2] = {std::vector<double>(n), std::vector<double>(n)};
    std::vector<double> cdn[2] = {std::vector<double>(n), std::vector<double>(n)};
    DDStateC s0;
    for (int v = 0; v < 5; ++v) s0.v[v] = in[v];
    DDPredictOut po = {cp1p.data(), cs1p.data(), YT.data(), Ycl.data(), Ycd.data()};
    for (int r = 0; r <= P->N; ++r)
        for (int j = 0; j <= P->M; ++j) HS_MODE(P->mode, (dd_node_predict<MODE>(g, c.mb, c.F, s0, po, 0, r, j)));
    DDStateC u;
    u.v[DD_CP] = cp1p.data(); u.v[DD_T] = in[DD_T]; u.v[DD_CL] = in[DD_CL]; u.v[DD_CD] = in[DD_CD]; u.v[DD_CS] = cs1p.data();
    DDRows R = {bb.data(), aW.data(), aE.data(), aS.data(), aN.data(), g.ld, g.mstride};
    int pp = 0;
    std::vector<double> cpc(n), csc(n);
    for (int pc = 0; pc < npc; ++pc) {
        for (int nw = 0; nw < nnewton; ++nw) {
            double* dst[3] = {Tn[pp].data(), cln[pp].data(), cdn[pp].data()};
            for (int q = 0; q < 3; ++q) {
                const int var = DD_T + q;
                double rho = 0.0;
                for (int r = 0; r <= P->N; ++r)
                    for (int j = 0; j <= P->M; ++j) {
                        double rr = 0.0;
                        if (var == DD_T) { HS_MODE(P->mode, (rr = dd_node_asm_T<MODE>(g, c.mb, c.F, u, YT.data(), R, 0, 0, r, j))); }
                        else if (var == DD_CL) { HS_MODE(P->mode, (rr = dd_node_asm_cl<MODE>(g, c.mb, c.F, u, dst[0], Ycl.data(), R, 0, 0, r, j))); }
                        else { HS_MODE(P->mode, (rr = dd_node_asm_cd<MODE>(g, c.mb, c.F, u, dst[0], dst[1], Ycd.data(), swap, R, 0, 0, r, j))); }
                        rho = fmax(rho, rr);
                    }
                const double res = rbsor(g, R, x, rho, sweeps);
                if (info) { info[q] = rho; info[3 + q] = res; }
                for (int r = 0; r <= P->N; ++r)
                    for (int j = 0; j <= P->M; ++j) {
                        const size_t p = (size_t)r * g.ld + j;
                        dst[q][p] = dd_newton_update(dd_is_interior(g, r, j), u.v[var][p], x[p], var == DD_T ? 1 : 0);
                    }
            }
            u.v[DD_T] = dst[0]; u.v[DD_CL] = dst[1]; u.v[DD_CD] = dst[2];
            pp ^= 1;
        }
        // correctors
        std::vector<double> xs(n), ys(n), as(n);
        for (int r = 0; r <= P->N; ++r)
            for (int j = 0; j <= P->M; ++j) {
                const size_t p = (size_t)r * g.ld + j;
                double cp1, yy, aa;
                HS_MODE(P->mode, (dd_node_correct_prepare<MODE>(g, c.mb, c.F, s0, u.v[DD_T], u.v[DD_CL], u.v[DD_CD], 0, r, j, &cp1, &yy, &aa)));
                cpc[p] = cp1; ys[p] = yy; as[p] = aa; xs[p] = in[DD_CS][p];
            }
        int used = 0;
        const bool closed = c.mb.m.react != DD_REACT_REGH;
        if (closed) {
            // CsTriple / HCsTriple: closed-form corrector (sources read again through the prepare routine)
            for (int r = 0; r <= P->N; ++r)
                for (int j = 0; j <= P->M; ++j) {
                    const size_t p = (size_t)r * g.ld + j;
                    double cp1, yy, aa, f0, f1;
                    HS_MODE(P->mode, (dd_node_correct_prepare<MODE>(g, c.mb, c.F, s0, u.v[DD_T], u.v[DD_CL], u.v[DD_CD], 0, r, j, &cp1, &yy, &aa, &f0, &f1)));
                    int bad = 0;
                    xs[p] = dd_node_correct_cs_closed(g, c.mb, s0, u.v[DD_CL], u.v[DD_CD], 0, r, j, f0, f1, &bad);
                    if (bad) return -5;
                }
        }
        for (int it = 0; it < (closed ? 0 : cap); ++it) {
            double mx = 0.0, mn = INFINITY;
            bool nan = false;
            for (size_t p = 0; p < n; ++p) {
                const double dx = dd_cs_newton_dx(xs[p], ys[p], as[p], c.mb.m.eta);
                xs[p] += dx;
                if (dx != dx) nan = true;
                mx = fmax(mx, fabs(dx));
                double ax = fabs(xs[p]);
                if (ax != ax) ax = 0.0;
                mn = fmin(mn, ax);
            }
            used = it + 1;
            if (rtol > 0.0 && !nan && mx < rtol * mn) break;
        }
        if (cs_iters) cs_iters[pc] = used;
        for (int r = 0; r <= P->N; ++r)
            for (int j = 0; j <= P->M; ++j) {
                const size_t p = (size_t)r * g.ld + j;
                csc[p] = closed ? xs[p] : xs[p] * (dd_is_interior(g, r, j) ? 1.0 : 0.0);
            }
        cp1p = cpc; cs1p = csc;
        u.v[DD_CP] = cp1p.data(); u.v[DD_CS] = cs1p.data();
    }
    for (size_t p = 0; p < n; ++p) {
        out[DD_CP][p] = cp1p[p]; out[DD_T][p] = u.v[DD_T][p]; out[DD_CL][p] = u.v[DD_CL][p];
        out[DD_CD][p] = u.v[DD_CD][p]; out[DD_CS][p] = cs1p[p];
    }
    return 0;
}


// combined error norms of B members from a norm series [K][B][8] (the arithmetic of k_combine_update /
// k_combine_final): out[B][6]
extern "C" int hs_combine(const double* series, int K, int B, const double* dt, int n_t, double* out) {
    std::vector<double> st(18);
    for (int m = 0; m < B; ++m) {
        for (int k = 0; k < K; ++k)
            dd_combine_fold(series + ((size_t)k * B + m) * 8, dt[n_t == 1 ? 0 : m], k == 0, st.data());
        for (int q = 0; q < 6; ++q) out[(size_t)m * 6 + q] = sqrt(st[q]);
    }
    return 0;
}
