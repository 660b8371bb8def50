// wavesim.cpp -- TEST ONLY: host emulation of the wavefront SOR kernel (csrc/dd_wave.cuh).
//
// The kernel's per-thread program (dd_wave_thread_step) is compiled for the host and run thread by thread
// between the barriers, CTA by CTA, exactly as dd_wave.cu drives it on the device; ws_reference is the plain
// global red-black SOR with the same dd_sor_* arithmetic.  tests/test_wave_hostsim.py compares the two bit for bit.
// `order` = 1 runs the threads of a step in reverse order: equal results show that no thread reads what another
// thread writes in the same step.
#include <math.h>
#include <stdlib.h>
#include <string.h>

#include <vector>

#include "dd_wave.cuh"

struct WSProblem {
    int N, M, row0, nrows, ld, ldR;
    int own0, own1, vr0, vr1;
    int cb, C, nwarps, nctas, sweeps, last_pass, zero_boundary, order;
    double rho, dt, DT;
    const double *rh, *rhp, *rk, *rkp;             // global 1-D metrics
    const double *bb, *aW, *aE, *aS, *aN;          // local rows, pitch ldR
    const double *xin, *vstar;                     // xin pitch ldR (nullable), vstar pitch ld
    double *xout, *vnew;                           // pitch ldR / ld
    double stats[4];                               // resid, xmax, vmax, bmax
    long long steps;                               // CTA time steps summed (cost model)
};

template <int CB, int C>
static void run(WSProblem& P) {
    constexpr int W = 64 * C;
    DDMember mb;
    memset(&mb, 0, sizeof(mb));
    mb.active = 1;
    mb.dt = P.dt;
    mb.m.DT = P.DT;
    DDSolveStats st;
    memset(&st, 0, sizeof(st));
    st.rho = P.rho;
    WaveArgs A;
    memset(&A, 0, sizeof(A));
    A.g.N = P.N; A.g.M = P.M; A.g.row0 = P.row0; A.g.nrows = P.nrows; A.g.ld = P.ld;
    A.g.mstride = (long long)P.nrows * P.ld;
    A.g.rh = P.rh; A.g.rhp = P.rhp; A.g.rk = P.rk; A.g.rkp = P.rkp;
    A.mem = &mb;
    A.bb = P.bb; A.aW = P.aW; A.aE = P.aE; A.aS = P.aS; A.aN = P.aN;
    A.xin = P.xin; A.xout = P.xout; A.vstar = P.vstar; A.vnew = P.vnew;
    A.stats = &st;
    A.zero_boundary = P.zero_boundary;
    A.ldR = P.ldR;
    A.mstrideR = (long long)P.nrows * P.ldR;
    A.own0 = P.own0; A.own1 = P.own1; A.vr0 = P.vr0; A.vr1 = P.vr1;
    A.sweeps = P.sweeps;
    A.halo = 2 * P.sweeps + 1;
    A.last_pass = P.last_pass;
    A.tj = W - 2 * A.halo - 2;
    A.nstrips = (P.M + 1 + A.tj - 1) / A.tj;
    A.flat_total = (long long)A.nstrips * (P.own1 - P.own0);
    A.flat_per_cta = (A.flat_total + P.nctas - 1) / P.nctas;
    A.rho_fix = -1.0;
    const int nwarps = P.nwarps, nthreads = 32 * nwarps;
    std::vector<double> smem(dd_wave_smem_doubles(C, nwarps));
    std::vector<WaveRegs<CB, C>> regs(nthreads);
    double rmax = 0, xmax = 0, vmax = 0, bmax = 0;
    P.steps = 0;
    for (int cta = 0; cta < P.nctas; ++cta) {
        WaveSmem sm;
        sm.x = smem.data();
        sm.vs = sm.x + (size_t)2 * nwarps * W;
        sm.scol = sm.vs + DD_WAVE_VS * W;
        long long f0 = (long long)cta * A.flat_per_cta;
        const long long f1 = f0 + A.flat_per_cta < A.flat_total ? f0 + A.flat_per_cta : A.flat_total;
        while (f0 < f1) {
            const WaveSeg sg = dd_wave_segment(A, f0, f1);
            f0 += sg.r1 - sg.r0;
            for (size_t k = 0; k < (size_t)2 * nwarps * W; ++k) sm.x[k] = 0.0;
            // poison the staging ring: a finish that reads a value nobody requested must show up
            for (int k = 0; k < DD_WAVE_VS * W; ++k) sm.vs[k] = NAN;
            const double fT = mb.dt * mb.m.DT;
            if (CB)
                for (int sj = 0; sj < W; ++sj) {
                    const int j = sg.cbase + sj;
                    const bool in = j >= 1 && j <= P.M - 1;
                    sm.scol[sj] = in ? fT * P.rkp[j] * P.rk[j] : 0.0;
                    sm.scol[W + sj] = in ? fT * P.rkp[j] * P.rk[j + 1] : 0.0;
                }
            double omega = 1.0;
            if (P.rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - P.rho * P.rho));
            for (int t = 0; t < nthreads; ++t) dd_wave_init_thread<CB, C>(regs[t], t >> 5);
            const int nsteps = dd_wave_steps(A, sg);
            P.steps += nsteps;
            for (int s = 0; s < nsteps; ++s)
                for (int tt = 0; tt < nthreads; ++tt) {
                    const int t = P.order ? nthreads - 1 - tt : tt;
                    dd_wave_thread_step<CB, C>(A, sg, regs[t], sm, t >> 5, t & 31, nwarps, omega, fT);
                }
            for (int t = 0; t < nthreads; ++t) {
                rmax = dd_nn_max(rmax, regs[t].rmax);
                xmax = dd_nn_max(xmax, regs[t].xmax);
                vmax = dd_nn_max(vmax, regs[t].vmax);
                bmax = dd_nn_max(bmax, regs[t].bmax);
            }
        }
    }
    P.stats[0] = rmax; P.stats[1] = xmax; P.stats[2] = vmax; P.stats[3] = bmax;
}

extern "C" int ws_wave(WSProblem* P) {
    if (2 * P->nwarps < 4 * P->sweeps + 4) return 2;
#define WS_CASE(CB_, C_) if (P->cb == CB_ && P->C == C_) { run<CB_, C_>(*P); return 0; }
    WS_CASE(1, 4) WS_CASE(1, 3) WS_CASE(1, 2) WS_CASE(1, 1) WS_CASE(0, 2) WS_CASE(0, 1)
#undef WS_CASE
    return 1;
}

// global red-black SOR on local rows [vr0, vr1) (zero outside), colour = parity of the global i + j;
// x has pitch ldR, starts from xin (or 0); writes x and, when vnew is given, v_new and the statistics
extern "C" int ws_reference(WSProblem* P, double* x) {
    const int ldR = P->ldR, nrows = P->nrows, M = P->M;
    double omega = 1.0;
    if (P->rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - P->rho * P->rho));
    const double fT = P->dt * P->DT;
    for (long long k = 0; k < (long long)nrows * ldR; ++k) x[k] = 0.0;
    if (P->xin)
        for (int i = P->vr0; i < P->vr1; ++i)
            for (int j = 0; j <= M; ++j) x[(size_t)i * ldR + j] = P->xin[(size_t)i * ldR + j];
    auto nb = [&](int i, int j) { return (i < P->vr0 || i >= P->vr1 || j < 0 || j > M) ? 0.0 : x[(size_t)i * ldR + j]; };
    auto gs = [&](int i, int j) {
        const size_t p = (size_t)i * ldR + j;
        const int gi = P->row0 + i;
        if (P->cb) {
            const bool ri = gi >= 1 && gi <= P->N - 1, cj = j >= 1 && j <= M - 1;
            const double rW = ri ? fT * P->rhp[gi] * P->rh[gi] : 0.0, rE = ri ? fT * P->rhp[gi] * P->rh[gi + 1] : 0.0;
            const double cS = cj ? fT * P->rkp[j] * P->rk[j] : 0.0, cN = cj ? fT * P->rkp[j] * P->rk[j + 1] : 0.0;
            return dd_sor_gsT(P->bb[p], P->aW[p], rW, rE, cS, cN, nb(i - 1, j), nb(i + 1, j), nb(i, j - 1), nb(i, j + 1));
        }
        return dd_sor_gs5(P->bb[p], P->aW[p], P->aE[p], P->aS[p], P->aN[p], nb(i - 1, j), nb(i + 1, j), nb(i, j - 1),
                          nb(i, j + 1));
    };
    for (int s = 0; s < P->sweeps; ++s)
        for (int colour = 0; colour < 2; ++colour)
            for (int i = P->vr0; i < P->vr1; ++i)
                for (int j = 0; j <= M; ++j) {
                    if (((P->row0 + i + j) & 1) != colour) continue;
                    const size_t p = (size_t)i * ldR + j;
                    x[p] = dd_sor_relax(x[p], gs(i, j), omega);
                }
    double rmax = 0, xmax = 0, vmax = 0, bmax = 0;
    if (P->vnew)
        for (int i = P->own0; i < P->own1; ++i)
            for (int j = 0; j <= M; ++j) {
                const size_t p = (size_t)i * ldR + j;
                const int gi = P->row0 + i;
                const bool inter = gi > 0 && gi < P->N && j > 0 && j < M;
                const double vn = dd_newton_update(inter, P->vstar[(size_t)i * P->ld + j], x[p], P->zero_boundary);
                P->vnew[(size_t)i * P->ld + j] = vn;
                rmax = dd_nn_max(rmax, gs(i, j) - x[p]);
                xmax = dd_nn_max(xmax, x[p]);
                vmax = dd_nn_max(vmax, vn);
                bmax = dd_nn_max(bmax, P->bb[p]);
            }
    P->stats[0] = rmax; P->stats[1] = xmax; P->stats[2] = vmax; P->stats[3] = bmax;
    return 0;
}
