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

#include "dd_lane.cuh"
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
    constexpr int W = 64 * C, PWP = 32 * C + 2;
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
    A.halo = 2 * P.sweeps + 2;
    A.last_pass = P.last_pass;
    A.tj = W - 2 * A.halo;
    A.nstrips = (P.M + 1 + A.tj - 1) / A.tj;
    A.nwo = P.nwarps;
    A.flat_total = (long long)A.nstrips * (P.own1 - P.own0);
    A.flat_per_cta = (A.flat_total + P.nctas - 1) / P.nctas;
    A.rho_fix = -1.0;
    const int nwo = P.nwarps, D = 2 * nwo, S4 = 4 * P.sweeps;
    const int nthreads = 32 * (nwo + DD_WAVE_NE);
    std::vector<double> smem(dd_wave_smem_doubles(CB, C, nwo));
    std::vector<WaveRegs<CB, C>> regs(32 * nwo);
    std::vector<WaveEpi> epi(32 * DD_WAVE_NE);
    unsigned hr = 0, hx = 0, hv = 0, hb = 0;
    P.steps = 0;
    const size_t nx = (size_t)D * 2 * PWP;
    for (int cta = 0; cta < P.nctas; ++cta) {
        WaveSmem sm;
        dd_wave_smem_carve(sm, smem.data(), 0u, CB, C, nwo);
        long long f0 = (long long)cta * A.flat_per_cta;
        const long long f1 = f0 + A.flat_per_cta < A.flat_total ? f0 + A.flat_per_cta : A.flat_total;
        while (f0 < f1) {
            const WaveSeg sg = dd_wave_segment(A, f0, f1);
            f0 += sg.r1 - sg.r0;
            for (size_t k = 0; k < nx; ++k) smem[k] = 0.0;
            const double fT = mb.dt * mb.m.DT;
            if (CB)
                for (int sj = 0; sj < W; ++sj) {
                    const int j = sg.cbase + sj;
                    const bool in = j >= 1 && j <= P.M - 1;
                    smem[nx + sj] = in ? fT * P.rkp[j] * P.rk[j] : 0.0;
                    smem[nx + W + sj] = in ? fT * P.rkp[j] * P.rk[j + 1] : 0.0;
                }
            double omega = 1.0;
            if (P.rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - P.rho * P.rho));
            for (int t = 0; t < 32 * nwo; ++t) dd_wave_init_thread<CB, C>(A, sg, regs[t], sm, t >> 5, t & 31);
            for (int t = 0; t < 32 * DD_WAVE_NE; ++t) dd_wave_epi_init(A, sg, epi[t], sm, t >> 5, t & 31, C);
            // poison the staging rings: a value nobody requested must show up
            for (size_t k = nx + 2 * W; k < smem.size(); ++k) smem[k] = NAN;
            const int nsteps = dd_wave_steps(A, sg);
            P.steps += nsteps;
            for (int s = 0; s < nsteps; ++s)
                for (int tt = 0; tt < nthreads; ++tt) {
                    const int t = P.order ? nthreads - 1 - tt : tt;
                    if (t < 32 * nwo)
                        dd_wave_thread_step<CB, C>(A, sg, regs[t], sm, t & 31, D, S4, omega, fT);
                    else
                        dd_wave_epi_step<C>(A, sg, epi[t - 32 * nwo], sm, s - 2 - DD_WAVE_LS - S4);
                }
            for (int t = 0; t < 32 * nwo; ++t) {
                hr = regs[t].hr > hr ? regs[t].hr : hr;
                hb = regs[t].hb > hb ? regs[t].hb : hb;
            }
            for (auto& e : epi) {
                hx = e.hx > hx ? e.hx : hx;
                hv = e.hv > hv ? e.hv : hv;
            }
        }
    }
    P.stats[0] = dd_wave_from_hi(hr, true); P.stats[1] = dd_wave_from_hi(hx, false);
    P.stats[2] = dd_wave_from_hi(hv, false); P.stats[3] = dd_wave_from_hi(hb, false);
}

extern "C" int ws_wave(WSProblem* P) {
    if (2 * P->nwarps < 4 * P->sweeps + 4) return 2;
#define WS_CASE(CB_, C_) if (P->cb == CB_ && P->C == C_) { run<CB_, C_>(*P); return 0; }
    WS_CASE(1, 4) WS_CASE(1, 3) WS_CASE(1, 2) WS_CASE(1, 1) WS_CASE(0, 2) WS_CASE(0, 1)
#undef WS_CASE
    return 1;
}

// ---- lane-private marching kernel (csrc/dd_lane.cuh): the 32 lanes of a warp stepped one after the other; what a
// lane receives by shuffle on the device is read here from its neighbour's registers BEFORE anybody's step ---------
template <int CB, int S, int XIN, int U>
static void lane_dispatch(int u, const WaveArgs& A, const WaveSeg& sg, std::vector<LaneRegs<CB, S, XIN>>& regs,
                          const LaneSmem& sm, int tau, double omega, double fT, int order) {
    if constexpr (U < 2 * S + 4) {
        if (u != U) {
            lane_dispatch<CB, S, XIN, U + 1>(u, A, sg, regs, sm, tau, omega, fT, order);
            return;
        }
        constexpr int o = (U + 1) & 1;
        double src[32][2 * S + 1], nb[32][2 * S + 1];
        for (int l = 0; l < 32; ++l) dd_lane_offer<CB, S, XIN, U>(regs[l], src[l]);
        for (int l = 0; l < 32; ++l) {
            // __shfl_down / __shfl_up by one lane: the edge lane gets its own value back
            const int from = o ? (l < 31 ? l + 1 : l) : (l > 0 ? l - 1 : l);
            for (int k = 0; k <= 2 * S; ++k) nb[l][k] = src[from][k];
        }
        for (int ll = 0; ll < 32; ++ll) {
            const int l = order ? 31 - ll : ll;
            dd_lane_step<CB, S, XIN, U>(A, sg, regs[l], sm, tau, l, omega, fT, nb[l]);
        }
    }
}

template <int CB, int S, int XIN>
static void run_lane(WSProblem& P) {
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
    A.sweeps = S;
    A.halo = dd_lane_halo(S, XIN);
    A.last_pass = P.last_pass;
    A.tj = 64 - 2 * A.halo;
    A.nstrips = (P.M + 1 + A.tj - 1) / A.tj;
    A.flat_total = (long long)A.nstrips * (P.own1 - P.own0);
    A.flat_per_cta = (A.flat_total + P.nctas - 1) / P.nctas;
    A.rho_fix = -1.0;
    A.nwo = P.nwarps > 1 ? P.nwarps : 0;  // rows per segment of the flat order (0: strip-major)
    constexpr int PP = 2 * S + 4;
    std::vector<double> ring(dd_lane_ring_doubles(CB, S, XIN));
    std::vector<LaneRegs<CB, S, XIN>> regs(32);
    unsigned hr = 0, hx = 0, hv = 0, hb = 0;
    P.steps = 0;
    const double fT = mb.dt * mb.m.DT;
    double omega = 1.0;
    if (P.rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - P.rho * P.rho));
    for (int cta = 0; cta < P.nctas; ++cta) {
        LaneSmem sm;
        sm.base = ring.data();
        long long f0 = (long long)cta * A.flat_per_cta;
        const long long f1 = f0 + A.flat_per_cta < A.flat_total ? f0 + A.flat_per_cta : A.flat_total;
        while (f0 < f1) {
            const WaveSeg sg = dd_lane_segment(A, f0, f1, dd_lane_warmup(S, XIN));
            f0 += sg.r1 - sg.r0;
            for (double& v : ring) v = NAN;  // a march must not depend on what the previous one left behind
            for (int l = 0; l < 32; ++l) dd_lane_init<CB, S, XIN>(A, sg, regs[l], sm, 0u, l, fT);
            for (int q = 0; q < DD_LANE_LS; ++q)
                for (int l = 0; l < 32; ++l) dd_lane_request<CB, S, XIN>(A, sg, regs[l], sm, q, l, q);
            const int nsteps = dd_lane_steps(A, sg);
            P.steps += nsteps;
            for (int tau = 0; tau < nsteps; ++tau)
                lane_dispatch<CB, S, XIN, 0>(tau % PP, A, sg, regs, sm, tau, omega, fT, P.order);
            for (auto& r : regs) {
                hr = r.hr > hr ? r.hr : hr;
                hb = r.hb > hb ? r.hb : hb;
                hx = r.hx > hx ? r.hx : hx;
                hv = r.hv > hv ? r.hv : hv;
            }
        }
    }
    P.stats[0] = dd_wave_from_hi(hr, true); P.stats[1] = dd_wave_from_hi(hx, false);
    P.stats[2] = dd_wave_from_hi(hv, false); P.stats[3] = dd_wave_from_hi(hb, false);
}

extern "C" int ws_lane(WSProblem* P) {
#define WS_LANE(CB_, S_)                                                                  \
    if (P->cb == CB_ && P->sweeps == S_) {                                                \
        if (P->xin) run_lane<CB_, S_, 1>(*P); else run_lane<CB_, S_, 0>(*P);              \
        return 0;                                                                         \
    }
    WS_LANE(1, 1) WS_LANE(1, 2) WS_LANE(1, 3) WS_LANE(1, 4) WS_LANE(1, 5)
    WS_LANE(0, 1) WS_LANE(0, 2) WS_LANE(0, 3) WS_LANE(0, 4) WS_LANE(0, 5)
#undef WS_LANE
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
            return dd_sor_dT(P->bb[p], P->aW[p], rW, rE, cS, cN, nb(i - 1, j), nb(i + 1, j), nb(i, j - 1), nb(i, j + 1),
                             x[p]);
        }
        return dd_sor_d5(P->bb[p], P->aW[p], P->aE[p], P->aS[p], P->aN[p], nb(i - 1, j), nb(i + 1, j), nb(i, j - 1),
                         nb(i, j + 1), x[p]);
    };
    for (int s = 0; s < P->sweeps; ++s)
        for (int colour = 0; colour < 2; ++colour)
            for (int i = P->vr0; i < P->vr1; ++i)
                for (int j = 0; j <= M; ++j) {
                    if (((P->row0 + i + j) & 1) != colour) continue;
                    const size_t p = (size_t)i * ldR + j;
                    x[p] = dd_sor_relax(x[p], gs(i, j), omega);
                }
    unsigned hr = 0, hx = 0, hv = 0, hb = 0;
    auto up = [](unsigned& m, double v) { const unsigned h = dd_wave_hi(v); m = h > m ? h : m; };
    if (P->vnew && P->last_pass)
        for (int i = P->own0; i < P->own1; ++i)
            for (int j = 0; j <= M; ++j) {
                const size_t p = (size_t)i * ldR + j;
                const int gi = P->row0 + i;
                const bool inter = gi > 0 && gi < P->N && j > 0 && j < M;
                const double vn = dd_newton_update(inter, P->vstar[(size_t)i * P->ld + j], x[p], P->zero_boundary);
                P->vnew[(size_t)i * P->ld + j] = vn;
                up(hr, gs(i, j));
                up(hx, x[p]);
                up(hv, vn);
                up(hb, P->bb[p]);
            }
    P->stats[0] = dd_wave_from_hi(hr, true); P->stats[1] = dd_wave_from_hi(hx, false);
    P->stats[2] = dd_wave_from_hi(hv, false); P->stats[3] = dd_wave_from_hi(hb, false);
    return 0;
}
