// dd_lane.cuh -- lane-private marching form of the red-black SOR solve (one lane's share of one time step).
//
// A WARP owns a strip of 64 columns (lane l: columns 2 l and 2 l + 1 of it) and marches down the rows on its own:
// no barrier, no data shared between warps, the only exchange between lanes is one shuffle per relaxation.  In
// time step tau row tau enters, and half-sweep (level) h = 1 .. 2 S is applied to row tau - h, in this order: row
// q at level h needs rows q + 1 (level h - 1, relaxed a moment ago in this very step), q - 1 (step tau - 2) and its
// own other colour (step tau - 1).  Row tau - 2 S - 1 leaves as v_new = v* + x.  Because a cell's colour is the
// parity of i + j and level h relaxes colour (h + 1) & 1, EVERY level of a step relaxes the same column of the
// lane's pair, o = (tau + 1) & 1: the structure of a step is static once the loop is unrolled over an even period.
//
// Where things live: the iterate of the 2 S + 3 rows in flight in REGISTERS (X[column parity][row slot], slots
// cycled with period P = 2 S + 4, all indices compile-time after unrolling P steps); the coefficients in a ring of
// P row slots in SHARED memory that is private to the lane (each lane reads back only what its own cp.async
// wrote: 16-byte copies of its column pair, requested LS rows ahead, zero-filled outside the grid), read with one
// 16-byte load per array at the odd levels, the half that belongs to the other column being carried in registers
// to the next step, where the same row is relaxed at the next (even) level.  Per relaxation: 2.5 shared loads,
// one shuffle, 6 fp64 instructions.
//
// Redundancy: 2 S + 2 columns on either side of a strip and 2 S + 1 rows before and after a march hold data that
// goes wrong one cell per half-sweep and never reaches an owned cell (same argument as for the tiles and the
// wavefront kernel; one more than 2 S because the residual looks one cell further).  Results are bit for bit those
// of the global red-black iteration with the arithmetic of dd_sor.cuh, whatever the strips, marches and slabs.
//
// This header is compiled for the device (dd_lane.cu) and, test only, for the host (tests/hostsim), where the 32
// lanes of a warp are stepped one after the other.
#pragma once

#include "dd_wave.cuh"

#define DD_LANE_LS 2  // a row's coefficients are requested this many steps before the row enters
// ... and prefetched into L2 this many rows before that.  Measured on B200 (T / cl / cd solves of the bench mesh, ms):
// no prefetch 0.217 / 0.230 / 0.171, 1 row 0.183 / 0.235 / 0.141, 2 rows 0.184 / 0.234 / 0.137, 3 rows 0.188 / 0.231 /
// 0.137, 6 rows 0.199 / 0.233 / 0.145, 12 rows 0.235 / 0.247 / 0.200: the lines must arrive shortly before the
// cp.async asks for them, earlier ones are evicted again by the stream of the other arrays.
#ifndef DD_LANE_PF
#define DD_LANE_PF ((S) >= 4 ? 3 : 2)
#endif

#define DD_LANE_NA(CB, XIN) (((CB) ? 2 : 5) + (XIN))  // staged arrays: bb, dinv | bb, aW, aE, aS, aN; + x of the previous pass
#define DD_LANE_VS 4  // slots of the v* ring (power of two > LS)

template <int CB, int S, int XIN>
struct LaneRegs {
    static constexpr int P = 2 * S + 4, NC = CB ? 2 : 5;
    double X[2][P];       // iterate: [column of the pair][row slot]
    double carry[S][NC];  // coefficients of the pair's other column, from the odd level to the next step's even level
    double RW[CB ? P : 1], RE[CB ? P : 1];  // const band: row factors dt DT / (hhat_i h_i), dt DT / (hhat_i h_{i+1})
    double kS[2], kN[2];  // const band: column factors of the lane's two columns
    unsigned hr, hb, hx, hv;  // high words of max |residual|, |bb|, |x|, |v_new| over its owned cells
    unsigned own;  // bit p: column p of the pair is an owned column of the strip; bit 2 + p: an interior column
    unsigned ring; // byte address of the lane's pair of (slot 0, array 0): shared window (device), offset (host)
};

struct LaneSmem {
    double* base;  // generic pointer to the warp's ring (host: the emulated array)
};

// the warp's shared memory: coefficient ring [P][NA][32 lanes][2] | v* ring [VS][32 lanes][2]
DD_HD size_t dd_lane_ring_doubles(int CB, int S, int XIN) {
    return (size_t)(2 * S + 4) * DD_LANE_NA(CB, XIN) * 64 + DD_LANE_VS * 64;
}
// columns of redundancy on either side of a strip / rows before and after a march.  From a zero iterate nothing is
// wrong before the second half-sweep, so the wrong data gets one cell less far.
DD_HD int dd_lane_halo(int S, int XIN) { return XIN ? 2 * S + 2 : 2 * S; }
DD_HD int dd_lane_warmup(int S, int XIN) { return XIN ? 2 * S + 1 : 2 * S; }

// one march (see dd_wave_segment): rows [r0, r1) of a strip, marched rows rs + q, q = 0 .. nq - 1, rs on an even
// global row.  The flat index runs over [member][row segment][strip][row of the segment] (A.nwo = rows per
// segment; 0: one segment): warps with neighbouring flat ranges march the SAME rows of NEIGHBOURING strips at about
// the same time, so the columns two strips share are read from HBM once and found in L2 the second time.
DD_HD WaveSeg dd_lane_segment(const WaveArgs& A, long long f0, long long f1, int wu) {
    WaveSeg s;
    const int R = A.own1 - A.own0;
    const int SR = A.nwo > 0 && A.nwo < R ? A.nwo : R;
    const int nseg = (R + SR - 1) / SR;
    const long long per_member = (long long)A.nstrips * R;
    s.member = (int)(f0 / per_member);
    const long long rem = f0 - (long long)s.member * per_member;
    const long long per_seg = (long long)SR * A.nstrips;
    int g = (int)(rem / per_seg);
    if (g > nseg - 1) g = nseg - 1;
    const long long rem2 = rem - (long long)g * per_seg;
    const int rows_g = g == nseg - 1 ? R - g * SR : SR;
    const int strip = (int)(rem2 / rows_g);
    const int r = (int)(rem2 - (long long)strip * rows_g);
    long long n = f1 - f0;
    if (n > rows_g - r) n = rows_g - r;
    s.r0 = A.own0 + g * SR + r;
    s.r1 = s.r0 + (int)n;
    s.c0 = strip * A.tj;
    s.tc = A.g.M + 1 - s.c0 < A.tj ? A.g.M + 1 - s.c0 : A.tj;
    s.cbase = s.c0 - A.halo;
    s.rs = s.r0 - wu > A.vr0 ? s.r0 - wu : A.vr0;
    s.rs -= (A.g.row0 + s.rs) & 1;
    const int re = s.r1 + wu < A.vr1 ? s.r1 + wu : A.vr1;
    s.nq = re - s.rs;
    s.mo = (long long)s.member * A.g.mstride;
    s.moR = (long long)s.member * A.mstrideR;
    return s;
}

#ifdef __CUDA_ARCH__
// 16-byte asynchronous copy global -> shared of which only the first n bytes are read (the rest is zero-filled)
__device__ __forceinline__ void dd_lane_cp16(const LaneSmem&, unsigned dst, const double* src, int n) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 16, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void dd_lane_cp8(const LaneSmem&, unsigned dst, const double* src, int n) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8, %2;" ::"r"(dst), "l"(src), "r"(n) : "memory");
}
__device__ __forceinline__ void dd_lane_lds2(const LaneSmem&, unsigned a, double& v0, double& v1) {
    asm volatile("ld.shared.v2.f64 {%0, %1}, [%2];" : "=d"(v0), "=d"(v1) : "r"(a));
}
__device__ __forceinline__ double dd_lane_lds(const LaneSmem&, unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void dd_lane_zero16(const LaneSmem&, unsigned a) {
    asm volatile("st.shared.v2.f64 [%0], {%1, %1};" ::"r"(a), "d"(0.0) : "memory");
}
__device__ __forceinline__ void dd_lane_prefetch(const double* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}
#define DD_LANE_COMMIT() asm volatile("cp.async.commit_group;" ::: "memory")
#define DD_LANE_WAIT() asm volatile("cp.async.wait_group %0;" ::"n"(DD_LANE_LS) : "memory")
#else
inline void dd_lane_cp16(const LaneSmem& sm, unsigned dst, const double* src, int n) {
    double* d = (double*)((char*)sm.base + dst);
    d[0] = n >= 8 ? src[0] : 0.0;
    d[1] = n >= 16 ? src[1] : 0.0;
}
inline void dd_lane_cp8(const LaneSmem& sm, unsigned dst, const double* src, int n) {
    *(double*)((char*)sm.base + dst) = n >= 8 ? src[0] : 0.0;
}
inline void dd_lane_lds2(const LaneSmem& sm, unsigned a, double& v0, double& v1) {
    const double* s = (const double*)((const char*)sm.base + a);
    v0 = s[0];
    v1 = s[1];
}
inline double dd_lane_lds(const LaneSmem& sm, unsigned a) { return *(const double*)((const char*)sm.base + a); }
inline void dd_lane_zero16(const LaneSmem& sm, unsigned a) {
    double* d = (double*)((char*)sm.base + a);
    d[0] = d[1] = 0.0;
}
inline void dd_lane_prefetch(const double*) {}
#define DD_LANE_COMMIT()
#define DD_LANE_WAIT()
#endif

// byte offset of the lane's pair of (row slot, array) from R.ring
#define DD_LANE_OFF(CB, XIN, slot, a) ((unsigned)((((slot) * DD_LANE_NA(CB, XIN)) + (a)) * 512))

template <int CB, int S, int XIN>
DD_HD void dd_lane_init(const WaveArgs& A, const WaveSeg& sg, LaneRegs<CB, S, XIN>& R, const LaneSmem& sm, unsigned ring0,
                        int lane, double fT) {
    constexpr int P = 2 * S + 4, NC = CB ? 2 : 5;
#pragma unroll
    for (int s = 0; s < P; ++s) {
        R.X[0][s] = R.X[1][s] = 0.0;
        if (CB) R.RW[s] = R.RE[s] = 0.0;
    }
#pragma unroll
    for (int k = 0; k < S; ++k)
#pragma unroll
        for (int a = 0; a < NC; ++a) R.carry[k][a] = 0.0;
    R.hr = R.hb = R.hx = R.hv = 0u;
    R.own = 0u;
    R.ring = ring0 + (unsigned)lane * 16u;
#pragma unroll
    for (int p = 0; p < 2; ++p) {
        const int j = sg.cbase + 2 * lane + p;
        if (j >= sg.c0 && j < sg.c0 + sg.tc) R.own |= 1u << p;
        const bool in = j >= 1 && j <= A.g.M - 1;
        if (in) R.own |= 4u << p;
        R.kS[p] = R.kN[p] = 0.0;
        if (CB && in) {
            R.kS[p] = fT * A.g.rkp[j] * A.g.rk[j];
            R.kN[p] = fT * A.g.rkp[j] * A.g.rk[j + 1];
        }
    }
    // rows before the first marched one are rows of zeros: every slot starts with zero coefficients
    for (int s = 0; s < P * DD_LANE_NA(CB, XIN) + DD_LANE_VS; ++s) dd_lane_zero16(sm, R.ring + (unsigned)s * 512u);
}

// ---- request: the coefficients (and the previous pass's x) of march row q into slot SL ---------------------------
template <int CB, int S, int XIN>
DD_HD void dd_lane_request(const WaveArgs& A, const WaveSeg& sg, const LaneRegs<CB, S, XIN>& R, const LaneSmem& sm, int q,
                           int lane, int slot) {
    constexpr int NA = DD_LANE_NA(CB, XIN);
    const int i = sg.rs + q;
    const int j0 = sg.cbase + 2 * lane;
    const bool rowok = q < sg.nq && i >= A.vr0;
    int n = 0;
    if (rowok && j0 >= 0 && j0 <= A.g.M) n = j0 < A.g.M ? 16 : 8;  // last column on an even index: no partner
    const long long o0 = n ? sg.moR + (long long)i * A.ldR + j0 : 0;
    const unsigned dst = R.ring + (unsigned)(slot * NA) * 512u;
    dd_lane_cp16(sm, dst, A.bb + o0, n);
    dd_lane_cp16(sm, dst + 512u, A.aW + o0, n);
    if (!CB) {
        dd_lane_cp16(sm, dst + 2 * 512u, A.aE + o0, n);
        dd_lane_cp16(sm, dst + 3 * 512u, A.aS + o0, n);
        dd_lane_cp16(sm, dst + 4 * 512u, A.aN + o0, n);
    }
    if (XIN) dd_lane_cp16(sm, dst + (unsigned)(NA - 1) * 512u, A.xin + o0, n);
    if (A.last_pass) {
        // v* of the row that leaves in the step in which row q enters (pitch g.ld: 8-byte alignment only)
        const int io = i - 2 * S - 1;
        const bool vrow = io >= sg.r0 && io < sg.r1;
        const long long ov = vrow ? sg.mo + (long long)io * A.g.ld + j0 : 0;
        const unsigned dv = R.ring + (unsigned)((2 * S + 4) * NA + (q & (DD_LANE_VS - 1))) * 512u;
        dd_lane_cp8(sm, dv, A.vstar + ov, vrow && (R.own & 1u) ? 8 : 0);
        dd_lane_cp8(sm, dv + 8u, A.vstar + ov + ((R.own & 2u) ? 1 : 0), vrow && (R.own & 2u) ? 8 : 0);
    }
    DD_LANE_COMMIT();
    // a few rows further down: into L2
    const int qp = q + DD_LANE_PF;
    if (DD_LANE_PF > 0 && qp < sg.nq && j0 >= 0 && j0 < A.g.M) {
        const long long op = sg.moR + (long long)(sg.rs + qp) * A.ldR + j0;
        dd_lane_prefetch(A.bb + op);
        dd_lane_prefetch(A.aW + op);
        if (!CB) {
            dd_lane_prefetch(A.aE + op);
            dd_lane_prefetch(A.aS + op);
            dd_lane_prefetch(A.aN + op);
        }
        if (XIN) dd_lane_prefetch(A.xin + op);
        const int ip = sg.rs + qp - 2 * S - 1;
        if (A.last_pass && ip >= sg.r0 && ip < sg.r1 && j0 >= sg.c0 && j0 < sg.c0 + sg.tc)
            dd_lane_prefetch(A.vstar + sg.mo + (long long)ip * A.g.ld + j0);
    }
}

// what the lane offers its neighbour in step U: the column that is NOT relaxed in this step, rows tau - 1 - k
template <int CB, int S, int XIN, int U>
DD_HD void dd_lane_offer(const LaneRegs<CB, S, XIN>& R, double* src) {
    constexpr int P = 2 * S + 4, o = (U + 1) & 1;
#pragma unroll
    for (int k = 0; k <= 2 * S; ++k) src[k] = R.X[1 - o][(U - 1 - k + 2 * P) % P];
}

// Gauss-Seidel value minus x (the residual for iterate value x) of cell (row slot SL, column o) -- c[]: bb, aW, aE,
// aS, aN | bb, dinv
template <int CB, int S, int XIN, int SL, int O>
DD_HD double dd_lane_gs(const LaneRegs<CB, S, XIN>& R, const double* c, double nb, double x) {
    constexpr int P = 2 * S + 4;
    const double xw = R.X[O][(SL + P - 1) % P], xe = R.X[O][(SL + 1) % P];
    const double xs = O ? R.X[0][SL] : nb, xn = O ? nb : R.X[1][SL];
    if (CB) return dd_sor_dT(c[0], c[1], R.RW[CB ? SL : 0], R.RE[CB ? SL : 0], R.kS[O], R.kN[O], xw, xe, xs, xn, x);
    return dd_sor_d5(c[0], c[1], c[CB ? 0 : 2], c[CB ? 0 : 3], c[CB ? 0 : 4], xw, xe, xs, xn, x);
}

// one relaxation of cell (row slot SL, column O)
template <int CB, int S, int XIN, int SL, int O>
DD_HD void dd_lane_relax(LaneRegs<CB, S, XIN>& R, const double* c, double nb, double omega) {
    const double x = R.X[O][SL];
    R.X[O][SL] = dd_sor_relax(x, dd_lane_gs<CB, S, XIN, SL, O>(R, c, nb, x), omega);
}

// level pairs K .. S - 1 of step U: odd level 2 K + 1 on row tau - 2 K - 1 (coefficients from the ring, the other
// column's half kept for the next step), then even level 2 K + 2 on row tau - 2 K - 2 (coefficients kept by the
// previous step).  fin_row: the row of the last level is an owned row of the last pass -- its gs - x_new is the
// residual of its colour-1 cells.
template <int CB, int S, int XIN, int U, int K>
DD_HD void dd_lane_pairs(LaneRegs<CB, S, XIN>& R, const LaneSmem& sm, const double* nb, double omega, bool fin_row) {
    if constexpr (K < S) {
        constexpr int P = 2 * S + 4, NC = CB ? 2 : 5, NA = DD_LANE_NA(CB, XIN), o = (U + 1) & 1;
        (void)NA;
        constexpr int s1 = (U + 2 * P - 2 * K - 1) % P, s2 = (U + 2 * P - 2 * K - 2) % P;
        double old[NC], use[NC];
#pragma unroll
        for (int a = 0; a < NC; ++a) old[a] = R.carry[K][a];
#pragma unroll
        for (int a = 0; a < NC; ++a) {
            double p0, p1;
            dd_lane_lds2(sm, R.ring + DD_LANE_OFF(CB, XIN, s1, a), p0, p1);
            use[a] = o ? p1 : p0;
            R.carry[K][a] = o ? p0 : p1;
        }
        dd_lane_relax<CB, S, XIN, s1, o>(R, use, nb[2 * K], omega);
        dd_lane_relax<CB, S, XIN, s2, o>(R, old, nb[2 * K + 1], omega);
        if (K == S - 1 && fin_row) {
            // the neighbours of these cells are final: the residual of the new value
            const double res = dd_lane_gs<CB, S, XIN, s2, o>(R, old, nb[2 * K + 1], R.X[o][s2]);
            const unsigned h = dd_wave_hi(res) & (0u - ((R.own >> o) & 1u));
            R.hr = h > R.hr ? h : R.hr;
        }
        dd_lane_pairs<CB, S, XIN, U, K + 1>(R, sm, nb, omega, fin_row);
    }
}

// ---- one time step of a lane; tau = U (mod P), nb[k]: the neighbour lane's offer for row tau - 1 - k ---------------
template <int CB, int S, int XIN, int U>
DD_HD void dd_lane_step(const WaveArgs& A, const WaveSeg& sg, LaneRegs<CB, S, XIN>& R, const LaneSmem& sm, int tau, int lane,
                        double omega, double fT, const double* nb) {
    constexpr int P = 2 * S + 4, NC = CB ? 2 : 5, NA = DD_LANE_NA(CB, XIN), o = (U + 1) & 1;
    (void)NA;
    // (1) row tau + LS is requested, row tau has arrived
    dd_lane_request<CB, S, XIN>(A, sg, R, sm, tau + DD_LANE_LS, lane, (U + DD_LANE_LS) % P);
    DD_LANE_WAIT();
    const int i = sg.rs + tau;
    // the row leaving in this step and v* of its cells (asked for early, used at the end of the step)
    constexpr int SO = (U + P - 2 * S - 1) % P;
    const int io = i - 2 * S - 1;
    const bool out = io >= sg.r0 && io < sg.r1;
    // (2) row tau enters: initial iterate, row factors
    if (XIN) {
        dd_lane_lds2(sm, R.ring + DD_LANE_OFF(CB, XIN, U, NA - 1), R.X[0][U], R.X[1][U]);
    } else {
        R.X[0][U] = R.X[1][U] = 0.0;
    }
    // const band: the row's metrics are asked for now and multiplied up after the levels (4): they are global
    // loads that mostly miss L1, and the row is first relaxed in the next step
    double m_rp = 0.0, m_rh0 = 0.0, m_rh1 = 0.0;
    if (CB) {
        const int gi = A.g.row0 + i;
        if (tau < sg.nq && i >= A.vr0 && gi >= 1 && gi <= A.g.N - 1) {
            m_rp = A.g.rhp[gi];
            m_rh0 = A.g.rh[gi];
            m_rh1 = A.g.rh[gi + 1];
        }
    }
    // (3) the 2 S levels, in pairs: odd level 2 k + 1 on row tau - 2 k - 1 (coefficients from the ring, the other
    //     column's half kept for the next step), even level 2 k + 2 on row tau - 2 k - 2 (coefficients kept by the
    //     previous step)
    const bool fin_row = A.last_pass && i - 2 * S >= sg.r0 && i - 2 * S < sg.r1;
    dd_lane_pairs<CB, S, XIN, U, 0>(R, sm, nb, omega, fin_row);
    if (CB) {
        // dt DT / (hhat_i h_i), dt DT / (hhat_i h_{i+1}): same products, in the same order, as in the tile kernels
        R.RW[CB ? U : 0] = fT * m_rp * m_rh0;
        R.RE[CB ? U : 0] = fT * m_rp * m_rh1;
    }
    // (4) row tau - 2 S - 1 leaves: residual of its colour-0 cells (column o), |bb|, v_new = v* + x
    if (out) {
        const double x0 = R.X[0][SO], x1 = R.X[1][SO];
        if (A.last_pass) {
            double c[NC], b0, b1;
            dd_lane_lds2(sm, R.ring + DD_LANE_OFF(CB, XIN, SO, 0), b0, b1);
            c[0] = o ? b1 : b0;
#pragma unroll
            for (int a = 1; a < NC; ++a) c[a] = dd_lane_lds(sm, R.ring + DD_LANE_OFF(CB, XIN, SO, a) + 8u * o);
            const double res = dd_lane_gs<CB, S, XIN, SO, o>(R, c, nb[2 * S], R.X[o][SO]);
            const unsigned m0 = 0u - (R.own & 1u), m1 = 0u - ((R.own >> 1) & 1u);
            const unsigned h = dd_wave_hi(res) & (o ? m1 : m0);
            R.hr = h > R.hr ? h : R.hr;
            const unsigned hb0 = dd_wave_hi(b0) & m0, hb1 = dd_wave_hi(b1) & m1;
            R.hb = hb0 > R.hb ? hb0 : R.hb;
            R.hb = hb1 > R.hb ? hb1 : R.hb;
            const int gi = A.g.row0 + io;
            const bool irow = gi > 0 && gi < A.g.N;
            double vs0, vs1;
            dd_lane_lds2(sm, R.ring + (unsigned)(P * NA + (tau & (DD_LANE_VS - 1))) * 512u, vs0, vs1);
            double* vp = A.vnew + sg.mo + (long long)io * A.g.ld + sg.cbase + 2 * lane;
            if (R.own & 1u) {
                const double vn = dd_newton_update(irow && (R.own & 4u), vs0, x0, A.zero_boundary);
                vp[0] = vn;
                const unsigned hx = dd_wave_hi(x0), hv = dd_wave_hi(vn);
                R.hx = hx > R.hx ? hx : R.hx;
                R.hv = hv > R.hv ? hv : R.hv;
            }
            if (R.own & 2u) {
                const double vn = dd_newton_update(irow && (R.own & 8u), vs1, x1, A.zero_boundary);
                vp[1] = vn;
                const unsigned hx = dd_wave_hi(x1), hv = dd_wave_hi(vn);
                R.hx = hx > R.hx ? hx : R.hx;
                R.hv = hv > R.hv ? hv : R.hv;
            }
        } else {
            double* xp = A.xout + sg.moR + (long long)io * A.ldR + sg.cbase + 2 * lane;
            if (R.own & 1u) xp[0] = x0;
            if (R.own & 2u) xp[1] = x1;
        }
    }
}

// number of time steps of a march: row r1 - 1 leaves in step (r1 - 1 - rs) + 2 S + 1; a whole number of periods
DD_HD int dd_lane_steps(const WaveArgs& A, const WaveSeg& sg) {
    const int P = 2 * A.sweeps + 4;
    const int n = (sg.r1 - 1 - sg.rs) + 2 * A.sweeps + 2;
    return (n + P - 1) / P * P;
}
