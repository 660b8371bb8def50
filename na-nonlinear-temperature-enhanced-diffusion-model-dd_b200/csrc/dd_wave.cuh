// dd_wave.cuh -- wavefront form of the red-black SOR solve (one thread's share of one time step).
//
// The tile solver of dd_solver.cu stages a tile plus a halo of 2 cells per sweep on every side and therefore
// re-reads (and re-relaxes) 1.8-2.5x the cells it owns.  Here a CTA owns a strip of W = 64 C columns and MARCHES
// down the rows: row q is relaxed at half-sweep h in time step tau = q + 2 h, so all 2 S half-sweeps of the pass
// are in flight at once, two rows apart (row q at level h needs rows q-1, q, q+1 at level h-1, which were
// produced in steps tau-3, tau-2, tau-1).  A row is read from HBM once, lives 4 S + 2 steps in a ring of
// D = 2 * warps slots and leaves as v_new = v* + x; only the strip's column halo (2 S + 1 cells per side) and a
// warm-up of 2 S + 1 rows per row segment are redundant.
//
// Work distribution: warp w owns ring slots 2 w and 2 w + 1 (rows of different parity: exactly one of them is
// relaxed per step), lane l owns packed columns l + 32 c (c < C) of them, both colours.  The coefficients of the
// owned cells stay in REGISTERS for the life of the row (loaded with one aligned 16-byte load per array and
// pair), only the iterate x lives in shared memory: [slot][colour][packed column], unit stride in every access.
// Per relaxation: 4 neighbour loads + own x + 1 store.
//
// Exactness: a cell's update is the same expression on the same operands as in the global red-black iteration
// (colour = parity of the GLOBAL i + j, dd_sor_* of dd_sor.cuh); rows before the first marched row and columns
// outside the strip read as 0, which is wrong data that moves one cell per half-sweep and never reaches an owned
// cell (same argument as for the tiles).  Results are therefore independent of strips, segments and slabs,
// bit for bit, and equal to those of the tile kernels.
//
// This header is compiled for the device (dd_wave.cu) and, test only, for the host (tests/hostsim), where the
// per-thread step is run thread by thread between the barriers.
#pragma once

#include "dd_nodeprog.cuh"
#include "dd_sor.cuh"

#define DD_WAVE_LC 1  // a row's coefficients are requested at age -LC (its iterate enters the ring at age 0,
                      // its first relaxation is at age 2); the ring needs D >= 4 S + 4 slots
#define DD_WAVE_VS 8  // ring of staged v* rows (requested 4 steps before the row's epilogue)

struct WaveArgs {
    DDGeom g;
    const DDMember* mem;
    const double *bb, *aW, *aE, *aS, *aN;  // const band: bb and dinv (in aW)
    const double* xin;                     // nullable: zero initial iterate
    double* xout;                          // passes that are not the last
    const double* vstar;
    double* vnew;
    DDSolveStats* stats;
    int zero_boundary;
    int ldR;
    long long mstrideR;
    int own0, own1;  // local rows whose result is written
    int vr0, vr1;    // local rows holding valid assembled rows
    int sweeps, halo, last_pass;
    int tj, nstrips;          // owned columns per strip, strips per member
    long long flat_total;     // members * nstrips * (own1 - own0): rows of all strips laid end to end
    long long flat_per_cta;
    double rho_fix;
};

// one march: rows [r0, r1) of strip (c0, tc) of one member
struct WaveSeg {
    int member, c0, tc, cbase, r0, r1, rs, nq;
    long long mo, moR;
};

DD_HD WaveSeg dd_wave_segment(const WaveArgs& A, long long f0, long long f1) {
    WaveSeg s;
    const int R = A.own1 - A.own0;
    const long long per_member = (long long)A.nstrips * R;
    s.member = (int)(f0 / per_member);
    const long long rem = f0 - (long long)s.member * per_member;
    const int strip = (int)(rem / R);
    const int r = (int)(rem - (long long)strip * R);
    long long n = f1 - f0;
    if (n > R - r) n = R - r;
    s.r0 = A.own0 + r;
    s.r1 = s.r0 + (int)n;
    s.c0 = strip * A.tj;
    s.tc = A.g.M + 1 - s.c0 < A.tj ? A.g.M + 1 - s.c0 : A.tj;
    s.cbase = s.c0 - A.halo - 1;
    s.rs = s.r0 - A.halo > A.vr0 ? s.r0 - A.halo : A.vr0;
    const int re = s.r1 + A.halo < A.vr1 ? s.r1 + A.halo : A.vr1;
    s.nq = re - s.rs;
    s.mo = (long long)s.member * A.g.mstride;
    s.moR = (long long)s.member * A.mstrideR;
    return s;
}

// registers of one thread: coefficients of its cells [slot k][chunk][colour], march counters, statistics
template <int CB, int C>
struct WaveRegs {
    double cb[2][C][2], cw[2][C][2];
    double ce[CB ? 1 : 2][CB ? 1 : C][2], cs[CB ? 1 : 2][CB ? 1 : C][2], cn[CB ? 1 : 2][CB ? 1 : C][2];
    double rW[2], rE[2];  // const band: row factors dt DT / (hhat_i h_i), dt DT / (hhat_i h_{i+1})
    int q[2], a[2];       // occupant row (march coordinates) of the slot and its age tau - q
    double rmax, xmax, vmax, bmax;
};

struct WaveSmem {
    double* x;     // [D][2][PW]
    double* vs;    // [DD_WAVE_VS][W]   v* of the rows about to leave
    double* scol;  // [2][W]            const band: column factors dt DT / (khat_j k_j), dt DT / (khat_j k_{j+1})
};

DD_HD size_t dd_wave_smem_doubles(int C, int nwarps) {
    const int W = 64 * C;
    return (size_t)(2 * nwarps) * W + (size_t)DD_WAVE_VS * W + 2 * (size_t)W;
}

#define DD_WAVE_PF 12  // L2 prefetch distance of the coefficient rows, in rows
#ifdef __CUDA_ARCH__
#define DD_WAVE_LD2(p) (*reinterpret_cast<const double2*>(p))
__device__ __forceinline__ void dd_wave_prefetch_l2(const void* p) { asm volatile("prefetch.global.L2 [%0];" ::"l"(p)); }
#else
struct dd_host_double2 {
    double x, y;
};
#define DD_WAVE_LD2(p) (dd_host_double2{(p)[0], (p)[1]})
#endif

template <int CB, int C>
DD_HD void dd_wave_init_thread(WaveRegs<CB, C>& R, int warp) {
#pragma unroll
    for (int k = 0; k < 2; ++k) {
#pragma unroll
        for (int c = 0; c < C; ++c)
#pragma unroll
            for (int o = 0; o < 2; ++o) {
                R.cb[k][c][o] = 0.0;
                R.cw[k][c][o] = 0.0;
                if (!CB) {
                    R.ce[k][c][o] = 0.0;
                    R.cs[k][c][o] = 0.0;
                    R.cn[k][c][o] = 0.0;
                }
            }
        R.rW[k] = R.rE[k] = 0.0;
        R.q[k] = 2 * warp + k;
        R.a[k] = -DD_WAVE_LC - (2 * warp + k);
    }
    R.rmax = R.xmax = R.vmax = R.bmax = 0.0;
}

// ---- load: the row's coefficients into the thread's registers -------------------------------------------------
template <int CB, int C, int K, int FLIP>
DD_HD void dd_wave_load(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, int lane, double fT) {
    constexpr int W = 64 * C;
    constexpr int e = FLIP ? 1 : 0, d = 1 - e;  // colours of the even / odd column of a pair
    const int q = R.q[K];
    const bool rowok = q < sg.nq;
    const int i = sg.rs + q;
    const long long orow = sg.moR + (long long)i * A.ldR;
#ifdef __CUDA_ARCH__
    // ask L2 for the coefficient rows DD_WAVE_PF steps ahead and for this row's v* (read at the row's epilogue):
    // the 16-byte loads below then find their lines on chip (prefetches hold no register and no scoreboard slot)
    if (q + DD_WAVE_PF < sg.nq) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j0 = sg.cbase + 2 * (lane + 32 * c);
            if (j0 >= 0 && j0 <= A.g.M && !(lane & 1)) {  // one request per 32-byte sector
                const long long o = orow + (long long)DD_WAVE_PF * A.ldR + j0;
                dd_wave_prefetch_l2(A.bb + o);
                dd_wave_prefetch_l2(A.aW + o);
                if (!CB) {
                    dd_wave_prefetch_l2(A.aE + o);
                    dd_wave_prefetch_l2(A.aS + o);
                    dd_wave_prefetch_l2(A.aN + o);
                }
                if (A.xin) dd_wave_prefetch_l2(A.xin + o);
            }
        }
    }
    if (A.last_pass && rowok && i >= sg.r0 && i < sg.r1) {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j0 = sg.cbase + 2 * (lane + 32 * c);
            if (j0 >= 0 && j0 <= A.g.M && !(lane & 1))
                dd_wave_prefetch_l2(A.vstar + sg.mo + (long long)i * A.g.ld + j0);
        }
    }
#endif
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int p = lane + 32 * c;
        const int sj0 = 2 * p, j0 = sg.cbase + sj0;
        const bool in0 = rowok && sj0 >= 1 && j0 >= 0 && j0 <= A.g.M;
        const bool in1 = rowok && sj0 + 1 <= W - 2 && j0 + 1 >= 0 && j0 + 1 <= A.g.M;
        double b0 = 0.0, b1 = 0.0, w0 = 0.0, w1 = 0.0, e0 = 0.0, e1 = 0.0, s0 = 0.0, s1 = 0.0, n0 = 0.0, n1 = 0.0;
        if (in0 || in1) {  // then j0 >= 0 or j0 == -1 ... the pair is addressable iff j0 >= 0 (j0 is even)
            if (j0 >= 0) {
                const long long o = orow + j0;
                const auto vb = DD_WAVE_LD2(A.bb + o);
                const auto vw = DD_WAVE_LD2(A.aW + o);
                b0 = vb.x; b1 = vb.y; w0 = vw.x; w1 = vw.y;
                if (!CB) {
                    const auto ve = DD_WAVE_LD2(A.aE + o);
                    const auto vs = DD_WAVE_LD2(A.aS + o);
                    const auto vn = DD_WAVE_LD2(A.aN + o);
                    e0 = ve.x; e1 = ve.y; s0 = vs.x; s1 = vs.y; n0 = vn.x; n1 = vn.y;
                }
            }
            if (!in0) b0 = w0 = e0 = s0 = n0 = 0.0;
            if (!in1) b1 = w1 = e1 = s1 = n1 = 0.0;
        }
        R.cb[K][c][e] = b0; R.cb[K][c][d] = b1;
        R.cw[K][c][e] = w0; R.cw[K][c][d] = w1;
        if (!CB) {
            R.ce[K][c][e] = e0; R.ce[K][c][d] = e1;
            R.cs[K][c][e] = s0; R.cs[K][c][d] = s1;
            R.cn[K][c][e] = n0; R.cn[K][c][d] = n1;
        }
    }
    if (CB) {
        const int gi = A.g.row0 + i;
        const bool ok = rowok && gi >= 1 && gi <= A.g.N - 1;
        R.rW[K] = ok ? fT * A.g.rhp[gi] * A.g.rh[gi] : 0.0;
        R.rE[K] = ok ? fT * A.g.rhp[gi] * A.g.rh[gi + 1] : 0.0;
    }
}

// ---- the row's initial iterate into its slot (zero, or the previous pass's x) -----------------------------------
template <int CB, int C, int K, int FLIP>
DD_HD void dd_wave_xinit(const WaveArgs& A, const WaveSeg& sg, const WaveRegs<CB, C>& R, const WaveSmem& sm, int slot,
                         int lane) {
    constexpr int PW = 32 * C, W = 64 * C;
    constexpr int e = FLIP ? 1 : 0, d = 1 - e;
    const int q = R.q[K];
    const bool rowok = q < sg.nq;
    const long long orow = sg.moR + (long long)(sg.rs + q) * A.ldR;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int p = lane + 32 * c;
        const int sj0 = 2 * p, j0 = sg.cbase + sj0;
        double x0 = 0.0, x1 = 0.0;
        if (A.xin && rowok && j0 >= 0 && j0 <= A.g.M) {
            const auto vx = DD_WAVE_LD2(A.xin + orow + j0);
            if (sj0 >= 1) x0 = vx.x;
            if (sj0 + 1 <= W - 2 && j0 + 1 <= A.g.M) x1 = vx.y;
        }
        sm.x[(slot * 2 + e) * PW + p] = x0;
        sm.x[(slot * 2 + d) * PW + p] = x1;
    }
}

// Gauss-Seidel value of cell (slot, colour CO, packed column p) from the other colour's current iterate
template <int CB, int C, int K, int CO>
DD_HD double dd_wave_gs(const WaveRegs<CB, C>& R, const WaveSmem& sm, int c, int p, int o, int slot, int slotW,
                        int slotE) {
    constexpr int PW = 32 * C, W = 64 * C;
    const double* xo = sm.x + (1 - CO) * PW;  // other colour's plane of slot 0
    const double xw = xo[slotW * 2 * PW + p], xe = xo[slotE * 2 * PW + p];
    const double xs = xo[slot * 2 * PW + p + o - 1], xn = xo[slot * 2 * PW + p + o];
    if (CB) {
        const int sj = 2 * p + o;
        return dd_sor_gsT(R.cb[K][c][CO], R.cw[K][c][CO], R.rW[K], R.rE[K], sm.scol[sj], sm.scol[W + sj], xw, xe, xs,
                          xn);
    }
    return dd_sor_gs5(R.cb[K][c][CO], R.cw[K][c][CO], R.ce[CB ? 0 : K][CB ? 0 : c][CO], R.cs[CB ? 0 : K][CB ? 0 : c][CO],
                      R.cn[CB ? 0 : K][CB ? 0 : c][CO], xw, xe, xs, xn);
}

// ---- one half-sweep of the slot's row: colour CO ---------------------------------------------------------------
template <int CB, int C, int K, int CO>
DD_HD void dd_wave_relax(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int slot,
                         int slotW, int slotE, int lane, double omega) {
    constexpr int PW = 32 * C;
    const int gi = A.g.row0 + sg.rs + R.q[K];
    const int o = (CO + gi) & 1;  // column parity of this colour's cells in this row (cbase is even)
    double xnew[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int p = lane + 32 * c;
        // columns 0 and W-1 of the strip are a ring of zeros that is never relaxed (no neighbour beyond it)
        const bool edge = (p == 0 && o == 0) || (p == PW - 1 && o == 1);
        xnew[c] = 0.0;
        if (!edge) {
            const double xv = sm.x[(slot * 2 + CO) * PW + p];
            xnew[c] = dd_sor_relax(xv, dd_wave_gs<CB, C, K, CO>(R, sm, c, p, o, slot, slotW, slotE), omega);
        }
    }
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int p = lane + 32 * c;
        const bool edge = (p == 0 && o == 0) || (p == PW - 1 && o == 1);
        if (!edge) sm.x[(slot * 2 + CO) * PW + p] = xnew[c];
    }
}

// ---- epilogue of the slot's row: residual statistics and v_new = v* + x (last pass) or x (other passes) ---------
template <int CB, int C, int K>
DD_HD void dd_wave_finish(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int slot,
                          int slotW, int slotE, int lane) {
    constexpr int PW = 32 * C, W = 64 * C;
    const int q = R.q[K], i = sg.rs + q, gi = A.g.row0 + i;
    if (i < sg.r0 || i >= sg.r1) return;
    const int H = A.halo;
    const double* vrow = sm.vs + (q & (DD_WAVE_VS - 1)) * W;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int p = lane + 32 * c;
#pragma unroll
        for (int co = 0; co < 2; ++co) {
            const int o = (co + gi) & 1, sj = 2 * p + o;
            if (sj < H + 1 || sj >= H + 1 + sg.tc) continue;
            const int j = sg.cbase + sj;
            const double x = sm.x[(slot * 2 + co) * PW + p];
            if (A.last_pass) {
                const double gs = co ? dd_wave_gs<CB, C, K, 1>(R, sm, c, p, o, slot, slotW, slotE)
                                     : dd_wave_gs<CB, C, K, 0>(R, sm, c, p, o, slot, slotW, slotE);
                const bool inter = gi > 0 && gi < A.g.N && j > 0 && j < A.g.M;
                const double vn = dd_newton_update(inter, vrow[sj], x, A.zero_boundary);
                A.vnew[sg.mo + (long long)i * A.g.ld + j] = vn;
                R.rmax = dd_nn_max(R.rmax, gs - x);
                R.xmax = dd_nn_max(R.xmax, x);
                R.vmax = dd_nn_max(R.vmax, vn);
                R.bmax = dd_nn_max(R.bmax, R.cb[K][c][co]);
            } else {
                A.xout[sg.moR + (long long)i * A.ldR + j] = x;
            }
        }
    }
}

// request v* of the slot's row into the staging ring (consumed four steps later by dd_wave_finish)
template <int CB, int C, int K>
DD_HD void dd_wave_request_vstar(const WaveArgs& A, const WaveSeg& sg, const WaveRegs<CB, C>& R, const WaveSmem& sm,
                                 int lane) {
    constexpr int W = 64 * C;
    const int q = R.q[K], i = sg.rs + q;
    if (!A.last_pass || i < sg.r0 || i >= sg.r1) return;
    double* vrow = sm.vs + (q & (DD_WAVE_VS - 1)) * W;
    const double* src = A.vstar + sg.mo + (long long)i * A.g.ld + sg.cbase;
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const int sj = 2 * (lane + 32 * c) + o;  // the thread stages exactly the cells it will finish
            if (sj < A.halo + 1 || sj >= A.halo + 1 + sg.tc) continue;
#ifdef __CUDA_ARCH__
            __pipeline_memcpy_async(vrow + sj, src + sj, 8);
#else
            vrow[sj] = src[sj];
#endif
        }
#ifdef __CUDA_ARCH__
    __pipeline_commit();
#endif
}

// ---- one time step of one thread (the caller synchronises the CTA between steps) --------------------------------
template <int CB, int C, int K>
DD_HD void dd_wave_slot_step(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int warp,
                             int lane, int nwarps, double omega, double fT) {
    const int D = 2 * nwarps, slot = 2 * warp + K;
    const int slotW = slot == 0 ? D - 1 : slot - 1, slotE = slot == D - 1 ? 0 : slot + 1;
    const int a = R.a[K], S4 = 4 * A.sweeps;
    const int flip = (A.g.row0 + sg.rs + R.q[K]) & 1;  // colour of the row's even columns
    if (a == -DD_WAVE_LC) {
        if (flip)
            dd_wave_load<CB, C, K, 1>(A, sg, R, lane, fT);
        else
            dd_wave_load<CB, C, K, 0>(A, sg, R, lane, fT);
    } else if (a == 0) {
        if (flip)
            dd_wave_xinit<CB, C, K, 1>(A, sg, R, sm, slot, lane);
        else
            dd_wave_xinit<CB, C, K, 0>(A, sg, R, sm, slot, lane);
    } else if (a >= 2 && a <= S4 && !(a & 1)) {
        if (R.q[K] < sg.nq) {
            if ((a >> 1) & 1)  // half-sweep h = a / 2 relaxes colour (h - 1) & 1
                dd_wave_relax<CB, C, K, 0>(A, sg, R, sm, slot, slotW, slotE, lane, omega);
            else
                dd_wave_relax<CB, C, K, 1>(A, sg, R, sm, slot, slotW, slotE, lane, omega);
        }
        if (a == S4 - 2) dd_wave_request_vstar<CB, C, K>(A, sg, R, sm, lane);
    } else if (a == S4 + 2) {
#ifdef __CUDA_ARCH__
        __pipeline_wait_prior(0);
#endif
        dd_wave_finish<CB, C, K>(A, sg, R, sm, slot, slotW, slotE, lane);
    }
    // advance the slot's clock; recycle it for row q + D once the current row and its neighbours are done with it
    R.a[K] = a + 1;
    if (R.a[K] == D - DD_WAVE_LC) {
        R.q[K] += D;
        R.a[K] = -DD_WAVE_LC;
    }
}

template <int CB, int C>
DD_HD void dd_wave_thread_step(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int warp,
                               int lane, int nwarps, double omega, double fT) {
    dd_wave_slot_step<CB, C, 0>(A, sg, R, sm, warp, lane, nwarps, omega, fT);
    dd_wave_slot_step<CB, C, 1>(A, sg, R, sm, warp, lane, nwarps, omega, fT);
}

// number of time steps of a march (steps tau = -LC .. tau_end)
DD_HD int dd_wave_steps(const WaveArgs& A, const WaveSeg& sg) {
    const int q_last = sg.r1 - 1 - sg.rs;  // last row with an epilogue
    return q_last + 4 * A.sweeps + 2 + DD_WAVE_LC + 1;
}
