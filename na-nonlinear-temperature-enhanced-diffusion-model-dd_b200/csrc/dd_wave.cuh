// dd_wave.cuh -- wavefront form of the red-black SOR solve (one thread's share of one time step).
//
// The tile solver of dd_solver.cu stages a tile plus a halo of 2 cells per sweep on every side and therefore
// re-reads (and re-relaxes) 1.8-2.5x the cells it owns.  Here a CTA owns a strip of W = 64 C columns and MARCHES
// down the rows: row q is relaxed at half-sweep h in time step tau = q + 2 h, so all 2 S half-sweeps of the pass
// are in flight at once, two rows apart (row q at level h needs rows q-1, q, q+1 at level h-1, which were
// produced in steps tau-3, tau-2, tau-1).  A row is read from HBM once, lives 4 S + 2 steps in a ring of
// D = 2 * warps >= 4 S + 4 slots and leaves as v_new = v* + x; only the strip's column halo (2 S + 2 cells per
// side) and a warm-up of 2 S + 1 rows per march are redundant.
//
// Work distribution: warp w owns ring slots 2 w and 2 w + 1, lane l owns packed columns l + 32 c (c < C) of them,
// both colours.  A march starts on an even global row, so slot 0 of every warp always holds even rows and slot 1
// odd rows: which column parity a colour has in a row is a compile-time property of the slot.  Exactly one of a
// warp's two rows is relaxed per step.  The coefficients of the owned cells stay in REGISTERS for the life of the
// row (requested LS steps ahead by 16-byte cp.async into a staging ring, copied into registers one step before
// the row enters the ring), only the iterate x lives in shared memory: [slot][colour][1 + packed column + 1],
// the two pad words are always 0 and stand for the columns outside the strip.  Per relaxation: 4 neighbour
// loads + own x + 1 store, all unit-stride, addressed by 32-bit shared-window addresses with immediate offsets.
//
// Exactness: a cell's update is the same expression on the same operands as in the global red-black iteration
// (colour = parity of the GLOBAL i + j, dd_sor_* of dd_sor.cuh); rows before the first marched row and columns
// outside the strip read as 0, which is wrong data that moves one cell per half-sweep and never reaches an owned
// cell (same argument as for the tiles).  Results are therefore independent of strips, marches and slabs,
// bit for bit, and equal to those of the tile kernels.
//
// This header is compiled for the device (dd_wave.cu) and, test only, for the host (tests/hostsim), where the
// per-thread step is run thread by thread between the barriers.
#pragma once

#include "dd_nodeprog.cuh"
#include "dd_sor.cuh"

#define DD_WAVE_LS 4   // a row's coefficients are requested (cp.async into a staging ring) this many steps before
                       // they are filled into the owner's registers
#define DD_WAVE_SR 8   // rows of the staging rings (>= LS + 1, power of two)
#define DD_WAVE_NE 4   // epilogue warps
#define DD_WAVE_EV 2   // columns per lane of an epilogue warp (strip's owned columns <= 32 * NE * EV)

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
    int sweeps, halo, last_pass;  // halo = 2 S + 2 columns per side (even: strips start on even columns)
    int tj, nstrips;              // owned columns per strip, strips per member
    int nwo;                      // owner warps (ring of D = 2 nwo slots); DD_WAVE_NE epilogue warps follow them
    long long flat_total;         // members * nstrips * (own1 - own0): rows of all strips laid end to end
    long long flat_per_cta;
    double rho_fix;
};

// one march: rows [r0, r1) of strip (c0, tc) of one member; marched rows rs + q, q = 0 .. nq - 1
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
    s.cbase = s.c0 - A.halo;
    const int wu = A.halo - 1;  // warm-up rows: 2 S + 1
    s.rs = s.r0 - wu > A.vr0 ? s.r0 - wu : A.vr0;
    s.rs -= (A.g.row0 + s.rs) & 1;  // start on an even global row (a row before vr0 is a row of zeros)
    const int re = s.r1 + wu < A.vr1 ? s.r1 + wu : A.vr1;
    s.nq = re - s.rs;
    s.mo = (long long)s.member * A.g.mstride;
    s.moR = (long long)s.member * A.mstrideR;
    return s;
}

// Statistics are maxima of magnitudes; they are kept as the HIGH WORDS of the doubles (sign cleared): one integer
// max per value, NaN (largest pattern) sticks.  The low word is dropped, i.e. a maximum is known to 2^-20
// relative: the residual is rounded up and the scales |x|, |v_new|, |bb| down when they are turned back into
// doubles, so the convergence test they feed stays a rigorous bound.
DD_HD unsigned dd_wave_hi(double v) {
#ifdef __CUDA_ARCH__
    return (unsigned)__double2hiint(v) & 0x7fffffffu;
#else
    union { double d; unsigned long long u; } c;
    c.d = v;
    return (unsigned)(c.u >> 32) & 0x7fffffffu;
#endif
}
DD_HD double dd_wave_from_hi(unsigned h, bool round_up) {
    union { double d; unsigned long long u; } c;
    c.u = (unsigned long long)(round_up && h != 0u ? h + 1u : h) << 32;
    return c.d;
}

// registers of an owner thread: coefficients of its cells [slot k][chunk][colour], march counters, statistics
template <int CB, int C>
struct WaveRegs {
    double cb[2][C][2], cw[2][C][2];
    double ce[CB ? 1 : 2][CB ? 1 : C][2], cs[CB ? 1 : 2][CB ? 1 : C][2], cn[CB ? 1 : 2][CB ? 1 : C][2];
    double rW[2], rE[2];             // const band: row factors dt DT / (hhat_i h_i), dt DT / (hhat_i h_{i+1})
    int q0, a0;                      // row (march coordinates) held by slot 0 and its age tau - q0, -1 <= a0 <= D - 2;
                                     // slot 1 holds row q0 + 1 (age a0 - 1), or still row q0 + 1 - D (age D - 2) while a0 = -1
    unsigned ax[2], axW[2], axE[2];  // shared-memory byte address of (slot, colour 0, packed column = lane) and
                                     // of the same element of the rows above and below
    unsigned tmask;                  // bit 2 c + o: the thread's cell (chunk c, column parity o) is an owned column
    unsigned hr, hb;                 // high words of max |residual|, max |bb| over its owned cells
};

// registers of an epilogue thread.  Everything that changes from step to step is advanced incrementally (pointers by
// one row, ring positions by one slot) instead of being recomputed: an epilogue warp's step is a serial chain of
// instructions that every other warp of the CTA waits for at the barrier.
struct WaveEpi {
    unsigned hx, hv;     // high words of max |x|, max |v_new|
    unsigned emask;      // bit e: the lane owns column t0 + 32 e; bit 8 + e: that column is an interior column
    unsigned xa;         // shared-memory byte address of x(slot of the row finished now, colour 0, lane's first column)
    unsigned xa_end;     // wrap-around bound of xa
    unsigned vsa, vra;   // staging ring: byte address of the lane's v* of the row finished now / requested now
    unsigned vs_end;     // wrap-around bound of the two
    int co;              // colour of the lane's columns in the row finished now (alternates from row to row)
    const double* src;   // v* of the row requested now (lane's first column)
    double* dst;         // v_new (last pass) or x (other passes) of the row finished now
};

// shared memory: x [D][2][PW + 2] | column factors [2][W] (const band) | coefficient staging [SR][NA][W] |
// row metrics staging [SR][4] | v* staging [SR][32 NE EV] | scratch
struct WaveSmem {
    double* base;    // generic pointer to the start (host: the emulated array)
    unsigned xaddr;  // byte address of x: shared-window address on the device, offset from `base` on the host
    unsigned kaddr;  // byte address of the column factors
    double* stage;   // coefficient staging ring
    double* mstage;  // row metrics staging ring
    double* vstage;  // v* staging ring (epilogue warps)
    unsigned vaddr;  // its byte address (shared window on the device, offset from `base` on the host)
};

#define DD_WAVE_NA(CB) ((CB) ? 3 : 6)  // staged arrays per row: bb, dinv | bb, aW, aE, aS, aN; + x of the previous pass

DD_HD size_t dd_wave_x_doubles(int C, int nwo) { return (size_t)(2 * nwo) * 2 * (32 * C + 2); }
DD_HD size_t dd_wave_smem_doubles(int CB, int C, int nwo) {
    const int W = 64 * C;
    return dd_wave_x_doubles(C, nwo) + 2 * (size_t)W + (size_t)DD_WAVE_SR * DD_WAVE_NA(CB) * W + DD_WAVE_SR * 4 +
           (size_t)DD_WAVE_SR * 32 * DD_WAVE_NE * DD_WAVE_EV + 64;
}
DD_HD void dd_wave_smem_carve(WaveSmem& sm, double* base, unsigned base_addr, int CB, int C, int nwo) {
    const int W = 64 * C;
    const size_t nx = dd_wave_x_doubles(C, nwo);
    sm.base = base;
    sm.xaddr = base_addr;
    sm.kaddr = base_addr + (unsigned)(nx * 8);
    sm.stage = base + nx + 2 * W;
    sm.mstage = sm.stage + (size_t)DD_WAVE_SR * DD_WAVE_NA(CB) * W;
    sm.vstage = sm.mstage + DD_WAVE_SR * 4;
    sm.vaddr = base_addr + (unsigned)((sm.vstage - base) * 8);
}
DD_HD unsigned* dd_wave_scratch(const WaveSmem& sm) {
    return reinterpret_cast<unsigned*>(sm.vstage + (size_t)DD_WAVE_SR * 32 * DD_WAVE_NE * DD_WAVE_EV);
}

#ifdef __CUDA_ARCH__
#define DD_WAVE_LD2(p) (*reinterpret_cast<const double2*>(p))
#define DD_WAVE_CP16(dst, src) __pipeline_memcpy_async(dst, src, 16)
#define DD_WAVE_CP8(dst, src) __pipeline_memcpy_async(dst, src, 8)
#define DD_WAVE_COMMIT() __pipeline_commit()
#define DD_WAVE_WAIT(n) __pipeline_wait_prior(n)
// 8-byte asynchronous copy global -> shared by shared-window address (no generic-address conversion)
__device__ __forceinline__ void dd_wave_cp8s(const WaveSmem&, unsigned dst, const double* src) {
    asm volatile("cp.async.ca.shared.global [%0], [%1], 8;" ::"r"(dst), "l"(src) : "memory");
}
__device__ __forceinline__ double dd_wave_lds(const WaveSmem&, unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void dd_wave_sts(const WaveSmem&, unsigned a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}
#else
struct dd_host_double2 {
    double x, y;
};
#define DD_WAVE_LD2(p) (dd_host_double2{(p)[0], (p)[1]})
#define DD_WAVE_CP16(dst, src) ((dst)[0] = (src)[0], (dst)[1] = (src)[1])
#define DD_WAVE_CP8(dst, src) ((dst)[0] = (src)[0])
#define DD_WAVE_COMMIT()
#define DD_WAVE_WAIT(n)
inline void dd_wave_cp8s(const WaveSmem& sm, unsigned dst, const double* src) { *(double*)((char*)sm.base + dst) = *src; }
inline double dd_wave_lds(const WaveSmem& sm, unsigned a) { return *(const double*)((const char*)sm.base + a); }
inline void dd_wave_sts(const WaveSmem& sm, unsigned a, double v) { *(double*)((char*)sm.base + a) = v; }
#endif

template <int CB, int C>
DD_HD void dd_wave_init_thread(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int warp,
                               int lane) {
    constexpr int PWP = 32 * C + 2;
    const int D = 2 * A.nwo;
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
        const int slot = 2 * warp + k;
        const int slotW = slot == 0 ? D - 1 : slot - 1, slotE = slot == D - 1 ? 0 : slot + 1;
        R.ax[k] = sm.xaddr + (unsigned)((slot * 2 * PWP + 1 + lane) * 8);
        R.axW[k] = sm.xaddr + (unsigned)((slotW * 2 * PWP + 1 + lane) * 8);
        R.axE[k] = sm.xaddr + (unsigned)((slotE * 2 * PWP + 1 + lane) * 8);
    }
    // The march starts at tau = -1 - LS with every slot holding a fictitious row of zeros, D rows before its first
    // real one: one rule then covers the start as well -- request row q + D at age D - 1 - LS, recycle at D - 2.
    R.q0 = 2 * warp - D;
    R.a0 = -1 - DD_WAVE_LS - R.q0;
    if (R.a0 > D - 2) {  // already past its recycling point: it is the real row's turn
        R.a0 -= D;
        R.q0 += D;
    }
    R.tmask = 0u;
#pragma unroll
    for (int c = 0; c < C; ++c)
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const int sj = 2 * (lane + 32 * c) + o;
            if (sj >= A.halo && sj < A.halo + sg.tc) R.tmask |= 1u << (2 * c + o);
        }
    R.hr = R.hb = 0u;
}

// ---- request: the coefficients of row q (the slot's next occupant) into the staging ring, asynchronously ----------
// cp.async is tracked by wait_group, not by register scoreboards: nothing waits for the data before dd_wave_fill,
// LS steps later (a load into registers would be waited for at the next branch).
template <int CB, int C, int K>
DD_HD void dd_wave_request(const WaveArgs& A, const WaveSeg& sg, const WaveSmem& sm, int q, int lane) {
    constexpr int W = 64 * C, NA = DD_WAVE_NA(CB);
    const int i = sg.rs + q;
    if (q < sg.nq && i >= A.vr0) {
        const int jl = sg.cbase + 2 * lane;  // even column of chunk 0
        const long long o0 = sg.moR + (long long)i * A.ldR + jl;
        double* st = sm.stage + (size_t)(q & (DD_WAVE_SR - 1)) * NA * W + 2 * lane;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j0 = jl + 64 * c;
            if (j0 >= 0 && j0 < A.g.M) {
                DD_WAVE_CP16(st + 64 * c, A.bb + o0 + 64 * c);
                DD_WAVE_CP16(st + W + 64 * c, A.aW + o0 + 64 * c);
                if (!CB) {
                    DD_WAVE_CP16(st + 2 * W + 64 * c, A.aE + o0 + 64 * c);
                    DD_WAVE_CP16(st + 3 * W + 64 * c, A.aS + o0 + 64 * c);
                    DD_WAVE_CP16(st + 4 * W + 64 * c, A.aN + o0 + 64 * c);
                }
                if (A.xin) DD_WAVE_CP16(st + (NA - 1) * W + 64 * c, A.xin + o0 + 64 * c);
            } else if (j0 == A.g.M) {
                // last grid column on an even index: its partner is the padding column of the row arrays
                DD_WAVE_CP8(st + 64 * c, A.bb + o0 + 64 * c);
                DD_WAVE_CP8(st + W + 64 * c, A.aW + o0 + 64 * c);
                if (!CB) {
                    DD_WAVE_CP8(st + 2 * W + 64 * c, A.aE + o0 + 64 * c);
                    DD_WAVE_CP8(st + 3 * W + 64 * c, A.aS + o0 + 64 * c);
                    DD_WAVE_CP8(st + 4 * W + 64 * c, A.aN + o0 + 64 * c);
                }
                if (A.xin) DD_WAVE_CP8(st + (NA - 1) * W + 64 * c, A.xin + o0 + 64 * c);
            }
        }
        if (CB && lane < 3) {
            // row metrics 1 / hhat_i, 1 / h_i, 1 / h_{i+1}: every lane copies them for itself (lanes 0..2 stage)
            const int gi = A.g.row0 + i;
            if (gi >= 1 && gi <= A.g.N - 1) {
                double* ms = sm.mstage + (q & (DD_WAVE_SR - 1)) * 4 + lane;
                const double* src = lane == 0 ? A.g.rhp + gi : (lane == 1 ? A.g.rh + gi : A.g.rh + gi + 1);
                DD_WAVE_CP8(ms, src);
            }
        }
    }
    DD_WAVE_COMMIT();
}

// ---- fill: the staged coefficients into the thread's registers ----------------------------------------------------
// Slot K holds rows of global parity K, so the even column of a pair has colour K.  The thread reads what it
// requested itself (no barrier needed); its other slot's request, one step younger, may still be in flight.
template <int CB, int C, int K>
DD_HD void dd_wave_fill(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int q, int lane,
                        double fT) {
    constexpr int W = 64 * C, NA = DD_WAVE_NA(CB);
    constexpr int e = K, d = 1 - K;  // colours of the even / odd column of a pair
    DD_WAVE_WAIT(K == 0 ? 1 : 0);
    const int i = sg.rs + q;
    const bool rowok = q < sg.nq && i >= A.vr0;
    const int jl = sg.cbase + 2 * lane;
    const double* st = sm.stage + (size_t)(q & (DD_WAVE_SR - 1)) * NA * W + 2 * lane;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const int j0 = jl + 64 * c;  // even
        double b0 = 0.0, b1 = 0.0, w0 = 0.0, w1 = 0.0, e0 = 0.0, e1 = 0.0, s0 = 0.0, s1 = 0.0, n0 = 0.0, n1 = 0.0;
        if (rowok && j0 >= 0 && j0 <= A.g.M) {
            const bool pair = j0 < A.g.M;
            const auto vb = DD_WAVE_LD2(st + 64 * c);
            const auto vw = DD_WAVE_LD2(st + W + 64 * c);
            b0 = vb.x; w0 = vw.x;
            b1 = pair ? vb.y : 0.0; w1 = pair ? vw.y : 0.0;
            if (!CB) {
                const auto ve = DD_WAVE_LD2(st + 2 * W + 64 * c);
                const auto vs = DD_WAVE_LD2(st + 3 * W + 64 * c);
                const auto vn = DD_WAVE_LD2(st + 4 * W + 64 * c);
                e0 = ve.x; s0 = vs.x; n0 = vn.x;
                e1 = pair ? ve.y : 0.0; s1 = pair ? vs.y : 0.0; n1 = pair ? vn.y : 0.0;
            }
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
        // dt DT / (hhat_i h_i), dt DT / (hhat_i h_{i+1}): same products, in the same order, as in the tile kernels
        const int gi = A.g.row0 + i;
        R.rW[K] = R.rE[K] = 0.0;
        if (rowok && gi >= 1 && gi <= A.g.N - 1) {
            const double* ms = sm.mstage + (q & (DD_WAVE_SR - 1)) * 4;
            const double rp = ms[0];
            R.rW[K] = fT * rp * ms[1];
            R.rE[K] = fT * rp * ms[2];
        }
    }
}

// ---- the row enters the ring: its initial iterate (zero, or the previous pass's x) into the slot -----------------
template <int CB, int C, int K>
DD_HD void dd_wave_xinit(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int q, int lane) {
    constexpr int PLB = (32 * C + 2) * 8, W = 64 * C, NA = DD_WAVE_NA(CB);
    constexpr int e = K, d = 1 - K;
    if (A.xin) {
        const int i = sg.rs + q;
        const bool rowok = q < sg.nq && i >= A.vr0;
        const double* st = sm.stage + (size_t)(q & (DD_WAVE_SR - 1)) * NA * W + (NA - 1) * W + 2 * lane;
#pragma unroll
        for (int c = 0; c < C; ++c) {
            const int j0 = sg.cbase + 2 * (lane + 32 * c);
            double x0 = 0.0, x1 = 0.0;
            if (rowok && j0 >= 0 && j0 <= A.g.M) {
                const auto vx = DD_WAVE_LD2(st + 64 * c);
                x0 = vx.x;
                if (j0 < A.g.M) x1 = vx.y;
            }
            dd_wave_sts(sm, R.ax[K] + e * PLB + c * 256, x0);
            dd_wave_sts(sm, R.ax[K] + d * PLB + c * 256, x1);
        }
    } else {
#pragma unroll
        for (int c = 0; c < C; ++c) {
            dd_wave_sts(sm, R.ax[K] + c * 256, 0.0);
            dd_wave_sts(sm, R.ax[K] + PLB + c * 256, 0.0);
        }
    }
}

// Gauss-Seidel value minus x (the residual for iterate value x) of the thread's cell (slot K, chunk c, colour CO)
// from the other colour's current iterate.
// In slot K the cells of colour CO sit on columns of parity O = CO ^ K: cell 2 p + O, neighbours 2 p + O -+ 1,
// i.e. packed columns p + O - 1 and p + O of the other colour.
template <int CB, int C, int K, int CO>
DD_HD double dd_wave_gs(const WaveRegs<CB, C>& R, const WaveSmem& sm, int lane, int c, double x) {
    constexpr int PLB = (32 * C + 2) * 8, O = CO ^ K, W = 64 * C;
    const int off = (1 - CO) * PLB + c * 256;
    const double xw = dd_wave_lds(sm, R.axW[K] + (unsigned)off), xe = dd_wave_lds(sm, R.axE[K] + (unsigned)off);
    const double xs = dd_wave_lds(sm, (unsigned)((int)R.ax[K] + off + (O - 1) * 8));
    const double xn = dd_wave_lds(sm, (unsigned)((int)R.ax[K] + off + O * 8));
    if (CB) {
        const unsigned ka = sm.kaddr + (unsigned)((2 * lane + O + 64 * c) * 8);
        return dd_sor_dT(R.cb[K][c][CO], R.cw[K][c][CO], R.rW[K], R.rE[K], dd_wave_lds(sm, ka),
                         dd_wave_lds(sm, ka + W * 8), xw, xe, xs, xn, x);
    }
    return dd_sor_d5(R.cb[K][c][CO], R.cw[K][c][CO], R.ce[CB ? 0 : K][CB ? 0 : c][CO], R.cs[CB ? 0 : K][CB ? 0 : c][CO],
                     R.cn[CB ? 0 : K][CB ? 0 : c][CO], xw, xe, xs, xn, x);
}

// ---- one half-sweep of the slot's row: colour CO ---------------------------------------------------------------
// FIN (last pass, last relaxation of colour 1): the neighbours are final, so gs - x_new is the true residual of
// these cells; colour 0's residual follows two steps later (dd_wave_resid0).
template <int CB, int C, int K, int CO, int FIN>
DD_HD void dd_wave_relax(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int q, int lane,
                         double omega) {
    constexpr int PLB = (32 * C + 2) * 8, O = CO ^ K;
    double xnew[C], gs[C];
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const double xv = dd_wave_lds(sm, R.ax[K] + CO * PLB + c * 256);
        gs[c] = dd_wave_gs<CB, C, K, CO>(R, sm, lane, c, xv);
        xnew[c] = dd_sor_relax(xv, gs[c], omega);
    }
#pragma unroll
    for (int c = 0; c < C; ++c) dd_wave_sts(sm, R.ax[K] + CO * PLB + c * 256, xnew[c]);
    if (FIN) {
        const int i = sg.rs + q;
        if (i >= sg.r0 && i < sg.r1) {
#pragma unroll
            for (int c = 0; c < C; ++c) {
                const unsigned m = 0u - ((R.tmask >> (2 * c + O)) & 1u);
                const unsigned h = dd_wave_hi(dd_wave_gs<CB, C, K, CO>(R, sm, lane, c, xnew[c])) & m;
                R.hr = h > R.hr ? h : R.hr;
            }
        }
    }
}

// residual of the row's colour-0 cells, once the colour-1 neighbours are final too, and |bb| of all its cells
template <int CB, int C, int K>
DD_HD void dd_wave_resid0(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int q, int lane) {
    constexpr int O = K;  // column parity of colour 0 in slot K
    const int i = sg.rs + q;
    if (!A.last_pass || i < sg.r0 || i >= sg.r1) return;
#pragma unroll
    for (int c = 0; c < C; ++c) {
        const unsigned m0 = 0u - ((R.tmask >> (2 * c + O)) & 1u), m1 = 0u - ((R.tmask >> (2 * c + 1 - O)) & 1u);
        const double x = dd_wave_lds(sm, R.ax[K] + c * 256);
        const unsigned h = dd_wave_hi(dd_wave_gs<CB, C, K, 0>(R, sm, lane, c, x)) & m0;
        R.hr = h > R.hr ? h : R.hr;
        const unsigned b0 = dd_wave_hi(R.cb[K][c][0]) & m0, b1 = dd_wave_hi(R.cb[K][c][1]) & m1;
        R.hb = b0 > R.hb ? b0 : R.hb;
        R.hb = b1 > R.hb ? b1 : R.hb;
    }
}

// ---- one time step of an owner thread -------------------------------------------------------------------------------
// Events of a row by age (tau - q):  D - 1 - LS of the PREVIOUS occupant: coefficients requested | -1: coefficients
// into the registers | 0: iterate enters the ring | 2, 4, .., 4 S: half-sweeps 1 .. 2 S (odd ones relax colour 0)
// | 4 S + 1 ..: the epilogue warps write the row | 4 S + 2 = D - 2: residual of colour 0, then the slot is recycled.
// The two slots of a warp are one step apart: in every step one of them has an even age (iterate, relaxation,
// residual) and the other an odd one (request or fill, mostly nothing).
template <int CB, int C, int K>
DD_HD void dd_wave_even(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int a, int q,
                        int lane, int S4, double omega) {
    if (a == 0) {
        dd_wave_xinit<CB, C, K>(A, sg, R, sm, q, lane);
    } else if (a <= S4) {
        if (q < sg.nq) {
            if (a & 2)
                dd_wave_relax<CB, C, K, 0, 0>(A, sg, R, sm, q, lane, omega);
            else if (a == S4 && A.last_pass)
                dd_wave_relax<CB, C, K, 1, 1>(A, sg, R, sm, q, lane, omega);
            else
                dd_wave_relax<CB, C, K, 1, 0>(A, sg, R, sm, q, lane, omega);
        }
    } else {
        dd_wave_resid0<CB, C, K>(A, sg, R, sm, q, lane);
    }
}

template <int CB, int C, int K>
DD_HD void dd_wave_odd(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int a, int q,
                       int lane, int D, double fT) {
    if (a == -1)
        dd_wave_fill<CB, C, K>(A, sg, R, sm, q, lane, fT);
    else if (a == D - 1 - DD_WAVE_LS)
        dd_wave_request<CB, C, K>(A, sg, sm, q + D, lane);
}

template <int CB, int C>
DD_HD void dd_wave_thread_step(const WaveArgs& A, const WaveSeg& sg, WaveRegs<CB, C>& R, const WaveSmem& sm, int lane,
                               int D, int S4, double omega, double fT) {
    const int a0 = R.a0, q0 = R.q0;
    if (a0 & 1) {
        // slot 1 has the even age: a0 - 1, or D - 2 (row q0 + 1 - D, about to leave) while slot 0 is being refilled
        const bool wrap = a0 < 0;
        dd_wave_even<CB, C, 1>(A, sg, R, sm, wrap ? D - 2 : a0 - 1, wrap ? q0 + 1 - D : q0 + 1, lane, S4, omega);
        dd_wave_odd<CB, C, 0>(A, sg, R, sm, a0, q0, lane, D, fT);
    } else {
        dd_wave_even<CB, C, 0>(A, sg, R, sm, a0, q0, lane, S4, omega);
        dd_wave_odd<CB, C, 1>(A, sg, R, sm, a0 - 1, q0 + 1, lane, D, fT);
    }
    R.a0 = a0 + 1;
    if (a0 == D - 2) {
        R.a0 = -1;
        R.q0 = q0 + D;
    }
}

// ---- epilogue warps ------------------------------------------------------------------------------------------------
// Row qf = tau - 4 S - 1 received its last relaxation in the previous step.  Epilogue warp ew owns a quarter of the
// strip's owned columns: lanes along consecutive columns, so v* is read and v_new = v* + x (last pass) or x
// (other passes) written with coalesced accesses; v* is requested LS steps ahead into a staging ring.
DD_HD void dd_wave_epi_init(const WaveArgs& A, const WaveSeg& sg, WaveEpi& E, const WaveSmem& sm, int ew, int lane,
                            int C) {
    const int PLB = (32 * C + 2) * 8, VW = 32 * DD_WAVE_NE * DD_WAVE_EV, D = 2 * A.nwo;
    const int per = (A.tj + DD_WAVE_NE - 1) / DD_WAVE_NE;  // <= 32 * EV
    const int t0 = ew * per + lane;                        // first owned column of this lane (strip-local, from c0)
    E.hx = E.hv = 0u;
    E.emask = 0u;
#pragma unroll
    for (int e = 0; e < DD_WAVE_EV; ++e)
        if (lane + 32 * e < per && t0 + 32 * e < sg.tc) {
            const int j = sg.c0 + t0 + 32 * e;
            E.emask |= 1u << e;
            if (j > 0 && j < A.g.M) E.emask |= 256u << e;
        }
    // state of the step tau = 4 S + 1 in which row q = 0 is finished; the request side runs LS rows ahead
    E.xa = sm.xaddr + (unsigned)((1 + ((A.halo + t0) >> 1)) * 8);
    E.xa_end = E.xa + (unsigned)(D * 2 * PLB);
    E.vsa = sm.vaddr + (unsigned)(t0 * 8);
    E.vra = E.vsa;  // row q = 0 is the first one requested (the pointers below start there, too)
    E.vs_end = E.vsa + (unsigned)(DD_WAVE_SR * VW * 8);
    E.co = (A.g.row0 + sg.rs + A.halo + t0) & 1;  // cbase is even
    const long long o = sg.mo + (long long)sg.rs * A.g.ld + sg.c0 + t0;
    E.src = A.vstar + o;
    E.dst = A.last_pass ? A.vnew + o : A.xout + sg.moR + (long long)sg.rs * A.ldR + sg.c0 + t0;
}

// tau0 = tau - (4 S + 1) = the row finished in this step (march coordinates)
template <int C>
DD_HD void dd_wave_epi_step(const WaveArgs& A, const WaveSeg& sg, WaveEpi& E, const WaveSmem& sm, int qf) {
    constexpr int PLB = (32 * C + 2) * 8, VW = 32 * DD_WAVE_NE * DD_WAVE_EV;
    const int qlo = sg.r0 - sg.rs, qhi = sg.r1 - sg.rs;  // owned rows in march coordinates
    if (A.last_pass) {
        // v* of the row finished LS steps from now (each lane stages the values it will use itself)
        const int qr = qf + DD_WAVE_LS;
        if (qr >= 0) {
            if (qr >= qlo && qr < qhi) {
#pragma unroll
                for (int e = 0; e < DD_WAVE_EV; ++e)
                    if (E.emask & (1u << e)) dd_wave_cp8s(sm, E.vra + 32 * 8 * e, E.src + 32 * e);
            }
            E.src += A.g.ld;
            E.vra += VW * 8;
            if (E.vra == E.vs_end) E.vra -= DD_WAVE_SR * VW * 8;
        }
        DD_WAVE_COMMIT();
        DD_WAVE_WAIT(DD_WAVE_LS);
    }
    if (qf < 0) return;
    if (qf >= qlo && qf < qhi) {
        const int gi = A.g.row0 + sg.rs + qf;
        const bool irow = gi > 0 && gi < A.g.N;
#pragma unroll
        for (int e = 0; e < DD_WAVE_EV; ++e) {
            if (!(E.emask & (1u << e))) continue;
            const double x = dd_wave_lds(sm, E.xa + (unsigned)(E.co * PLB + 16 * 8 * e));
            if (A.last_pass) {
                const double vs = dd_wave_lds(sm, E.vsa + 32 * 8 * e);
                const double vn = dd_newton_update(irow && (E.emask & (256u << e)), vs, x, A.zero_boundary);
                E.dst[32 * e] = vn;
                const unsigned hx = dd_wave_hi(x), hv = dd_wave_hi(vn);
                E.hx = hx > E.hx ? hx : E.hx;
                E.hv = hv > E.hv ? hv : E.hv;
            } else {
                E.dst[32 * e] = x;
            }
        }
    }
    // on to the next row
    E.co ^= 1;
    E.dst += A.last_pass ? A.g.ld : A.ldR;
    E.xa += 2 * PLB;
    if (E.xa == E.xa_end) E.xa -= (unsigned)(4 * A.nwo * PLB);
    E.vsa += VW * 8;
    if (E.vsa == E.vs_end) E.vsa -= DD_WAVE_SR * VW * 8;
}

// number of time steps of a march: tau = -1 - LS .. (last owned row) + 4 S + 2
DD_HD int dd_wave_steps(const WaveArgs& A, const WaveSeg& sg) {
    return (sg.r1 - 1 - sg.rs) + 4 * A.sweeps + 2 + 2 + DD_WAVE_LS;
}
