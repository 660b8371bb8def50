// dd_solver.cu -- matrix-free solve of the Newton systems (sm_100a).
//
// The reference factorises A = 2I - dt J (five diagonals, block-tridiagonal)
// with SuperLU (src/prob1base.py:2103,2130).  Here the system is kept in
// Jacobi-scaled stencil form x_c = bb + aW x_w + aE x_e + aS x_s + aN x_n and
// solved by red-black SOR with omega = 2 / (1 + sqrt(1 - rho^2)), rho = the
// Gershgorin ratio reduced on the device by the assemble kernel.
//
// One CTA owns a tile of the node grid, stages bb, the four bands and x in
// shared memory together with a halo of 2 cells per sweep, runs all sweeps of
// the pass on chip (the valid region shrinks by one ring per half-sweep, so the
// tile itself always equals the global red-black iteration bit for bit,
// independent of the tiling), and on the last pass also evaluates the true
// residual and writes v_new = v* + x.  Colour = parity of the GLOBAL (i + j).
//
// Boundary nodes carry a zero row (bb = a* = 0), so x stays 0 there and no
// special casing is needed; outside the grid the staging pads with zeros.
#include <cuda_pipeline.h>

#include <stdio.h>
#include <stdlib.h>
#include <string.h>

#include <map>
#include <mutex>
#include <tuple>

#include <cuda.h>
#include <cudaTypedefs.h>

#include "dd_kernels.cuh"
#include "dd_sor.cuh"

__device__ __forceinline__ void atomic_max_nn(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(fabs(v)));
}

// max of two non-negative doubles by bit pattern: NaN (largest pattern) is sticky
__device__ __forceinline__ double nn_max(double a, double b) {
    const unsigned long long x = (unsigned long long)__double_as_longlong(fabs(a));
    const unsigned long long y = (unsigned long long)__double_as_longlong(fabs(b));
    return __longlong_as_double((long long)(x > y ? x : y));
}

__device__ __forceinline__ double warp_max_nn(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
        b = other > b ? other : b;
    }
    return __longlong_as_double((long long)b);
}

// shared-memory access by 32-bit shared-window address: the sweep loop is integer-issue bound when the
// compiler rebuilds generic 64-bit addresses per array and per row (profiles/README.md, r01 solver capture)
__device__ __forceinline__ double lds64(unsigned a) {
    double v;
    asm volatile("ld.shared.f64 %0, [%1];" : "=d"(v) : "r"(a));
    return v;
}
__device__ __forceinline__ void sts64(unsigned a, double v) {
    asm volatile("st.shared.f64 [%0], %1;" ::"r"(a), "d"(v) : "memory");
}

struct SolveArgs {
    DDGeom g;
    const DDMember* mem;
    const double *bb, *aW, *aE, *aS, *aN;
    const double* xin;
    const double* vold;  // first pass only, register kernel: initial iterate vstar - vold (xin is null then)
    double* xout;
    const double* vstar;
    double* vnew;
    DDSolveStats* stats;
    int zero_boundary;
    int ldR;            // row pitch of bb / a* / xin / xout (even)
    long long mstrideR; // their member stride
    int own0, own1;  // local rows that are tiled (owned rows of a slab; all rows otherwise)
    int vr0, vr1;    // local rows holding valid assembled rows
    int sweeps, halo, tile_i, tile_j, tiles_i, tiles_j, last_pass;
    int nblocks;  // tiles x members (the pipelined kernel walks them with a persistent grid)
    double rho_fix;  // >= 0: use this Gershgorin ratio for omega (same on every rank of a slab mesh)
};

extern __shared__ double dd_smem[];

#ifdef DD_SOLVER_TIMING
// development instrumentation: cycles per phase summed over CTAs (staging, sweeps, epilogue, count)
__device__ unsigned long long dd_solver_cycles[4];
#define DD_TICK(k)                                                                   \
    do {                                                                             \
        __syncthreads();                                                             \
        if (threadIdx.x == 0) {                                                      \
            const long long now_ = clock64();                                        \
            atomicAdd(&dd_solver_cycles[k], (unsigned long long)(now_ - tick_));     \
            tick_ = now_;                                                            \
        }                                                                            \
    } while (0)
#else
#define DD_TICK(k)
#endif

// Shared-memory layout: every staged array is split by colour and packed along j, so that a half-sweep
// touches unit-stride words only (no bank conflicts): cell (si, sj) of colour c = (par0 + si + sj) & 1
// lives at [c][si][sj >> 1].  PW = SJ / 2 packed columns per row.
struct Packed {
    double* base;
    int plane;  // SI * PW
    __device__ __forceinline__ double* c(int colour) const { return base + colour * plane; }
};

template <int CONST_BAND>
__global__ void __launch_bounds__(512) k_rbsor_tile(SolveArgs A) {
    const DDGeom& g = A.g;
    const int tiles = A.tiles_i * A.tiles_j;
    const int member = blockIdx.x / tiles;
    const int t = blockIdx.x - member * tiles;
    const int ti = t / A.tiles_j, tj = t - ti * A.tiles_j;
    const int H = A.halo;
    // tile in LOCAL rows / columns; the node grid is nrows x (M+1)
    const int r0 = A.own0 + ti * A.tile_i, c0 = tj * A.tile_j;
    const int tr = min(A.tile_i, A.own1 - r0), tc = min(A.tile_j, g.M + 1 - c0);
    // staged region: tile +- H plus one ring of zeros; SJ rounded up to even
    const int SI = A.tile_i + 2 * H + 2, SJ = (A.tile_j + 2 * H + 3) & ~1, PW = SJ >> 1;
    const int rbase = r0 - H - 1, cbase = c0 - H - 1;
    const int par0 = (g.row0 + rbase + cbase) & 1;  // colour of smem cell (0, 0)
    const int plane = SI * PW;
    // CONST_BAND: AW holds dinv; AE/AS/AN are not staged, the geometry factors live in four 1-D arrays
    Packed X = {dd_smem, plane}, Bb = {dd_smem + 2 * plane, plane}, AW = {dd_smem + 4 * plane, plane},
           AE = {dd_smem + 6 * plane, plane}, AS = {dd_smem + 8 * plane, plane}, AN = {dd_smem + 10 * plane, plane};
    double* rowW = dd_smem + 6 * plane;  // [SI], [SI], [SJ], [SJ]  (CONST_BAND only)
    double* rowE = rowW + SI;
    double* colS = rowE + SI;
    double* colN = colS + SJ;
    const long long mo = member * g.mstride;
    const int nthreads = blockDim.x;
#ifdef DD_SOLVER_TIMING
    long long tick_ = clock64();
#endif

    // extent of real data inside the staged region (smem coordinates, half-open)
    const int vi0 = max(1, A.vr0 - rbase), vi1 = min(SI - 1, A.vr1 - rbase);
    const int vj0 = max(1, -cbase), vj1 = min(SJ - 1, g.M + 1 - cbase);
    const int li0 = vi0, li1 = min(vi1, tr + 2 * H + 1);
    const int lj0 = vj0, lj1 = min(vj1, tc + 2 * H + 1);

    // stage with 8-byte asynchronous copies (LDGSTS): every thread queues all its elements at once and waits
    // a single time, instead of paying one memory round trip per loop iteration.  One warp per staged row,
    // lanes along j (coalesced), destination scattered by colour; out-of-range elements are zero-filled.
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5, nwarps = nthreads >> 5;
    for (int si = warp; si < SI; si += nwarps) {
        const bool rowin = si >= li0 && si < li1;
        const long long moR = member * A.mstrideR;
        const long long orow = moR + (long long)(rbase + si) * A.ldR + cbase;
        for (int sj = lane; sj < SJ; sj += 32) {
            const bool in = rowin && sj >= lj0 && sj < lj1;
            const long long o = in ? orow + sj : moR;  // any valid address when nothing is read
            const size_t zf = in ? 0 : 8;
            const int col = (par0 + si + sj) & 1, q = col * plane + si * PW + (sj >> 1);
            __pipeline_memcpy_async(&Bb.base[q], &A.bb[o], 8, zf);
            __pipeline_memcpy_async(&AW.base[q], &A.aW[o], 8, zf);
            if (!CONST_BAND) {
                __pipeline_memcpy_async(&AE.base[q], &A.aE[o], 8, zf);
                __pipeline_memcpy_async(&AS.base[q], &A.aS[o], 8, zf);
                __pipeline_memcpy_async(&AN.base[q], &A.aN[o], 8, zf);
            }
            if (A.xin)
                __pipeline_memcpy_async(&X.base[q], &A.xin[o], 8, zf);
            else
                X.base[q] = 0.0;
        }
    }
    __pipeline_commit();
    if (CONST_BAND) {
        const DDMember& mb = A.mem[member];
        const double f = mb.dt * mb.m.DT;
        for (int si = threadIdx.x; si < SI; si += nthreads) {
            const int i = g.row0 + rbase + si;
            const bool in = i >= 1 && i <= g.N - 1;
            rowW[si] = in ? f * g.rhp[i] * g.rh[i] : 0.0;
            rowE[si] = in ? f * g.rhp[i] * g.rh[i + 1] : 0.0;
        }
        for (int sj = threadIdx.x; sj < SJ; sj += nthreads) {
            const int j = cbase + sj;
            const bool in = j >= 1 && j <= g.M - 1;
            colS[sj] = in ? f * g.rkp[j] * g.rk[j] : 0.0;
            colN[sj] = in ? f * g.rkp[j] * g.rk[j + 1] : 0.0;
        }
    }
    const double rho = A.rho_fix >= 0.0 ? A.rho_fix : A.stats[member].rho;
    __pipeline_wait_prior(0);
    DD_TICK(0);
    // omega_opt of SOR for a consistently ordered matrix whose Jacobi spectral radius is <= rho
    double omega = 1.0;
    if (rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - rho * rho));
    __syncthreads();

    // half-sweeps over rows 1 .. SI-2 (the zero ring is never updated).  Cells whose inputs are not yet
    // (or no longer) exact are updated too -- they only ever influence cells outside the final tile,
    // because invalidity travels one cell per half-sweep from the staged edge; rows that can no longer
    // matter are skipped.  After 2S half-sweeps the tile (+1 ring on the last pass) equals the global
    // red-black iteration exactly.
    const bool shrink_lo = !(g.row0 + rbase + li0 == 0), shrink_hi = !(g.row0 + rbase + li1 == g.N + 1);
    for (int hs = 1; hs <= 2 * A.sweeps; ++hs) {
        const int colour = (hs - 1) & 1;
        // element offsets inside dd_smem: arrays are 2 * plane doubles apart, colours plane doubles apart
        const int o_xc = colour * plane, o_xo = (1 - colour) * plane;
        const int o_bb = o_xc + 2 * plane, o_w = o_xc + 4 * plane, o_e = o_xc + 6 * plane, o_s = o_xc + 8 * plane,
                  o_n = o_xc + 10 * plane;
        const int ui0 = shrink_lo ? min(li0 + hs, H + 1) : li0;
        const int ui1 = shrink_hi ? max(li1 - hs, H + 1 + tr) : li1;
        // One warp per row, lanes along the packed columns (unit-stride shared-memory accesses).  The sweep is
        // latency bound (one dependent load -> 4 FMA -> store chain per cell), so every warp works on RU rows at
        // once: all loads first, then the arithmetic, then the stores.  Reads touch only the other colour, writes
        // only this one, so the batched order is exact.  Zero-ring columns (sj = 0, SJ-1) are skipped.
        constexpr int RU = 4;
        for (int sib = ui0 + warp; sib < ui1; sib += RU * nwarps) {
            for (int pk = lane; pk < PW; pk += 32) {
                double xw[RU], xe[RU], xs[RU], xn[RU], bv[RU], wv[RU], ev[RU], sv[RU], nv[RU], xv[RU];
                bool on[RU];
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int si = sib + u * nwarps;
                    const int o = (colour + par0 + si) & 1, sj = 2 * pk + o;
                    on[u] = si < ui1 && sj >= 1 && sj <= SJ - 2;
                    const int c = on[u] ? si * PW + pk : PW + 1;  // any in-range cell when masked off
                    xw[u] = dd_smem[o_xo + c - PW];
                    xe[u] = dd_smem[o_xo + c + PW];
                    xs[u] = dd_smem[o_xo + c + o - 1];
                    xn[u] = dd_smem[o_xo + c + o];
                    bv[u] = dd_smem[o_bb + c];
                    wv[u] = dd_smem[o_w + c];
                    xv[u] = dd_smem[o_xc + c];
                    if (!CONST_BAND) {
                        ev[u] = dd_smem[o_e + c];
                        sv[u] = dd_smem[o_s + c];
                        nv[u] = dd_smem[o_n + c];
                    }
                }
#pragma unroll
                for (int u = 0; u < RU; ++u) {
                    const int si = sib + u * nwarps;
                    double gs;
                    if (CONST_BAND) {
                        const int sj = 2 * pk + ((colour + par0 + si) & 1);
                        const int sic = on[u] ? si : 0, sjc = on[u] ? sj : 0;
                        gs = dd_sor_dT(bv[u], wv[u], rowW[sic], rowE[sic], colS[sjc], colN[sjc], xw[u], xe[u], xs[u],
                                       xn[u], xv[u]);
                    } else {
                        gs = dd_sor_d5(bv[u], wv[u], ev[u], sv[u], nv[u], xw[u], xe[u], xs[u], xn[u], xv[u]);
                    }
                    if (on[u]) dd_smem[o_xc + si * PW + pk] = dd_sor_relax(xv[u], gs, omega);
                }
            }
        }
        __syncthreads();
    }

    DD_TICK(1);
    // epilogue on the tile itself (smem coordinates H+1 .. H+1+tr); v* is fetched four cells ahead so that
    // the loads of a batch are in flight together
    double rmax = 0.0, xmax = 0.0, vmax = 0.0, bmax = 0.0;
    const int ncell = tr * tc;
    for (int base = threadIdx.x; base < ncell; base += 4 * nthreads) {
        double vs[4];
        long long ogs[4];
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * nthreads;
            vs[u] = 0.0;
            ogs[u] = 0;
            if (idx < ncell) {
                const int a = idx / tc, bcol = idx - a * tc;
                ogs[u] = mo + (long long)(r0 + a) * g.ld + (c0 + bcol);
                if (A.last_pass) vs[u] = A.vstar[ogs[u]];
            }
        }
#pragma unroll
        for (int u = 0; u < 4; ++u) {
            const int idx = base + u * nthreads;
            if (idx >= ncell) continue;
            const int a = idx / tc, bcol = idx - a * tc;
            const int si = H + 1 + a, sj = H + 1 + bcol;
            const int col = (par0 + si + sj) & 1, o = (sj & 1);
            const int pk = sj >> 1, p = si * PW + pk;
            const int r = r0 + a, j = c0 + bcol;
            const double x = X.c(col)[p];
            if (A.last_pass) {
                const double* xo = X.c(1 - col);
                const double bbv = Bb.c(col)[p];
                double res;
                if (CONST_BAND)
                    res = dd_sor_dT(bbv, AW.c(col)[p], rowW[si], rowE[si], colS[sj], colN[sj], xo[p - PW], xo[p + PW],
                                    xo[p - 1 + o], xo[p + o], x);
                else
                    res = dd_sor_d5(bbv, AW.c(col)[p], AE.c(col)[p], AS.c(col)[p], AN.c(col)[p], xo[p - PW], xo[p + PW],
                                    xo[p - 1 + o], xo[p + o], x);
                const bool inter = dd_is_interior(g, g.row0 + r, j);
                const double vn = dd_newton_update(inter, vs[u], x, A.zero_boundary);
                A.vnew[ogs[u]] = vn;
                rmax = nn_max(rmax, res);
                xmax = nn_max(xmax, x);
                vmax = nn_max(vmax, vn);
                bmax = nn_max(bmax, bbv);
            } else {
                A.xout[member * A.mstrideR + (long long)r * A.ldR + j] = x;
            }
        }
    }
    if (A.last_pass) {
        rmax = warp_max_nn(rmax);
        xmax = warp_max_nn(xmax);
        vmax = warp_max_nn(vmax);
        bmax = warp_max_nn(bmax);
        if ((threadIdx.x & 31) == 0) {
            atomic_max_nn(&A.stats[member].resid, rmax);
            atomic_max_nn(&A.stats[member].xmax, xmax);
            atomic_max_nn(&A.stats[member].vmax, vmax);
            atomic_max_nn(&A.stats[member].bmax, bmax);
        }
    }
    DD_TICK(2);
#ifdef DD_SOLVER_TIMING
    if (threadIdx.x == 0) atomicAdd(&dd_solver_cycles[3], 1ull);
#endif
}


// ---------------------------------------------------------------------------
// Register-resident variant.  The shared-memory kernel above is bound by shared-memory bandwidth (11 words
// per cell update).  Here every thread owns fixed cells for the whole pass -- rows {warp + 16 k}, packed
// column = lane, both colours -- and keeps their coefficients in registers, loaded straight from global
// memory; only x lives in shared memory (5 words per update).  Staged region: 16 * RPW rows x 64 columns.
// Same exactness argument as above (invalid data moves one cell per half-sweep).
// ---------------------------------------------------------------------------
#define DD_REG_WARPS 16
#define DD_REG_SJ 64
#define DD_REG_PW 32

// ---- bulk-copy (TMA, non-tensor form) pipeline of the persistent variant ------------------------------------
// The coefficient rows of the NEXT tile are fetched into shared memory by cp.async.bulk while the current tile
// sweeps: every staged row of a row array is one contiguous, 16-byte aligned segment (even pitch, even
// first column).  One mbarrier (arrival count 1 + transaction bytes) per CTA; its phase flips once per tile.
__device__ __forceinline__ unsigned dd_smem_u32(const void* p) { return (unsigned)__cvta_generic_to_shared(p); }
__device__ __forceinline__ void dd_mbar_init(unsigned long long* bar, unsigned count) {
    asm volatile("mbarrier.init.shared::cta.b64 [%0], %1;" ::"r"(dd_smem_u32(bar)), "r"(count) : "memory");
    asm volatile("fence.mbarrier_init.release.cluster;" ::: "memory");
}
__device__ __forceinline__ void dd_mbar_expect_tx(unsigned long long* bar, unsigned bytes) {
    asm volatile("mbarrier.arrive.expect_tx.shared::cta.b64 _, [%0], %1;" ::"r"(dd_smem_u32(bar)), "r"(bytes)
                 : "memory");
}
__device__ __forceinline__ void dd_mbar_wait(unsigned long long* bar, unsigned phase) {
    unsigned done = 0;
    while (!done) {
        asm volatile(
            "{\n"
            ".reg .pred p;\n"
            "mbarrier.try_wait.parity.shared::cta.b64 p, [%1], %2;\n"
            "selp.u32 %0, 1, 0, p;\n"
            "}\n"
            : "=r"(done)
            : "r"(dd_smem_u32(bar)), "r"(phase)
            : "memory");
    }
}
__device__ __forceinline__ void dd_bulk_g2s(void* dst, const void* src, unsigned bytes, unsigned long long* bar) {
    asm volatile("cp.async.bulk.shared::cluster.global.mbarrier::complete_tx::bytes [%0], [%1], %2, [%3];" ::"r"(
                     dd_smem_u32(dst)),
                 "l"(src), "r"(bytes), "r"(dd_smem_u32(bar))
                 : "memory");
}

// 2-D tiled TMA load: box (64 columns x SI rows) at (column c0, row c1) of a row array viewed as
// [members * nrows][ldR]; out-of-range elements are zero-filled and still counted in the transaction bytes
struct DDTileMaps {
    CUtensorMap m[5];  // bb, aW, aE, aS, aN
};
__device__ __forceinline__ void dd_tma_load_2d(void* dst, const CUtensorMap* map, int c0, int c1,
                                               unsigned long long* bar) {
    asm volatile(
        "cp.async.bulk.tensor.2d.shared::cluster.global.tile.mbarrier::complete_tx::bytes [%0], [%1, {%2, %3}], [%4];" ::
            "r"(dd_smem_u32(dst)),
        "l"(map), "r"(c0), "r"(c1), "r"(dd_smem_u32(bar))
        : "memory");
}

__device__ __forceinline__ void dd_prefetch_l2(const void* p) {
    asm volatile("prefetch.global.L2 [%0];" ::"l"(p));
}

// Coefficients of the thread's own cells into registers (index [k][colour]) and the initial iterate into the
// colour planes.  The two cells of a packed column are adjacent in memory and the row arrays have an even
// pitch, so one aligned 16-byte load fetches both (cbase is even: tile_j even, H odd).  FLIP = colour of the
// even column.  Cells outside the valid rows / the grid keep zero coefficients and a zero iterate.
template <int CONST_BAND, int RPW, int FLIP, int PIPE>
__device__ __forceinline__ void reg_load_cells(const SolveArgs& A, const DDGeom& g, const double* buf, long long moR,
                                               long long mo,
                                               int rbase, int colj, bool pair_ok, int li0, int li1, int lj0,
                                               int lj1, int lane, int warp, double (&cb)[RPW][2],
                                               double (&cw)[RPW][2], double (&ce)[RPW][2], double (&cs)[RPW][2],
                                               double (&cn)[RPW][2], double* sx) {
    constexpr int SI = DD_REG_WARPS * RPW, PW = DD_REG_PW, plane = SI * PW;
    constexpr int c0 = FLIP ? 1 : 0, c1 = 1 - c0;  // colours of the even / odd column
    const int sj0 = 2 * lane;
    const bool in0 = sj0 >= lj0 && sj0 < lj1, in1 = sj0 + 1 >= lj0 && sj0 + 1 < lj1;
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
        const int si = warp + DD_REG_WARPS * k;
        const int row = rbase + si;
        const bool rowok = pair_ok && si >= li0 && si < li1 && (in0 || in1);
        double2 vb = make_double2(0.0, 0.0), vw = vb, ve = vb, vs2 = vb, vn = vb, vx = vb;
        if (rowok) {
            const long long o = moR + (long long)row * A.ldR + colj;
            if (PIPE) {
                // staged by the bulk copies of the previous tile's iteration: [array][si][64 columns]
                const double* q = buf + si * DD_REG_SJ + 2 * lane;
                vb = *reinterpret_cast<const double2*>(q);
                vw = *reinterpret_cast<const double2*>(q + SI * DD_REG_SJ);
                if (!CONST_BAND) {
                    ve = *reinterpret_cast<const double2*>(q + 2 * SI * DD_REG_SJ);
                    vs2 = *reinterpret_cast<const double2*>(q + 3 * SI * DD_REG_SJ);
                    vn = *reinterpret_cast<const double2*>(q + 4 * SI * DD_REG_SJ);
                }
            } else {
                vb = *reinterpret_cast<const double2*>(A.bb + o);
                vw = *reinterpret_cast<const double2*>(A.aW + o);
                if (!CONST_BAND) {
                    ve = *reinterpret_cast<const double2*>(A.aE + o);
                    vs2 = *reinterpret_cast<const double2*>(A.aS + o);
                    vn = *reinterpret_cast<const double2*>(A.aN + o);
                }
            }
            if (A.xin) {
                vx = *reinterpret_cast<const double2*>(A.xin + o);
            } else if (A.vold) {
                // the previous step's increment as initial iterate (states have the grid's own row pitch)
                const int gi = g.row0 + row;
                if (gi >= 1 && gi <= g.N - 1) {
                    const long long og = mo + (long long)row * g.ld + colj;
                    if (colj >= 1 && colj <= g.M - 1) vx.x = A.vstar[og] - A.vold[og];
                    if (colj + 1 >= 1 && colj + 1 <= g.M - 1) vx.y = A.vstar[og + 1] - A.vold[og + 1];
                }
            }
            if (!(in0 && in1)) {  // only the pair that straddles the left / right end of the grid
                if (!in0) vb.x = vw.x = ve.x = vs2.x = vn.x = vx.x = 0.0;
                if (!in1) vb.y = vw.y = ve.y = vs2.y = vn.y = vx.y = 0.0;
            }
        }
        cb[k][c0] = vb.x;  cb[k][c1] = vb.y;
        cw[k][c0] = vw.x;  cw[k][c1] = vw.y;
        if (!CONST_BAND) {
            ce[k][c0] = ve.x;   ce[k][c1] = ve.y;
            cs[k][c0] = vs2.x;  cs[k][c1] = vs2.y;
            cn[k][c0] = vn.x;   cn[k][c1] = vn.y;
        }
        sx[c0 * plane + si * PW + lane] = vx.x;
        sx[c1 * plane + si * PW + lane] = vx.y;
    }
}

struct DDTileGeo {
    int member, r0, c0, tr, tc, rbase, cbase, li0, li1;
};

template <int RPW>
__device__ __forceinline__ DDTileGeo dd_tile_geo(const SolveArgs& A, int blk) {
    constexpr int SI = DD_REG_WARPS * RPW;
    DDTileGeo t;
    const int tiles = A.tiles_i * A.tiles_j;
    t.member = blk / tiles;
    const int q = blk - t.member * tiles;
    const int ti = q / A.tiles_j, tj = q - ti * A.tiles_j;
    t.r0 = A.own0 + ti * A.tile_i;
    t.c0 = tj * A.tile_j;
    t.tr = min(A.tile_i, A.own1 - t.r0);
    t.tc = min(A.tile_j, A.g.M + 1 - t.c0);
    t.rbase = t.r0 - A.halo - 1;
    t.cbase = t.c0 - A.halo - 1;
    // rows of the staged region that hold real data (smem coordinates, half-open)
    t.li0 = max(1, A.vr0 - t.rbase);
    t.li1 = min(min(SI - 1, A.vr1 - t.rbase), t.tr + 2 * A.halo + 1);
    return t;
}

// Starts the TMA loads of tile `blk`'s coefficient boxes into `buf` ([array][si][64]): one elected thread arms
// the barrier with the (constant) byte count and issues one instruction per array.
template <int CONST_BAND, int RPW>
__device__ __forceinline__ void dd_pipe_issue(const SolveArgs& A, const DDTileMaps& maps, int blk, double* buf,
                                              unsigned long long* bar) {
    constexpr int SI = DD_REG_WARPS * RPW, NARR = CONST_BAND ? 2 : 5;
    if (threadIdx.x != 0) return;
    const DDTileGeo t = dd_tile_geo<RPW>(A, blk);
    asm volatile("fence.proxy.async.shared::cta;" ::: "memory");  // earlier generic reads of buf vs the async writes
    dd_mbar_expect_tx(bar, (unsigned)(NARR * SI * DD_REG_SJ * sizeof(double)));
    const int row = t.member * A.g.nrows + t.rbase;
#pragma unroll
    for (int a = 0; a < NARR; ++a) dd_tma_load_2d(buf + (size_t)a * SI * DD_REG_SJ, &maps.m[a], t.cbase, row, bar);
}

template <int CONST_BAND, int RPW, int PIPE>
__global__ void __launch_bounds__(DD_REG_WARPS * 32, 1)
k_rbsor_reg(const __grid_constant__ SolveArgs A, const __grid_constant__ DDTileMaps maps) {
    const DDGeom& g = A.g;
    const int H = A.halo;
    constexpr int SI = DD_REG_WARPS * RPW, SJ = DD_REG_SJ, PW = DD_REG_PW, plane = SI * PW;
    constexpr int NARR = CONST_BAND ? 2 : 5;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double* sx = dd_smem;             // [colour][si][pk]
    double* buf = dd_smem + 2 * plane;  // PIPE: coefficient rows of one tile, [array][si][64]
    unsigned long long* bar = reinterpret_cast<unsigned long long*>(buf + NARR * SI * SJ);
    unsigned phase = 0;
    if (PIPE) {
        if (threadIdx.x == 0) dd_mbar_init(bar, 1);
        __syncthreads();
        if ((int)blockIdx.x < A.nblocks) dd_pipe_issue<CONST_BAND, RPW>(A, maps, blockIdx.x, buf, bar);
    }
    // PIPE: persistent grid, tile blk + gridDim.x is fetched while tile blk is swept; otherwise one tile per CTA
    for (int blk = blockIdx.x; blk < A.nblocks; blk += (PIPE ? (int)gridDim.x : A.nblocks)) {
    const DDTileGeo tg = dd_tile_geo<RPW>(A, blk);
    const int member = tg.member, r0 = tg.r0, c0 = tg.c0, tr = tg.tr, tc = tg.tc, rbase = tg.rbase, cbase = tg.cbase;
    const int par0 = (g.row0 + rbase + cbase) & 1;
    const long long mo = member * g.mstride;
#ifdef DD_SOLVER_TIMING
    long long tick_ = clock64();
#endif
    // real data inside the staged region (smem coordinates, half-open)
    const int li0 = tg.li0, li1 = tg.li1;
    const int lj0 = max(1, -cbase), lj1 = min(min(SJ - 1, g.M + 1 - cbase), tc + 2 * H + 1);

    // ---- load: coefficients of the thread's own cells into registers, x into shared memory -----------
    // The two cells of a packed column are adjacent in memory and the row arrays have an even pitch, so one
    // aligned 16-byte load fetches both colours (cbase is even: tile_j even, H odd).
    double cb[RPW][2], cw[RPW][2], ce[RPW][2], cs[RPW][2], cn[RPW][2];
    const long long moR = member * A.mstrideR;
    const int flip = (par0 + warp) & 1;  // colour of the even column of the pair (rows of a thread share parity)
    const int colj = cbase + 2 * lane;   // global column of the even cell
    const bool pair_ok = colj >= 0 && colj + 1 < A.ldR;
    // `flip` is uniform in the warp: one branch selects the variant that files the two cells of a pair under
    // their colours at compile time (no per-value selects)
    // Ask L2 for everything this CTA is going to read before the first load is issued: the loads below come in
    // register-limited groups, and only the first group should pay the DRAM latency.  (Prefetches hold no
    // registers and no scoreboard slot.)
    {
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
            const int si = warp + DD_REG_WARPS * k;
            if (pair_ok && si >= li0 && si < li1) {
                const long long o = moR + (long long)(rbase + si) * A.ldR + colj;
                if (!PIPE) {
                    dd_prefetch_l2(A.bb + o);
                    dd_prefetch_l2(A.aW + o);
                    if (!CONST_BAND) {
                        dd_prefetch_l2(A.aE + o);
                        dd_prefetch_l2(A.aS + o);
                        dd_prefetch_l2(A.aN + o);
                    }
                }
                if (A.xin) dd_prefetch_l2(A.xin + o);
            }
            if (A.last_pass) {
                const int a = warp + DD_REG_WARPS * k;
                if (a < tr && lane * 2 < tc) dd_prefetch_l2(A.vstar + mo + (long long)(r0 + a) * g.ld + c0 + lane * 2);
            }
        }
    }
    const double rho = A.rho_fix >= 0.0 ? A.rho_fix : A.stats[member].rho;  // requested before waiting for the staged coefficients
    if (PIPE) {
        dd_mbar_wait(bar, phase);  // this tile's coefficient rows have landed in buf
        phase ^= 1u;
    }
    if (flip)
        reg_load_cells<CONST_BAND, RPW, 1, PIPE>(A, g, buf, moR, mo, rbase, colj, pair_ok, li0, li1, lj0, lj1, lane,
                                                 warp, cb, cw, ce, cs, cn, sx);
    else
        reg_load_cells<CONST_BAND, RPW, 0, PIPE>(A, g, buf, moR, mo, rbase, colj, pair_ok, li0, li1, lj0, lj1, lane,
                                                 warp, cb, cw, ce, cs, cn, sx);
    // constant-band geometry factors of the thread's rows and of its two columns
    double rW[RPW], rE[RPW], cS2[2], cN2[2];
    if (CONST_BAND) {
        const DDMember& mb = A.mem[member];
        const double f = mb.dt * mb.m.DT;
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
            const int i = g.row0 + rbase + warp + DD_REG_WARPS * k;
            const bool ok = i >= 1 && i <= g.N - 1;
            rW[k] = ok ? f * g.rhp[i] * g.rh[i] : 0.0;
            rE[k] = ok ? f * g.rhp[i] * g.rh[i + 1] : 0.0;
        }
#pragma unroll
        for (int o = 0; o < 2; ++o) {
            const int j = cbase + 2 * lane + o;
            const bool ok = j >= 1 && j <= g.M - 1;
            cS2[o] = ok ? f * g.rkp[j] * g.rk[j] : 0.0;
            cN2[o] = ok ? f * g.rkp[j] * g.rk[j + 1] : 0.0;
        }
    }
    double omega = 1.0;
    if (rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - rho * rho));
    DD_TICK(0);
    __syncthreads();
    // buf has been consumed by every thread: fetch the next tile of this CTA while this one is swept
    if (PIPE && blk + (int)gridDim.x < A.nblocks) dd_pipe_issue<CONST_BAND, RPW>(A, maps, blk + gridDim.x, buf, bar);

    // ---- sweeps ---------------------------------------------------------------------------------------
    for (int sw = 0; sw < A.sweeps; ++sw) {
#pragma unroll
      for (int c = 0; c < 2; ++c) {  // colour is a compile-time index into the register arrays
        const double* xo = sx + (1 - c) * plane;
        double* xc = sx + c * plane;
        double xnew[RPW];
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
            const int si = warp + DD_REG_WARPS * k;
            const int o = (c + par0 + si) & 1;
            const int p = si * PW + lane;
            // rows 0 / SI-1 and columns 0 / SJ-1 form the zero ring: their cells carry zero coefficients and
            // stay 0; neighbours outside the array are never dereferenced
            const bool edge = si == 0 || si == SI - 1 || (lane == 0 && o == 0) || (lane == PW - 1 && o == 1);
            double xw = 0.0, xe = 0.0, xs = 0.0, xn = 0.0;
            if (!edge) {
                xw = xo[p - PW];
                xe = xo[p + PW];
                xs = xo[p + o - 1];
                xn = xo[p + o];
            }
            const double xv = xc[p];
            double gs;
            if (CONST_BAND)
                gs = dd_sor_dT(cb[k][c], cw[k][c], rW[k], rE[k], (o ? cS2[1] : cS2[0]), (o ? cN2[1] : cN2[0]), xw, xe,
                               xs, xn, xv);
            else
                gs = dd_sor_d5(cb[k][c], cw[k][c], ce[k][c], cs[k][c], cn[k][c], xw, xe, xs, xn, xv);
            xnew[k] = dd_sor_relax(xv, gs, omega);
        }
#pragma unroll
        for (int k = 0; k < RPW; ++k) xc[(warp + DD_REG_WARPS * k) * PW + lane] = xnew[k];
        __syncthreads();
      }
    }
    DD_TICK(1);

    // ---- epilogue ------------------------------------------------------------------------------------------
    // (1) owner threads: residual statistics of their cells inside the tile (coefficients are in registers);
    //     non-final passes store x with one aligned 16-byte store per pair
    double rmax = 0.0, xmax = 0.0, vmax = 0.0, bmax = 0.0;
#pragma unroll
    for (int k = 0; k < RPW; ++k) {
        const int si = warp + DD_REG_WARPS * k;
        const bool rowin = si >= H + 1 && si < H + 1 + tr;
        if (A.last_pass) {
#pragma unroll
            for (int c = 0; c < 2; ++c) {
                const int o = (c + par0 + si) & 1, sj = 2 * lane + o;
                if (!(rowin && sj >= H + 1 && sj < H + 1 + tc)) continue;
                const int p = si * PW + lane;
                const double* xo = sx + (1 - c) * plane;
                const double x = sx[c * plane + p];
                const double xw = xo[p - PW], xe = xo[p + PW], xs = xo[p + o - 1], xn = xo[p + o];
                double res;
                if (CONST_BAND)
                    res = dd_sor_dT(cb[k][c], cw[k][c], rW[k], rE[k], (o ? cS2[1] : cS2[0]), (o ? cN2[1] : cN2[0]), xw,
                                    xe, xs, xn, x);
                else
                    res = dd_sor_d5(cb[k][c], cw[k][c], ce[k][c], cs[k][c], cn[k][c], xw, xe, xs, xn, x);
                rmax = nn_max(rmax, res);
                xmax = nn_max(xmax, x);
                bmax = nn_max(bmax, cb[k][c]);
            }
        } else if (rowin && pair_ok) {
            const int sj0 = 2 * lane;
            const bool m0 = sj0 >= H + 1 && sj0 < H + 1 + tc, m1 = sj0 + 1 >= H + 1 && sj0 + 1 < H + 1 + tc;
            const long long o = moR + (long long)(rbase + si) * A.ldR + colj;
            const double x0 = sx[(flip ? 1 : 0) * plane + si * PW + lane];
            const double x1 = sx[(flip ? 0 : 1) * plane + si * PW + lane];
            if (m0 && m1)
                *reinterpret_cast<double2*>(A.xout + o) = make_double2(x0, x1);
            else if (m0)
                A.xout[o] = x0;
            else if (m1)
                A.xout[o + 1] = x1;
        }
    }
    // (2) v_new = v* + x: warp w takes tile rows w, w + 16, ..., its lanes consecutive columns (coalesced);
    //     all loads of v* are issued before the first store
    if (A.last_pass) {
        double vsv[RPW][2];
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
            const int a = warp + DD_REG_WARPS * k;
            const long long rowoff = mo + (long long)(r0 + a) * g.ld + c0;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int bcol = lane + 32 * cc;
                vsv[k][cc] = (a < tr && bcol < tc) ? A.vstar[rowoff + bcol] : 0.0;
            }
        }
#pragma unroll
        for (int k = 0; k < RPW; ++k) {
            const int a = warp + DD_REG_WARPS * k;
            const int si = H + 1 + a, gi = g.row0 + r0 + a;
            const bool irow = gi > 0 && gi < g.N;
            const long long rowoff = mo + (long long)(r0 + a) * g.ld + c0;
#pragma unroll
            for (int cc = 0; cc < 2; ++cc) {
                const int bcol = lane + 32 * cc;
                if (a < tr && bcol < tc) {
                    const int sj = H + 1 + bcol, gj = c0 + bcol;
                    const double x = sx[((par0 + si + sj) & 1) * plane + si * PW + (sj >> 1)];
                    const double vn = dd_newton_update(irow && gj > 0 && gj < g.M, vsv[k][cc], x, A.zero_boundary);
                    A.vnew[rowoff + bcol] = vn;
                    vmax = nn_max(vmax, vn);
                }
            }
        }
    }
    if (A.last_pass) {
        // one atomic per quantity and CTA (the four addresses are shared by every CTA of the member)
        rmax = warp_max_nn(rmax);
        xmax = warp_max_nn(xmax);
        vmax = warp_max_nn(vmax);
        bmax = warp_max_nn(bmax);
        __syncthreads();  // x planes are dead from here on: reuse the first words as scratch
        if (lane == 0) {
            sx[warp * 4 + 0] = rmax;
            sx[warp * 4 + 1] = xmax;
            sx[warp * 4 + 2] = vmax;
            sx[warp * 4 + 3] = bmax;
        }
        __syncthreads();
        if (threadIdx.x < 4) {
            double m = 0.0;
            for (int w = 0; w < DD_REG_WARPS; ++w) m = nn_max(m, sx[w * 4 + threadIdx.x]);
            double* dst = threadIdx.x == 0 ? &A.stats[member].resid
                        : threadIdx.x == 1 ? &A.stats[member].xmax
                        : threadIdx.x == 2 ? &A.stats[member].vmax : &A.stats[member].bmax;
            atomic_max_nn(dst, m);
        }
    }
    DD_TICK(2);
#ifdef DD_SOLVER_TIMING
    if (threadIdx.x == 0) atomicAdd(&dd_solver_cycles[3], 1ull);
#endif
    if (PIPE) __syncthreads();  // the x planes (and the statistics scratch in them) are reused by the next tile
    }
}

#ifdef DD_SOLVER_TIMING
extern "C" void dd_solver_timing_read(unsigned long long out[4], int reset) {
    cudaMemcpyFromSymbol(out, dd_solver_cycles, sizeof(unsigned long long) * 4);
    if (reset) {
        unsigned long long z[4] = {0, 0, 0, 0};
        cudaMemcpyToSymbol(dd_solver_cycles, z, sizeof(z));
    }
}
#endif

// initial iterate as an array (row pitch of the Newton rows) for the shared-memory kernel
__global__ void k_make_guess(DDGeom g, const double* __restrict__ vstar, const double* __restrict__ vold,
                             double* __restrict__ x0, int ldR, long long mstrideR, int r0, int r1) {
    const int j = blockIdx.x * blockDim.x + threadIdx.x, r = r0 + blockIdx.y;
    const long long member = blockIdx.z;
    if (j > g.M || r >= r1) return;
    const long long o = member * g.mstride + (long long)r * g.ld + j;
    x0[member * mstrideR + (long long)r * ldR + j] = dd_is_interior(g, g.row0 + r, j) ? vstar[o] - vold[o] : 0.0;
}

cudaError_t dd_launch_make_guess(const DDLaunch& L, const DDGeom& g, const DDRows& R, const double* vstar,
                                 const double* vold, double* x0) {
    const dim3 grid((unsigned)((g.M + 1 + 255) / 256), (unsigned)(L.vr1 - L.vr0), (unsigned)L.nmembers);
    if (grid.y == 0 || grid.y > 65535u || grid.z > 65535u) return cudaErrorInvalidConfiguration;
    k_make_guess<<<grid, 256, 0, L.stream>>>(g, vstar, vold, x0, R.ld, R.mstride, L.vr0, L.vr1);
    return cudaGetLastError();
}

// dynamic shared memory of the register-resident kernel: x planes (+ coefficient staging buffer and barrier)
static size_t dd_reg_smem_bytes(int const_band, int rpw, int pipe) {
    const size_t si = (size_t)DD_REG_WARPS * rpw;
    size_t bytes = 2 * si * DD_REG_PW * sizeof(double);
    if (pipe) bytes += (size_t)(const_band ? 2 : 5) * si * DD_REG_SJ * sizeof(double) + 16;
    return bytes;
}

cudaError_t dd_solver_configure() {
    cudaError_t e = cudaFuncSetAttribute(k_rbsor_tile<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_rbsor_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_rbsor_reg<1, 6, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 64 * 1024);
    if (e != cudaSuccess) return e;
    e = cudaFuncSetAttribute(k_rbsor_reg<1, 8, 0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 80 * 1024);
    if (e != cudaSuccess) return e;
#define DD_CFG_PIPE(CB, RPW)                                                                                  \
    e = cudaFuncSetAttribute(k_rbsor_reg<CB, RPW, 1>, cudaFuncAttributeMaxDynamicSharedMemorySize,            \
                             (int)dd_reg_smem_bytes(CB, RPW, 1));                                             \
    if (e != cudaSuccess) return e;
    DD_CFG_PIPE(1, 2) DD_CFG_PIPE(1, 3) DD_CFG_PIPE(1, 4) DD_CFG_PIPE(1, 6) DD_CFG_PIPE(1, 8)
    DD_CFG_PIPE(0, 2) DD_CFG_PIPE(0, 3) DD_CFG_PIPE(0, 4)
#undef DD_CFG_PIPE
    return cudaSuccess;
}

// ---- tensor maps of the row arrays (host) ------------------------------------------------------------------
static PFN_cuTensorMapEncodeTiled dd_encode_fn() {
    static PFN_cuTensorMapEncodeTiled fn = nullptr;
    static bool tried = false;
    if (!tried) {
        tried = true;
        void* p = nullptr;
        cudaDriverEntryPointQueryResult q;
        if (cudaGetDriverEntryPoint("cuTensorMapEncodeTiled", &p, cudaEnableDefault, &q) == cudaSuccess &&
            q == cudaDriverEntryPointSuccess)
            fn = (PFN_cuTensorMapEncodeTiled)p;
    }
    return fn;
}

// map of one row array [rows_total][ld] (doubles, even ld) with a box of 64 columns x box_rows rows
static bool dd_tile_map(const double* base, int ld, long long rows_total, int box_rows, CUtensorMap* out) {
    static std::mutex mu;
    static std::map<std::tuple<const double*, int, long long, int>, CUtensorMap> cache;
    std::lock_guard<std::mutex> lk(mu);
    const auto key = std::make_tuple(base, ld, rows_total, box_rows);
    auto it = cache.find(key);
    if (it != cache.end()) {
        *out = it->second;
        return true;
    }
    PFN_cuTensorMapEncodeTiled enc = dd_encode_fn();
    if (!enc || (ld & 1) || (reinterpret_cast<uintptr_t>(base) & 15)) return false;
    const cuuint64_t dims[2] = {(cuuint64_t)ld, (cuuint64_t)rows_total};
    const cuuint64_t strides[1] = {(cuuint64_t)ld * sizeof(double)};
    const cuuint32_t box[2] = {(cuuint32_t)DD_REG_SJ, (cuuint32_t)box_rows};
    const cuuint32_t estr[2] = {1, 1};
    CUtensorMap m;
    if (enc(&m, CU_TENSOR_MAP_DATA_TYPE_FLOAT64, 2, const_cast<double*>(base), dims, strides, box, estr,
            CU_TENSOR_MAP_INTERLEAVE_NONE, CU_TENSOR_MAP_SWIZZLE_NONE, CU_TENSOR_MAP_L2_PROMOTION_L2_256B,
            CU_TENSOR_MAP_FLOAT_OOB_FILL_NONE) != CUDA_SUCCESS)
        return false;
    if (cache.size() > 4096) cache.clear();  // freed and re-allocated work arrays leave stale keys behind
    cache[key] = m;
    *out = m;
    return true;
}

static thread_local char g_last_kernel[64] = "";
void dd_set_last_solver_kernel(const char* fmt, int a, int b, int c) { snprintf(g_last_kernel, sizeof(g_last_kernel), fmt, a, b, c); }
const char* dd_last_solver_kernel() { return g_last_kernel; }

cudaError_t dd_launch_solve_pass(const DDLaunch& L, const DDGeom& g, const DDMember* mem, const DDRows& R,
                                 const double* xin, const double* vold,
                                 double* xout, const double* vstar, double* vnew, int zero_boundary,
                                 DDSolveStats* stats, const DDSolvePlan& P) {
    if (vold && (xin || P.rpw <= 0)) return cudaErrorInvalidValue;  // see dd_launch_make_guess for the other kernel
    SolveArgs A;
    A.g = g;
    A.mem = mem;
    A.ldR = R.ld;
    A.mstrideR = R.mstride;
    A.bb = R.bb;
    A.aW = R.aW;
    A.aE = R.aE;
    A.aS = R.aS;
    A.aN = R.aN;
    A.xin = xin;
    A.vold = vold;
    A.xout = xout;
    A.vstar = vstar;
    A.vnew = vnew;
    A.stats = stats;
    A.zero_boundary = zero_boundary;
    A.sweeps = P.sweeps;
    A.halo = P.halo;
    A.tile_i = P.tile_i;
    A.tile_j = P.tile_j;
    A.own0 = L.own0;
    A.vr0 = L.vr0;
    A.vr1 = L.vr1;
    A.own1 = L.own1;
    A.tiles_i = (L.own1 - L.own0 + P.tile_i - 1) / P.tile_i;
    A.tiles_j = (g.M + 1 + P.tile_j - 1) / P.tile_j;
    A.last_pass = P.last_pass;
    A.rho_fix = P.rho_fix;
    const long long nblocks = (long long)A.tiles_i * A.tiles_j * L.nmembers;
    if (nblocks <= 0 || nblocks > 2147483647LL) return cudaErrorInvalidConfiguration;
    if (P.rpw > 0) {
        // register-resident kernel: 512 threads, x only in shared memory (2 colours x 16 rpw rows x 32 packed cols).
        // With more tiles than SMs: persistent grid, the next tile's coefficient rows arrive by bulk copies
        // while the current tile is swept (DD_SOLVER_PIPE=0 switches that off).
        static int sm_count = 0;
        if (sm_count == 0) {
            int dev = 0;
            cudaGetDevice(&dev);
            if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
                sm_count = 148;
        }
        static const bool pipe_off = getenv("DD_SOLVER_PIPE") && !strcmp(getenv("DD_SOLVER_PIPE"), "0");
        // (measured: pays for the five-array systems; the constant-band T system stages too little to gain)
        int pipe = (!pipe_off && nblocks > sm_count && !P.const_band) ? 1 : 0;
        DDTileMaps maps;
        memset(&maps, 0, sizeof(maps));
        if (pipe) {
            const double* arrs[5] = {R.bb, R.aW, R.aE, R.aS, R.aN};
            const long long rows_total = (long long)L.nmembers * g.nrows;
            for (int a = 0; a < (P.const_band ? 2 : 5) && pipe; ++a)
                if (!dd_tile_map(arrs[a], R.ld, rows_total, DD_REG_WARPS * P.rpw, &maps.m[a])) pipe = 0;
        }
        const size_t smem = dd_reg_smem_bytes(P.const_band, P.rpw, pipe);
        const unsigned grid = pipe ? (unsigned)sm_count : (unsigned)nblocks;
        A.nblocks = (int)nblocks;
        dd_set_last_solver_kernel("k_rbsor_reg<%d, %d, %d>", P.const_band, P.rpw, pipe);
#define DD_LAUNCH_REG(CB, RPW)                                                             \
    do {                                                                                   \
        if (pipe)                                                                          \
            k_rbsor_reg<CB, RPW, 1><<<grid, DD_REG_WARPS * 32, smem, L.stream>>>(A, maps); \
        else                                                                               \
            k_rbsor_reg<CB, RPW, 0><<<grid, DD_REG_WARPS * 32, smem, L.stream>>>(A, maps); \
    } while (0)
        if (P.const_band) {
            switch (P.rpw) {
                case 2: DD_LAUNCH_REG(1, 2); break;
                case 3: DD_LAUNCH_REG(1, 3); break;
                case 4: DD_LAUNCH_REG(1, 4); break;
                case 6: DD_LAUNCH_REG(1, 6); break;
                case 8: DD_LAUNCH_REG(1, 8); break;
                default: return cudaErrorInvalidValue;
            }
        } else {
            switch (P.rpw) {
                case 2: DD_LAUNCH_REG(0, 2); break;
                case 3: DD_LAUNCH_REG(0, 3); break;
                case 4: DD_LAUNCH_REG(0, 4); break;
                default: return cudaErrorInvalidValue;
            }
        }
#undef DD_LAUNCH_REG
    } else if (P.const_band) {
        dd_set_last_solver_kernel("k_rbsor_tile<%d>", 1, 0, 0);
        k_rbsor_tile<1><<<(unsigned)nblocks, P.threads, P.smem_bytes, L.stream>>>(A);
    } else {
        dd_set_last_solver_kernel("k_rbsor_tile<%d>", 0, 0, 0);
        k_rbsor_tile<0><<<(unsigned)nblocks, P.threads, P.smem_bytes, L.stream>>>(A);
    }
    return cudaGetLastError();
}
