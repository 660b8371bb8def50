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

#include "dd_kernels.cuh"

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

struct SolveArgs {
    DDGeom g;
    const DDMember* mem;
    const double *bb, *aW, *aE, *aS, *aN;
    const double* xin;
    double* xout;
    const double* vstar;
    double* vnew;
    DDSolveStats* stats;
    int zero_boundary;
    int own0, own1;  // local rows that are tiled (owned rows of a slab; all rows otherwise)
    int vr0, vr1;    // local rows holding valid assembled rows
    int sweeps, halo, tile_i, tile_j, tiles_i, tiles_j, last_pass;
};

extern __shared__ double dd_smem[];

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
        const long long orow = mo + (long long)(rbase + si) * g.ld + cbase;
        for (int sj = lane; sj < SJ; sj += 32) {
            const bool in = rowin && sj >= lj0 && sj < lj1;
            const long long o = in ? orow + sj : mo;  // any valid address when nothing is read
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
    const double rho = A.stats[member].rho;
    __pipeline_wait_prior(0);
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
        const double* xo = X.c(1 - colour);
        double* xc = X.c(colour);
        const double *bc = Bb.c(colour), *wc = AW.c(colour), *ec = AE.c(colour), *sc = AS.c(colour),
                     *nc = AN.c(colour);
        const int ui0 = shrink_lo ? min(li0 + hs, H + 1) : li0;
        const int ui1 = shrink_hi ? max(li1 - hs, H + 1 + tr) : li1;
        // one warp per row, lanes along the packed columns: unit-stride shared-memory accesses, no divisions.
        // Packed slots holding the zero ring columns (sj = 0, SJ-1) are skipped by the column guard.
        for (int si = ui0 + warp; si < ui1; si += nwarps) {
            const int o = (colour + par0 + si) & 1;  // sj = 2 pk + o
            const int rowp = si * PW;
            double rw = 0.0, re = 0.0;
            if (CONST_BAND) {
                rw = rowW[si];
                re = rowE[si];
            }
            for (int pk = lane; pk < PW; pk += 32) {
                const int sj = 2 * pk + o;
                if (sj >= 1 && sj <= SJ - 2) {
                    const int p = rowp + pk;
                    double gs;
                    if (CONST_BAND)
                        gs = bc[p] + wc[p] * (rw * xo[p - PW] + re * xo[p + PW] + colS[sj] * xo[p - 1 + o] +
                                              colN[sj] * xo[p + o]);
                    else
                        gs = bc[p] + wc[p] * xo[p - PW] + ec[p] * xo[p + PW] + sc[p] * xo[p - 1 + o] +
                             nc[p] * xo[p + o];
                    const double xv = xc[p];
                    xc[p] = xv + omega * (gs - xv);
                }
            }
        }
        __syncthreads();
    }

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
                    res = bbv + AW.c(col)[p] * (rowW[si] * xo[p - PW] + rowE[si] * xo[p + PW] +
                                                colS[sj] * xo[p - 1 + o] + colN[sj] * xo[p + o]) - x;
                else
                    res = bbv + AW.c(col)[p] * xo[p - PW] + AE.c(col)[p] * xo[p + PW] +
                          AS.c(col)[p] * xo[p - 1 + o] + AN.c(col)[p] * xo[p + o] - x;
                const bool inter = dd_is_interior(g, g.row0 + r, j);
                const double vn = dd_newton_update(inter, vs[u], x, A.zero_boundary);
                A.vnew[ogs[u]] = vn;
                rmax = nn_max(rmax, res);
                xmax = nn_max(xmax, x);
                vmax = nn_max(vmax, vn);
                bmax = nn_max(bmax, bbv);
            } else {
                A.xout[ogs[u]] = x;
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
}

cudaError_t dd_solver_configure() {
    cudaError_t e = cudaFuncSetAttribute(k_rbsor_tile<0>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
    if (e != cudaSuccess) return e;
    return cudaFuncSetAttribute(k_rbsor_tile<1>, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t dd_launch_solve_pass(const DDLaunch& L, const DDGeom& g, const DDMember* mem, const DDRows& R,
                                 const double* xin,
                                 double* xout, const double* vstar, double* vnew, int zero_boundary,
                                 DDSolveStats* stats, const DDSolvePlan& P) {
    SolveArgs A;
    A.g = g;
    A.mem = mem;
    A.bb = R.bb;
    A.aW = R.aW;
    A.aE = R.aE;
    A.aS = R.aS;
    A.aN = R.aN;
    A.xin = xin;
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
    const long long nblocks = (long long)A.tiles_i * A.tiles_j * L.nmembers;
    if (nblocks <= 0 || nblocks > 2147483647LL) return cudaErrorInvalidConfiguration;
    if (P.const_band)
        k_rbsor_tile<1><<<(unsigned)nblocks, P.threads, P.smem_bytes, L.stream>>>(A);
    else
        k_rbsor_tile<0><<<(unsigned)nblocks, P.threads, P.smem_bytes, L.stream>>>(A);
    return cudaGetLastError();
}
