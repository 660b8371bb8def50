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
    // staged region: tile +- H plus one ring of zeros; local row of smem row 0:
    const int SI = A.tile_i + 2 * H + 2, SJ = A.tile_j + 2 * H + 2;
    const int rbase = r0 - H - 1, cbase = c0 - H - 1;
    double* sx = dd_smem;
    double* sb = sx + (size_t)SI * SJ;
    double* sW = sb + (size_t)SI * SJ;
    double* sE = sW + (size_t)SI * SJ;
    double* sS = sE + (size_t)SI * SJ;
    double* sN = sS + (size_t)SI * SJ;
    const long long mo = member * g.mstride;
    const int nthreads = blockDim.x;

    // extent of real data inside the staged region (smem coordinates, half-open)
    const int vi0 = max(1, A.vr0 - rbase), vi1 = min(SI - 1, A.vr1 - rbase);
    const int vj0 = max(1, -cbase), vj1 = min(SJ - 1, g.M + 1 - cbase);
    // only stage what the sweeps can reach
    const int li0 = max(vi0, 1), li1 = min(vi1, tr + 2 * H + 1);
    const int lj0 = max(vj0, 1), lj1 = min(vj1, tc + 2 * H + 1);

    for (int idx = threadIdx.x; idx < SI * SJ; idx += nthreads) {
        const int si = idx / SJ, sj = idx - si * SJ;
        double x = 0.0, b = 0.0, w = 0.0, e = 0.0, s = 0.0, n = 0.0;
        if (si >= li0 && si < li1 && sj >= lj0 && sj < lj1) {
            const long long o = mo + (long long)(rbase + si) * g.ld + (cbase + sj);
            b = A.bb[o];
            w = A.aW[o];
            e = A.aE[o];
            s = A.aS[o];
            n = A.aN[o];
            if (A.xin) x = A.xin[o];
        }
        sx[idx] = x;
        sb[idx] = b;
        sW[idx] = w;
        sE[idx] = e;
        sS[idx] = s;
        sN[idx] = n;
    }
    const double rho = A.stats[member].rho;
    // omega_opt of SOR for a consistently ordered matrix whose Jacobi spectral radius is <= rho
    double omega = 1.0;
    if (rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - rho * rho));
    __syncthreads();

    // half-sweeps.  After half-sweep h (1-based) the updated colour is exact on the
    // region shrunk by h rings (sides on the physical grid edge do not shrink).
    const bool edge_lo_i = (rbase + li0 == 0);
    const bool edge_hi_i = (rbase + li1 == g.nrows) && (g.row0 + g.nrows == g.N + 1);
    const bool edge_lo_j = (cbase + lj0 == 0), edge_hi_j = (cbase + lj1 == g.M + 1);
    const bool top_is_grid_edge = edge_lo_i && (g.row0 == 0);
    for (int hs = 1; hs <= 2 * A.sweeps; ++hs) {
        const int colour = (hs - 1) & 1;
        const int ui0 = top_is_grid_edge ? li0 : li0 + hs, ui1 = edge_hi_i ? li1 : li1 - hs;
        const int uj0 = edge_lo_j ? lj0 : lj0 + hs, uj1 = edge_hi_j ? lj1 : lj1 - hs;
        const int ni = ui1 - ui0, nj = uj1 - uj0;
        if (ni > 0 && nj > 0) {
            const int halfw = (nj + 1) >> 1;
            for (int idx = threadIdx.x; idx < ni * halfw; idx += nthreads) {
                const int a = idx / halfw, bcol = idx - a * halfw;
                const int si = ui0 + a;
                // first column of this row with the right colour: global parity of (i + j)
                const int gi = g.row0 + rbase + si;
                int sj = uj0 + 2 * bcol;
                if (((gi + cbase + sj) & 1) != colour) sj += 1;
                if (sj < uj1) {
                    const int p = si * SJ + sj;
                    const double gs = sb[p] + sW[p] * sx[p - SJ] + sE[p] * sx[p + SJ] + sS[p] * sx[p - 1] +
                                      sN[p] * sx[p + 1];
                    sx[p] = sx[p] + omega * (gs - sx[p]);
                }
            }
        }
        __syncthreads();
    }

    // epilogue on the tile itself (smem coordinates H+1 .. H+1+tr)
    double rmax = 0.0, xmax = 0.0, vmax = 0.0, bmax = 0.0;
    for (int idx = threadIdx.x; idx < tr * tc; idx += nthreads) {
        const int a = idx / tc, bcol = idx - a * tc;
        const int si = H + 1 + a, sj = H + 1 + bcol;
        const int p = si * SJ + sj;
        const int r = r0 + a, j = c0 + bcol;
        const long long o = mo + (long long)r * g.ld + j;
        const double x = sx[p];
        if (A.last_pass) {
            const double res = sb[p] + sW[p] * sx[p - SJ] + sE[p] * sx[p + SJ] + sS[p] * sx[p - 1] +
                               sN[p] * sx[p + 1] - x;
            const bool inter = dd_is_interior(g, g.row0 + r, j);
            const double vn = dd_newton_update(inter, A.vstar[o], x, A.zero_boundary);
            A.vnew[o] = vn;
            rmax = nn_max(rmax, res);
            xmax = nn_max(xmax, x);
            vmax = nn_max(vmax, vn);
            bmax = nn_max(bmax, sb[p]);
        } else {
            A.xout[o] = x;
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
    return cudaFuncSetAttribute(k_rbsor_tile, cudaFuncAttributeMaxDynamicSharedMemorySize, 227 * 1024);
}

cudaError_t dd_launch_solve_pass(const DDLaunch& L, const DDGeom& g, const DDRows& R, const double* xin,
                                 double* xout, const double* vstar, double* vnew, int zero_boundary,
                                 DDSolveStats* stats, const DDSolvePlan& P) {
    SolveArgs A;
    A.g = g;
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
    k_rbsor_tile<<<(unsigned)nblocks, P.threads, P.smem_bytes, L.stream>>>(A);
    return cudaGetLastError();
}
