// dd_member.cu -- whole trajectories of SMALL grids on chip: one CTA per ensemble member, the time loop inside the
// kernel (sm_100a).
//
// Replaces, for grids whose working set fits in shared memory (up to about 36 x 36 nodes), the loop of the
// reference's trial driver (src/mms_trial_utils.py:126-140: integrator.step + collect_errors per step) and the
// step itself (src/prob1base.py:3117-3149).  A member's five fields are read from HBM once, every intermediate of
// the predictor-corrector step (predictors, Y_T / Y_cl / Y_cd, the three Newton systems, their iterates, T1 / cl1 /
// cd1) lives in shared memory, the per-step error norms are folded into the reference's max-integral combination
// on the spot, and the five fields are written back once at the end: 80 B of HBM traffic per node for the WHOLE
// trajectory instead of about 650 B per node and step through the mesh kernels.
//
// The arithmetic is not restated: every phase calls the node programs of dd_nodeprog.cuh / dd_physics.cuh (the same
// functions the mesh kernels and the test-only host build use) on a DDGeom / DDStateC that point into shared
// memory, the red-black SOR uses dd_sor.cuh, the norm combination dd_combine.cuh.  What is specific to this kernel:
//  * the linear solves iterate until the member's OWN residual bound is met (the member is alone in its CTA, so the
//    decision is local and needs no host round trip): planned sweeps from the Gershgorin ratio, then one sweep at
//    a time; a member that does not converge within max_sweeps raises a flag (DD_ERR_NOT_CONVERGED on the host);
//  * the cs-Newton exit test of the reference (max |dx| < rtol |x| at every node, src/prob1base.py:3661) is a
//    reduction over the CTA per iteration: the iteration count is the reference's, exactly.
#include <cuda_runtime.h>

#include <stdlib.h>
#include <string.h>

#include "dd_combine.cuh"
#include "dd_kernels.cuh"
#include "dd_member.cuh"
#include "dd_sor.cuh"

#define DD_MEMBER_THREADS 512
#define DD_MEMBER_WARPS (DD_MEMBER_THREADS / 32)
#define DD_MEMBER_ARRAYS 19  // 5 state + cp1p, cs1p, YT, Ycl, Ycd + bb, aW, aE, aS, aN + x + T1, cl1, cd1

extern __shared__ double dd_msmem[];

// ---- deterministic CTA reductions (fixed order: lanes by shuffle tree, then warps in order) ----------------------
__device__ __forceinline__ double cta_sum(double v, double* scratch) {
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) v += __shfl_xor_sync(0xffffffffu, v, o);
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = v;
    __syncthreads();
    double t = 0.0;
    for (int w = 0; w < DD_MEMBER_WARPS; ++w) t += scratch[w];
    return t;  // the same value in every thread
}

// max / min of non-negative doubles by bit pattern (NaN is the largest pattern: sticky in the max)
__device__ __forceinline__ double cta_max_nn(double v, double* scratch) {
    unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
        b = other > b ? other : b;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = __longlong_as_double((long long)b);
    __syncthreads();
    unsigned long long m = 0ull;
    for (int w = 0; w < DD_MEMBER_WARPS; ++w) {
        const unsigned long long o = (unsigned long long)__double_as_longlong(scratch[w]);
        m = o > m ? o : m;
    }
    return __longlong_as_double((long long)m);
}
__device__ __forceinline__ double cta_min_nn(double v, double* scratch) {
    unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
        b = other < b ? other : b;
    }
    __syncthreads();
    if ((threadIdx.x & 31) == 0) scratch[threadIdx.x >> 5] = __longlong_as_double((long long)b);
    __syncthreads();
    unsigned long long m = ~0ull;
    for (int w = 0; w < DD_MEMBER_WARPS; ++w) {
        const unsigned long long o = (unsigned long long)__double_as_longlong(scratch[w]);
        m = o < m ? o : m;
    }
    return __longlong_as_double((long long)m);
}

// SOR sweeps for a Gershgorin ratio rho (dd_capi.cu: sweeps_for_rho, same rule)
__device__ int member_sweeps_for_rho(double rho, int max_sweeps) {
    if (!(rho >= 0.0)) return max_sweeps;
    if (rho < 1e-300) return 2;
    double lam = 0.999;
    if (rho < 1.0) lam = 2.0 / (1.0 + sqrt(1.0 - rho * rho)) - 1.0;
    int k = 2;
    while (k < max_sweeps && (1.0 + k) * pow(lam, (double)k) > 1e-17) ++k;
    return k;
}

struct MemberSolveOut {
    int sweeps, ok;
    double rho, resid, bound;
};

// Solves the assembled system in shared memory and writes v_new = v* + x (T: boundary := 0; cl, cd keep v*).
// rowf / colf: dt DT / (hhat h) factors of the constant-band T system ([4][max(N, M) + 1]: W, E, S, N).
template <int CB>
__device__ MemberSolveOut member_solve(const DDGeom& g, const DDRows& R, double* x, const double* vstar, double* vnew,
                                       int zero_boundary, double rho_node, const double* geof, int gstride,
                                       const DDMemberArgs& A, double* scratch) {
    const int N = g.N, M = g.M, P = g.ld, n = (N + 1) * P;
    const double rho = cta_max_nn(rho_node, scratch);
    double omega = 1.0;
    if (rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - rho * rho));
    for (int k = threadIdx.x; k < n; k += DD_MEMBER_THREADS) x[k] = 0.0;
    __syncthreads();
    const int JH = (M + 2) / 2;  // column pairs per row
    const int ncell = (N - 1) * JH;
    const double* gW = geof;
    const double* gE = geof + gstride;
    const double* gS = geof + 2 * gstride;
    const double* gN = geof + 3 * gstride;
    auto gs_at = [&](int i, int j, int p) {
        // gs - x: the cell's residual (dd_sor.cuh)
        if (CB)
            return dd_sor_dT(R.bb[p], R.aW[p], gW[i], gE[i], gS[j], gN[j], x[p - P], x[p + P], x[p - 1], x[p + 1], x[p]);
        return dd_sor_d5(R.bb[p], R.aW[p], R.aE[p], R.aS[p], R.aN[p], x[p - P], x[p + P], x[p - 1], x[p + 1], x[p]);
    };
    auto sweep = [&]() {
        for (int c = 0; c < 2; ++c) {
            for (int k = threadIdx.x; k < ncell; k += DD_MEMBER_THREADS) {
                const int i = 1 + k / JH, j = 2 * (k - (i - 1) * JH) + ((i + c) & 1);  // colour = (i + j) & 1
                if (j >= 1 && j <= M - 1) {
                    const int p = i * P + j;
                    x[p] = dd_sor_relax(x[p], gs_at(i, j, p), omega);
                }
            }
            __syncthreads();
        }
    };
    MemberSolveOut out;
    out.rho = rho;
    int done = 0;
    int plan = A.fixed_sweeps > 0 ? A.fixed_sweeps : member_sweeps_for_rho(rho, A.max_sweeps);
    // first plan: the theoretical count shortened by what experience with these matrices allows (the bound is
    // verified below, a sweep at a time is added while it fails)
    if (A.fixed_sweeps <= 0 && plan > 3) plan -= 1;
    out.ok = 0;
    out.resid = out.bound = 0.0;
    for (;;) {
        for (; done < plan; ++done) sweep();
        double rmax = 0.0, xmax = 0.0, vmax = 0.0, bmax = 0.0;
        for (int k = threadIdx.x; k < (N - 1) * (M - 1); k += DD_MEMBER_THREADS) {
            const int i = 1 + k / (M - 1), j = 1 + (k - (i - 1) * (M - 1));
            const int p = i * P + j;
            rmax = dd_nn_max(rmax, gs_at(i, j, p));
            xmax = dd_nn_max(xmax, x[p]);
            vmax = dd_nn_max(vmax, vstar[p] + x[p]);
            bmax = dd_nn_max(bmax, R.bb[p]);
        }
        rmax = cta_max_nn(rmax, scratch);
        xmax = cta_max_nn(xmax, scratch);
        bmax = cta_max_nn(bmax, scratch);
        // |v_new| over all nodes the update writes: boundary nodes keep v* (cl, cd) or become 0 (T)
        if (!zero_boundary)
            for (int k = threadIdx.x; k < n; k += DD_MEMBER_THREADS) {
                const int i = k / P, j = k - i * P;
                if (j <= M && !(i > 0 && i < N && j > 0 && j < M)) vmax = dd_nn_max(vmax, vstar[k]);
            }
        vmax = cta_max_nn(vmax, scratch);
        // acceptance rule of k_summarise (dd_capi.cu): |x - x*| <= resid / (1 - rho) <= tol |v_new| + rounding floor
        const double gap = rho < 1.0 ? 1.0 - rho : 1e-4;
        const double allowed = A.solve_tol * gap * vmax + 16.0 * 2.220446049250313e-16 * (bmax + xmax);
        const bool nan_entry = !(rho < 1e300) || !(bmax < 1e300);  // blown-up state: accepted, NaN propagates
        out.resid = rmax;
        out.bound = vmax > 0.0 ? rmax / (gap * vmax) : 0.0;
        if (nan_entry || rmax <= allowed) {
            out.ok = 1;
            break;
        }
        if (A.fixed_sweeps > 0 || done >= A.max_sweeps) break;
        plan = done + 1;
    }
    out.sweeps = done;
    for (int k = threadIdx.x; k < n; k += DD_MEMBER_THREADS) {
        const int i = k / P, j = k - i * P;
        if (j <= M) vnew[k] = dd_newton_update(i > 0 && i < N && j > 0 && j < M, vstar[k], x[k], zero_boundary);
    }
    __syncthreads();
    return out;
}

// error norms of (state - exact at time slot `slot`) -> r[8] = H2[cp, T, cl, cd, cs], P2[T, cl, cd], in every thread
// (formulas of k_error_partial, dd_kernels.cu; e3: three scratch arrays for the errors of T, cl, cd)
template <int MODE>
__device__ void member_norms(const DDGeom& g, const DDMember& mb, const DDForcing& F, const DDStateC& s, int slot,
                             double* e3, double* scratch, double* r) {
    const int N = g.N, M = g.M, P = g.ld, n = (N + 1) * P;
    double acc[8] = {0, 0, 0, 0, 0, 0, 0, 0};
    for (int k = threadIdx.x; k < n; k += DD_MEMBER_THREADS) {
        const int i = k / P, j = k - i * P;
        if (j > M) continue;
        double u[DD_NVAR];
        dd_exact_values<MODE>(F, mb, slot, i, j, u);
        double e[DD_NVAR];
        for (int v = 0; v < DD_NVAR; ++v) e[v] = s.v[v][k] - u[v];
        for (int q = 0; q < 3; ++q) e3[q * n + k] = e[DD_T + q];
        if (i > 0 && i < N && j > 0 && j < M) {
            const double wgt = g.hp[i] * g.kp[j];
            for (int v = 0; v < DD_NVAR; ++v) acc[v] += e[v] * e[v] * wgt;
        }
    }
    __syncthreads();
    for (int k = threadIdx.x; k < n; k += DD_MEMBER_THREADS) {
        const int i = k / P, j = k - i * P;
        if (j > M) continue;
        const bool need_w = i >= 1 && j >= 1 && j <= M - 1;
        const bool need_s = j >= 1 && i >= 1 && i <= N - 1;
        for (int q = 0; q < 3; ++q) {
            const double ev = e3[q * n + k];
            double p = 0.0;
            if (need_w) {
                const double d = (ev - e3[q * n + k - P]) / g.h[i];
                p += d * d * g.h[i] * g.kp[j];
            }
            if (need_s) {
                const double d = (ev - e3[q * n + k - 1]) / g.k[j];
                p += d * d * g.hp[i] * g.k[j];
            }
            acc[5 + q] += p;
        }
    }
    for (int q = 0; q < 8; ++q) r[q] = cta_sum(acc[q], scratch);
    __syncthreads();
}

template <int MODE>
__global__ void __launch_bounds__(DD_MEMBER_THREADS, 1) k_member_run(const __grid_constant__ DDMemberArgs A) {
    const int N = A.g.N, M = A.g.M, P = A.pitch, n = (N + 1) * P;
    double* base = dd_msmem;
    double* arr[DD_MEMBER_ARRAYS];
    for (int k = 0; k < DD_MEMBER_ARRAYS; ++k) arr[k] = base + (size_t)k * n;
    const int gstride = (N > M ? N : M) + 2;
    double* geof = base + (size_t)DD_MEMBER_ARRAYS * n;   // [4][gstride]
    double* scratch = geof + 4 * gstride;                   // [32]
    double* cst = scratch + 32;                             // [18] combine state, [8] norms
    DDMember* smb = reinterpret_cast<DDMember*>(cst + 32);  // the member's parameters and time coefficients
    DDGeom g = A.g;
    g.ld = P;
    g.nrows = N + 1;
    g.row0 = 0;
    g.mstride = 0;
    const int tid = threadIdx.x;
    for (int member = blockIdx.x; member < A.nmembers; member += gridDim.x) {
        const DDMember& gmb = A.mem[member];
        if (!gmb.active) continue;
        __syncthreads();
        if (tid == 0) *smb = gmb;
        // state: slots 0..4 of arr; the other roles are assigned below and rotate from step to step
        const long long mo = (long long)member * A.g.mstride;
        for (int k = tid; k < n; k += DD_MEMBER_THREADS) {
            const int i = k / P, j = k - i * P;
            for (int v = 0; v < DD_NVAR; ++v) arr[v][k] = j <= M ? A.in.v[v][mo + (long long)i * A.g.ld + j] : 0.0;
        }
        __syncthreads();
        const DDMember& mb = *smb;
        if (tid == 0) {
            dd_time_coefs(MODE, *smb, smb->t0, 0, &smb->tc[0]);
            dd_time_coefs(MODE, *smb, smb->t0 + smb->dt, 1, &smb->tc[1]);
        }
        // constant-band factors of the T system: dt DT / (hhat_i h_i), / (hhat_i h_{i+1}), and the same in y
        {
            const double fT = gmb.dt * gmb.m.DT;
            for (int i = tid; i <= N; i += DD_MEMBER_THREADS) {
                const bool in = i >= 1 && i <= N - 1;
                geof[i] = in ? fT * g.rhp[i] * g.rh[i] : 0.0;
                geof[gstride + i] = in ? fT * g.rhp[i] * g.rh[i + 1] : 0.0;
            }
            for (int j = tid; j <= M; j += DD_MEMBER_THREADS) {
                const bool in = j >= 1 && j <= M - 1;
                geof[2 * gstride + j] = in ? fT * g.rkp[j] * g.rk[j] : 0.0;
                geof[3 * gstride + j] = in ? fT * g.rkp[j] * g.rk[j + 1] : 0.0;
            }
        }
        __syncthreads();
        // roles: s0 = state at t0; nw = {cp1, T1, cl1, cd1, cs1} (cp1p / cs1p while predicted); work arrays
        double* s0[DD_NVAR] = {arr[0], arr[1], arr[2], arr[3], arr[4]};
        double* nw[DD_NVAR] = {arr[5], arr[6], arr[7], arr[8], arr[9]};
        double *YT = arr[10], *Ycl = arr[11], *Ycd = arr[12];
        DDRows R;
        R.bb = arr[13]; R.aW = arr[14]; R.aE = arr[15]; R.aS = arr[16]; R.aN = arr[17];
        R.ld = P;
        R.mstride = 0;
        double* x = arr[18];
        double r8[8];
        MemberSolveOut so[3] = {};
        int used = A.cap, failed = 0;
        {
            DDStateC s;
            for (int v = 0; v < DD_NVAR; ++v) s.v[v] = s0[v];
            member_norms<MODE>(g, mb, A.F, s, 0, R.bb, scratch, r8);
            if (tid == 0) dd_combine_fold(r8, mb.dt, 1, cst);
        }
        for (int step = 0; step < A.nsteps; ++step) {
            DDStateC st0;
            for (int v = 0; v < DD_NVAR; ++v) st0.v[v] = s0[v];
            // ---- predictors, Y_T, Y_cl, Y_cd ----------------------------------------------------------------------
            DDPredictOut po;
            po.cp1p = nw[DD_CP]; po.cs1p = nw[DD_CS]; po.YT = YT; po.Ycl = Ycl; po.Ycd = Ycd;
            for (int k = tid; k < n; k += DD_MEMBER_THREADS) {
                const int i = k / P, j = k - i * P;
                if (j <= M) dd_node_predict<MODE>(g, mb, A.F, st0, po, 0, i, j);
            }
            __syncthreads();
            DDStateC u;  // linearisation state (cp1p, T*, cl*, cd*, cs1p)
            u.v[DD_CP] = nw[DD_CP]; u.v[DD_T] = s0[DD_T]; u.v[DD_CL] = s0[DD_CL]; u.v[DD_CD] = s0[DD_CD];
            u.v[DD_CS] = nw[DD_CS];
            // ---- T -------------------------------------------------------------------------------------------------
            double rho = 0.0;
            for (int k = tid; k < n; k += DD_MEMBER_THREADS) {
                const int i = k / P, j = k - i * P;
                if (j <= M) rho = fmax(rho, dd_node_asm_T_const<MODE>(g, mb, A.F, u, YT, R, 0, 0, i, j));
            }
            __syncthreads();
            so[0] = member_solve<1>(g, R, x, s0[DD_T], nw[DD_T], 1, rho, geof, gstride, A, scratch);
            // ---- cl ------------------------------------------------------------------------------------------------
            rho = 0.0;
            for (int k = tid; k < n; k += DD_MEMBER_THREADS) {
                const int i = k / P, j = k - i * P;
                if (j <= M) rho = fmax(rho, dd_node_asm_cl<MODE>(g, mb, A.F, u, nw[DD_T], Ycl, R, 0, 0, i, j));
            }
            __syncthreads();
            so[1] = member_solve<0>(g, R, x, s0[DD_CL], nw[DD_CL], 0, rho, geof, gstride, A, scratch);
            // ---- cd ------------------------------------------------------------------------------------------------
            rho = 0.0;
            for (int k = tid; k < n; k += DD_MEMBER_THREADS) {
                const int i = k / P, j = k - i * P;
                if (j <= M)
                    rho = fmax(rho, dd_node_asm_cd<MODE>(g, mb, A.F, u, nw[DD_T], nw[DD_CL], Ycd, A.cd_swap, R, 0, 0, i, j));
            }
            __syncthreads();
            so[2] = member_solve<0>(g, R, x, s0[DD_CD], nw[DD_CD], 0, rho, geof, gstride, A, scratch);
            if (!(so[0].ok && so[1].ok && so[2].ok)) failed = 1;
            // ---- correctors (cp: closed form; cs: Newton iterations with the reference's global exit test) -----------
            {
                constexpr int NPT = 3;  // nodes per thread (n <= NPT * threads is checked by the launcher)
                double xs[NPT], ys[NPT], as[NPT];
                int kk[NPT];
                for (int q = 0; q < NPT; ++q) {
                    const int k = tid + q * DD_MEMBER_THREADS;
                    kk[q] = -1;
                    xs[q] = ys[q] = as[q] = 0.0;
                    if (k < n) {
                        const int i = k / P, j = k - i * P;
                        if (j <= M) {
                            double cp1;
                            dd_node_correct_prepare<MODE>(g, mb, A.F, st0, nw[DD_T], nw[DD_CL], nw[DD_CD], 0, i, j, &cp1,
                                                          &ys[q], &as[q]);
                            xs[q] = s0[DD_CS][k];
                            kk[q] = k;
                            // (cp1p of this node has been read by nobody but this thread since the solves)
                            nw[DD_CP][k] = cp1;
                        }
                    }
                }
                const double eta = mb.m.eta;
                used = A.cap;
                for (int it = 0; it < A.cap; ++it) {
                    double mx = 0.0, mn = __longlong_as_double(0x7ff0000000000000LL);
                    for (int q = 0; q < NPT; ++q)
                        if (kk[q] >= 0) {
                            const double dx = dd_cs_newton_dx(xs[q], ys[q], as[q], eta);
                            xs[q] = xs[q] + dx;
                            mx = dd_nn_max(mx, dx);
                            double ax = fabs(xs[q]);
                            if (ax != ax) ax = 0.0;  // a NaN |x| fails the test exactly like 0 does
                            mn = ax < mn ? ax : mn;
                        }
                    if (A.rtol > 0.0) {
                        mx = cta_max_nn(mx, scratch);
                        mn = cta_min_nn(mn, scratch);
                        if (mx < A.rtol * mn) {
                            used = it + 1;
                            break;
                        }
                    }
                }
                __syncthreads();
                for (int q = 0; q < NPT; ++q)
                    if (kk[q] >= 0) {
                        const int i = kk[q] / P, j = kk[q] - i * P;
                        nw[DD_CS][kk[q]] = xs[q] * ((i > 0 && i < N && j > 0 && j < M) ? 1.0 : 0.0);
                    }
            }
            __syncthreads();
            // ---- the step is done: advance the time, fold the error norms of the new state ---------------------------
            if (tid == 0) {
                smb->t0 = smb->t0 + smb->dt;
                dd_time_coefs(MODE, *smb, smb->t0, 0, &smb->tc[0]);
                dd_time_coefs(MODE, *smb, smb->t0 + smb->dt, 1, &smb->tc[1]);
            }
            __syncthreads();
            for (int v = 0; v < DD_NVAR; ++v) {
                double* t = s0[v];
                s0[v] = nw[v];
                nw[v] = t;
            }
            if (A.combined) {
                DDStateC s;
                for (int v = 0; v < DD_NVAR; ++v) s.v[v] = s0[v];
                member_norms<MODE>(g, mb, A.F, s, 0, R.bb, scratch, r8);
                if (tid == 0) dd_combine_fold(r8, mb.dt, 0, cst);
            }
        }
        __syncthreads();
        // ---- results ---------------------------------------------------------------------------------------------------
        for (int k = tid; k < n; k += DD_MEMBER_THREADS) {
            const int i = k / P, j = k - i * P;
            if (j <= M)
                for (int v = 0; v < DD_NVAR; ++v) A.out.v[v][mo + (long long)i * A.g.ld + j] = s0[v][k];
        }
        if (tid == 0) {
            A.mem_rw[member].t0 = smb->t0;
            A.mem_rw[member].tc[0] = smb->tc[0];
            A.mem_rw[member].tc[1] = smb->tc[1];
            if (A.combined)
                for (int q = 0; q < 6; ++q) A.combined[(size_t)member * 6 + q] = __dsqrt_rn(cst[q]);
            double* so_out = A.stats + (size_t)member * 16;
            for (int q = 0; q < 3; ++q) {
                so_out[q] = (double)so[q].sweeps;
                so_out[3 + q] = so[q].rho;
                so_out[6 + q] = so[q].resid;
                so_out[9 + q] = so[q].bound;
            }
            so_out[12] = (double)used;
            so_out[13] = (double)failed;
        }
    }
}

size_t dd_member_smem_bytes(int N, int M, int* pitch) {
    int P = M + 1;
    if (!(P & 1)) ++P;  // odd pitch: rows a bank apart
    if (pitch) *pitch = P;
    const size_t n = (size_t)(N + 1) * P;
    const int gstride = (N > M ? N : M) + 2;
    return (DD_MEMBER_ARRAYS * n + 4 * (size_t)gstride + 32 + 32) * sizeof(double) + sizeof(DDMember) + 64;
}

bool dd_member_fits(int N, int M) {
    int P;
    const size_t bytes = dd_member_smem_bytes(N, M, &P);
    return N >= 2 && M >= 2 && bytes <= 227 * 1024 && (size_t)(N + 1) * P <= 3 * DD_MEMBER_THREADS;
}

cudaError_t dd_launch_member_run(cudaStream_t stream, int mode, DDMemberArgs A, int sm_count) {
    int P;
    const size_t smem = dd_member_smem_bytes(A.g.N, A.g.M, &P);
    A.pitch = P;
    const void* fn = mode == DD_FORCING_SEPARABLE ? (const void*)k_member_run<DD_FORCING_SEPARABLE>
                                                   : (const void*)k_member_run<DD_FORCING_EXPSIN>;
    if (mode != DD_FORCING_SEPARABLE && mode != DD_FORCING_EXPSIN) return cudaErrorInvalidValue;
    cudaError_t e = cudaFuncSetAttribute(fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
    if (e != cudaSuccess) return e;
    int grid = sm_count > 0 ? sm_count : 148;
    if (grid > A.nmembers) grid = A.nmembers;
    void* args[] = {&A};
    return cudaLaunchKernel(fn, dim3((unsigned)grid), dim3(DD_MEMBER_THREADS), args, smem, stream);
}
