// dd_wave.cu -- wavefront red-black SOR kernel (sm_100a) and its launcher; the per-thread program is in
// dd_wave.cuh.  Used for wide grids (the marching-kernel regime); narrow grids keep the tile kernels.
#include <cuda_pipeline.h>

#include <stdlib.h>
#include <string.h>

#include "dd_kernels.cuh"
#include "dd_wave.cuh"

extern __shared__ double dd_wsmem[];

__device__ __forceinline__ double wave_warp_max_nn(double v) {
    unsigned long long b = (unsigned long long)__double_as_longlong(fabs(v));
#pragma unroll
    for (int o = 16; o > 0; o >>= 1) {
        const unsigned long long other = __shfl_xor_sync(0xffffffffu, b, o);
        b = other > b ? other : b;
    }
    return __longlong_as_double((long long)b);
}

__device__ __forceinline__ void wave_atomic_max_nn(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(fabs(v)));
}

template <int CB, int C, int MAXT>
__global__ void __launch_bounds__(MAXT, 1) k_sor_wave(const __grid_constant__ WaveArgs A) {
    constexpr int W = 64 * C, PWP = 32 * C + 2;
    const int warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    const int nwo = A.nwo, D = 2 * nwo, S4 = 4 * A.sweeps;
    const int nx = D * 2 * PWP;  // doubles of the x ring
    const bool owner = warp < nwo;
    WaveSmem sm;
    dd_wave_smem_carve(sm, dd_wsmem, (unsigned)__cvta_generic_to_shared(dd_wsmem), CB, C, nwo);
    unsigned* scratch = dd_wave_scratch(sm);
    long long f0 = (long long)blockIdx.x * A.flat_per_cta;
    const long long f1 = f0 + A.flat_per_cta < A.flat_total ? f0 + A.flat_per_cta : A.flat_total;
    WaveRegs<CB, C> R;
    WaveEpi E;
    while (f0 < f1) {
        const WaveSeg sg = dd_wave_segment(A, f0, f1);
        f0 += sg.r1 - sg.r0;
        const DDMember& mb = A.mem[sg.member];
        if (!mb.active) continue;
        for (int k = threadIdx.x; k < nx; k += blockDim.x) dd_wsmem[k] = 0.0;
        const double fT = mb.dt * mb.m.DT;
        if (CB) {
            for (int sj = threadIdx.x; sj < W; sj += blockDim.x) {
                const int j = sg.cbase + sj;
                const bool in = j >= 1 && j <= A.g.M - 1;
                dd_wsmem[nx + sj] = in ? fT * A.g.rkp[j] * A.g.rk[j] : 0.0;
                dd_wsmem[nx + W + sj] = in ? fT * A.g.rkp[j] * A.g.rk[j + 1] : 0.0;
            }
        }
        const double rho = A.rho_fix >= 0.0 ? A.rho_fix : A.stats[sg.member].rho;
        double omega = 1.0;
        if (rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - rho * rho));
        if (owner)
            dd_wave_init_thread<CB, C>(A, sg, R, sm, warp, lane);
        else
            dd_wave_epi_init(A, sg, E, sm, warp - nwo, lane, C);
        __syncthreads();
        const int nsteps = dd_wave_steps(A, sg);
        if (owner) {
            for (int t = 0; t < nsteps; ++t) {
                dd_wave_thread_step<CB, C>(A, sg, R, sm, lane, D, S4, omega, fT);
                __syncthreads();
            }
        } else {
            for (int t = 0; t < nsteps; ++t) {
                dd_wave_epi_step<C>(A, sg, E, sm, t - 2 - DD_WAVE_LS - S4);
                __syncthreads();
            }
        }
        if (A.last_pass) {
            // one atomic per quantity and march: resid, |x|, |v_new|, |bb| (high words, see dd_wave_hi)
            const unsigned h0 = __reduce_max_sync(0xffffffffu, owner ? R.hr : 0u);
            const unsigned h1 = __reduce_max_sync(0xffffffffu, owner ? 0u : E.hx);
            const unsigned h2 = __reduce_max_sync(0xffffffffu, owner ? 0u : E.hv);
            const unsigned h3 = __reduce_max_sync(0xffffffffu, owner ? R.hb : 0u);
            if (lane == 0) {
                scratch[warp * 4 + 0] = h0;
                scratch[warp * 4 + 1] = h1;
                scratch[warp * 4 + 2] = h2;
                scratch[warp * 4 + 3] = h3;
            }
            __syncthreads();
            if (threadIdx.x < 4) {
                unsigned m = 0u;
                for (int w = 0; w < nwo + DD_WAVE_NE; ++w) m = max(m, scratch[w * 4 + threadIdx.x]);
                DDSolveStats* st = A.stats + sg.member;
                double* dst = threadIdx.x == 0 ? &st->resid : threadIdx.x == 1 ? &st->xmax
                            : threadIdx.x == 2 ? &st->vmax : &st->bmax;
                wave_atomic_max_nn(dst, dd_wave_from_hi(m, threadIdx.x == 0));
            }
        }
        __syncthreads();
    }
}

// ---- kernel variants: (const band, chunks per lane, thread limit) ------------------------------------------------
// Registers bound the CTA: 65536 / threads per thread (allotted per four warps), and an owner thread holds 8 C
// (const band) or 20 C doubles of coefficients.  The ring needs D = 2 * (owner warps) >= 4 S + 4 slots, which
// limits the sweeps of one pass; DD_WAVE_NE epilogue warps come on top.
struct WaveVariant {
    int cb, C, maxt;
    const void* fn;
};
static const WaveVariant kVariants[] = {
    {1, 3, 640, (const void*)k_sor_wave<1, 3, 640>},  // T, up to 7 sweeps per pass
    {1, 2, 640, (const void*)k_sor_wave<1, 2, 640>},
    {1, 2, 896, (const void*)k_sor_wave<1, 2, 896>},  // T, up to 11
    {0, 2, 384, (const void*)k_sor_wave<0, 2, 384>},  // cl / cd, up to 3
    {0, 2, 512, (const void*)k_sor_wave<0, 2, 512>},  // cl / cd, up to 5
    {0, 1, 768, (const void*)k_sor_wave<0, 1, 768>},  // cl / cd, up to 9
};

static int wave_owner_warps_for(int sweeps) { return 2 * sweeps + 2; }  // ring of D = 2 * warps >= 4 S + 4 slots

cudaError_t dd_wave_configure() {
    for (const WaveVariant& v : kVariants) {
        const size_t smem = dd_wave_smem_doubles(v.cb, v.C, v.maxt / 32 - DD_WAVE_NE) * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// largest number of sweeps one pass can take: (any variant, the widest variant)
void dd_wave_max_sweeps(int const_band, int* any, int* wide) {
    *any = *wide = 0;
    int wideC = 0;
    for (const WaveVariant& v : kVariants)
        if (v.cb == const_band) {
            const int s = (2 * (v.maxt / 32 - DD_WAVE_NE) - 4) / 4;
            if (s > *any) *any = s;
            if (v.C > wideC || (v.C == wideC && s > *wide)) {
                wideC = v.C;
                *wide = s;
            }
        }
}

// Which solves run on the wavefront kernel (when the lane kernel does not take them): DD_WAVE = 0 (none, the
// default), 1 (all), or a list of variables "T,cl,cd".  Measured on B200 at 8193 x 1025 nodes (profiles/README.md): cl 0.31 ms against 0.35 ms with the register-tile
// kernel (five sweeps in one pass, one CTA per SM, no row halo), T 0.47 against 0.35 (two tile passes), cd 0.23
// against 0.21 -- the wavefront's steps are lock-step sequences of a shared-memory phase, an fp64 phase and a
// barrier that cannot overlap, so it only wins where the tile kernel's halo redundancy is largest.
bool dd_wave_ok(const DDGeom& g, const DDLaunch& L, int var) {
    const char* on = getenv("DD_WAVE");  // read per call: the tests switch kernels inside one process
    if (!on || !*on) on = "0";  // superseded by the lane kernel (dd_lane.cu); kept for comparison
    if (*on == '0') return false;
    if (*on != '1') {
        static const char* names[3] = {"T", "cl", "cd"};
        const char* n = names[var - DD_T];
        const char* hit = strstr(on, n);
        // "cl" and "cd" both start with c: compare whole tokens
        bool found = false;
        while (hit) {
            const char after = hit[strlen(n)];
            if ((hit == on || hit[-1] == ',') && (after == 0 || after == ',')) found = true;
            hit = strstr(hit + 1, n);
        }
        if (!found) return false;
    }
    return g.M + 1 >= 4 * 31 && L.own1 - L.own0 >= 16;
}

cudaError_t dd_launch_solve_wave(const DDLaunch& L, const DDGeom& g, const DDMember* mem, const DDRows& R,
                                 const double* xin, double* xout, const double* vstar, double* vnew,
                                 int zero_boundary, DDSolveStats* stats, int const_band, int sweeps, int last_pass,
                                 double rho_fix) {
    if (sweeps < 1) return cudaErrorInvalidValue;
    const int nwo = wave_owner_warps_for(sweeps), nw = nwo + DD_WAVE_NE;
    const WaveVariant* v = nullptr;
    const char* wantC = getenv(const_band ? "DD_WAVE_C_T" : "DD_WAVE_C");  // development: force the strip width
    for (const WaveVariant& c : kVariants)
        if (c.cb == (const_band ? 1 : 0) && nw * 32 <= c.maxt && !(wantC && *wantC && atoi(wantC) != c.C)) {
            v = &c;
            break;
        }
    if (!v) return cudaErrorInvalidValue;
    static int sm_count = 0;
    if (sm_count == 0) {
        int dev = 0;
        cudaGetDevice(&dev);
        if (cudaDeviceGetAttribute(&sm_count, cudaDevAttrMultiProcessorCount, dev) != cudaSuccess || sm_count <= 0)
            sm_count = 148;
    }
    WaveArgs A;
    memset(&A, 0, sizeof(A));
    A.g = g;
    A.mem = mem;
    A.bb = R.bb; A.aW = R.aW; A.aE = R.aE; A.aS = R.aS; A.aN = R.aN;
    A.xin = xin;
    A.xout = xout;
    A.vstar = vstar;
    A.vnew = vnew;
    A.stats = stats;
    A.zero_boundary = zero_boundary;
    A.ldR = R.ld;
    A.mstrideR = R.mstride;
    A.own0 = L.own0; A.own1 = L.own1; A.vr0 = L.vr0; A.vr1 = L.vr1;
    A.sweeps = sweeps;
    A.halo = 2 * sweeps + 2;
    A.last_pass = last_pass;
    A.tj = 64 * v->C - 2 * A.halo;
    A.nwo = nwo;
    if (A.tj < 2 || A.tj > 32 * DD_WAVE_NE * DD_WAVE_EV) return cudaErrorInvalidValue;
    A.nstrips = (g.M + 1 + A.tj - 1) / A.tj;
    A.flat_total = (long long)L.nmembers * A.nstrips * (L.own1 - L.own0);
    A.rho_fix = rho_fix;
    // one CTA per SM (or several when registers and shared memory allow), each marching an equal share of the
    // rows of all strips laid end to end; a share is never shorter than a few pipeline depths
    const size_t smem = dd_wave_smem_doubles(v->cb, v->C, nwo) * sizeof(double);
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v->fn, nw * 32, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    long long ctas = (long long)sm_count * per_sm;
    const long long min_rows = 8LL * A.halo;
    if (ctas * min_rows > A.flat_total) ctas = (A.flat_total + min_rows - 1) / min_rows;
    if (ctas < 1) ctas = 1;
    A.flat_per_cta = (A.flat_total + ctas - 1) / ctas;
    ctas = (A.flat_total + A.flat_per_cta - 1) / A.flat_per_cta;
    dd_set_last_solver_kernel("k_sor_wave<%d, %d, %d>", v->cb, v->C, v->maxt);
    void* args[] = {&A};
    return cudaLaunchKernel(v->fn, dim3((unsigned)ctas), dim3((unsigned)(nw * 32)), args, smem, L.stream);
}
