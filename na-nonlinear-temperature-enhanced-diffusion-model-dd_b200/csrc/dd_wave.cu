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
    constexpr int W = 64 * C;
    const int nwarps = blockDim.x >> 5, warp = threadIdx.x >> 5, lane = threadIdx.x & 31;
    WaveSmem sm;
    sm.x = dd_wsmem;
    sm.vs = sm.x + (size_t)2 * nwarps * W;
    sm.scol = sm.vs + DD_WAVE_VS * W;
    long long f0 = (long long)blockIdx.x * A.flat_per_cta;
    const long long f1 = f0 + A.flat_per_cta < A.flat_total ? f0 + A.flat_per_cta : A.flat_total;
    WaveRegs<CB, C> R;
    while (f0 < f1) {
        const WaveSeg sg = dd_wave_segment(A, f0, f1);
        f0 += sg.r1 - sg.r0;
        const DDMember& mb = A.mem[sg.member];
        if (!mb.active) continue;
        for (int k = threadIdx.x; k < 2 * nwarps * W; k += blockDim.x) sm.x[k] = 0.0;
        const double fT = mb.dt * mb.m.DT;
        if (CB) {
            for (int sj = threadIdx.x; sj < W; sj += blockDim.x) {
                const int j = sg.cbase + sj;
                const bool in = j >= 1 && j <= A.g.M - 1;
                sm.scol[sj] = in ? fT * A.g.rkp[j] * A.g.rk[j] : 0.0;
                sm.scol[W + sj] = in ? fT * A.g.rkp[j] * A.g.rk[j + 1] : 0.0;
            }
        }
        const double rho = A.rho_fix >= 0.0 ? A.rho_fix : A.stats[sg.member].rho;
        double omega = 1.0;
        if (rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - rho * rho));
        dd_wave_init_thread<CB, C>(R, warp);
        __syncthreads();
        const int nsteps = dd_wave_steps(A, sg);
        for (int t = 0; t < nsteps; ++t) {
            dd_wave_thread_step<CB, C>(A, sg, R, sm, warp, lane, nwarps, omega, fT);
            __syncthreads();
        }
        if (A.last_pass) {
            // one atomic per quantity and march (the staging ring is free now: scratch for the CTA reduction)
            const double r0 = wave_warp_max_nn(R.rmax), r1 = wave_warp_max_nn(R.xmax), r2 = wave_warp_max_nn(R.vmax),
                         r3 = wave_warp_max_nn(R.bmax);
            if (lane == 0) {
                sm.vs[warp * 4 + 0] = r0;
                sm.vs[warp * 4 + 1] = r1;
                sm.vs[warp * 4 + 2] = r2;
                sm.vs[warp * 4 + 3] = r3;
            }
            __syncthreads();
            if (threadIdx.x < 4) {
                double m = 0.0;
                for (int w = 0; w < nwarps; ++w) m = dd_nn_max(m, sm.vs[w * 4 + threadIdx.x]);
                DDSolveStats* st = A.stats + sg.member;
                double* dst = threadIdx.x == 0 ? &st->resid : threadIdx.x == 1 ? &st->xmax
                            : threadIdx.x == 2 ? &st->vmax : &st->bmax;
                wave_atomic_max_nn(dst, m);
            }
        }
        __syncthreads();
    }
}

// ---- kernel variants: (const band, chunks per lane, thread limit) ------------------------------------------------
// Registers bound the CTA: 65536 / threads per thread, and a thread holds 8 C (const band) or 20 C doubles of
// coefficients (thread counts in multiples of 128: registers are allotted per four warps).  The ring needs
// D = 2 * warps >= 4 S + 4 slots, which limits the sweeps of one pass.
struct WaveVariant {
    int cb, C, maxt;
    const void* fn;
};
static const WaveVariant kVariants[] = {
    {1, 4, 512, (const void*)k_sor_wave<1, 4, 512>},  // T, up to 7 sweeps per pass
    {1, 3, 640, (const void*)k_sor_wave<1, 3, 640>},  // T, up to 9
    {1, 2, 768, (const void*)k_sor_wave<1, 2, 768>},  // T, up to 11
    {0, 2, 512, (const void*)k_sor_wave<0, 2, 512>},  // cl / cd, up to 7
    {0, 1, 768, (const void*)k_sor_wave<0, 1, 768>},  // cl / cd, up to 11
};

static int wave_warps_for(int sweeps) { return 2 * sweeps + 2; }  // ring of D = 2 * warps >= 4 S + 4 slots

cudaError_t dd_wave_configure() {
    for (const WaveVariant& v : kVariants) {
        const size_t smem = dd_wave_smem_doubles(v.C, v.maxt / 32) * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

// largest number of sweeps one pass of the widest fitting variant can take
int dd_wave_max_sweeps(int const_band) {
    int best = 0;
    for (const WaveVariant& v : kVariants)
        if (v.cb == const_band) {
            const int s = (2 * (v.maxt / 32) - 4) / 4;
            if (s > best) best = s;
        }
    return best;
}

bool dd_wave_ok(const DDGeom& g, const DDLaunch& L) {
    const char* off = getenv("DD_NO_WAVE");  // read per call: the tests switch kernels inside one process
    return !(off && *off && *off != '0') && g.M + 1 >= 4 * 31 && L.own1 - L.own0 >= 16;
}

cudaError_t dd_launch_solve_wave(const DDLaunch& L, const DDGeom& g, const DDMember* mem, const DDRows& R,
                                 const double* xin, double* xout, const double* vstar, double* vnew,
                                 int zero_boundary, DDSolveStats* stats, int const_band, int sweeps, int last_pass,
                                 double rho_fix) {
    if (sweeps < 1) return cudaErrorInvalidValue;
    const int nw = wave_warps_for(sweeps);
    const WaveVariant* v = nullptr;
    for (const WaveVariant& c : kVariants)
        if (c.cb == (const_band ? 1 : 0) && nw * 32 <= c.maxt) {
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
    A.halo = 2 * sweeps + 1;
    A.last_pass = last_pass;
    A.tj = 64 * v->C - 2 * A.halo - 2;
    if (A.tj < 2) return cudaErrorInvalidValue;
    A.nstrips = (g.M + 1 + A.tj - 1) / A.tj;
    A.flat_total = (long long)L.nmembers * A.nstrips * (L.own1 - L.own0);
    A.rho_fix = rho_fix;
    // one CTA per SM (or several when registers and shared memory allow), each marching an equal share of the
    // rows of all strips laid end to end; a share is never shorter than a few pipeline depths
    const size_t smem = dd_wave_smem_doubles(v->C, nw) * sizeof(double);
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v->fn, nw * 32, smem) != cudaSuccess || per_sm < 1)
        per_sm = 1;
    long long ctas = (long long)sm_count * per_sm;
    const long long min_rows = 8LL * A.halo;
    if (ctas * min_rows > A.flat_total) ctas = (A.flat_total + min_rows - 1) / min_rows;
    if (ctas < 1) ctas = 1;
    A.flat_per_cta = (A.flat_total + ctas - 1) / ctas;
    ctas = (A.flat_total + A.flat_per_cta - 1) / A.flat_per_cta;
    void* args[] = {&A};
    return cudaLaunchKernel(v->fn, dim3((unsigned)ctas), dim3((unsigned)(nw * 32)), args, smem, L.stream);
}
