// dd_lane.cu -- lane-private marching red-black SOR kernel (sm_100a) and its launcher; the per-lane program is in
// dd_lane.cuh.  One warp per CTA: a warp shares nothing with any other, so the CTA is only the unit of shared-memory
// allocation (the warp's private coefficient ring) and the SM keeps as many of them resident as the rings allow.
#include <stdlib.h>
#include <string.h>

#include "dd_kernels.cuh"
#include "dd_lane.cuh"

extern __shared__ double dd_lsmem[];

__device__ __forceinline__ void lane_atomic_max_nn(double* addr, double v) {
    atomicMax(reinterpret_cast<unsigned long long*>(addr), (unsigned long long)__double_as_longlong(fabs(v)));
}

// steps U .. P - 1 of one period of the march (every register index is a compile-time constant)
template <int CB, int S, int XIN, int U>
__device__ __forceinline__ void lane_period(const WaveArgs& A, const WaveSeg& sg, LaneRegs<CB, S, XIN>& R,
                                            const LaneSmem& sm, int base, int lane, double omega, double fT) {
    if constexpr (U < 2 * S + 4) {
        double src[2 * S + 1], nb[2 * S + 1];
        dd_lane_offer<CB, S, XIN, U>(R, src);
#pragma unroll
        for (int k = 0; k <= 2 * S; ++k)
            nb[k] = ((U + 1) & 1) ? __shfl_down_sync(0xffffffffu, src[k], 1) : __shfl_up_sync(0xffffffffu, src[k], 1);
        dd_lane_step<CB, S, XIN, U>(A, sg, R, sm, base + U, lane, omega, fT, nb);
        lane_period<CB, S, XIN, U + 1>(A, sg, R, sm, base, lane, omega, fT);
    }
}

template <int CB, int S, int XIN>
__global__ void __launch_bounds__(32) k_sor_lane(const __grid_constant__ WaveArgs A) {
    constexpr int P = 2 * S + 4;
    const int lane = threadIdx.x;
    LaneSmem sm;
    sm.base = dd_lsmem;
    const unsigned ring0 = (unsigned)__cvta_generic_to_shared(dd_lsmem);
    long long f0 = (long long)blockIdx.x * A.flat_per_cta;
    const long long f1 = f0 + A.flat_per_cta < A.flat_total ? f0 + A.flat_per_cta : A.flat_total;
    LaneRegs<CB, S, XIN> R;
    while (f0 < f1) {
        const WaveSeg sg = dd_lane_segment(A, f0, f1, dd_lane_warmup(S, XIN));
        f0 += sg.r1 - sg.r0;
        const DDMember& mb = A.mem[sg.member];
        if (!mb.active) continue;
        const double fT = mb.dt * mb.m.DT;
        const double rho = A.rho_fix >= 0.0 ? A.rho_fix : A.stats[sg.member].rho;
        double omega = 1.0;
        if (rho < 1.0) omega = 2.0 / (1.0 + sqrt(1.0 - rho * rho));
        dd_lane_init<CB, S, XIN>(A, sg, R, sm, ring0, lane, fT);
#pragma unroll
        for (int q = 0; q < DD_LANE_LS; ++q) dd_lane_request<CB, S, XIN>(A, sg, R, sm, q, lane, q);
        const int nsteps = dd_lane_steps(A, sg);
        for (int base = 0; base < nsteps; base += P) lane_period<CB, S, XIN, 0>(A, sg, R, sm, base, lane, omega, fT);
        // nothing of this march may still be in flight when the next one zeroes the ring
        asm volatile("cp.async.wait_all;" ::: "memory");
        if (A.last_pass) {
            // one atomic per quantity and march: resid, |x|, |v_new|, |bb| (high words, see dd_wave_hi)
            const unsigned h0 = __reduce_max_sync(0xffffffffu, R.hr);
            const unsigned h1 = __reduce_max_sync(0xffffffffu, R.hx);
            const unsigned h2 = __reduce_max_sync(0xffffffffu, R.hv);
            const unsigned h3 = __reduce_max_sync(0xffffffffu, R.hb);
            if (lane < 4) {
                DDSolveStats* st = A.stats + sg.member;
                double* dst = lane == 0 ? &st->resid : lane == 1 ? &st->xmax : lane == 2 ? &st->vmax : &st->bmax;
                const unsigned m = lane == 0 ? h0 : lane == 1 ? h1 : lane == 2 ? h2 : h3;
                lane_atomic_max_nn(dst, dd_wave_from_hi(m, lane == 0));
            }
        }
    }
}

// ---- kernel variants: (const band, sweeps of the pass) ------------------------------------------------------------
struct LaneVariant {
    int cb, S, xin;
    const void* fn;
};
#define DD_LANE_MAX_S 5
#define DD_LANE_V(CB, S) {CB, S, 0, (const void*)k_sor_lane<CB, S, 0>}, {CB, S, 1, (const void*)k_sor_lane<CB, S, 1>}
static const LaneVariant kLaneVariants[] = {
    DD_LANE_V(1, 1), DD_LANE_V(1, 2), DD_LANE_V(1, 3), DD_LANE_V(1, 4), DD_LANE_V(1, 5),
    DD_LANE_V(0, 1), DD_LANE_V(0, 2), DD_LANE_V(0, 3), DD_LANE_V(0, 4), DD_LANE_V(0, 5),
};

cudaError_t dd_lane_configure() {
    for (const LaneVariant& v : kLaneVariants) {
        const size_t smem = dd_lane_ring_doubles(v.cb, v.S, v.xin) * sizeof(double);
        cudaError_t e = cudaFuncSetAttribute(v.fn, cudaFuncAttributeMaxDynamicSharedMemorySize, (int)smem);
        if (e != cudaSuccess) return e;
    }
    return cudaSuccess;
}

int dd_lane_max_sweeps() { return DD_LANE_MAX_S; }

// How many sweeps the next pass of a solve with `left` sweeps to go takes: passes of nearly equal length
int dd_lane_pass_sweeps(int left) {
    int cap = DD_LANE_MAX_S;
    const char* e = getenv("DD_LANE_MAX_S");  // development: shorter passes
    if (e && *e && atoi(e) >= 1 && atoi(e) < cap) cap = atoi(e);
    const int passes = (left + cap - 1) / cap;
    return (left + passes - 1) / passes;
}

// Which solves run on the lane kernel: DD_LANE unset = those with enough work to fill the GPU with independent
// warps (a warp's march should be a few pipeline depths long: 2 x SMs x 80 rows of strips), 0 = none, 1 = every
// solve the kernel can take, or a list of variables "T,cl,cd" (forced like 1).
// Measured on B200 at 8193 x 1025 nodes (profiles/README.md): T 0.22 ms against 0.35 ms with the register-tile
// kernel, cl 0.24 against 0.31 (wavefront kernel), cd 0.14 against 0.21; at 257 x 257 the tiles are faster (a
// dozen warps cannot hide the march's latencies).
bool dd_lane_ok(const DDGeom& g, const DDLaunch& L, int var) {
    const char* on = getenv("DD_LANE");  // read per call: the tests switch kernels inside one process
    const bool fits = g.M + 1 >= 4 * 31 && L.own1 - L.own0 >= 16;
    if (!on || !*on) {
        const long long strips = (g.M + 1 + 43) / 44;
        return fits && (long long)L.nmembers * strips * (L.own1 - L.own0) >= 2LL * 148 * 80;
    }
    if (*on == '0') return false;
    if (*on != '1') {
        static const char* names[3] = {"T", "cl", "cd"};
        const char* n = names[var - DD_T];
        const char* hit = strstr(on, n);
        bool found = false;
        while (hit) {
            const char after = hit[strlen(n)];
            if ((hit == on || hit[-1] == ',') && (after == 0 || after == ',')) found = true;
            hit = strstr(hit + 1, n);
        }
        if (!found) return false;
    }
    return fits;
}

cudaError_t dd_launch_solve_lane(const DDLaunch& L, const DDGeom& g, const DDMember* mem, const DDRows& R,
                                 const double* xin, double* xout, const double* vstar, double* vnew,
                                 int zero_boundary, DDSolveStats* stats, int const_band, int sweeps, int last_pass,
                                 double rho_fix) {
    if (sweeps < 1 || sweeps > DD_LANE_MAX_S) return cudaErrorInvalidValue;
    const LaneVariant* v = nullptr;
    for (const LaneVariant& c : kLaneVariants)
        if (c.cb == (const_band ? 1 : 0) && c.S == sweeps && c.xin == (xin ? 1 : 0)) v = &c;
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
    A.halo = dd_lane_halo(sweeps, v->xin);
    A.last_pass = last_pass;
    A.tj = 64 - 2 * A.halo;
    A.nstrips = (g.M + 1 + A.tj - 1) / A.tj;
    A.flat_total = (long long)L.nmembers * A.nstrips * (L.own1 - L.own0);
    A.rho_fix = rho_fix;
    // as many warps as the SMs keep resident, each marching an equal share of the rows of all strips laid end to
    // end; a share is never shorter than a few pipeline depths
    const size_t smem = dd_lane_ring_doubles(v->cb, v->S, v->xin) * sizeof(double);
    int per_sm = 1;
    if (cudaOccupancyMaxActiveBlocksPerMultiprocessor(&per_sm, v->fn, 32, smem) != cudaSuccess || per_sm < 1) per_sm = 1;
    long long ctas = (long long)sm_count * per_sm;
    const long long min_rows = 8LL * A.halo;
    if (ctas * min_rows > A.flat_total) ctas = (A.flat_total + min_rows - 1) / min_rows;
    if (ctas < 1) ctas = 1;
    A.flat_per_cta = (A.flat_total + ctas - 1) / ctas;
    ctas = (A.flat_total + A.flat_per_cta - 1) / A.flat_per_cta;
    // Work order (see dd_lane_segment).  Strip-major (default): every warp marches one long piece of one strip.
    // DD_LANE_ORDER=segment lays the work out as [row segment][strip] so that neighbouring strips are marched at
    // the same time and share their halo columns through L2: measured on B200 at 8193 x 1025 it cuts the kernels'
    // HBM traffic by a third (cl 814 -> 550 MB, cd 709 -> 518, T 268 -> 190 per pass) but costs time (cl 0.234 ->
    // 0.276 ms, cd 0.145 -> 0.154, T 0.225 -> 0.236): a warp's share then spans two segments, i.e. two warm-ups and
    // two pipeline drains, and the kernel is bound by the shared-memory pipe, not by HBM.
    {
        const int R = L.own1 - L.own0;
        const char* order = getenv("DD_LANE_ORDER");
        long long nseg = (R + A.flat_per_cta - 1) / A.flat_per_cta;
        if (nseg < 1) nseg = 1;
        A.nwo = (order && !strcmp(order, "segment")) ? (int)((R + nseg - 1) / nseg) : 0;
    }
    dd_set_last_solver_kernel("k_sor_lane<%d, %d, %d>", v->cb, v->S, v->xin);
    void* args[] = {&A};
    return cudaLaunchKernel(v->fn, dim3((unsigned)ctas), dim3(32u), args, smem, L.stream);
}
