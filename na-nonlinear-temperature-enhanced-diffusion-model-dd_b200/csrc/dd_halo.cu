// dd_halo.cu -- halo exchange of a row slab by direct NVLink stores (sm_100a): one kernel per rank pushes the
// slab's boundary rows straight into the neighbours' halo rows (peer-mapped through CUDA IPC) and handshakes with
// them through flag words in peer memory.  It replaces NCCL's grouped send / recv pairs (44 us per exchange on
// B200 / NVLink 5, three exchanges per PC step) in the multi-process slab driver (ddmesh._DistComm).
//
// Protocol of exchange number `seq` (strictly increasing, the same on every rank) on rank r, flags F_r[0..3] in r's
// memory, written only by r's neighbours:
//   A  tell both neighbours "everything on my stream that read my halo rows is done, you may overwrite them":
//      F_up[1] = seq, F_down[0] = seq                                   (this kernel runs after those readers)
//   B  wait for F_r[0] >= seq (up is ready) and F_r[1] >= seq (down is ready)
//   C  copy my first G owned rows into up's bottom halo, my last G owned rows into down's top halo (16-byte stores)
//   D  the last block to finish: system fence, F_up[3] = seq, F_down[2] = seq ("delivered")
//   E  that block waits for F_r[2] >= seq (up has delivered) and F_r[3] >= seq; kernels launched after this one see
//      the neighbours' rows.
// Every rank runs its kernel on its own GPU, so the spin loops cannot starve each other; a wait that lasts longer
// than about a minute gives up and reports it (status word), which the driver turns into an error.
#include <string.h>

#include "dd_kernels.cuh"

struct HaloPushArgs {
    const double* src_top;   // my first G owned rows (null: no up neighbour)
    double* dst_up;          // up neighbour's bottom halo rows (peer memory)
    const double* src_bot;   // my last G owned rows (null: no down neighbour)
    double* dst_down;        // down neighbour's top halo rows (peer memory)
    long long count;         // doubles per block of rows (G * pitch)
    unsigned* my_flags;      // F_r
    unsigned* up_flags;      // F_up (peer memory)
    unsigned* down_flags;    // F_down (peer memory)
    unsigned seq;
    unsigned* done_blocks;   // local counter, zero on entry and on exit
    int* status;             // local: set to 1 when a wait timed out
};

__device__ __forceinline__ void halo_store_flag(unsigned* p, unsigned v) {
    asm volatile("st.release.sys.global.u32 [%0], %1;" ::"l"(p), "r"(v) : "memory");
}
__device__ __forceinline__ unsigned halo_load_flag(const unsigned* p) {
    unsigned v;
    asm volatile("ld.acquire.sys.global.u32 %0, [%1];" : "=r"(v) : "l"(p) : "memory");
    return v;
}
__device__ __forceinline__ bool halo_wait(const unsigned* p, unsigned seq, int* status) {
    const long long t0 = clock64();
    while ((int)(halo_load_flag(p) - seq) < 0) {
        if (clock64() - t0 > 120000000000LL) {  // about a minute: a neighbour may be busy with a one-off (a compile, a
                                                // check run on one rank); NCCL would wait for ever
            *status = 1;
            return false;
        }
        __nanosleep(64);
    }
    return true;
}

__global__ void __launch_bounds__(256) k_halo_push(const HaloPushArgs A) {
    __shared__ int ok_sh;
    if (threadIdx.x == 0) {
        // A (one block announces; the others only wait)
        if (blockIdx.x == 0) {
            if (A.src_top) halo_store_flag(A.up_flags + 1, A.seq);
            if (A.src_bot) halo_store_flag(A.down_flags + 0, A.seq);
        }
        // B
        bool ok = true;
        if (A.src_top) ok = halo_wait(A.my_flags + 0, A.seq, A.status) && ok;
        if (A.src_bot) ok = halo_wait(A.my_flags + 1, A.seq, A.status) && ok;
        ok_sh = ok ? 1 : 0;
    }
    __syncthreads();
    if (ok_sh) {
        // C: both blocks of rows, 16-byte accesses where the pair is aligned on both sides
        const long long n = A.count;
        const long long stride = (long long)gridDim.x * blockDim.x;
        for (int side = 0; side < 2; ++side) {
            const double* s = side ? A.src_bot : A.src_top;
            double* d = side ? A.dst_down : A.dst_up;
            if (!s) continue;
            const bool vec = ((((unsigned long long)s) | ((unsigned long long)d)) & 15ull) == 0ull;
            if (vec) {
                const double2* s2 = reinterpret_cast<const double2*>(s);
                double2* d2 = reinterpret_cast<double2*>(d);
                for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n / 2; k += stride) d2[k] = s2[k];
                if ((n & 1) && blockIdx.x == 0 && threadIdx.x == 0) d[n - 1] = s[n - 1];
            } else {
                for (long long k = (long long)blockIdx.x * blockDim.x + threadIdx.x; k < n; k += stride) d[k] = s[k];
            }
        }
    }
    // D, E: the last block to get here signals and waits
    __threadfence_system();
    __syncthreads();
    if (threadIdx.x == 0) {
        const unsigned prev = atomicAdd(A.done_blocks, 1u);
        if (prev == gridDim.x - 1) {
            __threadfence_system();
            if (A.src_top) halo_store_flag(A.up_flags + 3, A.seq);
            if (A.src_bot) halo_store_flag(A.down_flags + 2, A.seq);
            if (A.src_top) halo_wait(A.my_flags + 2, A.seq, A.status);
            if (A.src_bot) halo_wait(A.my_flags + 3, A.seq, A.status);
            *A.done_blocks = 0u;
            __threadfence_system();
        }
    }
}

cudaError_t dd_launch_halo_push(cudaStream_t stream, const double* src_top, double* dst_up, const double* src_bot,
                                double* dst_down, long long count, unsigned* my_flags, unsigned* up_flags,
                                unsigned* down_flags, unsigned seq, unsigned* done_blocks, int* status) {
    HaloPushArgs A;
    memset(&A, 0, sizeof(A));
    A.src_top = src_top; A.dst_up = dst_up; A.src_bot = src_bot; A.dst_down = dst_down;
    A.count = count;
    A.my_flags = my_flags; A.up_flags = up_flags; A.down_flags = down_flags;
    A.seq = seq;
    A.done_blocks = done_blocks;
    A.status = status;
    long long blocks = (count / 2 + 255) / 256;
    if (blocks > 64) blocks = 64;  // a few MB over NVLink: 64 CTAs saturate the links without taking the whole GPU
    if (blocks < 1) blocks = 1;
    k_halo_push<<<(unsigned)blocks, 256, 0, stream>>>(A);
    return cudaGetLastError();
}
