// Development probe: achievable HBM bandwidth of node-per-thread streaming kernels with R read arrays and
// W write arrays of the bench's field size (1025 x 8193 doubles), the access pattern of the pointwise kernels.
#include <cstdio>
#include <cuda_runtime.h>
#define CK(x) do { cudaError_t e = (x); if (e != cudaSuccess) { printf("%s: %s\n", #x, cudaGetErrorString(e)); return 1; } } while (0)
struct Ptrs { const double* r[12]; double* w[8]; };
template <int R, int W, int V>
__global__ void __launch_bounds__(256) k_stream(Ptrs p, long long n) {
    long long i = ((long long)blockIdx.x * blockDim.x + threadIdx.x) * V;
    if (i + V > n) return;
    double acc[V];
#pragma unroll
    for (int v = 0; v < V; ++v) acc[v] = 0.0;
#pragma unroll
    for (int a = 0; a < R; ++a)
#pragma unroll
        for (int v = 0; v < V; ++v) acc[v] += __ldg(p.r[a] + i + v);
#pragma unroll
    for (int a = 0; a < W; ++a)
#pragma unroll
        for (int v = 0; v < V; ++v) p.w[a][i + v] = acc[v] * (a + 1);
}
template <int R, int W, int V>
int run(Ptrs p, long long n) {
    const long long threads = n / V;
    const unsigned blocks = (unsigned)((threads + 255) / 256);
    cudaEvent_t a, b;
    cudaEventCreate(&a); cudaEventCreate(&b);
    for (int k = 0; k < 3; ++k) k_stream<R, W, V><<<blocks, 256>>>(p, n);
    cudaEventRecord(a);
    const int reps = 20;
    for (int k = 0; k < reps; ++k) k_stream<R, W, V><<<blocks, 256>>>(p, n);
    cudaEventRecord(b);
    CK(cudaEventSynchronize(b));
    float ms; cudaEventElapsedTime(&ms, a, b);
    const double gb = (double)(R + W) * n * 8 / 1e9;
    printf("R=%2d W=%d V=%d : %.3f ms/launch  %.0f GB/s\n", R, W, V, ms / reps, gb / (ms / reps * 1e-3));
    return 0;
}
int main() {
    const long long n = 1025LL * 8193;
    Ptrs p;
    for (int a = 0; a < 12; ++a) { double* d; CK(cudaMalloc(&d, n * 8 + 64)); CK(cudaMemset(d, 0, n * 8)); p.r[a] = d; }
    for (int a = 0; a < 8; ++a) { CK(cudaMalloc(&p.w[a], n * 8 + 64)); }
    run<1, 1, 1>(p, n); run<1, 1, 2>(p, n); run<1, 1, 4>(p, n);
    run<4, 2, 1>(p, n); run<4, 2, 2>(p, n);
    run<6, 5, 1>(p, n); run<6, 5, 2>(p, n);
    run<12, 5, 1>(p, n); run<12, 5, 2>(p, n);
    run<12, 2, 1>(p, n); run<12, 7, 1>(p, n);
    run<8, 0, 1>(p, n); run<0, 5, 1>(p, n);
    return 0;
}
