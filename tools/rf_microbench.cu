// Does a DFMA with three distinct 64-bit register operands issue every 2 cycles on sm_100a?
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 12;
// MODE 0: acc[k] = fma(z[k], w[k], acc[k])   (3 distinct, nothing shared between neighbours)
// MODE 1: acc[k] = fma(z0,   w[k], acc[k])   (one operand shared)
// MODE 2: acc[k] = fma(z0,   w0,   acc[k])   (two shared)
// MODE 3: acc[k] = z[k] * w[k] (DMUL 2 distinct) ; MODE 4: DADD acc[k] + z[k]
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(double *sink, long long *cycles, int iters, const double *g) {
    double acc[N], z[N], w[N];
#pragma unroll
    for (int i = 0; i < N; ++i) { acc[i] = g[i + threadIdx.x]; z[i] = g[64 + i + threadIdx.x]; w[i] = g[128 + i + threadIdx.x]; }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (MODE == 0) acc[i] = fma(z[i], w[i], acc[i]);
                if (MODE == 1) acc[i] = fma(z[0], w[i], acc[i]);
                if (MODE == 2) acc[i] = fma(z[0], w[0], acc[i]);
                if (MODE == 3) acc[i] = z[i] * acc[i];
                if (MODE == 4) acc[i] = z[i] + acc[i];
                if (MODE == 5) acc[i] = fma(z[i], w[(i + 1) % N], acc[i]);
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += acc[i] + z[i] + w[i];
    if (s == 1234.5) sink[0] = s;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}
int main() {
    double *sink, *g; long long *cyc, h;
    cudaMalloc(&sink, 64); cudaMalloc(&cyc, 64); cudaMalloc(&g, 4096 * 8);
    cudaMemset(g, 0, 4096 * 8);
    const int iters = 4000;
#define RUN(M, W) do { k<M><<<1, 32 * W>>>(sink, cyc, iters, g); cudaDeviceSynchronize(); k<M><<<1, 32 * W>>>(sink, cyc, iters, g); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); double per = (double)h / (iters * 4.0 * N * (W / 4.0)); \
    printf("mode %d warps %2d: %.3f cycles per instr per SMSP\n", M, W, per); } while (0)
    RUN(0, 4); RUN(0, 8); RUN(0, 16);
    RUN(1, 4); RUN(1, 8); RUN(1, 16);
    RUN(2, 4); RUN(2, 8); RUN(2, 16);
    RUN(3, 4); RUN(3, 16); RUN(4, 4); RUN(4, 16); RUN(5, 4); RUN(5, 16);
    return 0;
}
