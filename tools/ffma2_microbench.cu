// Does the packed FP32 FMA of sm_100a (PTX fma.rn.f32x2 -> SASS FFMA2) double the FP32 rate?
// cycles per warp-instruction per SM sub-partition for FFMA and FFMA2, operands all distinct or
// one shared between neighbours.  (Input for the complex64 phasor-stream variants, where the
// (re, im) pair of a phasor / accumulator is a natural f32x2.)
#include <cstdio>
#include <cuda_runtime.h>
constexpr int N = 12;
__device__ __forceinline__ unsigned long long pk(float a, float b) {
    return ((unsigned long long)__float_as_uint(b) << 32) | __float_as_uint(a);
}
template <int MODE>
__global__ void __launch_bounds__(512, 1) k(float *sink, long long *cycles, int iters, const float *g) {
    float acc[N], z[N], w[N];
    unsigned long long acc2[N], z2[N], w2[N];
#pragma unroll
    for (int i = 0; i < N; ++i) {
        acc[i] = g[i + threadIdx.x]; z[i] = g[64 + i + threadIdx.x]; w[i] = g[128 + i + threadIdx.x];
        acc2[i] = pk(acc[i], z[i]); z2[i] = pk(z[i], w[i]); w2[i] = pk(w[i], acc[i]);
    }
    __syncthreads();
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int rep = 0; rep < 4; ++rep) {
#pragma unroll
            for (int i = 0; i < N; ++i) {
                if (MODE == 0) acc[i] = fmaf(z[i], w[i], acc[i]);
                if (MODE == 1) acc[i] = fmaf(z[0], w[i], acc[i]);
                if (MODE == 2) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i]) : "l"(z2[i]), "l"(w2[i]));
                if (MODE == 3) asm volatile("fma.rn.f32x2 %0, %1, %2, %0;" : "+l"(acc2[i]) : "l"(z2[0]), "l"(w2[i]));
            }
        }
    }
    long long t1 = clock64();
    float s = 0;
#pragma unroll
    for (int i = 0; i < N; ++i) s += acc[i] + z[i] + w[i] + __uint_as_float((unsigned)acc2[i]) + __uint_as_float((unsigned)(acc2[i] >> 32));
    if (s == 1234.5f) sink[0] = s;
    if (threadIdx.x == 0) cycles[0] = t1 - t0;
}
int main() {
    float *sink, *g; long long *cyc, h;
    cudaMalloc(&sink, 64); cudaMalloc(&cyc, 64); cudaMalloc(&g, 4096 * 4);
    cudaMemset(g, 0, 4096 * 4);
    const int iters = 4000;
    const char *names[4] = {"FFMA  3 distinct", "FFMA  1 shared  ", "FFMA2 3 distinct", "FFMA2 1 shared  "};
#define RUN(M, W) do { k<M><<<1, 32 * W>>>(sink, cyc, iters, g); cudaDeviceSynchronize(); k<M><<<1, 32 * W>>>(sink, cyc, iters, g); cudaDeviceSynchronize(); \
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); double per = (double)h / (iters * 4.0 * N * (W / 4.0)); \
    printf("%s warps %2d: %.3f cycles per instruction per SMSP (%s)\n", names[M], W, per, cudaGetErrorString(cudaGetLastError())); } while (0)
    RUN(0, 8); RUN(0, 16); RUN(1, 16); RUN(2, 8); RUN(2, 16); RUN(3, 16);
    return 0;
}
