// FP64 pipe microbenchmark for B200: DFMA latency and throughput vs warps/ILP.
// Build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o fp64_microbench fp64_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int CHAINS>
__global__ void dfma_kernel(double *sink, long long *cycles, int iters, double a, double b) {
    double acc[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) acc[k] = threadIdx.x + k;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) acc[k] = fma(acc[k], a, b);
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += acc[k];
    if (s == 1234.5) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// rotation-like dependent pattern: z = z*d (DMUL + DFMA dependent), CHAINS independent
template <int CHAINS>
__global__ void rot_kernel(double *sink, long long *cycles, int iters, double dr, double di) {
    double zr[CHAINS], zi[CHAINS], accr[CHAINS], acci[CHAINS];
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) { zr[k] = 1.0 + threadIdx.x * 1e-9; zi[k] = k * 1e-9; accr[k] = acci[k] = 0; }
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
#pragma unroll
        for (int k = 0; k < CHAINS; ++k) {
            accr[k] = fma(zr[k], 1.5, accr[k]);
            acci[k] = fma(zi[k], 1.5, acci[k]);
            double t = zr[k] * dr - zi[k] * di;
            zi[k] = zr[k] * di + zi[k] * dr;
            zr[k] = t;
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < CHAINS; ++k) s += zr[k] + zi[k] + accr[k] + acci[k];
    if (s == 1234.5) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

__global__ void sincos_kernel(double *sink, long long *cycles, int iters, double x0) {
    double x = x0 + threadIdx.x * 0.37, acc = 0;
    __syncthreads();
    long long t0 = clock64();
    for (int i = 0; i < iters; ++i) {
        double s, c;
        sincos(x, &s, &c);
        acc += s * c;
        x += 1.2345;
    }
    long long t1 = clock64();
    if (acc == 1234.5) sink[0] = acc;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
    double *sink;
    long long *cyc, h;
    cudaMalloc(&sink, 64);
    cudaMalloc(&cyc, 64);
    const int iters = 4096;
    int warps_list[] = {1, 2, 4, 8, 16, 32};
    printf("== DFMA: cycles per DFMA warp-instr per SMSP-equivalent (1 block on 1 SM) ==\n");
#define RUN(KERN, CH, W, NINSTR, ...)                                                     \
    do {                                                                                  \
        KERN<CH><<<1, 32 * W>>>(sink, cyc, iters, __VA_ARGS__);                           \
        cudaDeviceSynchronize();                                                          \
        KERN<CH><<<1, 32 * W>>>(sink, cyc, iters, __VA_ARGS__);                           \
        cudaDeviceSynchronize();                                                          \
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);                                   \
        double per_sm = (double)h / ((double)iters * CH * NINSTR * W);                    \
        printf("  %-5s chains=%d warps=%2d : %8.3f cyc/warp-instr/SM  -> %6.1f lanes/clk/SM\n", \
               #KERN, CH, W, per_sm, 32.0 / per_sm);                                      \
    } while (0)
    for (int w : warps_list) {
        switch (w) {
            default: break;
        }
    }
    RUN(dfma_kernel, 1, 1, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 2, 1, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 4, 1, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 8, 1, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 1, 4, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 2, 4, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 4, 4, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 8, 4, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 1, 8, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 2, 8, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 4, 8, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 8, 8, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 1, 16, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 4, 16, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 8, 16, 1, 0.999, 1e-9);
    RUN(dfma_kernel, 8, 32, 1, 0.999, 1e-9);
    printf("== rotation+accumulate (6 DP instr per step per chain) ==\n");
    RUN(rot_kernel, 1, 1, 6, 0.9999, 0.01);
    RUN(rot_kernel, 2, 1, 6, 0.9999, 0.01);
    RUN(rot_kernel, 4, 1, 6, 0.9999, 0.01);
    RUN(rot_kernel, 1, 4, 6, 0.9999, 0.01);
    RUN(rot_kernel, 2, 4, 6, 0.9999, 0.01);
    RUN(rot_kernel, 4, 4, 6, 0.9999, 0.01);
    RUN(rot_kernel, 1, 8, 6, 0.9999, 0.01);
    RUN(rot_kernel, 2, 8, 6, 0.9999, 0.01);
    RUN(rot_kernel, 4, 8, 6, 0.9999, 0.01);
    RUN(rot_kernel, 2, 16, 6, 0.9999, 0.01);
    RUN(rot_kernel, 4, 16, 6, 0.9999, 0.01);
    printf("== sincos(double): cycles per call per warp ==\n");
    for (int w : {1, 4, 8, 16}) {
        for (double x0 : {1.0, 1000.0, 2.0e5}) {
            sincos_kernel<<<1, 32 * w>>>(sink, cyc, 1024, x0);
            cudaDeviceSynchronize();
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            printf("  warps=%2d x0=%8.0f : %8.1f cycles/sincos/warp (SM-time per warp-call %.1f)\n", w, x0,
                   (double)h / 1024, (double)h / 1024 / w);
        }
    }
    int clk;
    cudaDeviceGetAttribute(&clk, cudaDevAttrClockRate, 0);
    printf("clock rate attr: %d kHz\n", clk);
    return 0;
}
