// Isolates the consume loop of phasor_stream_kernel<1,false,false,double,16,16>:
// 2 rotation chains + 32 accumulators per thread, W broadcast from shared memory.
#include <cstdio>
#include <cuda_runtime.h>

struct C2 { double re, im; };
__device__ __forceinline__ C2 cmul(C2 a, C2 b) { return {a.re*b.re - a.im*b.im, a.re*b.im + a.im*b.re}; }

template <int VARIANT, int CH>
__global__ void __launch_bounds__(512, 1) consume_kernel(double *sink, long long *cycles, int iters, const double* gw) {
    __shared__ __align__(16) double wt[2 * 256];
    __shared__ __align__(16) C2 anch[2 * 512];
    for (int i = threadIdx.x; i < 512; i += blockDim.x) wt[i] = gw[i];
    for (int i = threadIdx.x; i < 1024; i += blockDim.x) anch[i] = {1.0 - 1e-9 * i, 1e-5 * i};
    __syncthreads();
    double are[CH], aim[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) are[j] = aim[j] = 0.0;
    const int warp = threadIdx.x >> 5;
    const int fo = (warp & 7) * CH;
    long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
        C2 za = anch[threadIdx.x], zb = anch[512 + threadIdx.x];
        const C2 da = anch[(threadIdx.x + 32) & 511], db = anch[512 + ((threadIdx.x + 64) & 511)];
        const double *wa = wt + (fo & 127), *wb = wt + 256 + (fo & 127);
#pragma unroll
        for (int j = 0; j < CH; j += 4) {
            double wva[4], wvb[4];
            if (VARIANT == 1) {
                double2 v0 = *reinterpret_cast<const double2*>(wa + j), v1 = *reinterpret_cast<const double2*>(wa + j + 2);
                double2 u0 = *reinterpret_cast<const double2*>(wb + j), u1 = *reinterpret_cast<const double2*>(wb + j + 2);
                wva[0]=v0.x; wva[1]=v0.y; wva[2]=v1.x; wva[3]=v1.y; wvb[0]=u0.x; wvb[1]=u0.y; wvb[2]=u1.x; wvb[3]=u1.y;
            } else {
#pragma unroll
                for (int g = 0; g < 4; ++g) { wva[g] = 1.25; wvb[g] = 0.75; }
            }
#pragma unroll
            for (int g = 0; g < 4; ++g) {
                if (VARIANT != 3) {
                    are[j+g] = fma(za.re, wva[g], are[j+g]); aim[j+g] = fma(za.im, wva[g], aim[j+g]);
                    are[j+g] = fma(zb.re, wvb[g], are[j+g]); aim[j+g] = fma(zb.im, wvb[g], aim[j+g]);
                }
                if (j + g + 1 < CH) { za = cmul(za, da); zb = cmul(zb, db); }
            }
        }
        if (VARIANT == 3) { are[0] += za.re + zb.re; aim[0] += za.im + zb.im; }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) s += are[j] + aim[j];
    if (s == 1234.5) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

int main() {
    double *sink, *gw; long long *cyc, h;
    cudaMalloc(&sink, 64); cudaMalloc(&cyc, 64); cudaMalloc(&gw, 512 * 8);
    cudaMemset(gw, 0, 512 * 8);
    const int iters = 2000;
    const int CH = 16;
#define RUN(V, W)  do { \
        consume_kernel<V, CH><<<1, 32 * W>>>(sink, cyc, iters, gw); cudaDeviceSynchronize(); \
        consume_kernel<V, CH><<<1, 32 * W>>>(sink, cyc, iters, gw); cudaDeviceSynchronize(); \
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost); \
        double ndp = (V == 3 ? 2.0 * 4 * (CH - 1) : 2.0 * (2 * CH + 4 * (CH - 1))) ; \
        double per = (double)h / (iters * ndp * W / 4.0); \
        printf("variant %d warps %2d: %.3f cycles per DP warp-instr per SMSP (ideal 2.0) -> %.1f%%  [%s]\n", V, W, per, 200.0 / per, cudaGetErrorString(cudaGetLastError())); \
    } while (0)
    RUN(1, 16); RUN(1, 8); RUN(1, 4);
    RUN(2, 16); RUN(2, 8); RUN(2, 4);
    RUN(3, 16); RUN(3, 8); RUN(3, 4);
    return 0;
}
