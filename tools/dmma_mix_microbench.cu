// What does a scalar FP64 instruction cost on an SM sub-partition whose FP64 pipe is busy with
// DMMAs?  (The antenna-mode DDE GEMM kernel needs ~160 scalar FP64 instructions per sub-partition
// and stage for the antenna phasors beside ~290 DMMAs.)
//
// One CTA on one SM.  Per sub-partition: ND warps issue DMMA.8x8x4 (4 independent accumulators)
// for a fixed count; NF warps issue DFMAs (CH independent chains) until the DMMA warps are done and
// count them.  Model: T = 16 * (DMMAs per sub-partition) + c * (DFMAs per sub-partition); prints c.
// Also: one warp interleaving 1 DFMA every K DMMAs in its own instruction stream.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_mix_microbench dmma_mix_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

__device__ __forceinline__ void dmma(double (&c)[2], double a, double b) {
    asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                 : "+d"(c[0]), "+d"(c[1])
                 : "d"(a), "d"(b));
}

template <int CH>
__global__ void __launch_bounds__(1024, 1)
    mix_kernel(int nd, int nf, int iters, const double *g, double *sink, long long *res) {
    __shared__ volatile int done;
    const int warp = threadIdx.x >> 5, wq = warp >> 2;  // wq: index of the warp inside its sub-partition
    if (threadIdx.x == 0) done = 0;
    __syncthreads();
    const long long t0 = clock64();
    if (wq < nd) {
        double a = g[threadIdx.x], b = g[threadIdx.x + 32], c[4][2];
        for (int k = 0; k < 4; ++k) c[k][0] = c[k][1] = g[k];
#pragma unroll 1
        for (int it = 0; it < iters; ++it) {
#pragma unroll
            for (int k = 0; k < 4; ++k) dmma(c[k], a, b);
        }
        const long long t1 = clock64();
        double s = 0;
        for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1];
        if (s == 1234.5) sink[0] = s;
        __syncwarp();
        if ((threadIdx.x & 31) == 0) {
            res[warp] = t1 - t0;
            atomicAdd((int *)&done, 1);
        }
    } else if (wq < nd + nf) {
        double x[CH], y = g[threadIdx.x], z = g[threadIdx.x + 64];
        for (int k = 0; k < CH; ++k) x[k] = g[k];
        long long n = 0;
        while (done < 4 * nd) {
#pragma unroll
            for (int r = 0; r < 8; ++r)
#pragma unroll
                for (int k = 0; k < CH; ++k) x[k] = fma(x[k], y, z);
            n += 8 * CH;
        }
        double s = 0;
        for (int k = 0; k < CH; ++k) s += x[k];
        if (s == 1234.5) sink[1] = s;
        if ((threadIdx.x & 31) == 0) res[32 + warp] = n;
    }
}

// one warp per sub-partition: K DMMAs then NFI independent DFMAs, in one instruction stream
template <int K, int NFI>
__global__ void __launch_bounds__(128, 1) inline_kernel(int iters, const double *g, double *sink, long long *res) {
    double a = g[threadIdx.x], b = g[threadIdx.x + 32], c[4][2], x[4], y = g[threadIdx.x + 64];
    for (int k = 0; k < 4; ++k) c[k][0] = c[k][1] = g[k], x[k] = g[k + 8];
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < K; ++k) dmma(c[k & 3], a, b);
#pragma unroll
        for (int k = 0; k < NFI; ++k) asm volatile("fma.rn.f64 %0, %0, %1, %1;\n" : "+d"(x[k & 3]) : "d"(y));
    }
    const long long t1 = clock64();
    double s = 0;
    for (int k = 0; k < 4; ++k) s += c[k][0] + c[k][1] + x[k];
    if (s == 1234.5) sink[0] = s;
    if (threadIdx.x == 0) res[0] = t1 - t0;
}

int main() {
    double *g, *sink;
    long long *res, h[64];
    cudaMalloc(&g, 8192 * 8), cudaMalloc(&sink, 64), cudaMalloc(&res, 64 * 8);
    cudaMemset(g, 0, 8192 * 8);
    const int iters = 4000;
    auto run = [&](int nd, int nf, int ch) {
        cudaMemset(res, 0, 64 * 8);
        for (int rep = 0; rep < 2; ++rep) {
            if (ch == 1) mix_kernel<1><<<1, 32 * 4 * (nd + nf)>>>(nd, nf, iters, g, sink, res);
            if (ch == 4) mix_kernel<4><<<1, 32 * 4 * (nd + nf)>>>(nd, nf, iters, g, sink, res);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, res, 64 * 8, cudaMemcpyDeviceToHost);
        long long T = 0, nfma = 0;
        for (int w = 0; w < 32; ++w) T = h[w] > T ? h[w] : T;
        for (int w = 0; w < 32; ++w) nfma += h[32 + w];
        const double dm = 4.0 * iters * nd;  // DMMAs per sub-partition
        const double fm = nfma / 4.0;        // DFMAs per sub-partition
        printf("DMMA warps/SMSP %d  DFMA warps/SMSP %d (chains %d): T %8lld cyc  DMMA %6.0f (%.1f cyc each alone-equivalent)  "
               "DFMA %8.0f  -> pipe share of DMMA %.1f%%, extra cycles per DFMA %.1f  (%s)\n",
               nd, nf, ch, T, dm, T / dm, fm, 100.0 * 16.0 * dm / T, fm > 0 ? (T - 16.0 * dm) / fm : 0.0,
               cudaGetErrorString(cudaGetLastError()));
    };
    run(1, 0, 1);
    run(2, 0, 1);
    run(3, 0, 1);
    run(1, 1, 1);
    run(1, 1, 4);
    run(2, 1, 1);
    run(2, 1, 4);
    run(2, 2, 1);
    run(2, 2, 4);
    run(3, 1, 4);
    run(1, 3, 4);
    auto runi = [&](auto kern, int K, int NFI) {
        for (int rep = 0; rep < 2; ++rep) {
            kern<<<1, 128>>>(iters, g, sink, res);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(h, res, 8, cudaMemcpyDeviceToHost);
        const double per = (double)h[0] / iters;
        printf("one warp/SMSP, %2d DMMA + %d DFMA per iteration: %.1f cycles per iteration (DMMA alone %.1f) -> %.1f extra per DFMA\n",
               K, NFI, per, 16.0 * K, NFI ? (per - 16.0 * K) / NFI : 0.0);
    };
    runi(inline_kernel<8, 0>, 8, 0);
    runi(inline_kernel<8, 1>, 8, 1);
    runi(inline_kernel<8, 2>, 8, 2);
    runi(inline_kernel<8, 4>, 8, 4);
    runi(inline_kernel<8, 8>, 8, 8);
    runi(inline_kernel<4, 4>, 4, 4);
    runi(inline_kernel<16, 16>, 16, 16);
    return 0;
}
