// FP64 tensor-core (DMMA) throughput on sm_100a: is mma.sync f64 worth using for the antenna-mode
// DDE predict (V = P Q^H as a GEMM over sources)?
//
// Measures, for 1 CTA on 1 SM and for a full-chip grid:
//   m8n8k4   (256 FMA per warp instruction),
//   m16n8k4 / m16n8k8 / m16n8k16 (the sm_90+ f64 shapes: 512 / 1024 / 2048 FMA),
// with NACC independent accumulator sets per warp, against the DFMA peak (64 lanes/clk/SM).
// Prints cycles per warp instruction per SM sub-partition and FMA lanes per clock per SM.
//
// build: nvcc -gencode arch=compute_100a,code=sm_100a -O3 -o dmma_microbench dmma_microbench.cu
#include <cstdio>
#include <cuda_runtime.h>

template <int SHAPE, int NACC>
__global__ void __launch_bounds__(1024, 1) dmma_kernel(double *sink, long long *cycles, int iters, const double *g) {
    // fragments: m8n8k4: A 1, B 1, C 2 doubles per thread; m16n8k4: A 2, B 1, C 4;
    // m16n8k8: A 4, B 2, C 4; m16n8k16: A 8, B 4, C 4
    constexpr int NA = SHAPE == 0 ? 1 : (SHAPE == 1 ? 2 : (SHAPE == 2 ? 4 : 8));
    constexpr int NB = SHAPE == 0 ? 1 : (SHAPE == 1 ? 1 : (SHAPE == 2 ? 2 : 4));
    constexpr int NC = SHAPE == 0 ? 2 : 4;
    double a[NA], b[NB], c[NACC][NC];
#pragma unroll
    for (int i = 0; i < NA; ++i) a[i] = g[threadIdx.x + 32 * i];
#pragma unroll
    for (int i = 0; i < NB; ++i) b[i] = g[threadIdx.x + 32 * i + 512];
#pragma unroll
    for (int k = 0; k < NACC; ++k)
#pragma unroll
        for (int i = 0; i < NC; ++i) c[k][i] = g[threadIdx.x + k + i];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int k = 0; k < NACC; ++k) {
            if constexpr (SHAPE == 0) {
                asm volatile("mma.sync.aligned.m8n8k4.row.col.f64.f64.f64.f64 {%0,%1}, {%2}, {%3}, {%0,%1};\n"
                             : "+d"(c[k][0]), "+d"(c[k][1])
                             : "d"(a[0]), "d"(b[0]));
            } else if constexpr (SHAPE == 1) {
                asm volatile(
                    "mma.sync.aligned.m16n8k4.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5}, {%6}, {%0,%1,%2,%3};\n"
                    : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
                    : "d"(a[0]), "d"(a[1]), "d"(b[0]));
            } else if constexpr (SHAPE == 2) {
                asm volatile(
                    "mma.sync.aligned.m16n8k8.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, {%4,%5,%6,%7}, {%8,%9}, "
                    "{%0,%1,%2,%3};\n"
                    : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
                    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(b[0]), "d"(b[1]));
            } else {
                asm volatile(
                    "mma.sync.aligned.m16n8k16.row.col.f64.f64.f64.f64 {%0,%1,%2,%3}, "
                    "{%4,%5,%6,%7,%8,%9,%10,%11}, {%12,%13,%14,%15}, {%0,%1,%2,%3};\n"
                    : "+d"(c[k][0]), "+d"(c[k][1]), "+d"(c[k][2]), "+d"(c[k][3])
                    : "d"(a[0]), "d"(a[1]), "d"(a[2]), "d"(a[3]), "d"(a[4]), "d"(a[5]), "d"(a[6]), "d"(a[7]),
                      "d"(b[0]), "d"(b[1]), "d"(b[2]), "d"(b[3]));
            }
        }
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int k = 0; k < NACC; ++k)
#pragma unroll
        for (int i = 0; i < NC; ++i) s += c[k][i];
    if (s == 1234.5) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

// the DFMA loop the DDE consumer runs today, for comparison on the same clock: 8 independent
// chains, one shared operand
__global__ void __launch_bounds__(1024, 1) dfma_kernel(double *sink, long long *cycles, int iters, const double *g) {
    double acc[8], x = g[threadIdx.x], y[8];
#pragma unroll
    for (int i = 0; i < 8; ++i) acc[i] = g[threadIdx.x + i], y[i] = g[threadIdx.x + 64 + i];
    __syncthreads();
    const long long t0 = clock64();
#pragma unroll 1
    for (int it = 0; it < iters; ++it) {
#pragma unroll
        for (int i = 0; i < 8; ++i) acc[i] = fma(x, y[i], acc[i]);
    }
    const long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int i = 0; i < 8; ++i) s += acc[i];
    if (s == 1234.5) sink[0] = s;
    if (threadIdx.x == 0 && blockIdx.x == 0) cycles[0] = t1 - t0;
}

template <int SHAPE, int NACC>
void run(int warps, double *sink, long long *cyc, const double *g, int nsm, double clock_ghz) {
    const int iters = 2000;
    const long long fma_per_instr = SHAPE == 0 ? 256 : (SHAPE == 1 ? 512 : (SHAPE == 2 ? 1024 : 2048));
    const char *names[] = {"m8n8k4", "m16n8k4", "m16n8k8", "m16n8k16"};
    long long h;
    // one CTA on one SM: cycle-accurate
    for (int rep = 0; rep < 2; ++rep) {
        dmma_kernel<SHAPE, NACC><<<1, 32 * warps>>>(sink, cyc, iters, g);
        cudaDeviceSynchronize();
    }
    cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
    const double instr = (double)iters * NACC * warps;
    const double lanes = instr * fma_per_instr / (double)h;
    // whole chip: wall clock
    cudaEvent_t e0, e1;
    cudaEventCreate(&e0), cudaEventCreate(&e1);
    dmma_kernel<SHAPE, NACC><<<nsm, 32 * warps>>>(sink, cyc, iters * 4, g);
    cudaEventRecord(e0);
    dmma_kernel<SHAPE, NACC><<<nsm, 32 * warps>>>(sink, cyc, iters * 4, g);
    cudaEventRecord(e1);
    cudaEventSynchronize(e1);
    float ms;
    cudaEventElapsedTime(&ms, e0, e1);
    const double tflops = 2.0 * instr * 4 * fma_per_instr * nsm / (ms * 1e-3) / 1e12;
    printf("%-9s acc=%d warps=%2d : %7.2f cycles/instr/SMSP  %6.1f FMA lanes/clk/SM  chip %6.2f TFLOP/s (err %s)\n",
           names[SHAPE], NACC, warps, (double)h / (instr / 4.0), lanes, tflops,
           cudaGetErrorString(cudaGetLastError()));
    (void)clock_ghz;
}

int main() {
    double *sink, *g;
    long long *cyc, h;
    cudaMalloc(&sink, 64), cudaMalloc(&cyc, 64), cudaMalloc(&g, 8192 * 8);
    cudaMemset(g, 0, 8192 * 8);
    cudaDeviceProp prop;
    cudaGetDeviceProperties(&prop, 0);
    int khz = 0;
    cudaDeviceGetAttribute(&khz, cudaDevAttrClockRate, 0);
    const int nsm = prop.multiProcessorCount;
    printf("%s, %d SMs, %d kHz\n", prop.name, nsm, khz);
    for (int w : {4, 8, 16, 32}) {
        for (int rep = 0; rep < 2; ++rep) {
            dfma_kernel<<<1, 32 * w>>>(sink, cyc, 4000, g);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        printf("DFMA 8 chains warps=%2d : %6.1f FMA lanes/clk/SM\n", w, 4000.0 * 8 * w * 32 / (double)h);
    }
    const double ghz = khz * 1e-6;
    for (int w : {1, 4, 8, 16}) run<0, 1>(w, sink, cyc, g, nsm, ghz);
    for (int w : {4, 8, 16}) run<0, 4>(w, sink, cyc, g, nsm, ghz);
    for (int w : {4, 8, 16}) run<0, 8>(w, sink, cyc, g, nsm, ghz);
    for (int w : {4, 8, 16}) run<1, 4>(w, sink, cyc, g, nsm, ghz);
    for (int w : {4, 8, 16}) run<2, 4>(w, sink, cyc, g, nsm, ghz);
    for (int w : {4, 8, 16}) run<3, 4>(w, sink, cyc, g, nsm, ghz);
    for (int w : {4, 16}) run<3, 8>(w, sink, cyc, g, nsm, ghz);
    return 0;
}
