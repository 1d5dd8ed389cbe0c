// Isolated consumer loop of the phasor-stream kernel with the three-term recurrence
// z_{j+1} = c z_j - z_{j-1}: cycles per warp-term per SM sub-partition for several source-level
// arrangements.  Pipe bound = 8 cycles (4 DFMA x 2 cycles).
#include <cstdio>
#include <cuda_runtime.h>
struct C2 { double re, im; };
template <int CH, int VAR>
__global__ void __launch_bounds__(512, 1) kt(double *out, const double *g, int yt, int iters, long long *cyc) {
    extern __shared__ __align__(16) unsigned char smem[];
    C2 *anch = (C2 *)smem;
    C2 *dstp = anch + 8 * 512;
    double *wt = (double *)(dstp + 8 * 512);
    for (int i = threadIdx.x; i < 8 * 512 * 2; i += blockDim.x) ((double *)anch)[i] = g[i & 1023];
    for (int i = threadIdx.x; i < 8 * 512 * 2; i += blockDim.x) ((double *)dstp)[i] = g[(i + 7) & 1023];
    for (int i = threadIdx.x; i < 8 * 256; i += blockDim.x) wt[i] = g[(i + 3) & 1023];
    __syncthreads();
    double are[CH], aim[CH];
#pragma unroll
    for (int j = 0; j < CH; ++j) are[j] = aim[j] = 0;
    const int fo = ((threadIdx.x >> 5) & 7) * CH;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
        if (VAR == 0 || VAR == 1 || VAR == 3) {
#pragma unroll 1
            for (int yl = 0; yl < yt; ++yl) {
                C2 z = anch[yl * 512 + threadIdx.x];
                const C2 d = dstp[yl * 512 + threadIdx.x];
                const double *w = wt + yl * 256 + fo;
                C2 zp = z;
                const double c2 = d.re + d.re;
                double2 wreg = *reinterpret_cast<const double2 *>(w);
#pragma unroll
                for (int j = 0; j < CH; j += 2) {
                    double2 wv;
                    if (VAR == 1) wv = wreg; else wv = *reinterpret_cast<const double2 *>(w + j);
#pragma unroll
                    for (int g2 = 0; g2 < 2; ++g2) {
                        const double ww = g2 ? wv.y : wv.x;
                        if (VAR == 3) {
                            // recurrence first, accumulate the OLD z afterwards (z.im shared b-slot)
                            C2 zn;
                            if (j + g2 == 0) { zn.re = z.re * d.re - z.im * d.im; zn.im = z.re * d.im + z.im * d.re; }
                            else { zn.re = fma(c2, z.re, -zp.re); zn.im = fma(c2, z.im, -zp.im); }
                            aim[j + g2] = fma(ww, z.im, aim[j + g2]);
                            are[j + g2] = fma(ww, z.re, are[j + g2]);
                            zp = z; z = zn;
                        } else {
                            are[j + g2] = fma(z.re, ww, are[j + g2]);
                            aim[j + g2] = fma(z.im, ww, aim[j + g2]);
                            C2 zn;
                            if (j + g2 == 0) { zn.re = z.re * d.re - z.im * d.im; zn.im = z.re * d.im + z.im * d.re; }
                            else { zn.re = fma(c2, z.re, -zp.re); zn.im = fma(c2, z.im, -zp.im); }
                            zp = z; z = zn;
                        }
                    }
                }
            }
        } else if (VAR == 2) {  // two y per iteration: two independent chains
#pragma unroll 1
            for (int yl = 0; yl < yt; yl += 2) {
                C2 za = anch[yl * 512 + threadIdx.x], zb = anch[(yl + 1) * 512 + threadIdx.x];
                const C2 da = dstp[yl * 512 + threadIdx.x], db = dstp[(yl + 1) * 512 + threadIdx.x];
                const double *wa = wt + yl * 256 + fo;
                const double *wb = wa + 256;
                C2 zpa = za, zpb = zb;
                const double ca = da.re + da.re, cb = db.re + db.re;
#pragma unroll
                for (int j = 0; j < CH; j += 2) {
                    const double2 wva = *reinterpret_cast<const double2 *>(wa + j);
                    const double2 wvb = *reinterpret_cast<const double2 *>(wb + j);
#pragma unroll
                    for (int g2 = 0; g2 < 2; ++g2) {
                        const double w1 = g2 ? wva.y : wva.x, w2 = g2 ? wvb.y : wvb.x;
                        are[j + g2] = fma(za.re, w1, are[j + g2]);
                        aim[j + g2] = fma(za.im, w1, aim[j + g2]);
                        are[j + g2] = fma(zb.re, w2, are[j + g2]);
                        aim[j + g2] = fma(zb.im, w2, aim[j + g2]);
                        C2 zn;
                        if (j + g2 == 0) { zn.re = za.re * da.re - za.im * da.im; zn.im = za.re * da.im + za.im * da.re; }
                        else { zn.re = fma(ca, za.re, -zpa.re); zn.im = fma(ca, za.im, -zpa.im); }
                        zpa = za; za = zn;
                        if (j + g2 == 0) { zn.re = zb.re * db.re - zb.im * db.im; zn.im = zb.re * db.im + zb.im * db.re; }
                        else { zn.re = fma(cb, zb.re, -zpb.re); zn.im = fma(cb, zb.im, -zpb.im); }
                        zpb = zb; zb = zn;
                    }
                }
            }
        }
    }
    long long t1 = clock64();
    double s = 0;
#pragma unroll
    for (int j = 0; j < CH; ++j) s += are[j] + aim[j];
    out[blockIdx.x * blockDim.x + threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
template <int VAR>
void run(double *out, double *g, long long *cyc) {
    const int smem = 8 * 512 * 16 * 2 + 8 * 256 * 8;
    cudaFuncSetAttribute(kt<16, VAR>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 300, yt = 8;
    long long h;
    for (int w : {4, 8, 16}) {
        for (int rep = 0; rep < 2; ++rep) {
            kt<16, VAR><<<1, 32 * w, smem>>>(out, g, yt, iters, cyc);
            cudaDeviceSynchronize();
        }
        cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
        double terms_per_smsp = (double)iters * yt * 16 * (w / 4.0);
        printf("var %d warps %2d: %.2f cycles per warp-term per SMSP (pipe bound 8)  [%s]\n", VAR, w,
               h / terms_per_smsp, cudaGetErrorString(cudaGetLastError()));
    }
}
int main() {
    double *out, *g;
    long long *cyc;
    cudaMalloc(&out, 1 << 20);
    cudaMalloc(&g, 1024 * 8);
    cudaMalloc(&cyc, 64);
    cudaMemset(g, 0, 1024 * 8);
    run<0>(out, g, cyc);
    run<1>(out, g, cyc);
    run<2>(out, g, cyc);
    run<3>(out, g, cyc);
    return 0;
}
