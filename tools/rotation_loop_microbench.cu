#include <cuda_runtime.h>
struct C2 { double re, im; };
template <int CH, int VAR>
__global__ void __launch_bounds__(512,1) k(double* out, const double* g, int yt, int iters) {
    extern __shared__ __align__(16) unsigned char smem[];
    C2* anch = (C2*)smem; C2* dstp = anch + 8*512; double* wt = (double*)(dstp + 8*512);
    for (int i = threadIdx.x; i < 8*512*2; i += blockDim.x) ((double*)anch)[i] = g[i & 1023];
    for (int i = threadIdx.x; i < 8*512*2; i += blockDim.x) ((double*)dstp)[i] = g[(i+7) & 1023];
    for (int i = threadIdx.x; i < 8*256; i += blockDim.x) wt[i] = g[(i+3) & 1023];
    __syncthreads();
    double are[CH], aim[CH];
    #pragma unroll
    for (int j=0;j<CH;++j) are[j]=aim[j]=0;
    const int fo = ((threadIdx.x >> 5) & 7) * CH;
    for (int it = 0; it < iters; ++it) {
      #pragma unroll 1
      for (int yl = 0; yl < yt; ++yl) {
        C2 z = anch[yl*512 + threadIdx.x];
        const C2 d = dstp[yl*512 + threadIdx.x];
        const double* w = wt + yl*256 + fo;
        #pragma unroll
        for (int j = 0; j < CH; j += 2) {
            const double2 wv = *reinterpret_cast<const double2*>(w + j);
            #pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
                const double ww = g2 ? wv.y : wv.x;
                if (VAR == 0) {
                    are[j+g2] = fma(z.re, ww, are[j+g2]);
                    aim[j+g2] = fma(z.im, ww, aim[j+g2]);
                    const double t1 = z.im * d.im, t2 = z.re * d.im;
                    const double nre = fma(z.re, d.re, -t1), nim = fma(z.im, d.re, t2);
                    z.re = nre; z.im = nim;
                } else {
                    const double t1 = z.im * d.im;
                    const double t2 = z.re * d.im;
                    are[j+g2] = fma(z.re, ww, are[j+g2]);
                    const double nre = fma(z.re, d.re, -t1);
                    const double nim = fma(z.im, d.re, t2);
                    aim[j+g2] = fma(z.im, ww, aim[j+g2]);
                    z.re = nre; z.im = nim;
                }
            }
        }
      }
    }
    double s = 0;
    #pragma unroll
    for (int j=0;j<CH;++j) s += are[j] + aim[j];
    out[blockIdx.x*blockDim.x+threadIdx.x] = s;
}
template __global__ void k<16,0>(double*, const double*, int, int);
template __global__ void k<16,1>(double*, const double*, int, int);
#include <cstdio>
template <int CH, int VAR>
__global__ void __launch_bounds__(512,1) kt(double* out, const double* g, int yt, int iters, long long* cyc) {
    extern __shared__ __align__(16) unsigned char smem[];
    C2* anch = (C2*)smem; C2* dstp = anch + 8*512; double* wt = (double*)(dstp + 8*512);
    for (int i = threadIdx.x; i < 8*512*2; i += blockDim.x) ((double*)anch)[i] = g[i & 1023];
    for (int i = threadIdx.x; i < 8*512*2; i += blockDim.x) ((double*)dstp)[i] = g[(i+7) & 1023];
    for (int i = threadIdx.x; i < 8*256; i += blockDim.x) wt[i] = g[(i+3) & 1023];
    __syncthreads();
    double are[CH], aim[CH];
    #pragma unroll
    for (int j=0;j<CH;++j) are[j]=aim[j]=0;
    const int fo = ((threadIdx.x >> 5) & 7) * CH;
    long long t0 = clock64();
    for (int it = 0; it < iters; ++it) {
      if (VAR == 0) {
      #pragma unroll 1
      for (int yl = 0; yl < yt; ++yl) {
        C2 z = anch[yl*512 + threadIdx.x];
        const C2 d = dstp[yl*512 + threadIdx.x];
        const double* w = wt + yl*256 + fo;
        #pragma unroll
        for (int j = 0; j < CH; j += 2) {
            const double2 wv = *reinterpret_cast<const double2*>(w + j);
            #pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
                const double ww = g2 ? wv.y : wv.x;
                are[j+g2] = fma(z.re, ww, are[j+g2]);
                aim[j+g2] = fma(z.im, ww, aim[j+g2]);
                const double t1 = z.im * d.im, t2 = z.re * d.im;
                const double nre = fma(z.re, d.re, -t1), nim = fma(z.im, d.re, t2);
                z.re = nre; z.im = nim;
            }
        }
      }
      } else {
      #pragma unroll 1
      for (int yl = 0; yl < yt; yl += 2) {
        C2 za = anch[yl*512 + threadIdx.x], zb = anch[(yl+1)*512 + threadIdx.x];
        const C2 da = dstp[yl*512 + threadIdx.x], db = dstp[(yl+1)*512 + threadIdx.x];
        const double* wa = wt + yl*256 + fo; const double* wb = wa + 256;
        #pragma unroll
        for (int j = 0; j < CH; j += 2) {
            const double2 wva = *reinterpret_cast<const double2*>(wa + j);
            const double2 wvb = *reinterpret_cast<const double2*>(wb + j);
            #pragma unroll
            for (int g2 = 0; g2 < 2; ++g2) {
                const double w1 = g2 ? wva.y : wva.x, w2 = g2 ? wvb.y : wvb.x;
                are[j+g2] = fma(za.re, w1, are[j+g2]); aim[j+g2] = fma(za.im, w1, aim[j+g2]);
                are[j+g2] = fma(zb.re, w2, are[j+g2]); aim[j+g2] = fma(zb.im, w2, aim[j+g2]);
                { const double t1 = za.im * da.im, t2 = za.re * da.im; const double nre = fma(za.re, da.re, -t1), nim = fma(za.im, da.re, t2); za.re = nre; za.im = nim; }
                { const double t1 = zb.im * db.im, t2 = zb.re * db.im; const double nre = fma(zb.re, db.re, -t1), nim = fma(zb.im, db.re, t2); zb.re = nre; zb.im = nim; }
            }
        }
      }
      }
    }
    long long t1 = clock64();
    double s = 0;
    #pragma unroll
    for (int j=0;j<CH;++j) s += are[j] + aim[j];
    out[blockIdx.x*blockDim.x+threadIdx.x] = s;
    if (threadIdx.x == 0) cyc[0] = t1 - t0;
}
int main() {
    double *out, *g; long long* cyc, h;
    cudaMalloc(&out, 1<<20); cudaMalloc(&g, 1024*8); cudaMalloc(&cyc, 64);
    cudaMemset(g, 0, 1024*8);
    const int smem = 8*512*16*2 + 8*256*8;
    cudaFuncSetAttribute(kt<16,0>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    cudaFuncSetAttribute(kt<16,1>, cudaFuncAttributeMaxDynamicSharedMemorySize, smem);
    const int iters = 300, yt = 8;
    for (int w : {4, 8, 16}) {
        for (int var = 0; var < 2; ++var) {
            for (int rep = 0; rep < 2; ++rep) {
                if (var == 0) kt<16,0><<<1, 32*w, smem>>>(out, g, yt, iters, cyc); else kt<16,1><<<1, 32*w, smem>>>(out, g, yt, iters, cyc);
                cudaDeviceSynchronize();
            }
            cudaMemcpy(&h, cyc, 8, cudaMemcpyDeviceToHost);
            double terms_per_smsp = (double)iters * yt * 16 * (w / 4.0);
            printf("var %d (chains=%d) warps %2d: %.2f cycles per warp-term per SMSP (ideal 12)  [%s]\n", var, var+1, w, h / terms_per_smsp, cudaGetErrorString(cudaGetLastError()));
        }
    }
    return 0;
}
