// Phasor-stream kernels: im_to_vis, vis_to_im (and the point-source part of the
// fused predict) for sm_100a.
//
// Mathematical object (africanus/dft/kernels.py:23-69, 83-148):
//     acc[x, f, c] += exp(i * cst * (X[x] . Y[y]) * nu_f) * W[y, f, c]      summed over y
// im_to_vis : owners x = rows (u,v,w), streamed y = sources (l,m,n), W = image, complex acc
// vis_to_im : owners x = sources,      streamed y = rows,            W = vis,   real part only
//
// B200 mapping (the path is FP64-pipe bound, bytes/term << 1, no GEMM is pretended):
//  * one lane owns one x (32 owners per warp), one warp owns a run of CH channels, so a
//    thread keeps CH*ncorr accumulators in registers for the whole y loop;
//  * y is streamed in tiles of kYT through shared memory: the W tile (read once per CTA,
//    broadcast to all 32 lanes) and, per (x, y) pair, the channel-run ANCHORS
//    exp(i*phi*nu_f0) plus the per-channel step exp(i*phi*dnu);
//  * inside a channel run the phasor advances by ONE complex rotation per channel
//    (2 DMUL + 2 DFMA) instead of a sincos; every run restarts from an anchor, so the
//    recurrence never runs longer than CH <= 32 steps.  The anchors themselves come from
//    3 sincos per (x, y) pair (first run, run-to-run step, channel step) and a rotation per
//    further run -- amortised over all channels of the CTA;
//  * the phase argument phi = cst*(l*u + m*v + n*w) is formed in FP64 in the reference's
//    operation order with explicitly rounded ops (no FMA contraction), and the anchor phase
//    is fl(phi*nu_f0) exactly as the reference computes it;
//  * non-equispaced channels (exact mode) take one sincos per term instead;
//  * two y's are processed per inner iteration so each thread has two independent
//    rotation chains in flight (DFMA latency hiding at 8 warps/SM).
#include "afr_dft.cuh"

namespace afr {

namespace {

constexpr int kThreads = 256;
constexpr int kWarps = kThreads / 32;
constexpr int kYT = 8;  // streamed items per shared-memory tile (even)

struct DftParams {
    const double *xc;      // (nx,3) owner coordinates
    const double *yc;      // (ny,3) streamed coordinates
    const double *w;       // (ny,nchan,wstride) real or complex
    const uint8_t *flags;  // (ny,nchan,wstride) or nullptr
    const double *freq;    // (nchan,)
    void *out;             // (nsplit,nx,nchan,wstride) in the accumulator type
    double cst;
    long long nx, ny;
    long long ysplit;            // y items per grid.z slice (multiple of kYT)
    long long out_split_stride;  // elements (of the output scalar type) between slices
    int nchan;
    int wstride;  // correlations in W / out
    int coff;     // first correlation handled by this launch
    int nck;      // channel runs per CTA: 1, 2, 4 or 8
    int f32dot;
};

template <int N>
__device__ __forceinline__ void load_vec(const double *src, double (&dst)[N]) {
    static_assert(N % 2 == 0, "16-byte granularity");
    const double2 *s = reinterpret_cast<const double2 *>(src);
#pragma unroll
    for (int i = 0; i < N / 2; ++i) {
        double2 v = s[i];
        dst[2 * i] = v.x;
        dst[2 * i + 1] = v.y;
    }
}

template <int N>
__device__ __forceinline__ void load_vec(const float *src, float (&dst)[N]) {
    static_assert(N % 4 == 0, "16-byte granularity");
    const float4 *s = reinterpret_cast<const float4 *>(src);
#pragma unroll
    for (int i = 0; i < N / 4; ++i) {
        float4 v = s[i];
        dst[4 * i] = v.x;
        dst[4 * i + 1] = v.y;
        dst[4 * i + 2] = v.z;
        dst[4 * i + 3] = v.w;
    }
}

__device__ __noinline__ C2<double> cis_noinline(double p) { return cis(p); }

// acc (+)= z * w for the NCORR correlations of one channel
template <int NCORR, bool WC, bool ADJ, typename ACC>
__device__ __forceinline__ void accumulate(ACC (&are)[NCORR], ACC (&aim)[ADJ ? 1 : NCORR],
                                           const C2<ACC> z, const ACC *wv) {
#pragma unroll
    for (int c = 0; c < NCORR; ++c) {
        if (WC) {
            const ACC wr = wv[2 * c], wi = wv[2 * c + 1];
            are[c] = fma(z.re, wr, are[c]);
            are[c] = fma(-z.im, wi, are[c]);
            if (!ADJ) {
                aim[c] = fma(z.re, wi, aim[c]);
                aim[c] = fma(z.im, wr, aim[c]);
            }
        } else {
            const ACC wr = wv[c];
            are[c] = fma(z.re, wr, are[c]);
            if (!ADJ) aim[c] = fma(z.im, wr, aim[c]);
        }
    }
}

template <int NCORR, bool WC, bool ADJ, typename ACC, int CH, bool EXACT>
__global__ void __launch_bounds__(kThreads) phasor_stream_kernel(const DftParams p) {
    constexpr int NV = NCORR * (WC ? 2 : 1);  // W scalars per channel
    constexpr int G = 4;                      // channels per 16-byte-aligned W group
    static_assert(CH % G == 0 && kYT % 2 == 0, "tiling");
    using CA = C2<ACC>;

    extern __shared__ __align__(16) unsigned char smem_raw[];

    const int tid = threadIdx.x, lane = tid & 31, warp = tid >> 5;
    const int nck = p.nck;
    const int xgw = (kWarps / nck) * 32;  // owners per CTA
    const int ft = nck * CH;              // channels per CTA
    const int cta_f0 = blockIdx.y * ft;
    const long long cta_x0 = (long long)blockIdx.x * xgw;

    // shared-memory carve-up
    CA *anch = reinterpret_cast<CA *>(smem_raw);              // [kYT][nck][xgw]
    CA *dstp = anch + (size_t)kYT * nck * xgw;                // [kYT][xgw]
    ACC *wt = reinterpret_cast<ACC *>(dstp + (size_t)kYT * xgw);  // [kYT][ft][NV]
    // exact mode re-uses the anchor region: double phi[kYT][xgw], double fq[ft]
    double *phis = reinterpret_cast<double *>(smem_raw);
    double *fq = phis + (size_t)kYT * xgw;

    // consumer role: lane -> owner, warp -> (owner group, channel run)
    const int ck = warp % nck;
    const int x_local = (warp / nck) * 32 + lane;
    const long long x = cta_x0 + x_local;
    const int fo = ck * CH;

    // producer role: fixed owner per thread, y strided
    const int px_local = tid % xgw;
    const int py0 = tid / xgw;
    const int pystep = kThreads / xgw;
    double px0, px1, px2;
    {
        long long pxi = cta_x0 + px_local;
        if (pxi >= p.nx) pxi = p.nx - 1;
        px0 = p.xc[3 * pxi];
        px1 = p.xc[3 * pxi + 1];
        px2 = p.xc[3 * pxi + 2];
    }

    // channel spacing for the recurrence
    double dnu = 0.0, nu0 = 0.0;
    if (!EXACT) {
        if (p.nchan > 1) dnu = (p.freq[p.nchan - 1] - p.freq[0]) / (double)(p.nchan - 1);
        nu0 = p.freq[cta_f0];
    } else {
        for (int i = tid; i < ft; i += kThreads)
            fq[i] = p.freq[min(cta_f0 + i, p.nchan - 1)];
    }

    ACC are[CH][NCORR];
    ACC aim[CH][ADJ ? 1 : NCORR];
#pragma unroll
    for (int j = 0; j < CH; ++j) {
#pragma unroll
        for (int c = 0; c < NCORR; ++c) are[j][c] = ACC(0);
#pragma unroll
        for (int c = 0; c < (ADJ ? 1 : NCORR); ++c) aim[j][c] = ACC(0);
    }

    const long long ys = (long long)blockIdx.z * p.ysplit;
    const long long ye = min(p.ny, ys + p.ysplit);
    const bool f32dot = p.f32dot != 0;

    for (long long y0 = ys; y0 < ye; y0 += kYT) {
        __syncthreads();  // previous tile fully consumed

        // ---- stage the W tile (flag-masked), zero-padded at the edges
        {
            const int per_y = ft * NV;
            const int total = kYT * per_y;
            for (int idx = tid; idx < total; idx += kThreads) {
                const int yl = idx / per_y;
                const int rem = idx - yl * per_y;
                const int fl = rem / NV;
                const int e = rem - fl * NV;
                const long long y = y0 + yl;
                const int f = cta_f0 + fl;
                double val = 0.0;
                if (y < ye && f < p.nchan) {
                    const long long base = (y * p.nchan + f) * p.wstride;
                    bool flagged = false;
                    if (p.flags != nullptr) {
                        for (int k = 0; k < p.wstride; ++k) flagged |= (p.flags[base + k] != 0);
                    }
                    if (!flagged) {
                        if (WC)
                            val = p.w[2 * (base + p.coff + (e >> 1)) + (e & 1)];
                        else
                            val = p.w[base + p.coff + e];
                    }
                }
                wt[idx] = (ACC)val;
            }
        }

        // ---- anchors for this tile's (x, y) pairs
        for (int yl = py0; yl < kYT; yl += pystep) {
            const long long y = y0 + yl;
            if (EXACT) {
                double phi = 0.0;
                if (y < ye) {
                    const double y0c = p.yc[3 * y], y1c = p.yc[3 * y + 1], y2c = p.yc[3 * y + 2];
                    phi = __dmul_rn(p.cst, phase_dot(px0, px1, px2, y0c, y1c, y2c, f32dot));
                }
                phis[yl * xgw + px_local] = phi;
            } else {
                CA zero;
                zero.re = ACC(0);
                zero.im = ACC(0);
                if (y < ye) {
                    const double y0c = p.yc[3 * y], y1c = p.yc[3 * y + 1], y2c = p.yc[3 * y + 2];
                    const double phi =
                        __dmul_rn(p.cst, phase_dot(px0, px1, px2, y0c, y1c, y2c, f32dot));
                    C2<double> a = cis(__dmul_rn(phi, nu0));
                    const C2<double> d = cis(__dmul_rn(phi, dnu));
                    CA dd;
                    dd.re = (ACC)d.re;
                    dd.im = (ACC)d.im;
                    dstp[yl * xgw + px_local] = dd;
                    C2<double> D;
                    D.re = 1.0;
                    D.im = 0.0;
                    if (nck > 1) D = cis(__dmul_rn(phi, (double)CH * dnu));
                    for (int k = 0; k < nck; ++k) {
                        CA aa;
                        aa.re = (ACC)a.re;
                        aa.im = (ACC)a.im;
                        anch[(yl * nck + k) * xgw + px_local] = aa;
                        a = cmul(a, D);
                    }
                } else {
                    dstp[yl * xgw + px_local] = zero;
                    for (int k = 0; k < nck; ++k) anch[(yl * nck + k) * xgw + px_local] = zero;
                }
            }
        }
        __syncthreads();  // tile ready

        // ---- consume: rotate + accumulate
        if (EXACT) {
#pragma unroll 1
            for (int yl = 0; yl < kYT; ++yl) {
                const double phi = phis[yl * xgw + x_local];
                const ACC *wrow = wt + (size_t)(yl * ft + fo) * NV;
#pragma unroll
                for (int j = 0; j < CH; j += G) {
                    ACC wv[G * NV];
                    load_vec<G * NV>(wrow + j * NV, wv);
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        const C2<double> zd = cis_noinline(__dmul_rn(phi, fq[fo + j + g]));
                        CA z;
                        z.re = (ACC)zd.re;
                        z.im = (ACC)zd.im;
                        accumulate<NCORR, WC, ADJ, ACC>(are[j + g], aim[j + g], z, wv + g * NV);
                    }
                }
            }
        } else {
#pragma unroll 1
            for (int yl = 0; yl < kYT; yl += 2) {
                CA za = anch[(yl * nck + ck) * xgw + x_local];
                CA zb = anch[((yl + 1) * nck + ck) * xgw + x_local];
                const CA da = dstp[yl * xgw + x_local];
                const CA db = dstp[(yl + 1) * xgw + x_local];
                const ACC *wa = wt + (size_t)(yl * ft + fo) * NV;
                const ACC *wb = wa + (size_t)ft * NV;
#pragma unroll
                for (int j = 0; j < CH; j += G) {
                    ACC wva[G * NV], wvb[G * NV];
                    load_vec<G * NV>(wa + j * NV, wva);
                    load_vec<G * NV>(wb + j * NV, wvb);
#pragma unroll
                    for (int g = 0; g < G; ++g) {
                        accumulate<NCORR, WC, ADJ, ACC>(are[j + g], aim[j + g], za, wva + g * NV);
                        accumulate<NCORR, WC, ADJ, ACC>(are[j + g], aim[j + g], zb, wvb + g * NV);
                        if (j + g + 1 < CH) {
                            za = cmul(za, da);
                            zb = cmul(zb, db);
                        }
                    }
                }
            }
        }
    }

    // ---- write the owner's channel run
    if (x < p.nx) {
        ACC *o = reinterpret_cast<ACC *>(p.out) + (size_t)blockIdx.z * p.out_split_stride;
#pragma unroll
        for (int j = 0; j < CH; ++j) {
            const int f = cta_f0 + fo + j;
            if (f < p.nchan) {
                const long long base = (x * p.nchan + f) * p.wstride + p.coff;
#pragma unroll
                for (int c = 0; c < NCORR; ++c) {
                    if (ADJ) {
                        o[base + c] = are[j][c];
                    } else {
                        C2<ACC> v;
                        v.re = are[j][c];
                        v.im = aim[j][c];
                        reinterpret_cast<C2<ACC> *>(o)[base + c] = v;
                    }
                }
            }
        }
    }
}

// out[i] = sum_k partial[k][i] in fixed k order (deterministic), correlations
// [coff, coff+ncorr) of every (x, f)
template <typename T>
__global__ void reduce_partials_kernel(const T *partial, T *out, long long n_xf, int wstride,
                                       int coff, int nc, int nsplit, long long split_stride) {
    const long long total = n_xf * nc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long xf = i / nc;
        const int c = (int)(i - xf * nc);
        const long long idx = xf * wstride + coff + c;
        T s = partial[idx];
        for (int k = 1; k < nsplit; ++k) s += partial[(size_t)k * split_stride + idx];
        out[idx] = s;
    }
}

__global__ void lm_to_lmn_kernel(const double *lm, long long nsrc, int mode, int lm_f32,
                                 double *lmn) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= nsrc) return;
    const double l = lm[2 * s], m = lm[2 * s + 1];
    double n;
    if (mode == kLmnPhaseClamp) {
        if (lm_f32) {  // rime/phase.py:23-25,42-43 with lm.dtype == float32
            const float lf = (float)l, mf = (float)m;
            float nf = __fsub_rn(__fsub_rn(1.0f, __fmul_rn(lf, lf)), __fmul_rn(mf, mf));
            nf = __fsub_rn(__fsqrt_rn(nf < 0.0f ? 0.0f : nf), 1.0f);
            n = (double)nf;
        } else {
            n = __dsub_rn(__dsub_rn(1.0, __dmul_rn(l, l)), __dmul_rn(m, m));
            n = __dsub_rn(__dsqrt_rn(n < 0.0 ? 0.0 : n), 1.0);
        }
    } else {  // dft/kernels.py:54 : l**2 in lm.dtype, the rest float64, no clamp
        double l2, m2;
        if (lm_f32) {
            const float lf = (float)l, mf = (float)m;
            l2 = (double)__fmul_rn(lf, lf);
            m2 = (double)__fmul_rn(mf, mf);
        } else {
            l2 = __dmul_rn(l, l);
            m2 = __dmul_rn(m, m);
        }
        n = __dsub_rn(__dsqrt_rn(__dsub_rn(__dsub_rn(1.0, l2), m2)), 1.0);
    }
    lmn[3 * s] = l;
    lmn[3 * s + 1] = m;
    lmn[3 * s + 2] = n;
}

template <int NCORR, bool WC, bool ADJ, typename ACC, int CH>
int launch_one(DftParams p, bool exact, cudaStream_t stream) {
    constexpr int NV = NCORR * (WC ? 2 : 1);
    // channel runs per CTA
    const int runs = (p.nchan + CH - 1) / CH;
    int nck = 1;
    while (nck < runs && nck < kWarps) nck *= 2;
    p.nck = nck;
    const int xgw = (kWarps / nck) * 32;
    const int ft = nck * CH;
    const long long gx = (p.nx + xgw - 1) / xgw;
    const int gy = (p.nchan + ft - 1) / ft;
    AFR_REQUIRE(gx <= 2147483647LL && gy <= 65535, "phasor_stream: grid too large");

    // split the streamed axis when owners alone cannot fill the machine
    const int sms = sm_count();
    long long nsplit = 1;
    const long long ctas = gx * gy;
    if (ctas < 4LL * sms) {
        nsplit = (4LL * sms + ctas - 1) / ctas;
        const long long max_split = (p.ny + 4 * kYT - 1) / (4 * kYT);
        if (nsplit > max_split) nsplit = max_split;
        if (nsplit > 1024) nsplit = 1024;
        if (nsplit < 1) nsplit = 1;
    }
    long long ysplit = (p.ny + nsplit - 1) / nsplit;
    ysplit = ((ysplit + kYT - 1) / kYT) * kYT;
    if (ysplit < kYT) ysplit = kYT;
    nsplit = p.ny > 0 ? (p.ny + ysplit - 1) / ysplit : 1;
    p.ysplit = ysplit;

    const size_t out_scalars = (size_t)p.nx * p.nchan * p.wstride * (ADJ ? 1 : 2);
    Scratch partial;
    void *final_out = p.out;
    if (nsplit > 1) {
        AFR_CUDA_OK(partial.alloc(out_scalars * sizeof(ACC) * nsplit, stream));
        p.out = partial.ptr;
        p.out_split_stride = (long long)out_scalars;
    } else {
        p.out_split_stride = 0;
    }

    size_t smem;
    if (exact)
        smem = (size_t)kYT * xgw * sizeof(double) + (size_t)ft * sizeof(double);
    else
        smem = (size_t)kYT * (nck + 1) * xgw * sizeof(C2<ACC>);
    // the W tile starts after the (non-exact) anchor region in both modes
    const size_t anchor_bytes = (size_t)kYT * (nck + 1) * xgw * sizeof(C2<ACC>);
    if (smem < anchor_bytes) smem = anchor_bytes;
    smem += (size_t)kYT * ft * NV * sizeof(ACC);

    dim3 grid((unsigned)gx, (unsigned)gy, (unsigned)nsplit);
    if (exact) {
        auto kern = phasor_stream_kernel<NCORR, WC, ADJ, ACC, CH, true>;
        AFR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        kern<<<grid, kThreads, smem, stream>>>(p);
    } else {
        auto kern = phasor_stream_kernel<NCORR, WC, ADJ, ACC, CH, false>;
        AFR_CUDA_OK(cudaFuncSetAttribute(kern, cudaFuncAttributeMaxDynamicSharedMemorySize,
                                         (int)smem));
        kern<<<grid, kThreads, smem, stream>>>(p);
    }
    AFR_CUDA_OK(cudaGetLastError());

    if (nsplit > 1) {
        const long long n_xf = p.nx * p.nchan;
        const int sc = ADJ ? 1 : 2;  // scalars per correlation
        const long long total = n_xf * NCORR * sc;
        int blocks = (int)((total + 255) / 256);
        if (blocks > 8 * sms) blocks = 8 * sms;
        if (blocks < 1) blocks = 1;
        reduce_partials_kernel<ACC><<<blocks, 256, 0, stream>>>(
            reinterpret_cast<const ACC *>(partial.ptr), reinterpret_cast<ACC *>(final_out), n_xf,
            p.wstride * sc, p.coff * sc, NCORR * sc, (int)nsplit, (long long)out_scalars);
        AFR_CUDA_OK(cudaGetLastError());
    }
    return 0;
}

template <bool WC, bool ADJ, typename ACC>
int launch_corr_blocks(DftParams p, bool exact, cudaStream_t stream) {
    // correlations are handled in blocks of 4, 2, 1 (ncorr = 3 -> 2 + 1, etc.)
    int c = 0;
    const int ncorr = p.wstride;
    while (c < ncorr) {
        p.coff = c;
        int rc;
        if (ncorr - c >= 4) {
            rc = launch_one<4, WC, ADJ, ACC, ADJ ? 16 : 8>(p, exact, stream);
            c += 4;
        } else if (ncorr - c >= 2) {
            rc = launch_one<2, WC, ADJ, ACC, ADJ ? 32 : 16>(p, exact, stream);
            c += 2;
        } else {
            rc = launch_one<1, WC, ADJ, ACC, 32>(p, exact, stream);
            c += 1;
        }
        if (rc) return rc;
    }
    return 0;
}

}  // namespace

int launch_lm_to_lmn(const double *lm, int64_t nsrc, int mode, bool lm_f32, double *lmn,
                     cudaStream_t stream) {
    if (nsrc <= 0) return 0;
    const int blocks = (int)((nsrc + 255) / 256);
    lm_to_lmn_kernel<<<blocks, 256, 0, stream>>>(lm, nsrc, mode, lm_f32 ? 1 : 0, lmn);
    AFR_CUDA_OK(cudaGetLastError());
    return 0;
}

int run_phasor_stream(const double *xc, int64_t nx, const double *yc, int64_t ny,
                      const double *w, bool w_complex, const uint8_t *flags, const double *freq,
                      int64_t nchan, int64_t ncorr, double cst, bool f32dot, bool adjoint,
                      bool exact, bool acc32, void *out, cudaStream_t stream) {
    AFR_REQUIRE(nchan <= 2147483647LL / 64 && ncorr <= 64, "phasor_stream: nchan/ncorr too large");
    const size_t out_bytes = (size_t)nx * nchan * ncorr * (adjoint ? 1 : 2) * (acc32 ? 4 : 8);
    if (nx <= 0 || nchan <= 0 || ncorr <= 0) return 0;
    if (ny <= 0) {  // empty sum: the reference returns zeros
        AFR_CUDA_OK(cudaMemsetAsync(out, 0, out_bytes, stream));
        return 0;
    }
    DftParams p{};
    p.xc = xc;
    p.yc = yc;
    p.w = w;
    p.flags = flags;
    p.freq = freq;
    p.out = out;
    p.cst = cst;
    p.nx = nx;
    p.ny = ny;
    p.nchan = (int)nchan;
    p.wstride = (int)ncorr;
    p.f32dot = f32dot ? 1 : 0;
    if (!acc32) {
        if (w_complex)
            return adjoint ? launch_corr_blocks<true, true, double>(p, exact, stream)
                           : launch_corr_blocks<true, false, double>(p, exact, stream);
        return adjoint ? launch_corr_blocks<false, true, double>(p, exact, stream)
                       : launch_corr_blocks<false, false, double>(p, exact, stream);
    }
    if (w_complex)
        return adjoint ? launch_corr_blocks<true, true, float>(p, exact, stream)
                       : launch_corr_blocks<true, false, float>(p, exact, stream);
    return adjoint ? launch_corr_blocks<false, true, float>(p, exact, stream)
                   : launch_corr_blocks<false, false, float>(p, exact, stream);
}

}  // namespace afr

// ---------------------------------------------------------------------------
// C ABI
// ---------------------------------------------------------------------------
using namespace afr;

extern "C" int afr_im_to_vis(const void *image, int image_complex, const double *uvw,
                             const double *lm, const double *freq, int64_t nsrc, int64_t nrow,
                             int64_t nchan, int64_t ncorr, int convention, int f32_flags,
                             int chan_mode, int out_c64, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(convention == AFR_FOURIER || convention == AFR_CASA,
                "convention not in ('fourier', 'casa')");
    AFR_REQUIRE(nsrc >= 0 && nrow >= 0 && nchan >= 0 && ncorr >= 0, "negative extent");
    // dft/kernels.py:34-39
    const double cst = convention == AFR_FOURIER ? -kTwoPiOverC : kTwoPiOverC;
    Scratch lmn;
    AFR_CUDA_OK(lmn.alloc(sizeof(double) * 3 * (size_t)nsrc, stream));
    int rc = launch_lm_to_lmn(lm, nsrc, kLmnDft, (f32_flags & AFR_F32_LM) != 0,
                              (double *)lmn.ptr, stream);
    if (rc) return rc;
    const bool f32dot = (f32_flags & AFR_F32_LM) && (f32_flags & AFR_F32_UVW);
    return run_phasor_stream(uvw, nrow, (const double *)lmn.ptr, nsrc, (const double *)image,
                             image_complex != 0, nullptr, freq, nchan, ncorr, cst, f32dot,
                             /*adjoint=*/false, chan_mode == AFR_CHAN_EXACT, out_c64 != 0, out,
                             stream);
}

extern "C" int afr_vis_to_im(const void *vis, int vis_complex, const double *uvw,
                             const double *lm, const double *freq, const uint8_t *flags,
                             int64_t nsrc, int64_t nrow, int64_t nchan, int64_t ncorr,
                             int convention, int f32_flags, int chan_mode, int out_f32, void *out,
                             void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(convention == AFR_FOURIER || convention == AFR_CASA,
                "convention not in ('fourier', 'casa')");
    AFR_REQUIRE(nsrc >= 0 && nrow >= 0 && nchan >= 0 && ncorr >= 0, "negative extent");
    // dft/kernels.py:110-115 : opposite sign to im_to_vis
    const double cst = convention == AFR_FOURIER ? kTwoPiOverC : -kTwoPiOverC;
    Scratch lmn;
    AFR_CUDA_OK(lmn.alloc(sizeof(double) * 3 * (size_t)nsrc, stream));
    int rc = launch_lm_to_lmn(lm, nsrc, kLmnDft, (f32_flags & AFR_F32_LM) != 0,
                              (double *)lmn.ptr, stream);
    if (rc) return rc;
    const bool f32dot = (f32_flags & AFR_F32_LM) && (f32_flags & AFR_F32_UVW);
    return run_phasor_stream((const double *)lmn.ptr, nsrc, uvw, nrow, (const double *)vis,
                             vis_complex != 0, flags, freq, nchan, ncorr, cst, f32dot,
                             /*adjoint=*/true, chan_mode == AFR_CHAN_EXACT, out_f32 != 0, out,
                             stream);
}
