// predict_vis / apply_gains (africanus/rime/predict.py:466-649) and the fused
// phase_delay (x) brightness -> predict_vis composition
// (africanus/rime/examples/predict.py:107-134,490,522-527) for sm_100a.
//
//   V[r,f] = G1[t,a1,f] (B[r,f] + sum_s E1[s,t,a1,f] X[s,r,f] E2[s,t,a2,f]^H) G2[t,a2,f]^H
//
// * predict_vis with a materialised source_coh is bound by the HBM read of X
//   (16*ncorr B/term): one thread owns one (row, chan) output, lanes run along
//   chan so every load/store is contiguous across the warp, the 2x2 chain stays in
//   registers, the source loop is sequential (the reference's accumulation order).
// * the fused kernels never materialise K or X:
//     - without DDEs the source sum is the phasor-stream kernel of afr_dft.cu with the
//       brightness as a complex "image" (CH-channel runs per thread, anchored rotation);
//     - with DDEs a warp owns (row, 32*kK channels): lane L handles channels
//       f0+L, f0+L+32, ... so DDE gathers stay coalesced along chan, and the phasor
//       advances by exp(i*phi*32*dnu) between a lane's channels.
//   base_vis and the DIE product are applied by the same epilogue as predict_vis.
#include "afr_dft.cuh"

namespace afr {
namespace {

template <typename T>
struct Cx {
    T re, im;
};

template <typename T>
__device__ __forceinline__ Cx<T> mul(Cx<T> a, Cx<T> b) {
    return {a.re * b.re - a.im * b.im, a.re * b.im + a.im * b.re};
}
template <typename T>
__device__ __forceinline__ Cx<T> mulc(Cx<T> a, Cx<T> b) {  // a * conj(b)
    return {a.re * b.re + a.im * b.im, a.im * b.re - a.re * b.im};
}
template <typename T>
__device__ __forceinline__ Cx<T> add(Cx<T> a, Cx<T> b) {
    return {a.re + b.re, a.im + b.im};
}

template <typename T, int N>
__device__ __forceinline__ void load_n(const Cx<T> *p, Cx<T> (&v)[N]) {
    // 16-byte vector loads where the element size allows
    if (sizeof(Cx<T>) == 16) {
        const double2 *q = reinterpret_cast<const double2 *>(p);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            double2 d = q[i];
            v[i].re = (T)d.x;
            v[i].im = (T)d.y;
        }
    } else if (N % 2 == 0) {
        const float4 *q = reinterpret_cast<const float4 *>(p);
#pragma unroll
        for (int i = 0; i < N / 2; ++i) {
            float4 d = q[i];
            v[2 * i].re = (T)d.x;
            v[2 * i].im = (T)d.y;
            v[2 * i + 1].re = (T)d.z;
            v[2 * i + 1].im = (T)d.w;
        }
    } else {
        const float2 *q = reinterpret_cast<const float2 *>(p);
#pragma unroll
        for (int i = 0; i < N; ++i) {
            float2 d = q[i];
            v[i].re = (T)d.x;
            v[i].im = (T)d.y;
        }
    }
}

template <typename T, int N>
__device__ __forceinline__ void store_n(Cx<T> *p, const Cx<T> (&v)[N]) {
#pragma unroll
    for (int i = 0; i < N; ++i) p[i] = v[i];
}

// out (+)= A * (X * C^H), association of predict.py:103-117
template <typename T>
__device__ __forceinline__ void jones3_2x2(const Cx<T> (&a)[4], const Cx<T> (&x)[4],
                                           const Cx<T> (&c)[4], Cx<T> (&o)[4], bool accumulate) {
    const Cx<T> xx = add(mulc(x[0], c[0]), mulc(x[1], c[1]));
    const Cx<T> xy = add(mulc(x[0], c[2]), mulc(x[1], c[3]));
    const Cx<T> yx = add(mulc(x[2], c[0]), mulc(x[3], c[1]));
    const Cx<T> yy = add(mulc(x[2], c[2]), mulc(x[3], c[3]));
    const Cx<T> r0 = add(mul(a[0], xx), mul(a[1], yx));
    const Cx<T> r1 = add(mul(a[0], xy), mul(a[1], yy));
    const Cx<T> r2 = add(mul(a[2], xx), mul(a[3], yx));
    const Cx<T> r3 = add(mul(a[2], xy), mul(a[3], yy));
    if (accumulate) {
        o[0] = add(o[0], r0);
        o[1] = add(o[1], r1);
        o[2] = add(o[2], r2);
        o[3] = add(o[3], r3);
    } else {
        o[0] = r0;
        o[1] = r1;
        o[2] = r2;
        o[3] = r3;
    }
}

// out += A * C^H  (predict.py:138-148)
template <typename T>
__device__ __forceinline__ void jones2_2x2(const Cx<T> (&a)[4], const Cx<T> (&c)[4],
                                           Cx<T> (&o)[4]) {
    o[0] = add(o[0], add(mulc(a[0], c[0]), mulc(a[1], c[1])));
    o[1] = add(o[1], add(mulc(a[0], c[2]), mulc(a[1], c[3])));
    o[2] = add(o[2], add(mulc(a[2], c[0]), mulc(a[3], c[1])));
    o[3] = add(o[3], add(mulc(a[2], c[2]), mulc(a[3], c[3])));
}

struct PredictParams {
    const int32_t *time_index, *ant1, *ant2;
    const void *dde1, *coh, *dde2, *die1, *bvis, *die2;
    const void *acc_init;  // optional (row,chan,C) pre-summed coherencies (fused path)
    void *out;
    long long nsrc, nrow, ntime, nant;
    long long nfc;  // chan (2x2 mode) or chan*ncorr (diagonal mode, element-wise)
};

// N = 4: (2,2) matrices; N = 1: element-wise ("diagonal" (1,) / (2,) Jones, flattened
// over chan*corr because every correlation is independent, predict.py:93-98)
template <typename T, int N>
__global__ void __launch_bounds__(256) predict_vis_kernel(const PredictParams p) {
    using C = Cx<T>;
    const C *dde1 = (const C *)p.dde1, *dde2 = (const C *)p.dde2, *coh = (const C *)p.coh;
    const C *die1 = (const C *)p.die1, *die2 = (const C *)p.die2, *bvis = (const C *)p.bvis;
    const C *init = (const C *)p.acc_init;
    C *out = (C *)p.out;
    const bool have_dde = dde1 != nullptr, have_coh = coh != nullptr;
    const long long total = p.nrow * p.nfc;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const long long r = i / p.nfc, f = i - r * p.nfc;
        const long long ti = p.time_index[r], a1 = p.ant1[r], a2 = p.ant2[r];
        C acc[N];
#pragma unroll
        for (int k = 0; k < N; ++k) acc[k] = {T(0), T(0)};
        if (init) load_n<T, N>(init + i * N, acc);
        if (have_dde || have_coh) {
            for (long long s = 0; s < p.nsrc; ++s) {
                C x[N], e1[N], e2[N];
                if (have_coh) load_n<T, N>(coh + ((s * p.nrow + r) * p.nfc + f) * N, x);
                if (have_dde) {
                    load_n<T, N>(dde1 + (((s * p.ntime + ti) * p.nant + a1) * p.nfc + f) * N, e1);
                    load_n<T, N>(dde2 + (((s * p.ntime + ti) * p.nant + a2) * p.nfc + f) * N, e2);
                }
                if constexpr (N == 4) {
                    if (have_dde && have_coh) {
                        jones3_2x2<T>((const C(&)[4])e1, (const C(&)[4])x, (const C(&)[4])e2,
                                      (C(&)[4])acc, true);
                    } else if (have_dde) {
                        jones2_2x2<T>((const C(&)[4])e1, (const C(&)[4])e2, (C(&)[4])acc);
                    } else {
#pragma unroll
                        for (int k = 0; k < N; ++k) acc[k] = add(acc[k], x[k]);
                    }
                } else {
                    if (have_dde && have_coh)
                        acc[0] = add(acc[0], mulc(mul(e1[0], x[0]), e2[0]));
                    else if (have_dde)
                        acc[0] = add(acc[0], mulc(e1[0], e2[0]));
                    else
                        acc[0] = add(acc[0], x[0]);
                }
            }
        }
        if (bvis) {
            C b[N];
            load_n<T, N>(bvis + i * N, b);
#pragma unroll
            for (int k = 0; k < N; ++k) acc[k] = add(acc[k], b[k]);
        }
        if (die1) {
            C g1[N], g2[N];
            load_n<T, N>(die1 + ((ti * p.nant + a1) * p.nfc + f) * N, g1);
            load_n<T, N>(die2 + ((ti * p.nant + a2) * p.nfc + f) * N, g2);
            if constexpr (N == 4) {
                C tmp[4];
#pragma unroll
                for (int k = 0; k < 4; ++k) tmp[k] = acc[k];
                jones3_2x2<T>((const C(&)[4])g1, (const C(&)[4])tmp, (const C(&)[4])g2,
                              (C(&)[4])acc, false);
            } else {
                acc[0] = mulc(mul(g1[0], acc[0]), g2[0]);
            }
        }
        store_n<T, N>(out + i * N, acc);
    }
}

// ---------------------------------------------------------------------------
// fused predict with DDEs
// ---------------------------------------------------------------------------
constexpr int kK = 8;  // channels per lane (stride 32)

struct FusedParams {
    const double *lmn, *uvw, *freq;
    const void *bright;  // (nsrc,nchan,C)
    const int32_t *time_index, *ant1, *ant2;
    const void *dde1, *dde2;
    void *out;  // (nrow,nchan,C) source sum only; epilogue applied by predict_vis_kernel
    double cst;
    long long nsrc, nrow, ntime, nant;
    int nchan;
    int ncorr;  // C
};

// N = 4 with MAT: 2x2 products; otherwise N element-wise correlations
template <typename T, int N, bool MAT, bool EXACT>
__global__ void __launch_bounds__(128) fused_dde_kernel(const FusedParams p) {
    using C = Cx<T>;
    const int lane = threadIdx.x & 31;
    const int segs = (p.nchan + 32 * kK - 1) / (32 * kK);
    const long long warp_global = (blockIdx.x * (long long)blockDim.x + threadIdx.x) >> 5;
    const long long nwarps = ((long long)gridDim.x * blockDim.x) >> 5;
    const C *dde1 = (const C *)p.dde1, *dde2 = (const C *)p.dde2, *bright = (const C *)p.bright;
    C *out = (C *)p.out;
    double dnu = 0.0;
    if (!EXACT && p.nchan > 1) dnu = (p.freq[p.nchan - 1] - p.freq[0]) / (double)(p.nchan - 1);

    for (long long wi = warp_global; wi < p.nrow * segs; wi += nwarps) {
        // consecutive warps take consecutive rows of the same channel segment so that
        // concurrently running warps share DDE (time, antenna) slices in L2
        const long long seg = wi / p.nrow, r = wi - seg * p.nrow;
        const int fbase = (int)seg * 32 * kK + lane;
        const long long ti = p.time_index[r], a1 = p.ant1[r], a2 = p.ant2[r];
        const double u = p.uvw[3 * r], v = p.uvw[3 * r + 1], w = p.uvw[3 * r + 2];
        C acc[kK][N];
#pragma unroll
        for (int k = 0; k < kK; ++k)
#pragma unroll
            for (int c = 0; c < N; ++c) acc[k][c] = {T(0), T(0)};

        for (long long s = 0; s < p.nsrc; ++s) {
            const double phi = __dmul_rn(
                p.cst, phase_dot(p.lmn[3 * s], p.lmn[3 * s + 1], p.lmn[3 * s + 2], u, v, w, false));
            C2<double> z = {1.0, 0.0}, d = {1.0, 0.0};
            if (!EXACT) {
                z = cis_fast(__dmul_rn(phi, p.freq[min(fbase, p.nchan - 1)]));
                d = cis_fast(__dmul_rn(phi, 32.0 * dnu));
            }
            const C *e1p = dde1 + (((s * p.ntime + ti) * p.nant + a1) * p.nchan) * N;
            const C *e2p = dde2 + (((s * p.ntime + ti) * p.nant + a2) * p.nchan) * N;
            const C *bp = bright + (s * p.nchan) * N;
#pragma unroll
            for (int k = 0; k < kK; ++k) {
                const int f = fbase + 32 * k;
                if (f < p.nchan) {
                    if (EXACT) z = cis_fast(__dmul_rn(phi, p.freq[f]));
                    const C zz = {(T)z.re, (T)z.im};
                    C b[N], e1[N], e2[N], x[N];
                    load_n<T, N>(bp + (long long)f * N, b);
                    load_n<T, N>(e1p + (long long)f * N, e1);
                    load_n<T, N>(e2p + (long long)f * N, e2);
#pragma unroll
                    for (int c = 0; c < N; ++c) x[c] = mul(zz, b[c]);  // K * brightness
                    if constexpr (MAT) {
                        jones3_2x2<T>((const C(&)[4])e1, (const C(&)[4])x, (const C(&)[4])e2,
                                      (C(&)[4])acc[k], true);
                    } else {
#pragma unroll
                        for (int c = 0; c < N; ++c)
                            acc[k][c] = add(acc[k][c], mulc(mul(e1[c], x[c]), e2[c]));
                    }
                }
                if (!EXACT) z = cmul(z, d);
            }
        }
#pragma unroll
        for (int k = 0; k < kK; ++k) {
            const int f = fbase + 32 * k;
            if (f < p.nchan) store_n<T, N>(out + (r * p.nchan + f) * N, acc[k]);
        }
    }
}

int grid_for(long long total, int threads) {
    long long blocks = (total + threads - 1) / threads;
    const long long cap = (long long)sm_count() * 16;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T>
int launch_predict(const PredictParams &p, int jones_mode, cudaStream_t stream) {
    const long long total = p.nrow * p.nfc;
    if (total <= 0) return 0;
    if (jones_mode == AFR_JONES_2X2)
        predict_vis_kernel<T, 4><<<grid_for(total, 256), 256, 0, stream>>>(p);
    else
        predict_vis_kernel<T, 1><<<grid_for(total, 256), 256, 0, stream>>>(p);
    AFR_LAUNCH_OK();
    return 0;
}

template <typename T>
int launch_fused_dde(const FusedParams &p, int jones_mode, bool exact, cudaStream_t stream) {
    const int segs = (p.nchan + 32 * kK - 1) / (32 * kK);
    const long long warps = p.nrow * segs;
    if (warps <= 0) return 0;
    const int grid = grid_for(warps * 32, 128);
#define AFR_LAUNCH_FUSED(N, MAT)                                                        \
    do {                                                                                \
        if (exact)                                                                      \
            fused_dde_kernel<T, N, MAT, true><<<grid, 128, 0, stream>>>(p);             \
        else                                                                            \
            fused_dde_kernel<T, N, MAT, false><<<grid, 128, 0, stream>>>(p);            \
    } while (0)
    if (jones_mode == AFR_JONES_2X2) {
        AFR_LAUNCH_FUSED(4, true);
    } else if (p.ncorr == 1) {
        AFR_LAUNCH_FUSED(1, false);
    } else if (p.ncorr == 2) {
        AFR_LAUNCH_FUSED(2, false);
    } else if (p.ncorr == 4) {
        AFR_LAUNCH_FUSED(4, false);
    } else {
        return fail("afr_predict_fused: diagonal Jones with DDEs supports ncorr in (1, 2, 4)");
    }
#undef AFR_LAUNCH_FUSED
    AFR_LAUNCH_OK();
    return 0;
}

}  // namespace
}  // namespace afr

using namespace afr;

extern "C" int afr_predict_vis(const int32_t *time_index, const int32_t *antenna1,
                               const int32_t *antenna2, const void *dde1, const void *source_coh,
                               const void *dde2, const void *die1, const void *base_vis,
                               const void *die2, int64_t nsrc, int64_t nrow, int64_t ntime,
                               int64_t nant, int64_t nchan, int64_t ncorr, int jones_mode,
                               int is_c64, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    // africanus/rime/predict.py:403-407
    AFR_REQUIRE((dde1 == nullptr) == (dde2 == nullptr),
                "Both dde1_jones and dde2_jones must be present or absent");
    AFR_REQUIRE((die1 == nullptr) == (die2 == nullptr),
                "Both die1_jones and die2_jones must be present or absent");
    AFR_REQUIRE(dde1 || source_coh || die1 || base_vis, "No Jones Matrices were supplied");
    AFR_REQUIRE(jones_mode == AFR_JONES_DIAG || (jones_mode == AFR_JONES_2X2 && ncorr == 4),
                "Jones Matrix Correlations were mismatched");
    AFR_REQUIRE(nsrc >= 0 && nrow >= 0 && nchan >= 0 && ncorr >= 1, "bad extent");
    PredictParams p{};
    p.time_index = time_index;
    p.ant1 = antenna1;
    p.ant2 = antenna2;
    p.dde1 = dde1;
    p.coh = source_coh;
    p.dde2 = dde2;
    p.die1 = die1;
    p.bvis = base_vis;
    p.die2 = die2;
    p.acc_init = nullptr;
    p.out = out;
    p.nsrc = nsrc;
    p.nrow = nrow;
    p.ntime = ntime;
    p.nant = nant;
    p.nfc = jones_mode == AFR_JONES_2X2 ? nchan : nchan * ncorr;
    return is_c64 ? launch_predict<float>(p, jones_mode, stream)
                  : launch_predict<double>(p, jones_mode, stream);
}

extern "C" int afr_predict_fused(const double *lm, const double *uvw, const double *freq,
                                 const void *brightness, const int32_t *time_index,
                                 const int32_t *antenna1, const int32_t *antenna2,
                                 const void *dde1, const void *dde2, const void *die1,
                                 const void *base_vis, const void *die2, int64_t nsrc,
                                 int64_t nrow, int64_t ntime, int64_t nant, int64_t nchan,
                                 int64_t ncorr, int jones_mode, int convention, int chan_mode,
                                 int out_c64, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(convention == AFR_FOURIER || convention == AFR_CASA,
                "convention not in ('fourier', 'casa')");
    AFR_REQUIRE((dde1 == nullptr) == (dde2 == nullptr),
                "Both dde1_jones and dde2_jones must be present or absent");
    AFR_REQUIRE((die1 == nullptr) == (die2 == nullptr),
                "Both die1_jones and die2_jones must be present or absent");
    AFR_REQUIRE(jones_mode == AFR_JONES_DIAG || (jones_mode == AFR_JONES_2X2 && ncorr == 4),
                "Jones Matrix Correlations were mismatched");
    AFR_REQUIRE(nsrc >= 0 && nrow >= 0 && nchan >= 0 && ncorr >= 1 && nchan < (1LL << 30),
                "bad extent");
    if (nrow == 0 || nchan == 0) return 0;
    const bool exact = chan_mode == AFR_CHAN_EXACT;
    // rime/phase.py:29-34
    const double cst = convention == AFR_FOURIER ? -kTwoPiOverC : kTwoPiOverC;

    Scratch lmn;
    AFR_CUDA_OK(lmn.alloc(sizeof(double) * 3 * (size_t)nsrc, stream));
    int rc = launch_lm_to_lmn(lm, nsrc, kLmnPhaseClamp, false, (double *)lmn.ptr, stream);
    if (rc) return rc;

    const size_t elem = out_c64 ? 8 : 16;
    if (nsrc == 0) {
        AFR_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)nrow * nchan * ncorr * elem, stream));
    } else if (dde1 == nullptr) {
        // point-source sum: phasor-stream kernel with the brightness as complex "image"
        // (complex64 brightness feeds the FP32 accumulator variant directly)
        rc = run_phasor_stream(uvw, nrow, (const double *)lmn.ptr, nsrc, brightness, true, nullptr,
                               freq, nchan, ncorr, cst, false, /*adjoint=*/false, exact,
                               out_c64 != 0, out, stream);
        if (rc) return rc;
    } else {
        FusedParams f{};
        f.lmn = (const double *)lmn.ptr;
        f.uvw = uvw;
        f.freq = freq;
        f.bright = brightness;
        f.time_index = time_index;
        f.ant1 = antenna1;
        f.ant2 = antenna2;
        f.dde1 = dde1;
        f.dde2 = dde2;
        f.out = out;
        f.cst = cst;
        f.nsrc = nsrc;
        f.nrow = nrow;
        f.ntime = ntime;
        f.nant = nant;
        f.nchan = (int)nchan;
        f.ncorr = (int)ncorr;
        rc = out_c64 ? launch_fused_dde<float>(f, jones_mode, exact, stream)
                     : launch_fused_dde<double>(f, jones_mode, exact, stream);
        if (rc) return rc;
    }

    if (base_vis != nullptr || die1 != nullptr) {
        // epilogue: out = G1 (out + B) G2^H  (predict.py:329-373), in place
        PredictParams p{};
        p.time_index = time_index;
        p.ant1 = antenna1;
        p.ant2 = antenna2;
        p.die1 = die1;
        p.bvis = base_vis;
        p.die2 = die2;
        p.acc_init = out;
        p.out = out;
        p.nsrc = 0;
        p.nrow = nrow;
        p.ntime = ntime;
        p.nant = nant;
        p.nfc = jones_mode == AFR_JONES_2X2 ? nchan : nchan * ncorr;
        rc = out_c64 ? launch_predict<float>(p, jones_mode, stream)
                     : launch_predict<double>(p, jones_mode, stream);
        if (rc) return rc;
    }
    return 0;
}
