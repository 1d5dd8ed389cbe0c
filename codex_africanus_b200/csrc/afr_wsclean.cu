// WSClean component-list predict (africanus/rime/wsclean_predict.py:11-116) and the WSClean
// spectral model (africanus/model/wsclean/spec_model.py:76-124) for sm_100a.
//
//   vis[r,f] = sum_s spectrum[s,f] * exp(+i 2pi/c (u l + v m + w n) nu_f) * shape[s,r,f]
//   shape = 1 for POINT sources, exp(-(fu1^2 + fv1^2)) for GAUSSIAN ones (:48-76)
//
// POINT sources are exactly an im_to_vis with the "casa" sign, the spectrum as a real
// single-correlation image and no n clamp, so they run on the phasor-stream kernel of
// afr_dft.cu (FP64-pipe bound, 11 flop/term) with the GAUSSIAN rows of the image zeroed
// (zero pixels contribute nothing, dft/kernels.py:64).  GAUSSIAN sources carry a real taper
// that goes with nu^2: a second kernel (warp <-> row, lane <-> channel, sources in index
// order as the reference) adds them onto the same output, advancing phasor and taper by
// recurrences along equispaced channels.
#include "afr_dft.cuh"

namespace afr {
namespace {

// x ** n in numba's int_power order (binary square-and-multiply from the low bit; it differs
// from a left-to-right product from the 4th power on)
__device__ __forceinline__ double ipow(double x, int n) {
    double r = 1.0, a = x;
    while (n != 0) {
        if (n & 1) r = __dmul_rn(r, a);
        n >>= 1;
        a = __dmul_rn(a, a);
    }
    return r;
}

// spec_model.py:94-124; point_only (optional): the same with GAUSSIAN rows zeroed
__global__ void wsclean_spectra_kernel(const double *flux, const double *coeffs,
                                       const uint8_t *log_poly, const double *ref_freq,
                                       const double *freq, const uint8_t *is_gauss, long long nsrc,
                                       int ncoeffs, long long nchan, double *out, double *point_only) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= nsrc * nchan) return;
    const long long s = i / nchan, f = i - s * nchan;
    const double nu = freq[f], rf = ref_freq[s];
    double acc;
    if (log_poly[s]) {
        const double lg = log(nu / rf);
        acc = 0.0;
        for (int c = 0; c < ncoeffs; ++c)
            acc = __dadd_rn(acc, __dmul_rn(coeffs[s * ncoeffs + c], ipow(lg, c + 1)));
        acc = __dmul_rn(flux[s], exp(acc));
    } else {
        const double x = __dsub_rn(nu / rf, 1.0);
        acc = flux[s];
        for (int c = 0; c < ncoeffs; ++c)
            acc = __dadd_rn(acc, __dmul_rn(coeffs[s * ncoeffs + c], ipow(x, c + 1)));
    }
    if (out) out[i] = acc;
    if (point_only) point_only[i] = (is_gauss && is_gauss[s]) ? 0.0 : acc;
}

// per source: l, m, n, el, em, er  (wsclean_predict.py:29-31,48-52)
__global__ void gauss_params_kernel(const double *lm, const double *gauss_shape, long long nsrc,
                                    double *prm) {
    const long long s = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (s >= nsrc) return;
    const double l = lm[2 * s], m = lm[2 * s + 1];
    const double n = __dsub_rn(__dsqrt_rn(__dsub_rn(__dsub_rn(1.0, __dmul_rn(l, l)), __dmul_rn(m, m))), 1.0);
    const double emaj = gauss_shape[3 * s], emin = gauss_shape[3 * s + 1], angle = gauss_shape[3 * s + 2];
    double sn, cs;
    sincos(angle, &sn, &cs);
    prm[6 * s] = l;
    prm[6 * s + 1] = m;
    prm[6 * s + 2] = n;
    prm[6 * s + 3] = __dmul_rn(emaj, sn);
    prm[6 * s + 4] = __dmul_rn(emaj, cs);
    prm[6 * s + 5] = emin / (emaj == 0.0 ? 1.0 : emaj);
}

constexpr int kGK = 8;  // channels per lane

// out[r,f] += sum over GAUSSIAN sources; warp <-> (row, block of 32*kGK channels), lane L owns
// channels f0+L, f0+L+32, ...  (coalesced loads of the spectrum and of the output).
//
// UNIFORM (equispaced channels): along a lane's channels (stride 32) the phasor advances by
// the three-term recurrence z_{j+1} = 2 Re(d) z_j - z_{j-1}, d = exp(i phi 32 dnu), and the
// taper exp(-g sigma_j^2), sigma_j = sigma_0 + j dsig, by the second-order product recurrence
// s_{j+1} = s_j r_j, r_{j+1} = r_j q with r_0 = exp(-g (2 sigma_0 dsig + dsig^2)),
// q = exp(-2 g dsig^2): 2 sincos + 3 exp per (row, source, lane) instead of one of each per
// term (25 instead of 65 FP64 instructions per term at 8 channels per lane; relative error of
// the products <= 8 * 3 eps).  The recurrence is used only while g sigma^2 < 600 over the
// lane's channels (no factor can underflow or overflow); otherwise, and for non-equispaced
// channels, every term takes its own sincos and exp.
template <bool UNIFORM>
__global__ void __launch_bounds__(256) wsclean_gauss_kernel(const double *uvw, const double *prm,
                                                            const uint8_t *is_gauss,
                                                            const double *spectrum, const double *freq,
                                                            long long nsrc, long long nrow,
                                                            long long nchan, double gauss_scale,
                                                            double *out) {
    const int lane = threadIdx.x & 31;
    const long long segs = (nchan + 32 * kGK - 1) / (32 * kGK);
    const long long nwork = nrow * segs;
    double dnu32 = 0.0;  // frequency step between a lane's consecutive channels
    if (UNIFORM && nchan > 1) dnu32 = 32.0 * ((freq[nchan - 1] - freq[0]) / (double)(nchan - 1));
    const double dsig = __dmul_rn(dnu32, gauss_scale);
    for (long long wk = blockIdx.x * (long long)(blockDim.x >> 5) + (threadIdx.x >> 5); wk < nwork;
         wk += (long long)gridDim.x * (blockDim.x >> 5)) {
        const long long r = wk / segs, seg = wk - r * segs;
        const double u = uvw[3 * r], v = uvw[3 * r + 1], w = uvw[3 * r + 2];
        const long long fl0 = seg * 32 * kGK + lane;  // this lane's first channel
        double are[kGK], aim[kGK];
#pragma unroll
        for (int j = 0; j < kGK; ++j) are[j] = aim[j] = 0.0;
        const double nu_first = fl0 < nchan ? freq[fl0] : 0.0;
        const double sig0 = __dmul_rn(nu_first, gauss_scale);
        // largest sigma^2 among this lane's channels (channels may run either way)
        const double sig_last = sig0 + (kGK - 1) * dsig;
        const double smax2 = fmax(sig0 * sig0, sig_last * sig_last);
        for (long long s = 0; s < nsrc; ++s) {
            if (!is_gauss[s]) continue;  // warp-uniform
            const double *q = prm + 6 * s;
            const double l = q[0], m = q[1], n = q[2], el = q[3], em = q[4], er = q[5];
            // real_phase = two_pi_over_c * (u*l + v*m + w*n)
            const double real_phase =
                __dmul_rn(kTwoPiOverC, __dadd_rn(__dadd_rn(__dmul_rn(u, l), __dmul_rn(v, m)), __dmul_rn(w, n)));
            const double u1 = __dmul_rn(__dsub_rn(__dmul_rn(u, em), __dmul_rn(v, el)), er);
            const double v1 = __dadd_rn(__dmul_rn(u, el), __dmul_rn(v, em));
            const double *sp_row = spectrum + s * nchan;
            const double g = fma(u1, u1, v1 * v1);
            if (UNIFORM && g * smax2 < 600.0) {
                C2<double> z = cis_fast(__dmul_rn(real_phase, nu_first));
                const C2<double> d = cis_fast(__dmul_rn(real_phase, dnu32));
                C2<double> zp = z;
                const double c2 = d.re + d.re;
                double sh = exp(-g * sig0 * sig0);
                double rr = exp(-g * (2.0 * sig0 * dsig + dsig * dsig));
                const double qq = exp(-2.0 * g * dsig * dsig);
#pragma unroll
                for (int j = 0; j < kGK; ++j) {
                    const long long f = fl0 + 32 * j;
                    if (f < nchan) {
                        const double amp = __dmul_rn(sp_row[f], sh);
                        are[j] = fma(z.re, amp, are[j]);
                        aim[j] = fma(z.im, amp, aim[j]);
                    }
                    C2<double> zn;
                    if (j == 0) {
                        zn = cmul(z, d);
                    } else {
                        zn.re = fma(c2, z.re, -zp.re);
                        zn.im = fma(c2, z.im, -zp.im);
                    }
                    zp = z;
                    z = zn;
                    sh *= rr;
                    rr *= qq;
                }
            } else {
#pragma unroll
                for (int j = 0; j < kGK; ++j) {
                    const long long f = fl0 + 32 * j;
                    if (f < nchan) {
                        const double nu = freq[f];
                        const double sf = __dmul_rn(nu, gauss_scale);
                        const C2<double> z = cis_fast(__dmul_rn(real_phase, nu));
                        const double fu1 = __dmul_rn(u1, sf), fv1 = __dmul_rn(v1, sf);
                        const double shape = exp(-__dadd_rn(__dmul_rn(fu1, fu1), __dmul_rn(fv1, fv1)));
                        const double sp = sp_row[f];
                        are[j] = __dadd_rn(are[j], __dmul_rn(__dmul_rn(z.re, sp), shape));
                        aim[j] = __dadd_rn(aim[j], __dmul_rn(__dmul_rn(z.im, sp), shape));
                    }
                }
            }
        }
#pragma unroll
        for (int j = 0; j < kGK; ++j) {
            const long long f = fl0 + 32 * j;
            if (f < nchan) {
                double2 *o = reinterpret_cast<double2 *>(out) + r * nchan + f;
                double2 cur = *o;
                cur.x += are[j];
                cur.y += aim[j];
                *o = cur;
            }
        }
    }
}

}  // namespace
}  // namespace afr

using namespace afr;

extern "C" int afr_wsclean_spectra(const double *flux, const double *coeffs, const uint8_t *log_poly,
                                   const double *ref_freq, const double *freq, int64_t nsrc,
                                   int64_t ncoeffs, int64_t nchan, double *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(nsrc >= 0 && ncoeffs >= 0 && nchan >= 0, "negative extent");
    const long long total = nsrc * nchan;
    if (total == 0) return 0;
    wsclean_spectra_kernel<<<(unsigned)((total + 255) / 256), 256, 0, stream>>>(
        flux, coeffs, log_poly, ref_freq, freq, nullptr, nsrc, (int)ncoeffs, nchan, out, nullptr);
    AFR_LAUNCH_OK();
    return 0;
}

extern "C" int afr_wsclean_predict(const double *uvw, const double *lm, const uint8_t *is_gauss,
                                   const double *gauss_shape, const double *flux, const double *coeffs,
                                   const uint8_t *log_poly, const double *ref_freq, const double *freq,
                                   int64_t nsrc, int64_t ncoeffs, int64_t nrow, int64_t nchan,
                                   int64_t ngauss, int chan_mode, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(nsrc >= 0 && ncoeffs >= 0 && nrow >= 0 && nchan >= 0, "negative extent");
    if (nrow == 0 || nchan == 0) return 0;
    if (nsrc == 0) {
        AFR_CUDA_OK(cudaMemsetAsync(out, 0, (size_t)nrow * nchan * 16, stream));
        return 0;
    }
    Scratch spec, spec_point, lmn, prm;
    const size_t nsf = (size_t)nsrc * nchan;
    AFR_CUDA_OK(spec.alloc(sizeof(double) * nsf, stream));
    AFR_CUDA_OK(spec_point.alloc(sizeof(double) * nsf, stream));
    wsclean_spectra_kernel<<<(unsigned)((nsf + 255) / 256), 256, 0, stream>>>(
        flux, coeffs, log_poly, ref_freq, freq, is_gauss, nsrc, (int)ncoeffs, nchan, (double *)spec.ptr,
        (double *)spec_point.ptr);
    AFR_LAUNCH_OK();
    // POINT sources: phasor-stream kernel, sign +2pi/c (wsclean_predict.py:40), n unclamped (:31)
    AFR_CUDA_OK(lmn.alloc(sizeof(double) * 3 * (size_t)nsrc, stream));
    int rc = launch_lm_to_lmn(lm, nsrc, kLmnDft, false, (double *)lmn.ptr, stream);
    if (rc) return rc;
    rc = run_phasor_stream(uvw, nrow, (const double *)lmn.ptr, nsrc, spec_point.ptr, false, nullptr, freq,
                           nchan, 1, kTwoPiOverC, false, /*adjoint=*/false, chan_mode == AFR_CHAN_EXACT,
                           false, out, stream);
    if (rc) return rc;
    if (ngauss > 0) {
        AFR_CUDA_OK(prm.alloc(sizeof(double) * 6 * (size_t)nsrc, stream));
        gauss_params_kernel<<<(unsigned)((nsrc + 255) / 256), 256, 0, stream>>>(lm, gauss_shape, nsrc,
                                                                                (double *)prm.ptr);
        AFR_LAUNCH_OK();
        // wsclean_predict.py:12-14
        const double fwhm = 2.0 * sqrt(2.0 * log(2.0));
        const double gauss_scale = (1.0 / fwhm) * sqrt(2.0) * 3.141592653589793 / kLightSpeed;
        const long long segs = (nchan + 32 * kGK - 1) / (32 * kGK);
        const long long warps = nrow * segs;
        long long blocks = (warps + 7) / 8;
        const long long cap = 8LL * sm_count();
        if (blocks > cap) blocks = cap;
        if (chan_mode == AFR_CHAN_EXACT)
            wsclean_gauss_kernel<false><<<(unsigned)blocks, 256, 0, stream>>>(
                uvw, (const double *)prm.ptr, is_gauss, (const double *)spec.ptr, freq, nsrc, nrow, nchan,
                gauss_scale, (double *)out);
        else
            wsclean_gauss_kernel<true><<<(unsigned)blocks, 256, 0, stream>>>(
                uvw, (const double *)prm.ptr, is_gauss, (const double *)spec.ptr, freq, nsrc, nrow, nchan,
                gauss_scale, (double *)out);
        AFR_LAUNCH_OK();
    }
    return 0;
}
