// beam_cube_dde + freq_grid_interp (africanus/rime/fast_beam_cubes.py:57-240, :10-54)
// for sm_100a.
//
// Bound by the 8 scattered corner gathers from the beam cube and the coalesced store of
// the (source,time,ant,chan,corr) output.  One thread produces one output element (all of
// its correlations); threads run along chan (the output's fastest axis apart from corr) so
// stores are contiguous and the per-channel interpolation data is a coalesced read.  The
// coordinate arithmetic uses explicitly rounded operations in the reference's order so the
// grid cell chosen by floor() is the reference's.
#include <algorithm>
#include <cstdlib>
#include <vector>

#include "afr_common.cuh"

namespace afr {
namespace {

// fast_beam_cubes.py:10-54
__global__ void freq_grid_interp_kernel(const double *freq, const double *bfm, long long nchan,
                                        long long nud, double *fd) {
    const long long f = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (f >= nchan) return;
    const double nu = freq[f];
    long long lo = 0, hi = nud - 1;
    while (lo <= hi) {
        const long long mid = lo + (hi - lo) / 2;
        const double bf = bfm[mid];
        if (bf < nu) {
            lo = mid + 1;
        } else if (bf > nu) {
            hi = mid - 1;
        } else {
            lo = mid;
            break;
        }
    }
    if (hi < lo) lo = hi;
    hi = lo + 1;
    double scale, wlo, glo;
    if (lo == -1) {
        scale = __ddiv_rn(nu, bfm[0]);
        wlo = 1.0;
        glo = 0.0;
    } else if (hi == nud) {
        scale = __ddiv_rn(nu, bfm[nud - 1]);
        wlo = 0.0;
        glo = (double)(nud - 2);
    } else {
        scale = 1.0;
        const double flo = bfm[lo], fhi = bfm[hi];
        wlo = __ddiv_rn(__dsub_rn(fhi, nu), __dsub_rn(fhi, flo));
        glo = (double)lo;
    }
    fd[3 * f] = scale;
    fd[3 * f + 1] = wlo;
    fd[3 * f + 2] = glo;
}

struct BeamParams {
    const void *beam;  // (lw,mh,nud,ncorr) complex
    const double *fd;  // (nchan,3)
    const double *lm, *pa, *perr, *ascale;
    void *out;  // (nsrc,ntime,nant,nchan,ncorr)
    const void *babs;     // |beam| per element, same layout as beam without the re/im axis
    const double *pa_sc;  // (ntime,nant,2) sin, cos of the parallactic angles
    const void *feed;     // optional (ntime,nant,2,2) complex feed rotation applied on the right
    double lower_l, lower_m, lscale, mscale, lmaxf, mmaxf;
    long long lw, mh, nud, nsrc, ntime, nant, nchan;
    int ncorr, coff;
    const uint8_t *row_flag;  // (ntime,nant) 1: pointing errors / antenna scaling constant along chan
};

// |re + i im| as the reference's np.abs (hypot).  hypot's overflow/underflow guards cost
// ~10x a square root; they only matter when re^2 + im^2 leaves the normal range, so take the
// direct sqrt (correctly rounded sqrt of a 1-ulp sum: within 1 ulp of hypot) inside it.
__device__ __forceinline__ double habs(double re, double im) {
    const double t = fma(re, re, im * im);
    if (t > 1e-290 && t < 1e290) return sqrt(t);
    return hypot(re, im);
}
__device__ __forceinline__ float habs(float re, float im) {
    const float t = fmaf(re, re, im * im);
    if (t > 1e-30f && t < 1e30f) return sqrtf(t);
    return hypotf(re, im);
}

// Per output element the reference takes |v| of 8 corners x ncorr beam values and one
// sincos of the parallactic angle; both depend only on the cube element / on (time, ant), and
// an output array has ~10^3 times more elements than the cube.  Two tiny pre-passes evaluate
// them once with the same functions (bit-identical results): the main kernel then spends its
// FP64 time on the interpolation itself (measured 0.47 -> 1.03 TB/s of output on the
// 257x257x64 cube, 4096 channels).
template <typename T>
__global__ void beam_abs_kernel(const T *beam, long long n, T *babs) {
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < n;
         i += (long long)gridDim.x * blockDim.x)
        babs[i] = habs(beam[2 * i], beam[2 * i + 1]);
}
__global__ void pa_sincos_kernel(const double *pa, long long n, double *sc) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double sn, cs;
    sincos(pa[i], &sn, &cs);
    sc[2 * i] = sn;
    sc[2 * i + 1] = cs;
}

// 16-byte (complex128) / 8-byte (complex64) vector accesses: one request per complex value and
// per pair of |beam| values instead of one per scalar (ncu: 107 -> ~55 global-load requests per warp
// and output element; the L1 wavefront pipe, 69 % busy, and lg-throttle stalls were the limit)
template <typename T>
struct Vec2;
template <>
struct Vec2<double> {
    using type = double2;
};
template <>
struct Vec2<float> {
    using type = float2;
};

// ROT (NC == 4 only): the epilogue right-multiplies the interpolated 2x2 Jones by the feed
// rotation L[t,a] -- einsum("stafij,tajk->stafik", beam_dde, feed_rot) of
// africanus/rime/examples/predict.py:469-472 -- before the one store, so the rotated DDE costs
// no extra pass over the (source,time,ant,chan,2,2) array.
// grid coordinates of one output element (fast_beam_cubes.py:130-151), explicitly rounded
// operations in the reference's order
__device__ __forceinline__ void beam_coords(const BeamParams &p, long long s, long long t, long long a,
                                            long long f, double sin_pa, double cos_pa, double &vl,
                                            double &vm) {
    const double l = p.lm[2 * s], m = p.lm[2 * s + 1];
    const double fscale = p.fd[3 * f];
    const double sl = __dmul_rn(l, fscale), sm = __dmul_rn(m, fscale);
    const double2 pe = *reinterpret_cast<const double2 *>(p.perr + ((t * p.nant + a) * p.nchan + f) * 2);
    const double tl = __dadd_rn(sl, pe.x), tm = __dadd_rn(sm, pe.y);
    vl = __dsub_rn(__dmul_rn(tl, cos_pa), __dmul_rn(tm, sin_pa));
    vm = __dadd_rn(__dmul_rn(tl, sin_pa), __dmul_rn(tm, cos_pa));
    const double2 as = *reinterpret_cast<const double2 *>(p.ascale + (a * p.nchan + f) * 2);
    vl = __dmul_rn(vl, as.x);
    vm = __dmul_rn(vm, as.y);
    vl = __dmul_rn(p.lscale, __dsub_rn(vl, p.lower_l));
    vm = __dmul_rn(p.mscale, __dsub_rn(vm, p.lower_m));
    vl = fmax(0.0, fmin(vl, p.lmaxf));
    vm = fmax(0.0, fmin(vm, p.mmaxf));
}

// one output element, all NC correlations of this launch: 8 corner gathers per element
template <typename T, int NC, bool ROT>
__device__ __forceinline__ void beam_element(const BeamParams &p, long long i, T *stage = nullptr) {
    const T *beam = (const T *)p.beam;
    const T *babs = (const T *)p.babs;
    T *out = (T *)p.out;
    {
        const long long f = i % p.nchan;
        long long rest = i / p.nchan;
        const long long a = rest % p.nant;
        rest /= p.nant;
        const long long t = rest % p.ntime;
        const long long s = rest / p.ntime;

        const double sin_pa = p.pa_sc[2 * (t * p.nant + a)], cos_pa = p.pa_sc[2 * (t * p.nant + a) + 1];
        const double nudw = p.fd[3 * f + 1];
        const double inv_nud = __dsub_rn(1.0, nudw);
        const long long gc0 = (long long)(int)p.fd[3 * f + 2], gc1 = gc0 + 1;
        double vl, vm;
        beam_coords(p, s, t, a, f, sin_pa, cos_pa, vl, vm);
        // :154-163
        const long long gl0 = (long long)(int)floor(vl), gm0 = (long long)(int)floor(vm);
        const long long gl1 = min(gl0 + 1, p.lw - 1), gm1 = min(gm0 + 1, p.mh - 1);
        const double ld = __dsub_rn(vl, (double)gl0), md = __dsub_rn(vm, (double)gm0);
        const double oml = __dsub_rn(1.0, ld), omm = __dsub_rn(1.0, md);

        // eight corners in the reference's order (:169-225)
        const long long gls[8] = {gl0, gl1, gl0, gl1, gl0, gl1, gl0, gl1};
        const long long gms[8] = {gm0, gm0, gm1, gm1, gm0, gm0, gm1, gm1};
        const double w4[4] = {__dmul_rn(oml, omm), __dmul_rn(ld, omm), __dmul_rn(oml, md),
                              __dmul_rn(ld, md)};
        T csr[NC], csi[NC], asum[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) csr[c] = csi[c] = asum[c] = T(0);
#pragma unroll
        for (int k = 0; k < 8; ++k) {
            const long long gc = k < 4 ? gc0 : gc1;
            const double wt = __dmul_rn(w4[k & 3], k < 4 ? nudw : inv_nud);
            const long long e = ((gls[k] * p.mh + gms[k]) * p.nud + gc) * p.ncorr + p.coff;
            using V2 = typename Vec2<T>::type;
            const V2 *b = reinterpret_cast<const V2 *>(beam + e * 2);
            T abv[NC];
            if (NC >= 2 && (p.ncorr & 1) == 0) {  // e is even: blocks of 4 / 2 start at even offsets
#pragma unroll
                for (int c = 0; c < NC; c += 2) {
                    const V2 v = *reinterpret_cast<const V2 *>(babs + e + c);
                    abv[c] = v.x;
                    abv[(c + 1) % NC] = v.y;
                }
            } else {
#pragma unroll
                for (int c = 0; c < NC; ++c) abv[c] = babs[e + c];
            }
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const V2 bv = b[c];
                const T br = bv.x, bi = bv.y;
                const T ab = abv[c];
                // accumulators live in the beam's precision, weights in float64 (:106-108)
                asum[c] = (T)__dadd_rn((double)asum[c], __dmul_rn(wt, (double)ab));
                csr[c] = (T)__dadd_rn((double)csr[c], __dmul_rn(wt, (double)br));
                csi[c] = (T)__dadd_rn((double)csi[c], __dmul_rn(wt, (double)bi));
            }
        }
        // `stage`: the NC values of this element go there instead of to the output array
        T *o = stage ? stage : out + (i * p.ncorr + p.coff) * 2;
        T er[NC], ei[NC];
#pragma unroll
        for (int c = 0; c < NC; ++c) {  // :227-238
            const T div = habs(csr[c], csi[c]);
            const T k = (div == T(0)) ? asum[c] : asum[c] / div;
            er[c] = csr[c] * k;
            ei[c] = csi[c] * k;
        }
        if (ROT && NC >= 2) {
            // the NC / 2 rows of the 2x2 Jones held by this launch (coff = 0: rows 0, 1 or row 0; 2: row 1)
            const T *L = (const T *)p.feed + (t * p.nant + a) * 8;
#pragma unroll
            for (int r = 0; r < NC / 2; ++r)
#pragma unroll
                for (int k = 0; k < 2; ++k) {  // out[r,k] = E[r,0] L[0,k] + E[r,1] L[1,k]
                    const T l0r = L[2 * k], l0i = L[2 * k + 1], l1r = L[2 * (2 + k)], l1i = L[2 * (2 + k) + 1];
                    const T e0r = er[(2 * r) % NC], e0i = ei[(2 * r) % NC];
                    const T e1r = er[(2 * r + 1) % NC], e1i = ei[(2 * r + 1) % NC];
                    typename Vec2<T>::type w;
                    w.x = (e0r * l0r - e0i * l0i) + (e1r * l1r - e1i * l1i);
                    w.y = (e0r * l0i + e0i * l0r) + (e1r * l1i + e1i * l1r);
                    reinterpret_cast<typename Vec2<T>::type *>(o)[2 * r + k] = w;
                }
        } else {
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                typename Vec2<T>::type w;
                w.x = er[c];
                w.y = ei[c];
                reinterpret_cast<typename Vec2<T>::type *>(o)[c] = w;
            }
        }
    }
}

// out-of-line copy for the channels / rows of beam_cube_dde_planes_kernel that have their own grid position
template <typename T, int NC, bool ROT>
__device__ __noinline__ void beam_element_call(const BeamParams &p, long long i, T *stage) {
    beam_element<T, NC, ROT>(p, i, stage);
}

template <typename T, int NC, bool ROT = false>
__global__ void __launch_bounds__(256) beam_cube_dde_kernel(const BeamParams p) {
    const long long total = p.nsrc * p.ntime * p.nant * p.nchan;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x)
        beam_element<T, NC, ROT>(p, i);
}

// ---------------------------------------------------------------------------------------
// Plane-interpolated kernel: one CTA per (source, time, antenna) row of the output.
//
// What bounded the one-element-per-thread kernel was the L1 -> register path (ncu: 768 B of corner
// values ingested per 64 B written, 8 scattered gathers per element) and, behind it, 320 FP64
// instructions per element.  But for one (source, time, antenna) the grid position (l, m) is the
// SAME for every in-band channel as long as pointing errors and antenna scaling do not change along
// the channel axis (they almost never do; out-of-band channels scale lm per channel), so the eight
// corners of every channel are the same four (l, m) corners at two of the nud frequency planes:
//   stage 1: S[g] = sum_{k<4} w4[k] beam[corner k][g], A[g] = sum_k w4[k] |beam[corner k][g]| for all
//            nud planes -- four CONTIGUOUS (nud x ncorr) reads per row instead of 8 gathers per element
//            (0.1 B ingested per byte written instead of 12);
//   stage 2: per channel E = nu_w S[g] + (1 - nu_w) S[g+1] (same for A), the amplitude
//            renormalisation (:227-238), the optional feed rotation, one coalesced store.
// ~160 FP64 instructions per element: the kernel becomes store-bound.
// The sum over the 8 corners is re-associated (4 spatial corners per plane first, then the two
// planes), where the reference accumulates w4[k] nu_w[plane] v over the 8 corners in turn: results
// differ from the element kernel by a few ulp (tested <= 1e-12 of the largest value; the project
// gate is 1e-10).  Rows whose pointing errors / antenna scaling change along the channel axis
// (row_flag, a device pre-pass) and out-of-band channels take the element path.
// ---------------------------------------------------------------------------------------
__global__ void __launch_bounds__(256) beam_row_flag_kernel(const double *perr, const double *ascale,
                                                            long long nant, long long nchan, uint8_t *flag) {
    // flag[t * nant + a] = 1: point_errors[t,a,:,:] and antenna_scaling[a,:,:] are constant along chan
    const long long ta = blockIdx.x, a = ta % nant;
    const double2 *pe = reinterpret_cast<const double2 *>(perr) + ta * nchan;
    const double2 *as = reinterpret_cast<const double2 *>(ascale) + a * nchan;
    const double2 p0 = pe[0], a0 = as[0];
    bool ok = true;
    for (long long f = threadIdx.x; f < nchan; f += blockDim.x) {
        const double2 pv = pe[f], av = as[f];
        ok = ok && pv.x == p0.x && pv.y == p0.y && av.x == a0.x && av.y == a0.y;
    }
    const int all = __syncthreads_and(ok ? 1 : 0);
    if (threadIdx.x == 0) flag[ta] = (uint8_t)all;
}

template <typename T, int NC, bool ROT>
__global__ void __launch_bounds__(256) beam_cube_dde_planes_kernel(const __grid_constant__ BeamParams p) {
    using V2 = typename Vec2<T>::type;
    extern __shared__ __align__(16) unsigned char plane_smem[];
    T *sre = reinterpret_cast<T *>(plane_smem);  // [nud][NC] real part of S, then imaginary part, then A
    T *sim = sre + p.nud * NC, *sab = sim + p.nud * NC;
    const T *beam = (const T *)p.beam;
    const T *babs = (const T *)p.babs;
    V2 *out = (V2 *)p.out;
    for (long long sta = blockIdx.x; sta < p.nsrc * p.ntime * p.nant; sta += gridDim.x) {
        const long long a = sta % p.nant, t = (sta / p.nant) % p.ntime, s = sta / (p.nant * p.ntime);
        const bool planes = p.row_flag[t * p.nant + a] != 0;  // CTA-uniform
        if (planes) {
            const double sin_pa = p.pa_sc[2 * (t * p.nant + a)], cos_pa = p.pa_sc[2 * (t * p.nant + a) + 1];
            // grid position of an in-band channel (lm scale 1): the row's pointing error / scaling,
            // the operations of beam_coords (fast_beam_cubes.py:130-151)
            const double l = p.lm[2 * s], m = p.lm[2 * s + 1];
            const double2 pe = *reinterpret_cast<const double2 *>(p.perr + ((t * p.nant + a) * p.nchan) * 2);
            const double2 as = *reinterpret_cast<const double2 *>(p.ascale + (a * p.nchan) * 2);
            const double tl = __dadd_rn(__dmul_rn(l, 1.0), pe.x), tm = __dadd_rn(__dmul_rn(m, 1.0), pe.y);
            double vl = __dsub_rn(__dmul_rn(tl, cos_pa), __dmul_rn(tm, sin_pa));
            double vm = __dadd_rn(__dmul_rn(tl, sin_pa), __dmul_rn(tm, cos_pa));
            vl = __dmul_rn(vl, as.x), vm = __dmul_rn(vm, as.y);
            vl = __dmul_rn(p.lscale, __dsub_rn(vl, p.lower_l));
            vm = __dmul_rn(p.mscale, __dsub_rn(vm, p.lower_m));
            vl = fmax(0.0, fmin(vl, p.lmaxf)), vm = fmax(0.0, fmin(vm, p.mmaxf));
            // :154-163
            const long long gl0 = (long long)(int)floor(vl), gm0 = (long long)(int)floor(vm);
            const long long gl1 = min(gl0 + 1, p.lw - 1), gm1 = min(gm0 + 1, p.mh - 1);
            const double ld = __dsub_rn(vl, (double)gl0), md = __dsub_rn(vm, (double)gm0);
            const double oml = __dsub_rn(1.0, ld), omm = __dsub_rn(1.0, md);
            const double w4[4] = {__dmul_rn(oml, omm), __dmul_rn(ld, omm), __dmul_rn(oml, md), __dmul_rn(ld, md)};
            const long long corner[4] = {gl0 * p.mh + gm0, gl1 * p.mh + gm0, gl0 * p.mh + gm1, gl1 * p.mh + gm1};
            // ---- stage 1: thread <-> (plane, correlation); the (nud x ncorr) block of a corner is contiguous
            for (int idx = threadIdx.x; idx < (int)p.nud * NC; idx += blockDim.x) {
                const int g = idx / NC, c = idx - g * NC;
                double re = 0.0, im = 0.0, ab = 0.0;
#pragma unroll
                for (int k = 0; k < 4; ++k) {
                    const long long e = (corner[k] * p.nud + g) * p.ncorr + p.coff + c;
                    const V2 bv = reinterpret_cast<const V2 *>(beam)[e];
                    re = (double)(T)__dadd_rn(re, __dmul_rn(w4[k], (double)bv.x));
                    im = (double)(T)__dadd_rn(im, __dmul_rn(w4[k], (double)bv.y));
                    ab = (double)(T)__dadd_rn(ab, __dmul_rn(w4[k], (double)babs[e]));
                }
                sre[idx] = (T)re, sim[idx] = (T)im, sab[idx] = (T)ab;
            }
        }
        __syncthreads();
        // ---- stage 2: thread <-> channel
        for (long long f = threadIdx.x; f < p.nchan; f += blockDim.x) {
            const long long i = sta * p.nchan + f;
            if (!planes || p.fd[3 * f] != 1.0) {  // per-channel grid position: the element path
                beam_element_call<T, NC, ROT>(p, i, nullptr);
                continue;
            }
            const double nudw = p.fd[3 * f + 1], inv = __dsub_rn(1.0, nudw);
            const int g = (int)p.fd[3 * f + 2];
            T er[NC], ei[NC];
#pragma unroll
            for (int c = 0; c < NC; ++c) {
                const int lo = g * NC + c, hi = lo + NC;
                // (the plane sums are already re-associated against the reference's corner order, so the
                // two-plane combination may contract to an FMA: 2 instead of 3 instructions per value)
                const T csr = (T)fma(nudw, (double)sre[lo], inv * (double)sre[hi]);
                const T csi = (T)fma(nudw, (double)sim[lo], inv * (double)sim[hi]);
                const T asum = (T)fma(nudw, (double)sab[lo], inv * (double)sab[hi]);
                // :227-238  corr_sum * (absc_sum / |corr_sum|): one reciprocal square root (MUFU seed + two
                // Newton steps, <= 2 ulp) instead of a square root and a division (~30 of the ~40 FP64
                // instructions of a correlation) while re^2 + im^2 is inside the normal range
                T kk;
                if constexpr (sizeof(T) == 8) {
                    const double tt = fma((double)csr, (double)csr, (double)csi * (double)csi);
                    if (tt > 1e-290 && tt < 1e290) {
                        kk = (T)((double)asum * rsqrt(tt));
                    } else {
                        const T div = habs(csr, csi);
                        kk = (div == T(0)) ? asum : asum / div;
                    }
                } else {
                    const T div = habs(csr, csi);
                    kk = (div == T(0)) ? asum : asum / div;
                }
                er[c] = csr * kk, ei[c] = csi * kk;
            }
            V2 *o = out + i * p.ncorr + p.coff;
            if (ROT && NC == 4) {
                const T *L = (const T *)p.feed + (t * p.nant + a) * 8;
#pragma unroll
                for (int r = 0; r < 2; ++r)
#pragma unroll
                    for (int k = 0; k < 2; ++k) {  // out[r,k] = E[r,0] L[0,k] + E[r,1] L[1,k]
                        const T l0r = L[2 * k], l0i = L[2 * k + 1], l1r = L[2 * (2 + k)], l1i = L[2 * (2 + k) + 1];
                        const T e0r = er[(2 * r) % NC], e0i = ei[(2 * r) % NC];
                        const T e1r = er[(2 * r + 1) % NC], e1i = ei[(2 * r + 1) % NC];
                        V2 w;
                        w.x = (e0r * l0r - e0i * l0i) + (e1r * l1r - e1i * l1i);
                        w.y = (e0r * l0i + e0i * l0r) + (e1r * l1i + e1i * l1r);
                        o[2 * r + k] = w;
                    }
            } else {
#pragma unroll
                for (int c = 0; c < NC; ++c) {
                    V2 w;
                    w.x = er[c], w.y = ei[c];
                    o[c] = w;
                }
            }
        }
        __syncthreads();  // the planes are overwritten by the next row
    }
}

// Stage 1 of beam_cube_dde_planes_kernel on its own: the four spatial corners of every frequency
// plane reduced per (source, time, antenna) row and written out, planes[row][g][re 0..3 | im 0..3 |
// |.| 0..3] (2x2 complex128 beams).  43x smaller than the (source,time,ant,chan,2,2) Jones array at
// 4096 channels; the predict kernel's producers combine the two planes of a channel themselves
// (afr_rime_ws.cu, SAMPLE), so the interpolated Jones never exist in memory.
__global__ void __launch_bounds__(256) beam_plane_reduce_kernel(const __grid_constant__ BeamParams p, double *planes) {
    const double *beam = (const double *)p.beam;
    const double *babs = (const double *)p.babs;
    for (long long sta = blockIdx.x; sta < p.nsrc * p.ntime * p.nant; sta += gridDim.x) {
        const long long a = sta % p.nant, t = (sta / p.nant) % p.ntime, s = sta / (p.nant * p.ntime);
        const double sin_pa = p.pa_sc[2 * (t * p.nant + a)], cos_pa = p.pa_sc[2 * (t * p.nant + a) + 1];
        // the operations of beam_cube_dde_planes_kernel (fast_beam_cubes.py:130-163), lm scale 1
        const double l = p.lm[2 * s], m = p.lm[2 * s + 1];
        const double2 pe = *reinterpret_cast<const double2 *>(p.perr + ((t * p.nant + a) * p.nchan) * 2);
        const double2 as = *reinterpret_cast<const double2 *>(p.ascale + (a * p.nchan) * 2);
        const double tl = __dadd_rn(__dmul_rn(l, 1.0), pe.x), tm = __dadd_rn(__dmul_rn(m, 1.0), pe.y);
        double vl = __dsub_rn(__dmul_rn(tl, cos_pa), __dmul_rn(tm, sin_pa));
        double vm = __dadd_rn(__dmul_rn(tl, sin_pa), __dmul_rn(tm, cos_pa));
        vl = __dmul_rn(vl, as.x), vm = __dmul_rn(vm, as.y);
        vl = __dmul_rn(p.lscale, __dsub_rn(vl, p.lower_l));
        vm = __dmul_rn(p.mscale, __dsub_rn(vm, p.lower_m));
        vl = fmax(0.0, fmin(vl, p.lmaxf)), vm = fmax(0.0, fmin(vm, p.mmaxf));
        const long long gl0 = (long long)(int)floor(vl), gm0 = (long long)(int)floor(vm);
        const long long gl1 = min(gl0 + 1, p.lw - 1), gm1 = min(gm0 + 1, p.mh - 1);
        const double ld = __dsub_rn(vl, (double)gl0), md = __dsub_rn(vm, (double)gm0);
        const double oml = __dsub_rn(1.0, ld), omm = __dsub_rn(1.0, md);
        const double w4[4] = {__dmul_rn(oml, omm), __dmul_rn(ld, omm), __dmul_rn(oml, md), __dmul_rn(ld, md)};
        const long long corner[4] = {gl0 * p.mh + gm0, gl1 * p.mh + gm0, gl0 * p.mh + gm1, gl1 * p.mh + gm1};
        double *row = planes + sta * p.nud * 12;
        for (int idx = threadIdx.x; idx < (int)p.nud * 4; idx += blockDim.x) {
            const int g = idx >> 2, c = idx & 3;
            double re = 0.0, im = 0.0, ab = 0.0;
#pragma unroll
            for (int k = 0; k < 4; ++k) {
                const long long e = (corner[k] * p.nud + g) * 4 + c;
                const double2 bv = reinterpret_cast<const double2 *>(beam)[e];
                re = __dadd_rn(re, __dmul_rn(w4[k], bv.x));
                im = __dadd_rn(im, __dmul_rn(w4[k], bv.y));
                ab = __dadd_rn(ab, __dmul_rn(w4[k], babs[e]));
            }
            row[g * 12 + c] = re, row[g * 12 + 4 + c] = im, row[g * 12 + 8 + c] = ab;
        }
    }
}

// ok[0] = 1 when every (time, antenna) row keeps one grid position along the channel axis and every
// channel is inside the cube's frequency range: the conditions under which a channel's Jones is the
// combination of two planes
__global__ void beam_planes_ok_kernel(const uint8_t *row_flag, long long nrows, const double *fd, long long nchan,
                                      int *ok) {
    bool good = true;
    for (long long i = threadIdx.x; i < nrows; i += blockDim.x) good = good && row_flag[i] != 0;
    for (long long f = threadIdx.x; f < nchan; f += blockDim.x) good = good && fd[3 * f] == 1.0;
    const int all = __syncthreads_and(good ? 1 : 0);
    if (threadIdx.x == 0) ok[0] = all;
}

// africanus/rime/feeds.py:13-48: (n,) parallactic angles -> (n,2,2) complex.
// linear [[cos, sin], [-sin, cos]]; circular diag(exp(-i pa), exp(+i pa))
template <typename T>
__global__ void feed_rotation_kernel(const double *pa, long long n, int circular, T *out) {
    const long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x;
    if (i >= n) return;
    double sn, cs;
    sincos(pa[i], &sn, &cs);
    T *o = out + 8 * i;
    const T c = (T)cs, s = (T)sn, z = T(0);
    if (circular) {
        o[0] = c, o[1] = -s, o[2] = z, o[3] = z, o[4] = z, o[5] = z, o[6] = c, o[7] = s;
    } else {
        o[0] = c, o[1] = z, o[2] = s, o[3] = z, o[4] = -s, o[5] = z, o[6] = c, o[7] = z;
    }
}

int grid_for(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

template <typename T>
int launch_beam(BeamParams p, cudaStream_t stream) {
    const long long total = p.nsrc * p.ntime * p.nant * p.nchan;
    if (total <= 0) return 0;
    // AFR_BEAM_PLANES=0 selects the one-element-per-thread kernel (tests compare the two)
    const size_t plane_bytes = (size_t)p.nud * 4 * 3 * sizeof(T);
    const bool planes = p.row_flag != nullptr && plane_bytes <= 48 * 1024 && p.nchan >= 64 &&
                        !(getenv("AFR_BEAM_PLANES") && atoi(getenv("AFR_BEAM_PLANES")) == 0);
    const long long rows = p.nsrc * p.ntime * p.nant;
    const int grid = planes ? (int)std::min<long long>(rows, 64LL * sm_count()) : grid_for(total);
    int c = 0;
    while (c < p.ncorr) {  // correlations are independent: blocks of 4, 2, 1
        p.coff = c;
        const int nc = p.ncorr - c >= 4 ? 4 : (p.ncorr - c >= 2 ? 2 : 1);
        const size_t sm = (size_t)p.nud * nc * 3 * sizeof(T);
        if (planes && nc == 4 && p.feed)
            beam_cube_dde_planes_kernel<T, 4, true><<<grid, 256, sm, stream>>>(p);
        else if (planes && nc == 4)
            beam_cube_dde_planes_kernel<T, 4, false><<<grid, 256, sm, stream>>>(p);
        else if (planes && nc == 2)
            beam_cube_dde_planes_kernel<T, 2, false><<<grid, 256, sm, stream>>>(p);
        else if (planes)
            beam_cube_dde_planes_kernel<T, 1, false><<<grid, 256, sm, stream>>>(p);
        else if (nc == 4 && p.feed)
            beam_cube_dde_kernel<T, 4, true><<<grid, 256, 0, stream>>>(p);
        else if (nc == 4)
            beam_cube_dde_kernel<T, 4><<<grid, 256, 0, stream>>>(p);
        else if (nc == 2)
            beam_cube_dde_kernel<T, 2><<<grid, 256, 0, stream>>>(p);
        else
            beam_cube_dde_kernel<T, 1><<<grid, 256, 0, stream>>>(p);
        AFR_LAUNCH_OK();
        c += nc;
    }
    return 0;
}

}  // namespace
}  // namespace afr

using namespace afr;

extern "C" int afr_freq_grid_interp(const double *freq, const double *beam_freq_map,
                                    int64_t nchan, int64_t nud, double *freq_data,
                                    void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(nud >= 1 && nchan >= 0, "afr_freq_grid_interp: bad extent");
    if (nchan == 0) return 0;
    freq_grid_interp_kernel<<<(int)((nchan + 255) / 256), 256, 0, stream>>>(freq, beam_freq_map,
                                                                          nchan, nud, freq_data);
    AFR_LAUNCH_OK();
    return 0;
}

extern "C" int afr_feed_rotation(const double *parallactic_angles, int64_t n, int feed_type,
                                 int is_c64, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(n >= 0, "negative extent");
    AFR_REQUIRE(feed_type == AFR_FEED_LINEAR || feed_type == AFR_FEED_CIRCULAR, "Invalid feed_type");
    if (n == 0) return 0;
    if (is_c64)
        feed_rotation_kernel<float><<<(int)((n + 255) / 256), 256, 0, stream>>>(parallactic_angles, n,
                                                                                feed_type, (float *)out);
    else
        feed_rotation_kernel<double><<<(int)((n + 255) / 256), 256, 0, stream>>>(parallactic_angles, n,
                                                                                 feed_type, (double *)out);
    AFR_LAUNCH_OK();
    return 0;
}

extern "C" int afr_beam_cube_dde(const void *beam, const double *ext_host_or_dev,
                                 const double *beam_freq_map, const double *lm,
                                 const double *parallactic_angles, const double *point_errors,
                                 const double *antenna_scaling, const double *freq, int64_t lw,
                                 int64_t mh, int64_t nud, int64_t ncorr, int64_t nsrc,
                                 int64_t ntime, int64_t nant, int64_t nchan, int is_c64,
                                 void *out, void *stream_) {
    return afr_beam_cube_dde_rot(beam, ext_host_or_dev, beam_freq_map, lm, parallactic_angles,
                                 point_errors, antenna_scaling, freq, nullptr, lw, mh, nud, ncorr, nsrc,
                                 ntime, nant, nchan, is_c64, out, stream_);
}

extern "C" int afr_beam_cube_dde_rot(const void *beam, const double *ext_host_or_dev,
                                     const double *beam_freq_map, const double *lm,
                                     const double *parallactic_angles, const double *point_errors,
                                     const double *antenna_scaling, const double *freq,
                                     const void *feed_rotation, int64_t lw, int64_t mh, int64_t nud,
                                     int64_t ncorr, int64_t nsrc, int64_t ntime, int64_t nant,
                                     int64_t nchan, int is_c64, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(feed_rotation == nullptr || ncorr == 4,
                "a feed rotation needs 2x2 correlations (beam of shape (lw,mh,nud,2,2))");
    // the kernel reads / writes complex values and (l, m) pairs as one vector each
    const uintptr_t cal = is_c64 ? 8 : 16;
    AFR_REQUIRE(reinterpret_cast<uintptr_t>(beam) % cal == 0 && reinterpret_cast<uintptr_t>(out) % cal == 0 &&
                    reinterpret_cast<uintptr_t>(point_errors) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(antenna_scaling) % 16 == 0,
                "afr_beam_cube_dde: beam / out must be aligned to one complex value, point_errors / "
                "antenna_scaling to 16 bytes");
    // fast_beam_cubes.py:74-75
    AFR_REQUIRE(lw >= 2 && mh >= 2 && nud >= 2, "beam_lw, beam_mh and beam_nud must be >= 2");
    AFR_REQUIRE(ncorr >= 1 && nsrc >= 0 && ntime >= 0 && nant >= 0 && nchan >= 0, "bad extent");
    if (nsrc == 0 || ntime == 0 || nant == 0 || nchan == 0) return 0;
    Scratch fd;
    AFR_CUDA_OK(fd.alloc(sizeof(double) * 3 * (size_t)nchan, stream));
    int rc = afr_freq_grid_interp(freq, beam_freq_map, nchan, nud, (double *)fd.ptr, stream_);
    if (rc) return rc;
    // the four extents are needed on the host to form lscale/mscale (:83-92)
    double ext[4];
    AFR_CUDA_OK(cudaMemcpyAsync(ext, ext_host_or_dev, sizeof(ext), cudaMemcpyDefault, stream));
    AFR_CUDA_OK(cudaStreamSynchronize(stream));
    // which (time, antenna) rows keep one grid position along the channel axis
    Scratch rflag;
    AFR_CUDA_OK(rflag.alloc((size_t)(ntime * nant), stream));
    beam_row_flag_kernel<<<(unsigned)(ntime * nant), 256, 0, stream>>>(point_errors, antenna_scaling, nant, nchan,
                                                                      (uint8_t *)rflag.ptr);
    AFR_LAUNCH_OK();
    // pre-passes: |beam| and sin/cos of the parallactic angles
    const long long nbeam = lw * mh * nud * ncorr;
    Scratch babs, pasc;
    AFR_CUDA_OK(babs.alloc((size_t)nbeam * (is_c64 ? 4 : 8), stream));
    AFR_CUDA_OK(pasc.alloc(sizeof(double) * 2 * (size_t)(ntime * nant), stream));
    {
        const int blocks = (int)std::min<long long>((nbeam + 255) / 256, 32LL * sm_count());
        if (is_c64)
            beam_abs_kernel<float><<<blocks, 256, 0, stream>>>((const float *)beam, nbeam, (float *)babs.ptr);
        else
            beam_abs_kernel<double><<<blocks, 256, 0, stream>>>((const double *)beam, nbeam, (double *)babs.ptr);
        AFR_LAUNCH_OK();
        pa_sincos_kernel<<<(int)((ntime * nant + 255) / 256), 256, 0, stream>>>(
            parallactic_angles, ntime * nant, (double *)pasc.ptr);
        AFR_LAUNCH_OK();
    }
    BeamParams p{};
    p.beam = beam;
    p.babs = babs.ptr;
    p.pa_sc = (const double *)pasc.ptr;
    p.feed = feed_rotation;
    p.fd = (const double *)fd.ptr;
    p.lm = lm;
    p.pa = parallactic_angles;
    p.perr = point_errors;
    p.ascale = antenna_scaling;
    p.out = out;
    p.lower_l = ext[0];
    p.lower_m = ext[2];
    p.lmaxf = (double)(lw - 1);
    p.mmaxf = (double)(mh - 1);
    p.lscale = p.lmaxf / (ext[1] - ext[0]);
    p.mscale = p.mmaxf / (ext[3] - ext[2]);
    p.lw = lw;
    p.mh = mh;
    p.nud = nud;
    p.nsrc = nsrc;
    p.ntime = ntime;
    p.nant = nant;
    p.nchan = nchan;
    p.ncorr = (int)ncorr;
    p.row_flag = (const uint8_t *)rflag.ptr;
    return is_c64 ? launch_beam<float>(p, stream) : launch_beam<double>(p, stream);
}

// Plane-reduced beam for in-kernel sampling: planes (nsrc,ntime,nant,nud,12) float64,
// fd (nchan,3) the frequency-grid table of freq_grid_interp, ok (device int) = 1 when every channel of
// every row is the combination of two planes.  2x2 complex128 beams.
extern "C" int afr_beam_plane_reduce(const void *beam, const double *ext_host_or_dev, const double *beam_freq_map,
                                     const double *lm, const double *parallactic_angles,
                                     const double *point_errors, const double *antenna_scaling,
                                     const double *freq, int64_t lw, int64_t mh, int64_t nud, int64_t nsrc,
                                     int64_t ntime, int64_t nant, int64_t nchan, double *planes, double *fd,
                                     int *ok, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(reinterpret_cast<uintptr_t>(beam) % 16 == 0 && reinterpret_cast<uintptr_t>(planes) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(point_errors) % 16 == 0 &&
                    reinterpret_cast<uintptr_t>(antenna_scaling) % 16 == 0,
                "afr_beam_plane_reduce: beam / planes / point_errors / antenna_scaling must be 16-byte aligned");
    AFR_REQUIRE(lw >= 2 && mh >= 2 && nud >= 2, "beam_lw, beam_mh and beam_nud must be >= 2");
    AFR_REQUIRE(nsrc >= 0 && ntime >= 1 && nant >= 1 && nchan >= 1, "bad extent");
    int rc = afr_freq_grid_interp(freq, beam_freq_map, nchan, nud, fd, stream_);
    if (rc) return rc;
    double ext[4];
    AFR_CUDA_OK(cudaMemcpyAsync(ext, ext_host_or_dev, sizeof(ext), cudaMemcpyDefault, stream));
    AFR_CUDA_OK(cudaStreamSynchronize(stream));
    Scratch rflag, babs, pasc;
    AFR_CUDA_OK(rflag.alloc((size_t)(ntime * nant), stream));
    beam_row_flag_kernel<<<(unsigned)(ntime * nant), 256, 0, stream>>>(point_errors, antenna_scaling, nant, nchan,
                                                                      (uint8_t *)rflag.ptr);
    AFR_LAUNCH_OK();
    beam_planes_ok_kernel<<<1, 256, 0, stream>>>((const uint8_t *)rflag.ptr, ntime * nant, fd, nchan, ok);
    AFR_LAUNCH_OK();
    if (nsrc == 0) return 0;
    const long long nbeam = lw * mh * nud * 4;
    AFR_CUDA_OK(babs.alloc((size_t)nbeam * 8, stream));
    AFR_CUDA_OK(pasc.alloc(sizeof(double) * 2 * (size_t)(ntime * nant), stream));
    beam_abs_kernel<double><<<(int)std::min<long long>((nbeam + 255) / 256, 32LL * sm_count()), 256, 0, stream>>>(
        (const double *)beam, nbeam, (double *)babs.ptr);
    AFR_LAUNCH_OK();
    pa_sincos_kernel<<<(int)((ntime * nant + 255) / 256), 256, 0, stream>>>(parallactic_angles, ntime * nant,
                                                                           (double *)pasc.ptr);
    AFR_LAUNCH_OK();
    BeamParams p{};
    p.beam = beam;
    p.babs = babs.ptr;
    p.pa_sc = (const double *)pasc.ptr;
    p.lm = lm;
    p.perr = point_errors;
    p.ascale = antenna_scaling;
    p.lower_l = ext[0];
    p.lower_m = ext[2];
    p.lmaxf = (double)(lw - 1);
    p.mmaxf = (double)(mh - 1);
    p.lscale = p.lmaxf / (ext[1] - ext[0]);
    p.mscale = p.mmaxf / (ext[3] - ext[2]);
    p.lw = lw, p.mh = mh, p.nud = nud, p.nsrc = nsrc, p.ntime = ntime, p.nant = nant, p.nchan = nchan;
    p.ncorr = 4;
    const long long rows = nsrc * ntime * nant;
    beam_plane_reduce_kernel<<<(unsigned)std::min<long long>(rows, 64LL * sm_count()), 256, 0, stream>>>(p, planes);
    AFR_LAUNCH_OK();
    return 0;
}
