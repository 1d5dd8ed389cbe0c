/*
 * afr_oracle.c -- TEST INFRASTRUCTURE ONLY (not product code).
 *
 * A plain-C, scalar, CPU restatement of the codex-africanus RIME / DFT hot
 * path, used as the parity checker for the sm_100a CUDA kernels in
 * codex_africanus_b200/csrc.  Only tests/, __graft_entry__.smoke() and the
 * cpu_baseline / --impl reference legs of bench.py may load this library; the
 * product path never does (it fails loudly when the CUDA library is missing).
 *
 * Parity pinning: every entry point here is checked in tests/test_oracle.py
 * against golden vectors produced by importing the reference's own numba
 * implementation (oracle/gen_golden.py, fixtures in tests/golden/), and against
 * the reference's known-answer tests (beam 0.470255+0.4786j, freq-interp
 * table, bit-exact phasor, FFT identity, adjointness).
 *
 * Each function cites the reference lines (relative to /root/reference/) whose
 * arithmetic -- evaluation order, rounding points, dtype promotion -- it
 * restates.  The reference is numba without fastmath, i.e. IEEE arithmetic with
 * no FMA contraction: this file must be compiled with -ffp-contract=off and
 * without -ffast-math (see oracle/Makefile).
 *
 * dtype handling: the numpy-facing wrapper (oracle/__init__.py) widens every
 * real input to float64 (exact) and tells us through *_f32 flags which inputs
 * were float32, so that the handful of operations the reference performs in
 * float32 because BOTH operands are float32 can be rounded the same way.
 *
 * Optional OpenMP parallelism (used only by bench.py's CPU-baseline legs) is
 * over the outermost INDEPENDENT axis, so per-element summation order is the
 * reference's sequential order regardless of the thread count.
 */
#include <math.h>
#include <stdint.h>
#include <stdlib.h>
#include <string.h>

#ifdef _OPENMP
#include <omp.h>
#endif

#define AFR_C 2.99792458e8 /* africanus/constants/consts.py:6 */

static double two_pi_over_c(void) {
    /* africanus/constants/consts.py:8 : 2 * math.pi / c */
    return 2.0 * 3.141592653589793 / AFR_C;
}

int orc_max_threads(void) {
#ifdef _OPENMP
    return omp_get_max_threads();
#else
    return 1;
#endif
}

/* ------------------------------------------------------------------------ */
/* helpers that round like the reference does for float32 operands           */
/* ------------------------------------------------------------------------ */
static inline double sq_as(double x, int is_f32) {
    /* x**2 evaluated in the dtype of x */
    if (is_f32) {
        float xf = (float)x;
        return (double)(xf * xf);
    }
    return x * x;
}

/* l*u + m*v in the promoted dtype of (lm, uvw) */
static inline double lu_plus_mv(double l, double u, double m, double v, int both_f32) {
    if (both_f32) {
        float a = (float)l * (float)u;
        float b = (float)m * (float)v;
        return (double)(a + b);
    }
    return l * u + m * v;
}

/* ------------------------------------------------------------------------ */
/* phase_delay : africanus/rime/phase.py:20-63                               */
/* ------------------------------------------------------------------------ */
/*
 * All-float64 working precision (any input is float64 => complex128 output,
 * phase.py:26).  lm_f32: constants `one`, `zero` and -2pi/c are typed as
 * lm.dtype (phase.py:23-25) so n and the constant are float32-rounded.
 * The general mixed case is reproduced by the operand-wise promotion below.
 * sign = +1 -> "fourier" (constant = -2pi/c), -1 -> "casa" (phase.py:29-34).
 * out: (nsrc, nrow, nchan) complex128 interleaved.
 */
int orc_phase_delay_f64(const double *lm, const double *uvw, const double *freq,
                        int64_t nsrc, int64_t nrow, int64_t nchan, int sign,
                        int lm_f32, int uvw_f32, int freq_f32, double *out) {
    if (lm_f32 && uvw_f32 && freq_f32) return -1; /* use the _f32 entry */
    double cst = -two_pi_over_c();
    if (lm_f32) cst = (double)(float)cst; /* phase.py:25 */
    if (sign < 0) cst = -cst;

#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nsrc; ++s) {
        double l = lm[2 * s], m = lm[2 * s + 1];
        double n;
        if (lm_f32) { /* phase.py:42-43 in float32 */
            float lf = (float)l, mf = (float)m;
            float nf = 1.0f - lf * lf - mf * mf;
            nf = sqrtf(nf < 0.0f ? 0.0f : nf) - 1.0f;
            n = (double)nf;
        } else {
            n = 1.0 - l * l - m * m;
            n = sqrt(n < 0.0 ? 0.0 : n) - 1.0;
        }
        for (int64_t r = 0; r < nrow; ++r) {
            double u = uvw[3 * r], v = uvw[3 * r + 1], w = uvw[3 * r + 2];
            /* phase.py:49 : constant * (l*u + m*v + n*w), operand-wise promotion */
            double real_phase;
            if (lm_f32 && uvw_f32) {
                float a = (float)l * (float)u + (float)m * (float)v + (float)n * (float)w;
                real_phase = (double)((float)cst * a);
            } else {
                real_phase = cst * (l * u + m * v + n * w);
            }
            double *o = out + 2 * ((s * nrow + r) * nchan);
            for (int64_t f = 0; f < nchan; ++f) {
                double p = real_phase * freq[f]; /* phase.py:53 */
                o[2 * f] = cos(p);               /* phase.py:58-59 */
                o[2 * f + 1] = sin(p);
            }
        }
    }
    return 0;
}

/* all-float32 inputs: whole phase in float32, complex64 out (phase.py:23-26) */
int orc_phase_delay_f32(const float *lm, const float *uvw, const float *freq,
                        int64_t nsrc, int64_t nrow, int64_t nchan, int sign,
                        float *out) {
    float cst = (float)(-two_pi_over_c());
    if (sign < 0) cst = -cst;
#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nsrc; ++s) {
        float l = lm[2 * s], m = lm[2 * s + 1];
        float n = 1.0f - l * l - m * m;
        n = sqrtf(n < 0.0f ? 0.0f : n) - 1.0f;
        for (int64_t r = 0; r < nrow; ++r) {
            float u = uvw[3 * r], v = uvw[3 * r + 1], w = uvw[3 * r + 2];
            float real_phase = cst * (l * u + m * v + n * w);
            float *o = out + 2 * ((s * nrow + r) * nchan);
            for (int64_t f = 0; f < nchan; ++f) {
                float p = real_phase * freq[f];
                o[2 * f] = cosf(p);
                o[2 * f + 1] = sinf(p);
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* im_to_vis : africanus/dft/kernels.py:23-69                                */
/* ------------------------------------------------------------------------ */
/*
 * image: (nsrc, nchan, ncorr) float64, or complex128 interleaved when
 * image_complex.  out: (nrow, nchan, ncorr) complex128 interleaved; when
 * out_c64 the accumulator is rounded to float32 after every += as the
 * reference's complex64 output array does (kernels.py:45,65).
 * sign +1 = "fourier" (-2pi/c), -1 = "casa" (kernels.py:34-39).
 * n = sqrt(1 - l^2 - m^2) - 1 with NO clamp (kernels.py:54).
 * Loop nest r -> s -> nu -> c (kernels.py:48-65); zero pixels skipped (:64).
 */
int orc_im_to_vis(const double *image, int image_complex, const double *uvw,
                  const double *lm, const double *freq, int64_t nsrc, int64_t nrow,
                  int64_t nchan, int64_t ncorr, int sign, int lm_f32, int uvw_f32,
                  int out_c64, double *out) {
    const double cst = sign < 0 ? two_pi_over_c() : -two_pi_over_c();
    const int both_f32 = lm_f32 && uvw_f32;
    double *nn = (double *)malloc(sizeof(double) * (size_t)(nsrc > 0 ? nsrc : 1));
    if (!nn) return -2;
    for (int64_t s = 0; s < nsrc; ++s) {
        double l = lm[2 * s], m = lm[2 * s + 1];
        nn[s] = sqrt(1.0 - sq_as(l, lm_f32) - sq_as(m, lm_f32)) - 1.0;
    }
    memset(out, 0, sizeof(double) * 2 * (size_t)(nrow * nchan * ncorr));

#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nrow; ++r) {
        double u = uvw[3 * r], v = uvw[3 * r + 1], w = uvw[3 * r + 2];
        double *o = out + 2 * (r * nchan * ncorr);
        for (int64_t s = 0; s < nsrc; ++s) {
            double l = lm[2 * s], m = lm[2 * s + 1], n = nn[s];
            double real_phase = cst * (lu_plus_mv(l, u, m, v, both_f32) + n * w);
            for (int64_t f = 0; f < nchan; ++f) {
                /* p = real_phase*freq*1j ; exp(p) = cos + i sin (kernels.py:61-65) */
                double p = real_phase * freq[f];
                double cp = cos(p), sp = sin(p);
                for (int64_t c = 0; c < ncorr; ++c) {
                    int64_t ii = (s * nchan + f) * ncorr + c;
                    double ire, iim;
                    if (image_complex) {
                        ire = image[2 * ii];
                        iim = image[2 * ii + 1];
                    } else {
                        ire = image[ii];
                        iim = 0.0;
                    }
                    if (ire == 0.0 && iim == 0.0) continue; /* kernels.py:64 */
                    /* complex product (cp + i sp)(ire + i iim) */
                    double tre = cp * ire - sp * iim;
                    double tim = cp * iim + sp * ire;
                    double *a = o + 2 * (f * ncorr + c);
                    if (out_c64) {
                        a[0] = (double)(float)(a[0] + tre);
                        a[1] = (double)(float)(a[1] + tim);
                    } else {
                        a[0] += tre;
                        a[1] += tim;
                    }
                }
            }
        }
    }
    free(nn);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* vis_to_im : africanus/dft/kernels.py:83-148                               */
/* ------------------------------------------------------------------------ */
/*
 * vis: (nrow, nchan, ncorr) float64 or complex128 interleaved; flags uint8 of
 * the same shape.  out: (nsrc, nchan, ncorr) float64 (rounded through float32
 * after each += when out_f32).  Sign is OPPOSITE to im_to_vis for the same
 * convention (kernels.py:110-115).  A (row, chan) sample is skipped if any of
 * its correlations is flagged (kernels.py:136-137).  Loop nest s -> r -> nu -> c.
 */
int orc_vis_to_im(const double *vis, int vis_complex, const double *uvw,
                  const double *lm, const double *freq, const uint8_t *flags,
                  int64_t nsrc, int64_t nrow, int64_t nchan, int64_t ncorr, int sign,
                  int lm_f32, int uvw_f32, int out_f32, double *out) {
    const double cst = sign < 0 ? -two_pi_over_c() : two_pi_over_c();
    const int both_f32 = lm_f32 && uvw_f32;
    /* any-flag mask per (row, chan) */
    uint8_t *anyf = (uint8_t *)malloc((size_t)(nrow * nchan > 0 ? nrow * nchan : 1));
    if (!anyf) return -2;
    for (int64_t i = 0; i < nrow * nchan; ++i) {
        uint8_t a = 0;
        for (int64_t c = 0; c < ncorr; ++c) a |= flags[i * ncorr + c];
        anyf[i] = a;
    }
    memset(out, 0, sizeof(double) * (size_t)(nsrc * nchan * ncorr));

#pragma omp parallel for schedule(static)
    for (int64_t s = 0; s < nsrc; ++s) {
        double l = lm[2 * s], m = lm[2 * s + 1];
        double n = sqrt(1.0 - sq_as(l, lm_f32) - sq_as(m, lm_f32)) - 1.0;
        double *o = out + s * nchan * ncorr;
        for (int64_t r = 0; r < nrow; ++r) {
            double u = uvw[3 * r], v = uvw[3 * r + 1], w = uvw[3 * r + 2];
            double real_phase = cst * (lu_plus_mv(l, u, m, v, both_f32) + n * w);
            for (int64_t f = 0; f < nchan; ++f) {
                if (anyf[r * nchan + f]) continue;
                double p = real_phase * freq[f];
                double cp = cos(p), sp = sin(p);
                for (int64_t c = 0; c < ncorr; ++c) {
                    int64_t vi = (r * nchan + f) * ncorr + c;
                    double vre, vim;
                    if (vis_complex) {
                        vre = vis[2 * vi];
                        vim = vis[2 * vi + 1];
                    } else {
                        vre = vis[vi];
                        vim = 0.0;
                    }
                    double t = cp * vre - sp * vim; /* kernels.py:141-144 */
                    double *a = o + f * ncorr + c;
                    if (out_f32)
                        *a = (double)(float)(*a + t);
                    else
                        *a += t;
                }
            }
        }
    }
    free(anyf);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* predict_vis : africanus/rime/predict.py:56-373, 574-617                   */
/* ------------------------------------------------------------------------ */
/*
 * Templated over the real component type T (float -> complex64 arithmetic,
 * double -> complex128).  Complex numbers are interleaved (re, im).
 * mode 0: "diagonal" Jones, trailing corr dim (ncorr,) multiplied element-wise
 *         (predict.py:93-98); mode 1: (2,2) matrices, ncorr == 4.
 * Absent terms are passed as NULL.  Order of operations:
 *   source sum, s-major sequential (predict.py:193-252)
 *     2x2: T = X * E2^H first, then out += E1 * T (predict.py:103-117)
 *   out += base_vis (predict.py:329-339)
 *   out = G1 * out * G2^H (predict.py:342-373)
 * ti = time_index[r] - min(time_index) (predict.py:597).
 */
#define CMUL_RE(ar, ai, br, bi) ((ar) * (br) - (ai) * (bi))
#define CMUL_IM(ar, ai, br, bi) ((ar) * (bi) + (ai) * (br))

#define DEFINE_PREDICT(NAME, T)                                                          \
    /* out2x2 (+)= A * (B * C^H) with the reference's association */                     \
    static inline void NAME##_mul3_2x2(const T *a, const T *b, const T *c, T *o,         \
                                       int accumulate) {                                 \
        /* c^H entries: conj of c[0,0], c[0,1], c[1,0], c[1,1] (predict.py:103-106) */   \
        T c00r = c[0], c00i = -c[1], c01r = c[2], c01i = -c[3];                          \
        T c10r = c[4], c10i = -c[5], c11r = c[6], c11i = -c[7];                          \
        /* xx = b00*c00H + b01*c01H ; xy = b00*c10H + b01*c11H ; ... (:108-111) */        \
        T xxr = CMUL_RE(b[0], b[1], c00r, c00i) + CMUL_RE(b[2], b[3], c01r, c01i);       \
        T xxi = CMUL_IM(b[0], b[1], c00r, c00i) + CMUL_IM(b[2], b[3], c01r, c01i);       \
        T xyr = CMUL_RE(b[0], b[1], c10r, c10i) + CMUL_RE(b[2], b[3], c11r, c11i);       \
        T xyi = CMUL_IM(b[0], b[1], c10r, c10i) + CMUL_IM(b[2], b[3], c11r, c11i);       \
        T yxr = CMUL_RE(b[4], b[5], c00r, c00i) + CMUL_RE(b[6], b[7], c01r, c01i);       \
        T yxi = CMUL_IM(b[4], b[5], c00r, c00i) + CMUL_IM(b[6], b[7], c01r, c01i);       \
        T yyr = CMUL_RE(b[4], b[5], c10r, c10i) + CMUL_RE(b[6], b[7], c11r, c11i);       \
        T yyi = CMUL_IM(b[4], b[5], c10r, c10i) + CMUL_IM(b[6], b[7], c11r, c11i);       \
        /* out = a * [[xx,xy],[yx,yy]] (:113-122) */                                     \
        T r0 = CMUL_RE(a[0], a[1], xxr, xxi) + CMUL_RE(a[2], a[3], yxr, yxi);            \
        T i0 = CMUL_IM(a[0], a[1], xxr, xxi) + CMUL_IM(a[2], a[3], yxr, yxi);            \
        T r1 = CMUL_RE(a[0], a[1], xyr, xyi) + CMUL_RE(a[2], a[3], yyr, yyi);            \
        T i1 = CMUL_IM(a[0], a[1], xyr, xyi) + CMUL_IM(a[2], a[3], yyr, yyi);            \
        T r2 = CMUL_RE(a[4], a[5], xxr, xxi) + CMUL_RE(a[6], a[7], yxr, yxi);            \
        T i2 = CMUL_IM(a[4], a[5], xxr, xxi) + CMUL_IM(a[6], a[7], yxr, yxi);            \
        T r3 = CMUL_RE(a[4], a[5], xyr, xyi) + CMUL_RE(a[6], a[7], yyr, yyi);            \
        T i3 = CMUL_IM(a[4], a[5], xyr, xyi) + CMUL_IM(a[6], a[7], yyr, yyi);            \
        if (accumulate) {                                                                \
            o[0] += r0; o[1] += i0; o[2] += r1; o[3] += i1;                              \
            o[4] += r2; o[5] += i2; o[6] += r3; o[7] += i3;                              \
        } else {                                                                         \
            o[0] = r0; o[1] = i0; o[2] = r1; o[3] = i1;                                  \
            o[4] = r2; o[5] = i2; o[6] = r3; o[7] = i3;                                  \
        }                                                                                \
    }                                                                                    \
    /* out2x2 += A * C^H (predict.py:138-148) */                                         \
    static inline void NAME##_mul2_2x2(const T *a, const T *c, T *o) {                   \
        T c00r = c[0], c00i = -c[1], c01r = c[2], c01i = -c[3];                          \
        T c10r = c[4], c10i = -c[5], c11r = c[6], c11i = -c[7];                          \
        o[0] += CMUL_RE(a[0], a[1], c00r, c00i) + CMUL_RE(a[2], a[3], c01r, c01i);       \
        o[1] += CMUL_IM(a[0], a[1], c00r, c00i) + CMUL_IM(a[2], a[3], c01r, c01i);       \
        o[2] += CMUL_RE(a[0], a[1], c10r, c10i) + CMUL_RE(a[2], a[3], c11r, c11i);       \
        o[3] += CMUL_IM(a[0], a[1], c10r, c10i) + CMUL_IM(a[2], a[3], c11r, c11i);       \
        o[4] += CMUL_RE(a[4], a[5], c00r, c00i) + CMUL_RE(a[6], a[7], c01r, c01i);       \
        o[5] += CMUL_IM(a[4], a[5], c00r, c00i) + CMUL_IM(a[6], a[7], c01r, c01i);       \
        o[6] += CMUL_RE(a[4], a[5], c10r, c10i) + CMUL_RE(a[6], a[7], c11r, c11i);       \
        o[7] += CMUL_IM(a[4], a[5], c10r, c10i) + CMUL_IM(a[6], a[7], c11r, c11i);       \
    }                                                                                    \
    int NAME(const int64_t *time_index, const int64_t *ant1, const int64_t *ant2,        \
             const T *dde1, const T *coh, const T *dde2, const T *die1, const T *bvis,   \
             const T *die2, int64_t nsrc, int64_t nrow, int64_t ntime, int64_t nant,     \
             int64_t nchan, int64_t ncorr, int mode, T *out) {                           \
        if (mode == 1 && ncorr != 4) return -1;                                          \
        const int have_dde = dde1 != NULL && dde2 != NULL;                               \
        const int have_coh = coh != NULL;                                                \
        const int have_die = die1 != NULL && die2 != NULL;                               \
        int64_t tmin = nrow > 0 ? time_index[0] : 0;                                     \
        for (int64_t r = 1; r < nrow; ++r)                                               \
            if (time_index[r] < tmin) tmin = time_index[r];                              \
        const int64_t nc2 = 2 * ncorr;                                                   \
        memset(out, 0, sizeof(T) * (size_t)(nrow * nchan * nc2));                        \
        _Pragma("omp parallel for schedule(static)")                                     \
        for (int64_t r = 0; r < nrow; ++r) {                                             \
            int64_t ti = time_index[r] - tmin, a1 = ant1[r], a2 = ant2[r];               \
            for (int64_t s = 0; s < nsrc; ++s) {                                         \
                for (int64_t f = 0; f < nchan; ++f) {                                    \
                    T *o = out + (r * nchan + f) * nc2;                                  \
                    const T *x = have_coh ? coh + ((s * nrow + r) * nchan + f) * nc2     \
                                          : NULL;                                        \
                    const T *e1 = have_dde ? dde1 + (((s * ntime + ti) * nant + a1)      \
                                                     * nchan + f) * nc2 : NULL;          \
                    const T *e2 = have_dde ? dde2 + (((s * ntime + ti) * nant + a2)      \
                                                     * nchan + f) * nc2 : NULL;          \
                    if (have_dde && have_coh) {                                          \
                        if (mode == 1) {                                                 \
                            NAME##_mul3_2x2(e1, x, e2, o, 1);                            \
                        } else {                                                         \
                            for (int64_t c = 0; c < ncorr; ++c) {                        \
                                /* (a1*bl)*conj(a2)  (predict.py:95) */                  \
                                T tr = CMUL_RE(e1[2*c], e1[2*c+1], x[2*c], x[2*c+1]);    \
                                T tm = CMUL_IM(e1[2*c], e1[2*c+1], x[2*c], x[2*c+1]);    \
                                o[2*c]   += CMUL_RE(tr, tm, e2[2*c], -e2[2*c+1]);        \
                                o[2*c+1] += CMUL_IM(tr, tm, e2[2*c], -e2[2*c+1]);        \
                            }                                                            \
                        }                                                                \
                    } else if (have_dde) {                                               \
                        if (mode == 1) {                                                 \
                            NAME##_mul2_2x2(e1, e2, o);                                  \
                        } else {                                                         \
                            for (int64_t c = 0; c < ncorr; ++c) {                        \
                                o[2*c]   += CMUL_RE(e1[2*c], e1[2*c+1], e2[2*c], -e2[2*c+1]); \
                                o[2*c+1] += CMUL_IM(e1[2*c], e1[2*c+1], e2[2*c], -e2[2*c+1]); \
                            }                                                            \
                        }                                                                \
                    } else if (have_coh) {                                               \
                        for (int64_t c = 0; c < nc2; ++c) o[c] += x[c];                  \
                    }                                                                    \
                }                                                                        \
            }                                                                            \
            for (int64_t f = 0; f < nchan; ++f) {                                        \
                T *o = out + (r * nchan + f) * nc2;                                      \
                if (bvis) {                                                              \
                    const T *b = bvis + (r * nchan + f) * nc2;                           \
                    for (int64_t c = 0; c < nc2; ++c) o[c] += b[c];                      \
                }                                                                        \
                if (have_die) {                                                          \
                    const T *g1 = die1 + ((ti * nant + a1) * nchan + f) * nc2;           \
                    const T *g2 = die2 + ((ti * nant + a2) * nchan + f) * nc2;           \
                    if (mode == 1) {                                                     \
                        T tmp[8];                                                        \
                        memcpy(tmp, o, sizeof(tmp));                                     \
                        NAME##_mul3_2x2(g1, tmp, g2, o, 0);                              \
                    } else {                                                             \
                        for (int64_t c = 0; c < ncorr; ++c) {                            \
                            T tr = CMUL_RE(g1[2*c], g1[2*c+1], o[2*c], o[2*c+1]);        \
                            T tm = CMUL_IM(g1[2*c], g1[2*c+1], o[2*c], o[2*c+1]);        \
                            o[2*c]   = CMUL_RE(tr, tm, g2[2*c], -g2[2*c+1]);             \
                            o[2*c+1] = CMUL_IM(tr, tm, g2[2*c], -g2[2*c+1]);             \
                        }                                                                \
                    }                                                                    \
                }                                                                        \
            }                                                                            \
        }                                                                                \
        return 0;                                                                        \
    }

DEFINE_PREDICT(orc_predict_vis_c128, double)
DEFINE_PREDICT(orc_predict_vis_c64, float)

/* ------------------------------------------------------------------------ */
/* freq_grid_interp : africanus/rime/fast_beam_cubes.py:10-54                */
/* ------------------------------------------------------------------------ */
/* freq_data: (nchan, 3) = (scale, lower-weight, lower grid index) */
int orc_freq_grid_interp(const double *freq, const double *beam_freq_map, int64_t nchan,
                         int64_t nud, double *freq_data) {
    for (int64_t f = 0; f < nchan; ++f) {
        double nu = freq[f];
        int64_t lo = 0, hi = nud - 1;
        while (lo <= hi) { /* binary search (:20-30) */
            int64_t mid = lo + (hi - lo) / 2;
            double bf = beam_freq_map[mid];
            if (bf < nu)
                lo = mid + 1;
            else if (bf > nu)
                hi = mid - 1;
            else {
                lo = mid;
                break;
            }
        }
        if (hi < lo) lo = hi; /* lower = min(lower, upper) (:33) */
        hi = lo + 1;
        double *fd = freq_data + 3 * f;
        if (lo == -1) { /* below the cube (:37-40) */
            fd[0] = nu / beam_freq_map[0];
            fd[1] = 1.0;
            fd[2] = 0.0;
        } else if (hi == nud) { /* at/above the top (:41-44) */
            fd[0] = nu / beam_freq_map[nud - 1];
            fd[1] = 0.0;
            fd[2] = (double)(nud - 2);
        } else { /* inside (:45-52) */
            fd[0] = 1.0;
            double flo = beam_freq_map[lo], fhi = beam_freq_map[hi];
            fd[1] = (fhi - nu) / (fhi - flo);
            fd[2] = (double)lo;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* beam_cube_dde : africanus/rime/fast_beam_cubes.py:57-240                  */
/* ------------------------------------------------------------------------ */
/*
 * All real inputs float64.  beam: (lw, mh, nud, ncorr) complex (T components),
 * out: (nsrc, ntime, nant, nchan, ncorr) complex T.  For T = float the
 * accumulators are stored as float32 after each update, as the reference's
 * complex64 / float32 scratch arrays are (:106-108), while weights stay float64.
 */
#define DEFINE_BEAM(NAME, T, HYPOT)                                                      \
    int NAME(const T *beam, const double *ext, const double *beam_freq_map,              \
             const double *lm, const double *pa, const double *perr,                     \
             const double *ascale, const double *freq, int64_t lw, int64_t mh,           \
             int64_t nud, int64_t ncorr, int64_t nsrc, int64_t ntime, int64_t nant,      \
             int64_t nchan, T *out) {                                                    \
        if (lw < 2 || mh < 2 || nud < 2) return -1; /* :74-75 */                         \
        if (ncorr > 16) return -3;                                                       \
        double lower_l = ext[0], upper_l = ext[1], lower_m = ext[2], upper_m = ext[3];   \
        double lmaxf = (double)(lw - 1), mmaxf = (double)(mh - 1);                       \
        int64_t lmaxi = lw - 1, mmaxi = mh - 1;                                          \
        double lscale = lmaxf / (upper_l - lower_l);                                     \
        double mscale = mmaxf / (upper_m - lower_m);                                     \
        double *fd = (double *)malloc(sizeof(double) * 3 * (size_t)(nchan > 0 ? nchan : 1)); \
        if (!fd) return -2;                                                              \
        orc_freq_grid_interp(freq, beam_freq_map, nchan, nud, fd);                       \
        _Pragma("omp parallel for schedule(static) collapse(2)")                         \
        for (int64_t t = 0; t < ntime; ++t) {                                            \
            for (int64_t a = 0; a < nant; ++a) {                                         \
                double sin_pa = sin(pa[t * nant + a]);                                   \
                double cos_pa = cos(pa[t * nant + a]);                                   \
                for (int64_t s = 0; s < nsrc; ++s) {                                     \
                    double l = lm[2 * s], m = lm[2 * s + 1];                             \
                    for (int64_t f = 0; f < nchan; ++f) {                                \
                        double fscale = fd[3 * f], nudw = fd[3 * f + 1];                 \
                        double inv_nud = 1.0 - nudw;                                     \
                        int64_t gc0 = (int64_t)(int32_t)fd[3 * f + 2], gc1 = gc0 + 1;    \
                        double sl = l * fscale, sm = m * fscale;          /* :130-131 */ \
                        const double *pe = perr + ((t * nant + a) * nchan + f) * 2;      \
                        double tl = sl + pe[0], tm = sm + pe[1];          /* :134-135 */ \
                        double vl = tl * cos_pa - tm * sin_pa;            /* :138-139 */ \
                        double vm = tl * sin_pa + tm * cos_pa;                           \
                        const double *as = ascale + (a * nchan + f) * 2;                 \
                        vl *= as[0];                                      /* :142-143 */ \
                        vm *= as[1];                                                     \
                        vl = lscale * (vl - lower_l);                     /* :146-147 */ \
                        vm = mscale * (vm - lower_m);                                    \
                        vl = fmax(0.0, fmin(vl, lmaxf));                  /* :150-151 */ \
                        vm = fmax(0.0, fmin(vm, mmaxf));                                 \
                        int64_t gl0 = (int64_t)(int32_t)floor(vl);        /* :154-155 */ \
                        int64_t gm0 = (int64_t)(int32_t)floor(vm);                       \
                        int64_t gl1 = gl0 + 1 < lmaxi ? gl0 + 1 : lmaxi;  /* :158-159 */ \
                        int64_t gm1 = gm0 + 1 < mmaxi ? gm0 + 1 : mmaxi;                 \
                        double ld = vl - (double)gl0, md = vm - (double)gm0;             \
                        T csr[16], csi[16], asum[16];                                    \
                        for (int64_t c = 0; c < ncorr; ++c) csr[c] = csi[c] = asum[c] = 0; \
                        /* 8 corners in the reference's order (:169-225) */              \
                        const int64_t gls[8] = {gl0, gl1, gl0, gl1, gl0, gl1, gl0, gl1}; \
                        const int64_t gms[8] = {gm0, gm0, gm1, gm1, gm0, gm0, gm1, gm1}; \
                        const int64_t gcs[8] = {gc0, gc0, gc0, gc0, gc1, gc1, gc1, gc1}; \
                        const double wts[8] = {                                          \
                            (1.0 - ld) * (1.0 - md) * nudw, ld * (1.0 - md) * nudw,      \
                            (1.0 - ld) * md * nudw, ld * md * nudw,                      \
                            (1.0 - ld) * (1.0 - md) * inv_nud, ld * (1.0 - md) * inv_nud,\
                            (1.0 - ld) * md * inv_nud, ld * md * inv_nud};               \
                        for (int k = 0; k < 8; ++k) {                                    \
                            const T *b = beam + (((gls[k] * mh + gms[k]) * nud + gcs[k]) \
                                                 * ncorr) * 2;                           \
                            double wt = wts[k];                                          \
                            for (int64_t c = 0; c < ncorr; ++c) {                        \
                                T br = b[2 * c], bi = b[2 * c + 1];                      \
                                T ab = HYPOT(br, bi);                                    \
                                asum[c] = (T)((double)asum[c] + wt * (double)ab);        \
                                csr[c] = (T)((double)csr[c] + wt * (double)br);          \
                                csi[c] = (T)((double)csi[c] + wt * (double)bi);          \
                            }                                                            \
                        }                                                                \
                        T *o = out + ((((s * ntime + t) * nant + a) * nchan + f)         \
                                      * ncorr) * 2;                                      \
                        for (int64_t c = 0; c < ncorr; ++c) {             /* :227-238 */ \
                            T div = HYPOT(csr[c], csi[c]);                               \
                            T k = (div == (T)0) ? asum[c] : asum[c] / div;               \
                            /* complex * real promoted to complex: (a+bi)(k+0i) */       \
                            T zero = (T)0;                                               \
                            o[2 * c] = csr[c] * k - csi[c] * zero;                       \
                            o[2 * c + 1] = csr[c] * zero + csi[c] * k;                   \
                        }                                                                \
                    }                                                                    \
                }                                                                        \
            }                                                                            \
        }                                                                                \
        free(fd);                                                                        \
        return 0;                                                                        \
    }

DEFINE_BEAM(orc_beam_cube_dde_c128, double, hypot)
DEFINE_BEAM(orc_beam_cube_dde_c64, float, hypotf)

/* ------------------------------------------------------------------------ */
/* composed (un-fused) point/full predict                                     */
/* africanus/rime/examples/predict.py:107-134,490,522-527 ; the recipe is      */
/* asserted in africanus/experimental/rime/fused/tests/test_rime.py:175-209    */
/* ------------------------------------------------------------------------ */
/*
 * V[r,f] = G1 (B[r,f] + sum_s E1 (K[s,r,f] * Bright[s,f]) E2^H) G2^H without
 * materialising K or the (s,r,f,corr) coherency: per (s,r,f) the phasor is
 * computed as phase_delay does (clamped n, float64), multiplied into the
 * brightness (the einsum "srf,sfij->srfij"), then fed through the predict_vis
 * chain above.  complex128 only.
 */
int orc_fused_predict_c128(const double *lm, const double *uvw, const double *freq,
                           const double *bright, const int64_t *time_index,
                           const int64_t *ant1, const int64_t *ant2, const double *dde1,
                           const double *dde2, const double *die1, const double *bvis,
                           const double *die2, int64_t nsrc, int64_t nrow, int64_t ntime,
                           int64_t nant, int64_t nchan, int64_t ncorr, int mode, int sign,
                           double *out) {
    if (mode == 1 && ncorr != 4) return -1;
    if (ncorr > 16) return -3;
    double cst = -two_pi_over_c();
    if (sign < 0) cst = -cst;
    const int have_dde = dde1 != NULL && dde2 != NULL;
    const int have_die = die1 != NULL && die2 != NULL;
    int64_t tmin = nrow > 0 ? time_index[0] : 0;
    for (int64_t r = 1; r < nrow; ++r)
        if (time_index[r] < tmin) tmin = time_index[r];
    const int64_t nc2 = 2 * ncorr;
    double *nn = (double *)malloc(sizeof(double) * (size_t)(nsrc > 0 ? nsrc : 1));
    if (!nn) return -2;
    for (int64_t s = 0; s < nsrc; ++s) {
        double l = lm[2 * s], m = lm[2 * s + 1];
        double n = 1.0 - l * l - m * m;
        nn[s] = sqrt(n < 0.0 ? 0.0 : n) - 1.0;
    }
    memset(out, 0, sizeof(double) * (size_t)(nrow * nchan * nc2));

#pragma omp parallel for schedule(static)
    for (int64_t r = 0; r < nrow; ++r) {
        int64_t ti = time_index[r] - tmin, a1 = ant1[r], a2 = ant2[r];
        double u = uvw[3 * r], v = uvw[3 * r + 1], w = uvw[3 * r + 2];
        for (int64_t s = 0; s < nsrc; ++s) {
            double l = lm[2 * s], m = lm[2 * s + 1], n = nn[s];
            double real_phase = cst * (l * u + m * v + n * w);
            for (int64_t f = 0; f < nchan; ++f) {
                double p = real_phase * freq[f];
                double kr = cos(p), ki = sin(p);
                const double *b = bright + (s * nchan + f) * nc2;
                double x[32];
                for (int64_t c = 0; c < ncorr; ++c) { /* K * brightness */
                    x[2 * c] = CMUL_RE(kr, ki, b[2 * c], b[2 * c + 1]);
                    x[2 * c + 1] = CMUL_IM(kr, ki, b[2 * c], b[2 * c + 1]);
                }
                double *o = out + (r * nchan + f) * nc2;
                if (have_dde) {
                    const double *e1 = dde1 + (((s * ntime + ti) * nant + a1) * nchan + f) * nc2;
                    const double *e2 = dde2 + (((s * ntime + ti) * nant + a2) * nchan + f) * nc2;
                    if (mode == 1) {
                        orc_predict_vis_c128_mul3_2x2(e1, x, e2, o, 1);
                    } else {
                        for (int64_t c = 0; c < ncorr; ++c) {
                            double tr = CMUL_RE(e1[2*c], e1[2*c+1], x[2*c], x[2*c+1]);
                            double tm = CMUL_IM(e1[2*c], e1[2*c+1], x[2*c], x[2*c+1]);
                            o[2*c]   += CMUL_RE(tr, tm, e2[2*c], -e2[2*c+1]);
                            o[2*c+1] += CMUL_IM(tr, tm, e2[2*c], -e2[2*c+1]);
                        }
                    }
                } else {
                    for (int64_t c = 0; c < nc2; ++c) o[c] += x[c];
                }
            }
        }
        for (int64_t f = 0; f < nchan; ++f) {
            double *o = out + (r * nchan + f) * nc2;
            if (bvis) {
                const double *b = bvis + (r * nchan + f) * nc2;
                for (int64_t c = 0; c < nc2; ++c) o[c] += b[c];
            }
            if (have_die) {
                const double *g1 = die1 + ((ti * nant + a1) * nchan + f) * nc2;
                const double *g2 = die2 + ((ti * nant + a2) * nchan + f) * nc2;
                if (mode == 1) {
                    double tmp[8];
                    memcpy(tmp, o, sizeof(tmp));
                    orc_predict_vis_c128_mul3_2x2(g1, tmp, g2, o, 0);
                } else {
                    for (int64_t c = 0; c < ncorr; ++c) {
                        double tr = CMUL_RE(g1[2*c], g1[2*c+1], o[2*c], o[2*c+1]);
                        double tm = CMUL_IM(g1[2*c], g1[2*c+1], o[2*c], o[2*c+1]);
                        o[2*c]   = CMUL_RE(tr, tm, g2[2*c], -g2[2*c+1]);
                        o[2*c+1] = CMUL_IM(tr, tm, g2[2*c], -g2[2*c+1]);
                    }
                }
            }
        }
    }
    free(nn);
    return 0;
}

/* ------------------------------------------------------------------------ */
/* WSClean component model: africanus/model/wsclean/spec_model.py:76-124 and */
/* africanus/rime/wsclean_predict.py:11-83                                   */
/* ------------------------------------------------------------------------ */
/*
 * spectra (spec_model.py:94-124): out (nsrc, nchan) float64.
 *   log_poly[s] != 0 : out = I * exp(sum_c coeffs[s,c] * log(nu/rf)^(c+1))
 *   else             : out = I + sum_c coeffs[s,c] * (nu/rf - 1)^(c+1)
 * x ** (c+1) with an integer exponent is evaluated by numba's int_power
 * (numba/cpython/numbers.py): binary square-and-multiply, low bit first.  It
 * differs from a left-to-right product from the 4th power on (checked against
 * the reference with 6 coefficients, tests/golden/brightness.npz).
 */
static double ipow_(double x, int n) {
    double r = 1.0, a = x;
    while (n != 0) {
        if (n & 1) r *= a;
        n >>= 1;
        a *= a;
    }
    return r;
}

int orc_wsclean_spectra(const double *flux, const double *coeffs, const uint8_t *log_poly,
                        const double *ref_freq, const double *freq, int64_t nsrc,
                        int64_t ncoeffs, int64_t nchan, double *out) {
    for (int64_t s = 0; s < nsrc; ++s) {
        const double rf = ref_freq[s];
        for (int64_t f = 0; f < nchan; ++f) {
            const double nu = freq[f];
            double acc;
            if (log_poly[s]) {
                acc = 0.0;
                for (int64_t c = 0; c < ncoeffs; ++c)
                    acc += coeffs[s * ncoeffs + c] * ipow_(log(nu / rf), (int)(c + 1));
                acc = flux[s] * exp(acc);
            } else {
                acc = flux[s];
                for (int64_t c = 0; c < ncoeffs; ++c) {
                    double term = coeffs[s * ncoeffs + c];
                    term *= ipow_((nu / rf) - 1.0, (int)(c + 1));
                    acc += term;
                }
            }
            out[s * nchan + f] = acc;
        }
    }
    return 0;
}

/*
 * wsclean_predict_main (wsclean_predict.py:11-83): vis (nrow, nchan, 1)
 * complex128.  Phase sign +2pi/c (the "casa" convention), n without clamp;
 * GAUSSIAN sources multiply every term by exp(-(fu1^2 + fv1^2)) with
 * (emaj, emin, angle) -> el, em, er as in :48-52 and scaled_freq = nu * gauss_scale.
 * Sources are accumulated in index order.
 */
int orc_wsclean_predict(const double *uvw, const double *lm, const uint8_t *is_gauss,
                        const double *gauss_shape, const double *freq, const double *spectrum,
                        int64_t nsrc, int64_t nrow, int64_t nchan, double *out) {
    const double fwhm = 2.0 * sqrt(2.0 * log(2.0));
    const double fwhminv = 1.0 / fwhm;
    const double gauss_scale = fwhminv * sqrt(2.0) * 3.141592653589793 / 2.99792458e8;
    const double tpc = two_pi_over_c();
    memset(out, 0, sizeof(double) * 2 * (size_t)(nrow * nchan));
    for (int64_t s = 0; s < nsrc; ++s) {
        const double l = lm[2 * s], m = lm[2 * s + 1];
        const double n = sqrt(1.0 - l * l - m * m) - 1.0;
        double el = 0, em = 0, er = 0;
        if (is_gauss[s]) {
            const double emaj = gauss_shape[3 * s], emin = gauss_shape[3 * s + 1],
                         angle = gauss_shape[3 * s + 2];
            el = emaj * sin(angle);
            em = emaj * cos(angle);
            er = emin / (emaj == 0.0 ? 1.0 : emaj);
        }
        for (int64_t r = 0; r < nrow; ++r) {
            const double u = uvw[3 * r], v = uvw[3 * r + 1], w = uvw[3 * r + 2];
            const double real_phase = tpc * (u * l + v * m + w * n);
            const double u1 = (u * em - v * el) * er;
            const double v1 = u * el + v * em;
            for (int64_t f = 0; f < nchan; ++f) {
                const double p = real_phase * freq[f];
                double re = cos(p) * spectrum[s * nchan + f];
                double im = sin(p) * spectrum[s * nchan + f];
                if (is_gauss[s]) {
                    const double sf = freq[f] * gauss_scale;
                    const double fu1 = u1 * sf, fv1 = v1 * sf;
                    const double shape = exp(-(fu1 * fu1 + fv1 * fv1));
                    re *= shape;
                    im *= shape;
                }
                out[2 * (r * nchan + f)] += re;
                out[2 * (r * nchan + f) + 1] += im;
            }
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* Brightness from Stokes parameters:                       */
/* africanus/model/spectral/spec_model.py:106-211 and                        */
/* africanus/model/coherency/conversion.py:17-48,218-243                     */
/* ------------------------------------------------------------------------ */
/*
 * spectral_model (spec_model.py:160-211): out (nsrc, nchan, npol) float64.
 * base[p]: 0 "std"   out = stokes * prod_i (nu/rf) ** spi[s,i,p]      (libm pow)
 *          1 "log"   out = stokes * exp  (sum_i spi[s,i,p] * log  (nu/rf) ** (i+1))
 *          2 "log10" out = stokes * 10 ** (sum_i spi[s,i,p] * log10(nu/rf) ** (i+1))
 * Integer powers as numba's int_power (ipow_ above); 10 ** x is pow(10.0, x).
 */
int orc_spectral_model(const double *stokes, const double *spi, const double *ref_freq,
                       const double *freq, const int *base, int64_t nsrc, int64_t nspi,
                       int64_t npol, int64_t nchan, double *out) {
    for (int64_t p = 0; p < npol; ++p) {
        if (base[p] < 0 || base[p] > 2) return 1;
        for (int64_t s = 0; s < nsrc; ++s) {
            const double rf = ref_freq[s];
            const double st = stokes[s * npol + p];
            for (int64_t f = 0; f < nchan; ++f) {
                double v;
                if (base[p] == 0) {
                    const double ratio = freq[f] / rf;
                    v = st;
                    for (int64_t i = 0; i < nspi; ++i) v *= pow(ratio, spi[(s * nspi + i) * npol + p]);
                } else {
                    const double lr = base[p] == 1 ? log(freq[f] / rf) : log10(freq[f] / rf);
                    double acc = 0.0;
                    for (int64_t i = 0; i < nspi; ++i)
                        acc += spi[(s * nspi + i) * npol + p] * ipow_(lr, (int)(i + 1));
                    v = st * (base[p] == 1 ? exp(acc) : pow(10.0, acc));
                }
                out[(s * nchan + f) * npol + p] = v;
            }
        }
    }
    return 0;
}

/*
 * convert (conversion.py:218-243) after the schema has been resolved to one
 * (source_one, source_two, op) triple per output element (conversion.py:145-215).
 * in (n, nin) complex128 (real inputs widened), out (n, nout) complex128.
 * A source index of -1 is the DataSource.Default zero.  ops, conversion.py:19-48:
 *   0: a + b        (RR, XX)          1: a - b          (LL, YY)
 *   2: a + b*1j     (RL, XY)          3: a - b*1j       (LR, YX)
 *   4: (a + b) / 2  (I, Q)            5: (a - b) / 2    (Q, V)
 *   6: (a - b) / 2j (U, V)
 */
int orc_convert(const double *in, int64_t n, int64_t nin, const int *src1, const int *src2,
                const int *op, int64_t nout, double *out) {
    for (int64_t i = 0; i < n; ++i) {
        for (int64_t o = 0; o < nout; ++o) {
            double ar = 0, ai = 0, br = 0, bi = 0, re, im;
            if (src1[o] >= 0) { ar = in[2 * (i * nin + src1[o])]; ai = in[2 * (i * nin + src1[o]) + 1]; }
            if (src2[o] >= 0) { br = in[2 * (i * nin + src2[o])]; bi = in[2 * (i * nin + src2[o]) + 1]; }
            switch (op[o]) {
            case 0: re = ar + br; im = ai + bi; break;
            case 1: re = ar - br; im = ai - bi; break;
            case 2: re = ar - bi; im = ai + br; break;
            case 3: re = ar + bi; im = ai - br; break;
            case 4: re = (ar + br) * 0.5; im = (ai + bi) * 0.5; break;
            case 5: re = (ar - br) * 0.5; im = (ai - bi) * 0.5; break;
            case 6: re = (ai - bi) * 0.5; im = -((ar - br) * 0.5); break;
            default: return 1;
            }
            out[2 * (i * nout + o)] = re;
            out[2 * (i * nout + o) + 1] = im;
        }
    }
    return 0;
}

/* ------------------------------------------------------------------------ */
/* feed_rotation: africanus/rime/feeds.py:13-48.  out (n,2,2) complex.       */
/* feed_type 0 linear [[cos, sin], [-sin, cos]]; 1 circular diag(e^-ipa, e^+ipa) */
/* ------------------------------------------------------------------------ */
int orc_feed_rotation_f64(const double *pa, int64_t n, int feed_type, double *out) {
    if (feed_type != 0 && feed_type != 1) return 1;
    for (int64_t i = 0; i < n; ++i) {
        const double c = cos(pa[i]), s = sin(pa[i]);
        double *o = out + 8 * i;
        if (feed_type == 0) {
            o[0] = c; o[1] = 0.0; o[2] = s; o[3] = 0.0; o[4] = -s; o[5] = 0.0; o[6] = c; o[7] = 0.0;
        } else {
            o[0] = c; o[1] = -s; o[2] = 0.0; o[3] = 0.0; o[4] = 0.0; o[5] = 0.0; o[6] = c; o[7] = s;
        }
    }
    return 0;
}

int orc_feed_rotation_f32(const float *pa, int64_t n, int feed_type, float *out) {
    if (feed_type != 0 && feed_type != 1) return 1;
    for (int64_t i = 0; i < n; ++i) {
        const float c = cosf(pa[i]), s = sinf(pa[i]);
        float *o = out + 8 * i;
        if (feed_type == 0) {
            o[0] = c; o[1] = 0.0f; o[2] = s; o[3] = 0.0f; o[4] = -s; o[5] = 0.0f; o[6] = c; o[7] = 0.0f;
        } else {
            o[0] = c; o[1] = -s; o[2] = 0.0f; o[3] = 0.0f; o[4] = 0.0f; o[5] = 0.0f; o[6] = c; o[7] = s;
        }
    }
    return 0;
}
