// Source brightness from Stokes parameters, sm_100a:
//   africanus.model.spectral.spectral_model   africanus/model/spectral/spec_model.py:106-211
//   africanus.model.coherency.convert         africanus/model/coherency/conversion.py:17-48,145-243
// and their composition stokes -> (source, chan, corr) complex brightness in one kernel, so the
// predict's per-source input is (stokes, spi, ref_freq): O(10) numbers instead of nchan*4 complex.
//
// Bound: HBM store of the result (8*npol B per element for the spectral model, 16*nout B for
// the brightness); the transcendental work (nspi pow, or one log + one exp, per element) hides
// behind it.  The brightness is (source, chan) work while the predict that consumes it is
// (source, row, chan) work, so this is never the dominant kernel.
#include <math.h>

#include "afr_common.cuh"

namespace afr {
namespace {

constexpr int kMaxPol = 16;  // polarisations per spectral_model call / elements per schema

struct Bases {
    int b[kMaxPol];
};
struct Mapping {  // one (source_one, source_two, op) triple per output element
    int s1[kMaxPol], s2[kMaxPol], op[kMaxPol];
};

// x ** n, n a positive integer, in numba's int_power order (numba/cpython/numbers.py):
// binary square-and-multiply starting at the low bit
__device__ __forceinline__ double ipow(double x, int n) {
    double r = 1.0, a = x;
    while (n != 0) {
        if (n & 1) r = __dmul_rn(r, a);
        n >>= 1;
        a = __dmul_rn(a, a);
    }
    return r;
}

// spec_model.py:181-208 for one (source, chan, pol)
__device__ __forceinline__ double spectral_value(int base, double st, const double *spi_sp, long long pol_stride,
                                                 int nspi, double nu, double rf) {
    if (base == 0) {
        const double ratio = nu / rf;
        double v = st;
        for (int i = 0; i < nspi; ++i) v = __dmul_rn(v, pow(ratio, spi_sp[i * pol_stride]));
        return v;
    }
    const double lr = base == 1 ? log(nu / rf) : log10(nu / rf);
    double acc = 0.0;
    for (int i = 0; i < nspi; ++i) acc = __dadd_rn(acc, __dmul_rn(spi_sp[i * pol_stride], ipow(lr, i + 1)));
    return __dmul_rn(st, base == 1 ? exp(acc) : pow(10.0, acc));
}

// out (nsrc, nchan, npol): thread per element, pol fastest (coalesced stores)
__global__ void spectral_model_kernel(const double *__restrict__ stokes, const double *__restrict__ spi,
                                      const double *__restrict__ ref_freq, const double *__restrict__ freq,
                                      Bases bases, long long nsrc, int nspi, int npol, long long nchan,
                                      double *__restrict__ out) {
    const long long total = nsrc * nchan * npol;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int p = (int)(i % npol);
        const long long sf = i / npol;
        const long long s = sf / nchan, f = sf - s * nchan;
        out[i] = spectral_value(bases.b[p], stokes[s * npol + p], spi + (s * nspi) * npol + p, npol, nspi,
                                freq[f], ref_freq[s]);
    }
}

// conversion.py:19-48 on one pair of (possibly defaulted) inputs
__device__ __forceinline__ double2 convert_op(int op, double2 a, double2 b) {
    switch (op) {
    case 0: return make_double2(__dadd_rn(a.x, b.x), __dadd_rn(a.y, b.y));                      // a + b
    case 1: return make_double2(__dsub_rn(a.x, b.x), __dsub_rn(a.y, b.y));                      // a - b
    case 2: return make_double2(__dsub_rn(a.x, b.y), __dadd_rn(a.y, b.x));                      // a + b*1j
    case 3: return make_double2(__dadd_rn(a.x, b.y), __dsub_rn(a.y, b.x));                      // a - b*1j
    case 4: return make_double2(0.5 * __dadd_rn(a.x, b.x), 0.5 * __dadd_rn(a.y, b.y));          // (a + b)/2
    case 5: return make_double2(0.5 * __dsub_rn(a.x, b.x), 0.5 * __dsub_rn(a.y, b.y));          // (a - b)/2
    default: return make_double2(0.5 * __dsub_rn(a.y, b.y), -(0.5 * __dsub_rn(a.x, b.x)));      // (a - b)/2j
    }
}

// out (n, nout) complex128 from in (n, nin) real or complex float64
template <bool IN_COMPLEX>
__global__ void convert_kernel(const double *__restrict__ in, long long n, int nin, Mapping map, int nout,
                               double2 *__restrict__ out) {
    const long long total = n * nout;
    for (long long i = blockIdx.x * (long long)blockDim.x + threadIdx.x; i < total;
         i += (long long)gridDim.x * blockDim.x) {
        const int o = (int)(i % nout);
        const long long e = i / nout;
        double2 a = make_double2(0.0, 0.0), b = a;
        const int s1 = map.s1[o], s2 = map.s2[o];
        if (IN_COMPLEX) {
            const double2 *row = reinterpret_cast<const double2 *>(in) + e * nin;
            if (s1 >= 0) a = row[s1];
            if (s2 >= 0) b = row[s2];
        } else {
            const double *row = in + e * nin;
            if (s1 >= 0) a.x = row[s1];
            if (s2 >= 0) b.x = row[s2];
        }
        out[i] = convert_op(map.op[o], a, b);
    }
}

// spectral_value with the logarithm of nu/rf hoisted out of the polarisation loop: the kernel below
// is thread <-> (source, chan), and pow(ratio, spi) per (polarisation, index) made it FP64-pipe bound
// (4 pow = ~600 FP64 instructions per 64 bytes stored: 1.36 TB/s).  x**y = exp(y*ln x) with one ln per
// thread costs a quarter of that; |y*ln x| stays O(1) for spectral indices, so the result is within a
// few ulp of pow's (the parity gate is 1e-10).  Non-positive or non-finite ratios keep pow's own
// special cases.
__device__ __forceinline__ double spectral_value_ln(int base, double st, const double *spi_sp, long long pol_stride,
                                                    int nspi, double ratio, double ln_ratio, double lg_ratio,
                                                    bool regular) {
    if (base == 0) {
        double v = st;
        if (regular)
            for (int i = 0; i < nspi; ++i) v = __dmul_rn(v, exp(__dmul_rn(spi_sp[i * pol_stride], ln_ratio)));
        else
            for (int i = 0; i < nspi; ++i) v = __dmul_rn(v, pow(ratio, spi_sp[i * pol_stride]));
        return v;
    }
    const double lr = base == 1 ? ln_ratio : lg_ratio;
    double acc = 0.0;
    for (int i = 0; i < nspi; ++i) acc = __dadd_rn(acc, __dmul_rn(spi_sp[i * pol_stride], ipow(lr, i + 1)));
    return __dmul_rn(st, base == 1 ? exp(acc) : pow(10.0, acc));
}

// stokes (nsrc, npol) -> brightness (nsrc, nchan, nout) complex128 (or complex64): the spectral
// model of each polarisation in registers, then the schema mapping.  Thread per (source, chan);
// a CTA's 256 x nout results are staged in shared memory (STAGED: nout <= 4) and leave as
// contiguous 16-byte stores -- per-thread stores would be nout segments at a 16*nout-byte stride.
constexpr int kSbThreads = 256;
template <typename OUT2, bool STAGED>
__global__ void __launch_bounds__(kSbThreads)
stokes_brightness_kernel(const double *__restrict__ stokes, const double *__restrict__ spi,
                         const double *__restrict__ ref_freq, const double *__restrict__ freq, Bases bases,
                         Mapping map, long long nsrc, int nspi, int npol, long long nchan, int nout,
                         int need_lg, OUT2 *__restrict__ out) {
    __shared__ OUT2 stage[STAGED ? kSbThreads * 4 : 1];
    const long long total = nsrc * nchan;
    const long long nblk = (total + kSbThreads - 1) / kSbThreads;
    for (long long blk = blockIdx.x; blk < nblk; blk += gridDim.x) {
        const long long i0 = blk * kSbThreads, i = i0 + threadIdx.x;
        if (i < total) {
            const long long s = i / nchan, f = i - s * nchan;
            const double nu = freq[f], rf = ref_freq[s];
            const double ratio = nu / rf;
            const bool regular = ratio > 0.0 && ratio < 1.7976931348623157e308;
            const double ln_ratio = log(ratio);
            const double lg_ratio = need_lg ? log10(ratio) : 0.0;
            double sm[4];
#pragma unroll
            for (int p = 0; p < 4; ++p)
                sm[p] = p < npol ? spectral_value_ln(bases.b[p], stokes[s * npol + p], spi + (s * nspi) * npol + p,
                                                     npol, nspi, ratio, ln_ratio, lg_ratio, regular)
                                 : 0.0;
            for (int o = 0; o < nout; ++o) {
                const int s1 = map.s1[o], s2 = map.s2[o];
                double2 a = make_double2(0.0, 0.0), b = a;
#pragma unroll
                for (int p = 0; p < 4; ++p) {  // register select (no local-memory indexing)
                    if (s1 == p) a.x = sm[p];
                    if (s2 == p) b.x = sm[p];
                }
                const double2 v = convert_op(map.op[o], a, b);
                OUT2 w;
                w.x = v.x;
                w.y = v.y;
                if (STAGED)
                    stage[threadIdx.x * nout + o] = w;
                else
                    out[i * nout + o] = w;
            }
        }
        if (STAGED) {
            __syncthreads();
            const long long left = total - i0;
            const int n = (int)(left < kSbThreads ? left : kSbThreads) * nout;
            OUT2 *dst = out + i0 * nout;
            for (int j = threadIdx.x; j < n; j += kSbThreads) dst[j] = stage[j];
            __syncthreads();
        }
    }
}

inline unsigned grid_for(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = 16LL * sm_count();
    return (unsigned)(blocks < cap ? (blocks > 0 ? blocks : 1) : cap);
}

int fill_bases(const int *base, int64_t npol, Bases &b) {
    AFR_REQUIRE(npol >= 1 && npol <= kMaxPol, "spectral model: 1 <= npol <= 16 required");
    for (int p = 0; p < kMaxPol; ++p) b.b[p] = 0;
    for (int64_t p = 0; p < npol; ++p) {
        AFR_REQUIRE(base[p] >= 0 && base[p] <= 2, "Invalid base");
        b.b[p] = base[p];
    }
    return 0;
}

int fill_mapping(const int *src1, const int *src2, const int *op, int64_t nin, int64_t nout, Mapping &m) {
    AFR_REQUIRE(nout >= 1 && nout <= kMaxPol, "convert: 1 <= output schema elements <= 16 required");
    for (int o = 0; o < kMaxPol; ++o) m.s1[o] = m.s2[o] = -1, m.op[o] = 0;
    for (int64_t o = 0; o < nout; ++o) {
        AFR_REQUIRE(src1[o] >= -1 && src1[o] < nin && src2[o] >= -1 && src2[o] < nin,
                    "convert: input index outside the input schema");
        AFR_REQUIRE(op[o] >= 0 && op[o] <= 6, "convert: unknown operation");
        m.s1[o] = src1[o];
        m.s2[o] = src2[o];
        m.op[o] = op[o];
    }
    return 0;
}

}  // namespace
}  // namespace afr

using namespace afr;

extern "C" int afr_spectral_model(const double *stokes, const double *spi, const double *ref_freq,
                                  const double *freq, const int *base, int64_t nsrc, int64_t nspi,
                                  int64_t npol, int64_t nchan, double *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(nsrc >= 0 && nspi >= 0 && nchan >= 0, "negative extent");
    Bases b;
    if (int rc = fill_bases(base, npol, b)) return rc;
    const long long total = nsrc * nchan * npol;
    if (total == 0) return 0;
    spectral_model_kernel<<<grid_for(total), 256, 0, stream>>>(stokes, spi, ref_freq, freq, b, nsrc, (int)nspi,
                                                               (int)npol, nchan, out);
    AFR_LAUNCH_OK();
    return 0;
}

extern "C" int afr_convert(const void *in, int in_complex, int64_t n, int64_t nin, const int *src1,
                           const int *src2, const int *op, int64_t nout, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(n >= 0 && nin >= 1, "convert: bad extents");
    Mapping m;
    if (int rc = fill_mapping(src1, src2, op, nin, nout, m)) return rc;
    const long long total = n * nout;
    if (total == 0) return 0;
    if (in_complex)
        convert_kernel<true><<<grid_for(total), 256, 0, stream>>>((const double *)in, n, (int)nin, m, (int)nout,
                                                                  (double2 *)out);
    else
        convert_kernel<false><<<grid_for(total), 256, 0, stream>>>((const double *)in, n, (int)nin, m, (int)nout,
                                                                   (double2 *)out);
    AFR_LAUNCH_OK();
    return 0;
}

extern "C" int afr_stokes_brightness(const double *stokes, const double *spi, const double *ref_freq,
                                     const double *freq, const int *base, int64_t nsrc, int64_t nspi,
                                     int64_t npol, int64_t nchan, const int *src1, const int *src2,
                                     const int *op, int64_t nout, int is_c64, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(nsrc >= 0 && nspi >= 0 && nchan >= 0, "negative extent");
    AFR_REQUIRE(npol >= 1 && npol <= 4, "stokes brightness: 1 <= npol <= 4 required");
    Bases b;
    Mapping m;
    if (int rc = fill_bases(base, npol, b)) return rc;
    if (int rc = fill_mapping(src1, src2, op, npol, nout, m)) return rc;
    const long long total = nsrc * nchan;
    if (total == 0) return 0;
    int need_lg = 0;
    for (int64_t p = 0; p < npol; ++p) need_lg |= base[p] == 2;
    const long long nblk = (total + kSbThreads - 1) / kSbThreads;
    const long long cap = 8LL * sm_count();
    const unsigned grid = (unsigned)(nblk < cap ? nblk : cap);
#define AFR_SB_LAUNCH(T, STAGED)                                                                             \
    stokes_brightness_kernel<T, STAGED><<<grid, kSbThreads, 0, stream>>>(                                    \
        stokes, spi, ref_freq, freq, b, m, nsrc, (int)nspi, (int)npol, nchan, (int)nout, need_lg, (T *)out)
    if (is_c64) {
        if (nout <= 4) AFR_SB_LAUNCH(float2, true); else AFR_SB_LAUNCH(float2, false);
    } else {
        if (nout <= 4) AFR_SB_LAUNCH(double2, true); else AFR_SB_LAUNCH(double2, false);
    }
#undef AFR_SB_LAUNCH
    AFR_LAUNCH_OK();
    return 0;
}
