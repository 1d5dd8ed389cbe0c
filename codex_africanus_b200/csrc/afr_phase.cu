// phase_delay (africanus/rime/phase.py:11-63) for sm_100a.
//
// Stand-alone, the K term is bound by the HBM store of its (source,row,chan) output
// (16 B/term complex128, 8 B/term complex64), so the kernel is organised around the
// store: a warp walks one (source,row) line of the output with lane <-> channel, so every
// warp store is a fully coalesced 512-byte (256-byte for complex64) segment.  At 16 B/term the FP64 pipe has room for one
// branch-free sincos (cis_fast, 22 FP64 instructions) per element, so each element's
// phase is fl(fl(phi)*nu_f) exactly as the reference rounds it -- no recurrence here.
#include "afr_dft.cuh"

namespace afr {
namespace {

struct PhaseParams {
    const double *lmn;   // (nsrc,3), n already clamped (rime/phase.py:42-43)
    const double *uvw;   // (nrow,3)
    const double *freq;  // (nchan,)
    double *out;         // (nsrc,nrow,nchan) complex128
    double cst;          // signed, lm.dtype-rounded (phase.py:25)
    long long nsrc, nrow;
    int nchan;
    int all_f32_coords;  // lm and uvw were float32: l*u + m*v + n*w and cst*(.) in float32
};

__device__ __forceinline__ double real_phase(const PhaseParams &p, long long s, long long r) {
    const double l = p.lmn[3 * s], m = p.lmn[3 * s + 1], n = p.lmn[3 * s + 2];
    const double u = p.uvw[3 * r], v = p.uvw[3 * r + 1], w = p.uvw[3 * r + 2];
    if (p.all_f32_coords) {
        const float a = __fadd_rn(__fadd_rn(__fmul_rn((float)l, (float)u), __fmul_rn((float)m, (float)v)),
                                  __fmul_rn((float)n, (float)w));
        return (double)__fmul_rn((float)p.cst, a);
    }
    return __dmul_rn(p.cst, phase_dot(l, m, n, u, v, w, false));
}

constexpr int kRowsPerBlock = 64;  // rows of one source handled by a 256-thread block

// block <-> (source, block of kRowsPerBlock rows); warp <-> one row at a time; lane <-> channel
// (f = lane, lane+32, ...), so a warp store is 512 contiguous bytes and no thread divides.
__global__ void __launch_bounds__(256) phase_delay_f64_kernel(const PhaseParams p) {
    const long long row_blocks = (p.nrow + kRowsPerBlock - 1) / kRowsPerBlock;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    double2 *out = reinterpret_cast<double2 *>(p.out);
    for (long long blk = blockIdx.x; blk < p.nsrc * row_blocks; blk += gridDim.x) {
        const long long s = blk / row_blocks;
        const long long r0 = (blk - s * row_blocks) * kRowsPerBlock;
        const long long r1 = min(p.nrow, r0 + kRowsPerBlock);
        for (long long r = r0 + warp; r < r1; r += 8) {
            const double phi = real_phase(p, s, r);
            double2 *o = out + (s * p.nrow + r) * p.nchan;
            for (int f = lane; f < p.nchan; f += 32) {
                const C2<double> z = cis_fast(__dmul_rn(phi, p.freq[f]));
                o[f] = make_double2(z.re, z.im);
            }
        }
    }
}

// all-float32 inputs: the reference does the entire phase in float32 (phase.py:23-26);
// p is rounded to float32 per channel, so no recurrence is possible.
__global__ void __launch_bounds__(256)
    phase_delay_f32_kernel(const float *lm, const float *uvw, const float *freq, float2 *out,
                           float cst, long long nsrc, long long nrow, int nchan) {
    const long long row_blocks = (nrow + kRowsPerBlock - 1) / kRowsPerBlock;
    const int lane = threadIdx.x & 31, warp = threadIdx.x >> 5;
    for (long long blk = blockIdx.x; blk < nsrc * row_blocks; blk += gridDim.x) {
        const long long s = blk / row_blocks;
        const long long r0 = (blk - s * row_blocks) * kRowsPerBlock;
        const long long r1 = min(nrow, r0 + kRowsPerBlock);
        const float l = lm[2 * s], m = lm[2 * s + 1];
        float n = __fsub_rn(__fsub_rn(1.0f, __fmul_rn(l, l)), __fmul_rn(m, m));
        n = __fsub_rn(__fsqrt_rn(n < 0.0f ? 0.0f : n), 1.0f);
        for (long long r = r0 + warp; r < r1; r += 8) {
            const float u = uvw[3 * r], v = uvw[3 * r + 1], w = uvw[3 * r + 2];
            const float rp = __fmul_rn(
                cst, __fadd_rn(__fadd_rn(__fmul_rn(l, u), __fmul_rn(m, v)), __fmul_rn(n, w)));
            float2 *o = out + (s * nrow + r) * nchan;
            for (int f = lane; f < nchan; f += 32) {
                float sn, cs;
                sincosf(__fmul_rn(rp, freq[f]), &sn, &cs);
                o[f] = make_float2(cs, sn);
            }
        }
    }
}

int grid_for(long long total) {
    long long blocks = (total + 255) / 256;
    const long long cap = (long long)sm_count() * 32;
    if (blocks > cap) blocks = cap;
    if (blocks < 1) blocks = 1;
    return (int)blocks;
}

}  // namespace
}  // namespace afr

using namespace afr;

extern "C" int afr_phase_delay_f64(const double *lm, const double *uvw, const double *freq,
                                   int64_t nsrc, int64_t nrow, int64_t nchan, int convention,
                                   int f32_flags, int chan_mode, void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(convention == AFR_FOURIER || convention == AFR_CASA,
                "convention not in ('fourier', 'casa')");
    AFR_REQUIRE(nsrc >= 0 && nrow >= 0 && nchan >= 0 && nchan < (1LL << 30), "bad extent");
    if (nsrc == 0 || nrow == 0 || nchan == 0) return 0;
    const bool lm_f32 = (f32_flags & AFR_F32_LM) != 0;
    const bool uvw_f32 = (f32_flags & AFR_F32_UVW) != 0;
    Scratch lmn;
    AFR_CUDA_OK(lmn.alloc(sizeof(double) * 3 * (size_t)nsrc, stream));
    int rc = launch_lm_to_lmn(lm, nsrc, kLmnPhaseClamp, lm_f32, (double *)lmn.ptr, stream);
    if (rc) return rc;
    PhaseParams p{};
    p.lmn = (const double *)lmn.ptr;
    p.uvw = uvw;
    p.freq = freq;
    p.out = (double *)out;
    double cst = -kTwoPiOverC;             // phase.py:25,30
    if (lm_f32) cst = (double)(float)cst;  // constant typed as lm.dtype
    if (convention == AFR_CASA) cst = -cst;
    p.cst = cst;
    p.nsrc = nsrc;
    p.nrow = nrow;
    p.nchan = (int)nchan;
    p.all_f32_coords = (lm_f32 && uvw_f32) ? 1 : 0;
    (void)chan_mode;  // every element takes its own sincos: both modes are exact
    const long long blocks = nsrc * ((nrow + kRowsPerBlock - 1) / kRowsPerBlock);
    phase_delay_f64_kernel<<<grid_for(blocks * 256), 256, 0, stream>>>(p);
    AFR_LAUNCH_OK();
    return 0;
}

extern "C" int afr_phase_delay_f32(const float *lm, const float *uvw, const float *freq,
                                   int64_t nsrc, int64_t nrow, int64_t nchan, int convention,
                                   void *out, void *stream_) {
    cudaStream_t stream = (cudaStream_t)stream_;
    AFR_REQUIRE(convention == AFR_FOURIER || convention == AFR_CASA,
                "convention not in ('fourier', 'casa')");
    AFR_REQUIRE(nsrc >= 0 && nrow >= 0 && nchan >= 0 && nchan < (1LL << 30), "bad extent");
    if (nsrc == 0 || nrow == 0 || nchan == 0) return 0;
    float cst = (float)(-kTwoPiOverC);
    if (convention == AFR_CASA) cst = -cst;
    const long long blocks = nsrc * ((nrow + kRowsPerBlock - 1) / kRowsPerBlock);
    phase_delay_f32_kernel<<<grid_for(blocks * 256), 256, 0, stream>>>(lm, uvw, freq, (float2 *)out, cst,
                                                               nsrc, nrow, (int)nchan);
    AFR_LAUNCH_OK();
    return 0;
}
