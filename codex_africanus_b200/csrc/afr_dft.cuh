// Internal interface of the "phasor-stream" kernel family (afr_dft.cu), shared
// by im_to_vis, vis_to_im and the fused point-source predict.
#pragma once

#include "afr_common.cuh"

namespace afr {

// How n is derived from (l, m):
enum LmnMode {
    kLmnDft = 0,        // n = sqrt(1 - l^2 - m^2) - 1, no clamp (dft/kernels.py:54,122)
    kLmnPhaseClamp = 1  // n = sqrt(max(0, 1 - l^2 - m^2)) - 1   (rime/phase.py:42-43)
};

// lm (nsrc,2) -> lmn (nsrc,3) on `stream`
int launch_lm_to_lmn(const double *lm, int64_t nsrc, int mode, bool lm_f32, double *lmn,
                     cudaStream_t stream);

// acc[x,f,c] (+)= sum_y exp(i * cst * (xc[x] . yc[y]) * freq[f]) * w[y,f,c]
//   adjoint = false: complex accumulators, out (nx,nchan,ncorr) complex
//   adjoint = true : only the real part is kept, out (nx,nchan,ncorr) real
// w: (ny,nchan,ncorr) real (or complex when w_complex) in the ACCUMULATOR precision: float64 /
// complex128, or float32 / complex64 when acc32; flags (same shape) zero a (y,f) sample when
// any correlation is flagged.  acc32: rotate/accumulate in FP32 and write complex64/float32
// (the phase argument and anchors stay FP64).
int run_phasor_stream(const double *xc, int64_t nx, const double *yc, int64_t ny, const void *w,
                      bool w_complex, const uint8_t *flags, const double *freq, int64_t nchan,
                      int64_t ncorr, double cst, bool f32dot, bool adjoint, bool exact,
                      bool acc32, void *out, cudaStream_t stream);

}  // namespace afr
