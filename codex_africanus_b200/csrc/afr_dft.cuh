// Internal interface of the "phasor-stream" kernel family (afr_dft.cu), shared
// by im_to_vis, vis_to_im and the fused point-source predict.
#pragma once

#include <vector>

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

// ---- warp-specialised full-RIME predict with DDEs (afr_rime_ws.cu)
struct DdeWsParams {
    const double *lmn, *uvw, *freq;
    const double *bright;      // (nsrc,nchan,2,2) complex128
    const int32_t *ant1, *ant2;
    const int32_t *row_start;  // (ntime+1) first row of each timestep (rows sorted by time)
    const int32_t *perm;       // (nrow) rows ordered by (time, antenna tile)
    const double *dde1, *dde2; // (nsrc,ntime,nant,nchan,2,2) complex128
    const double *ant_uvw;     // (ntime,nant,3) per-antenna coordinates (antenna mode) or nullptr
    double *out;               // (nrow,nchan,2,2) complex128
    double cst;
    long long nsrc, nrow, ntime, nant;
    int nchan;
    int same_dde;
    int arrive_all;  // AFR_SANITIZE=1: every consumer lane arrives on the "empty" barriers
    // in-kernel beam sampling (antenna mode, E1 = E2 = the beam): instead of dde1 / dde2 the producers
    // fetch the two frequency planes of each channel from the plane-reduced beam and form the Jones
    const double *planes;  // (nsrc,ntime,nant,nud,12) or nullptr
    const double *fd;      // (nchan,3) frequency-grid table (scale, weight of the lower plane, lower plane)
    const double *feed;    // optional (ntime,nant,2,2) complex128 feed rotation applied on the right of the beam Jones
    int nud;
};
size_t dde_ws_smem_bytes(int64_t nant, int ft, bool ant, bool sample = false);
int dde_ws_row_tile_channels(int64_t nant);
// per-timestep antenna coordinates from baseline uvw; ok[0] is cleared when the rows of any
// timestep are not differences of per-antenna coordinates
int launch_antenna_uvw(const double *uvw, const int32_t *ant1, const int32_t *ant2,
                       const int32_t *row_start, int64_t ntime, int64_t nant, const double *lmn,
                       int64_t nsrc, const double *freq, int64_t nchan, double cst, double *ant_uvw,
                       int *ok, cudaStream_t stream);
int launch_row_tile_order(const int32_t *time_index, const int32_t *ant1, const int32_t *ant2,
                          int64_t nrow, int64_t ntime, bool pair_order, int32_t *perm,
                          cudaStream_t stream);
int launch_fused_dde_ws(const DdeWsParams &p, int max_rows_per_time, bool exact, bool ant_mode,
                        cudaStream_t stream);

// ---- antenna-mode DDE predict as a complex FP64 GEMM on the DMMA pipe (afr_rime_mma.cu)
// One pass: the consumer tiles (8 x 8 antennas each) of a panel of ni x nj antenna groups of 8,
// starting at groups gi0 (rows, antenna 1) and gj0 (columns, antenna 2).
constexpr int kDdeMmaMaxTiles = 36;
struct DdeMmaPass {
    int gi0, ni, gj0, nj, ntiles;
    uint8_t tile_m[kDdeMmaMaxTiles], tile_n[kDdeMmaMaxTiles];  // tile position inside the panel
    uint8_t mask[kDdeMmaMaxTiles];  // which 4 x 4-antenna blocks of the tile hold baselines
};
struct DdeMmaParams {
    const double *lmn, *freq;
    const double *bright;       // (nsrc,nchan,2,2) complex128
    const double *dde1, *dde2;  // (nsrc,ntime,nant,nchan,2,2) complex128
    const double *ant_uvw;      // (ntime,nant,3) per-antenna coordinates
    const int32_t *rowmap;      // (ntime,nant,nant) row of baseline (t, a1, a2) or -1
    const DdeMmaPass *passes;
    double *out;                // (nrow,nchan,2,2) complex128
    double cst;
    long long nsrc, ntime, nant;
    int nchan;
    int same_dde;
    int arrive_all;
    int ablate;  // diagnostics (AFR_MMA_ABLATE): 1 = skip the transform, 2 = skip the DMMAs
};
int launch_baseline_map(const int32_t *time_index, const int32_t *ant1, const int32_t *ant2, int64_t nrow,
                        int64_t ntime, int64_t nant, int32_t *rowmap, uint8_t *used4, int *dup,
                        cudaStream_t stream);
std::vector<DdeMmaPass> dde_mma_passes(const std::vector<uint8_t> &used4, int64_t nant);
int launch_fused_dde_mma(DdeMmaParams p, const std::vector<DdeMmaPass> &passes, cudaStream_t stream);

}  // namespace afr
