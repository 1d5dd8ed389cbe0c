/*
 * africanus_b200.h -- C ABI of libafricanus_b200.so
 *
 * B200 (sm_100a) implementation of the codex-africanus RIME / DFT hot path.
 * The reference (ratt-ru/codex-africanus 0.4.4) has no native ABI: its
 * interface for this path is a set of Python callables
 *     africanus/rime/__init__.py:3-10   phase_delay, predict_vis, apply_gains, beam_cube_dde
 *     africanus/dft/__init__.py:3       im_to_vis, vis_to_im
 * and alternative backends are sibling modules exporting the same names
 * (africanus/rime/cuda/__init__.py:3-6).  The functions below are what a
 * ctypes/cffi binding for such a sibling module binds; each cites the
 * reference callable it stands behind.  See INTEGRATION.md for the stub.
 *
 * Conventions
 *  - every array pointer is a DEVICE pointer on the current CUDA device, dense
 *    C-contiguous in the reference's documented layout; complex = interleaved
 *    (re, im).  Inputs are read-only; outputs are caller-allocated.
 *  - `stream` is a cudaStream_t passed as void* (NULL = legacy default stream).
 *    Calls are asynchronous with respect to the host unless stated otherwise.
 *  - real coordinate inputs (uvw, lm, frequency ...) are float64.  Where the
 *    reference rounds an intermediate to float32 because the caller's arrays
 *    were float32, the AFR_F32_* flags request the same rounding (the caller
 *    widens the values exactly and says what they were).
 *  - return value 0 = success; non-zero = error, text in afr_last_error()
 *    (thread-local).  The library is re-entrant and keeps no global mutable
 *    state besides that thread-local string.
 */
#ifndef AFRICANUS_B200_H
#define AFRICANUS_B200_H

#include <stdint.h>

#ifdef __cplusplus
extern "C" {
#endif

#define AFR_VERSION 100

/* sign convention: africanus/rime/phase.py:29-34, africanus/dft/kernels.py:34-39,110-115 */
#define AFR_FOURIER 1
#define AFR_CASA (-1)

/* dtype-origin flags (bitmask) */
#define AFR_F32_LM 1   /* lm was float32   */
#define AFR_F32_UVW 2  /* uvw was float32  */
#define AFR_F32_FREQ 4 /* frequency was float32 */

/* channel mode */
#define AFR_CHAN_EXACT 0   /* one sincos per (source,row,chan): any frequency array */
#define AFR_CHAN_UNIFORM 1 /* equispaced channels: anchored complex-rotation recurrence */

/* Jones correlation mode: africanus/rime/predict.py:10-12 */
#define AFR_JONES_DIAG 0 /* trailing (ncorr,) multiplied element-wise */
#define AFR_JONES_2X2 1  /* trailing (2,2) matrix products, ncorr == 4 */

/* ---- library / device ------------------------------------------------- */
int afr_version(void);
const char *afr_last_error(void);
int afr_device_count(void);
/* number of CUDA kernels this library has launched in this process so far */
unsigned long long afr_kernel_launches(void);
/* Which kernel the last afr_predict_fused call of this host thread ran (diagnostics/tests). */
#define AFR_PATH_NONE 0
#define AFR_PATH_POINT 1        /* no DDEs: phasor-stream kernel                              */
#define AFR_PATH_DDE_WS_ANT 2   /* warp-specialised DDE kernel, phasors folded per antenna    */
#define AFR_PATH_DDE_WS_ROW 3   /* warp-specialised DDE kernel, per-row phasors               */
#define AFR_PATH_DDE_TILED 4    /* antenna-tiled single-role kernel                           */
#define AFR_PATH_DDE_GATHER 5   /* gather kernel (unsorted rows, diagonal Jones, complex64)   */
#define AFR_PATH_DDE_MMA_ANT 6  /* antenna phasors, source sum as a complex GEMM on the FP64 tensor pipe */
#define AFR_PATH_DDE_WS_ANT_SAMPLED 7 /* AFR_PATH_DDE_WS_ANT with the beam sampled inside the kernel */
int afr_last_fused_path(void);
/* Which schedule of the phasor-stream kernel the last afr_im_to_vis / afr_vis_to_im / point-source
 * afr_predict_fused / afr_wsclean_predict launch of this host thread used (its last correlation
 * block): bit 0 warp-specialised (16 consumer + 4 producer warps), bit 1 one sincos per term
 * (non-equispaced channels), bit 2 W tile by TMA bulk copies, bit 3 FP32 accumulators, bit 4 eight
 * producer warps (few-channel adjoint), bit 5 consumers on the FP64 tensor pipe (DMMA; 2x2 complex
 * brightness, forward, equispaced channels),
 * bits 8-15 channel runs per CTA, bits 16-31 slices of the streamed axis.  Environment overrides
 * (AFR_WS, AFR_POINT_MMA, AFR_SANITIZE) show up here, so a benchmark can report the path it measured. */
int afr_last_dft_path(void);
/* Return the scratch memory this library keeps cached in `device`'s default CUDA memory pool
 * (up to 2 GiB between calls) to the driver. */
int afr_trim_scratch(int device);
/* Select the CUDA device used by subsequent calls on this host thread.  The library
 * links its own (static) CUDA runtime, whose current-device state is separate from any
 * other runtime in the process (e.g. PyTorch's): bindings call this before each entry
 * point with the device that owns the pointers. */
int afr_set_device(int device);
/* sm_count, compute capability and SM clock (kHz) of `device` */
int afr_device_info(int device, int *sm_count, int *cc_major, int *cc_minor, int *clock_khz);
/* 1 if `freq` (HOST pointer) is equispaced to within rtol*max|freq| */
int afr_freq_is_uniform(const double *freq_host, int64_t nchan, double rtol);

/* Measured pipe peaks for the roofline denominators: runs a dependent-free
 * DFMA (fp64 != 0) or FFMA chain on every SM for `iters` iterations and
 * returns achieved FLOP/s (FMA = 2).  Synchronous. */
int afr_measure_fma_peak(int fp64, int iters, double *flops_per_s, void *stream);

/* ---- africanus.dft.im_to_vis  (africanus/dft/kernels.py:14-69) --------- */
/* image (nsrc,nchan,ncorr) float64 or complex128 (image_complex);
 * uvw (nrow,3); lm (nsrc,2); freq (nchan,); out (nrow,nchan,ncorr) complex128.
 * out_c64: out is complex64, image is float32 / complex64, and the rotation and
 * accumulation run in FP32 (phase argument and anchors stay FP64). */
int afr_im_to_vis(const void *image, int image_complex, const double *uvw, const double *lm,
                  const double *freq, int64_t nsrc, int64_t nrow, int64_t nchan, int64_t ncorr,
                  int convention, int f32_flags, int chan_mode, int out_c64, void *out,
                  void *stream);

/* ---- africanus.dft.vis_to_im  (africanus/dft/kernels.py:72-148) -------- */
/* vis (nrow,nchan,ncorr) float64 or complex128 (vis_complex); flags uint8 same shape
 * (NULL = nothing flagged); out (nsrc,nchan,ncorr) float64.
 * out_f32: out is float32 and vis is float32 / complex64 (FP32 rotation/accumulation). */
int afr_vis_to_im(const void *vis, int vis_complex, const double *uvw, const double *lm,
                  const double *freq, const uint8_t *flags, int64_t nsrc, int64_t nrow,
                  int64_t nchan, int64_t ncorr, int convention, int f32_flags, int chan_mode,
                  int out_f32, void *out, void *stream);

/* ---- africanus.rime.phase_delay  (africanus/rime/phase.py:11-63) ------- */
/* out (nsrc,nrow,nchan) complex128.  f32_flags reproduce the lm.dtype-typed
 * constants (phase.py:23-25).  All-float32 inputs use afr_phase_delay_f32. */
int afr_phase_delay_f64(const double *lm, const double *uvw, const double *freq, int64_t nsrc,
                        int64_t nrow, int64_t nchan, int convention, int f32_flags,
                        int chan_mode, void *out, void *stream);
/* all inputs float32, whole phase in float32, out complex64 (phase.py:23-26) */
int afr_phase_delay_f32(const float *lm, const float *uvw, const float *freq, int64_t nsrc,
                        int64_t nrow, int64_t nchan, int convention, void *out, void *stream);

/* ---- africanus.rime.predict_vis / apply_gains (africanus/rime/predict.py:466-649) */
/* time_index/antenna1/antenna2 (nrow,) int32, time_index ALREADY minus its minimum
 * (predict.py:597).  dde{1,2} (nsrc,ntime,nant,nchan,C), source_coh (nsrc,nrow,nchan,C),
 * die{1,2} (ntime,nant,nchan,C), base_vis (nrow,nchan,C), out (nrow,nchan,C); C = ncorr
 * complex values; any of the inputs may be NULL subject to the reference's pairing
 * rules (predict.py:403-407).  is_c64: all arrays complex64, else complex128. */
int afr_predict_vis(const int32_t *time_index, const int32_t *antenna1, const int32_t *antenna2,
                    const void *dde1, const void *source_coh, const void *dde2, const void *die1,
                    const void *base_vis, const void *die2, int64_t nsrc, int64_t nrow,
                    int64_t ntime, int64_t nant, int64_t nchan, int64_t ncorr, int jones_mode,
                    int is_c64, void *out, void *stream);

/* ---- fused phase_delay (x) brightness -> predict_vis -------------------- */
/* The composition africanus/rime/examples/predict.py:107-134,490,522-527
 * (asserted in africanus/experimental/rime/fused/tests/test_rime.py:175-209)
 * without materialising the (source,row,chan[,corr]) intermediates:
 *   V[r,f] = G1 (B[r,f] + sum_s E1 (K[s,r,f] Bright[s,f]) E2^H) G2^H
 * brightness (nsrc,nchan,C) complex; other arrays as afr_predict_vis.
 * out_c64: brightness/dde/die/base_vis/out are complex64 and the Jones chain
 * runs in FP32, while the phase argument stays FP64. */
int afr_predict_fused(const double *lm, const double *uvw, const double *freq,
                      const void *brightness, const int32_t *time_index,
                      const int32_t *antenna1, const int32_t *antenna2, const void *dde1,
                      const void *dde2, const void *die1, const void *base_vis, const void *die2,
                      int64_t nsrc, int64_t nrow, int64_t ntime, int64_t nant, int64_t nchan,
                      int64_t ncorr, int jones_mode, int convention, int chan_mode, int out_c64,
                      void *out, void *stream);

/* ---- africanus.rime.beam_cube_dde (africanus/rime/fast_beam_cubes.py:57-240) */
/* beam (lw,mh,nud,ncorr) complex; extents (2,2); beam_freq_map (nud,); lm (nsrc,2);
 * parallactic_angles (ntime,nant); point_errors (ntime,nant,nchan,2);
 * antenna_scaling (nant,nchan,2); freq (nchan,); out (nsrc,ntime,nant,nchan,ncorr).
 * beam / out aligned to one complex value, point_errors / antenna_scaling to 16 bytes. */
int afr_beam_cube_dde(const void *beam, const double *beam_lm_extents,
                      const double *beam_freq_map, const double *lm,
                      const double *parallactic_angles, const double *point_errors,
                      const double *antenna_scaling, const double *freq, int64_t lw, int64_t mh,
                      int64_t nud, int64_t ncorr, int64_t nsrc, int64_t ntime, int64_t nant,
                      int64_t nchan, int is_c64, void *out, void *stream);
/* The same with the feed rotation of africanus/rime/examples/predict.py:469-472 applied in the
 * kernel's epilogue: out = einsum("stafij,tajk->stafik", beam_cube_dde(...), feed_rotation).
 * feed_rotation (ntime,nant,2,2) complex of the beam's precision, or NULL (= afr_beam_cube_dde);
 * requires ncorr == 4. */
int afr_beam_cube_dde_rot(const void *beam, const double *beam_lm_extents,
                          const double *beam_freq_map, const double *lm,
                          const double *parallactic_angles, const double *point_errors,
                          const double *antenna_scaling, const double *freq,
                          const void *feed_rotation, int64_t lw, int64_t mh, int64_t nud,
                          int64_t ncorr, int64_t nsrc, int64_t ntime, int64_t nant, int64_t nchan,
                          int is_c64, void *out, void *stream);
/* ---- SURVEY 8f-1 proper: the beam sampled inside the predict kernel
 * (africanus/experimental/rime/fused/terms/cube_dde.py:96-313 samples the cube inside the reference's
 * fused loop; africanus/rime/examples/predict.py:390-401 materialises beam_cube_dde instead).
 * afr_beam_plane_reduce: the four spatial corners of every frequency plane reduced per (source, time,
 * antenna): planes (nsrc,ntime,nant,nud,12) float64 = re[4] | im[4] | abs[4] per plane (2x2 complex128
 * beam), fd (nchan,3) = freq_grid_interp, ok[0] (device int) = 1 when every channel of every row is the
 * combination of two planes (pointing errors / antenna scaling constant along chan, all channels inside
 * the cube's frequency range).
 * afr_predict_fused_planes: afr_predict_fused with dde1 = dde2 = the beam Jones formed by the kernel's
 * producers from `planes` (never materialised), times the optional feed rotation (ntime,nant,2,2) complex128
 * on its right (as afr_beam_cube_dde_rot).  used[0] (HOST int) = 1 when it ran, 0 when the path does
 * not apply (uvw not differences of antenna coordinates within the admission bound, rows not ordered by
 * time, antenna tile too large): nothing was written and the caller takes the chunked route. */
int afr_beam_plane_reduce(const void *beam, const double *beam_lm_extents, const double *beam_freq_map,
                          const double *lm, const double *parallactic_angles, const double *point_errors,
                          const double *antenna_scaling, const double *freq, int64_t lw, int64_t mh,
                          int64_t nud, int64_t nsrc, int64_t ntime, int64_t nant, int64_t nchan,
                          double *planes, double *fd, int *ok, void *stream);
int afr_predict_fused_planes(const double *lm, const double *uvw, const double *freq, const void *brightness,
                             const int32_t *time_index, const int32_t *antenna1, const int32_t *antenna2,
                             const double *planes, const double *fd, int64_t nud, const void *feed_rotation,
                             const void *die1, const void *base_vis, const void *die2, int64_t nsrc, int64_t nrow,
                             int64_t ntime, int64_t nant, int64_t nchan, int convention, int *used, void *out,
                             void *stream);
/* feed_rotation (africanus/rime/feeds.py:13-71): parallactic_angles (n,) float64 -> out (n,2,2)
 * complex128 (is_c64 0) / complex64 (1).  AFR_FEED_LINEAR [[cos,sin],[-sin,cos]],
 * AFR_FEED_CIRCULAR diag(exp(-i pa), exp(+i pa)). */
#define AFR_FEED_LINEAR 0
#define AFR_FEED_CIRCULAR 1
int afr_feed_rotation(const double *parallactic_angles, int64_t n, int feed_type, int is_c64,
                      void *out, void *stream);
/* freq_grid_interp (fast_beam_cubes.py:10-54): freq_data (nchan,3) float64 */
int afr_freq_grid_interp(const double *freq, const double *beam_freq_map, int64_t nchan,
                         int64_t nud, double *freq_data, void *stream);

/* ---- africanus.rime.wsclean_predict (africanus/rime/wsclean_predict.py:11-116) and
 * africanus.model.wsclean.spectra (africanus/model/wsclean/spec_model.py:76-124) ---------- */
/* out (nsrc,nchan) float64 = I + sum_c coeffs[c] (nu/ref - 1)^(c+1), or, where log_poly[s],
 * I exp(sum_c coeffs[c] log(nu/ref)^(c+1)).  flux (nsrc,), coeffs (nsrc,ncoeffs),
 * log_poly (nsrc,) uint8, ref_freq (nsrc,), freq (nchan,). */
int afr_wsclean_spectra(const double *flux, const double *coeffs, const uint8_t *log_poly,
                        const double *ref_freq, const double *freq, int64_t nsrc, int64_t ncoeffs,
                        int64_t nchan, double *out, void *stream);
/* out (nrow,nchan,1) complex128.  is_gauss (nsrc,) uint8: 0 = "POINT", 1 = "GAUSSIAN";
 * gauss_shape (nsrc,3) = (emaj, emin, angle); ngauss = number of GAUSSIAN sources (0 skips the
 * taper kernel); chan_mode as in afr_im_to_vis.  Phase sign +2pi/c, n not clamped. */
int afr_wsclean_predict(const double *uvw, const double *lm, const uint8_t *is_gauss,
                        const double *gauss_shape, const double *flux, const double *coeffs,
                        const uint8_t *log_poly, const double *ref_freq, const double *freq,
                        int64_t nsrc, int64_t ncoeffs, int64_t nrow, int64_t nchan, int64_t ngauss,
                        int chan_mode, void *out, void *stream);

/* ---- brightness from Stokes parameters ---------------------------------
 * africanus.model.spectral.spectral_model (africanus/model/spectral/spec_model.py:106-211) and
 * africanus.model.coherency.convert (africanus/model/coherency/conversion.py:145-243) ------ */
/* out (nsrc,nchan,npol) float64.  stokes (nsrc,npol), spi (nsrc,nspi,npol), ref_freq (nsrc,),
 * freq (nchan,) device float64; base: HOST array of npol ints, 0 "std" (stokes prod_i (nu/ref)^spi_i),
 * 1 "log" (stokes exp(sum_i spi_i log(nu/ref)^(i+1))), 2 "log10"; npol <= 16. */
int afr_spectral_model(const double *stokes, const double *spi, const double *ref_freq,
                       const double *freq, const int *base, int64_t nsrc, int64_t nspi,
                       int64_t npol, int64_t nchan, double *out, void *stream);
/* out (n,nout) complex128 from in (n,nin) float64 (in_complex 0) or complex128 (1).  src1, src2,
 * op: HOST arrays of nout ints, the schema resolved as in conversion.py:145-215: output o =
 * op[o](in[src1[o]], in[src2[o]]), index -1 = the implicit-Stokes zero; ops (conversion.py:19-48)
 * 0 a+b, 1 a-b, 2 a+b*1j, 3 a-b*1j, 4 (a+b)/2, 5 (a-b)/2, 6 (a-b)/2j; nout <= 16. */
int afr_convert(const void *in, int in_complex, int64_t n, int64_t nin, const int *src1,
                const int *src2, const int *op, int64_t nout, void *out, void *stream);
/* the two composed: out (nsrc,nchan,nout) complex128 (is_c64 0) or complex64 (1) =
 * convert(spectral_model(stokes, spi, ref_freq, freq, base)); npol <= 4 Stokes inputs.  This is
 * the `brightness` argument of afr_predict_fused (rime/examples/predict.py:107-134). */
int afr_stokes_brightness(const double *stokes, const double *spi, const double *ref_freq,
                          const double *freq, const int *base, int64_t nsrc, int64_t nspi,
                          int64_t npol, int64_t nchan, const int *src1, const int *src2,
                          const int *op, int64_t nout, int is_c64, void *out, void *stream);

#ifdef __cplusplus
}
#endif
#endif /* AFRICANUS_B200_H */
