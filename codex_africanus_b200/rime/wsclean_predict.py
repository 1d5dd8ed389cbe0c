"""``wsclean_predict`` and the WSClean spectral model on B200 --
africanus/rime/wsclean_predict.py:11-116, africanus/model/wsclean/spec_model.py:76-124."""
import numpy as np
import torch

from .. import _plumbing as pl


def _log_poly_array(log_poly, nsrc):
    """bool or (source,) bool array -> (source,) uint8 (spec_model.py:30-64)."""
    if isinstance(log_poly, (bool, np.bool_)):
        return np.full(nsrc, 1 if log_poly else 0, np.uint8)
    if pl.is_torch(log_poly):
        log_poly = log_poly.detach().cpu().numpy()
    lp = np.asarray(log_poly)
    if lp.ndim != 1 or lp.dtype != np.bool_:
        raise ValueError("log_poly must be ndarray or bool")
    if lp.shape[0] != nsrc:
        raise ValueError("coeffs.shape[0] != log_poly.shape[0]")
    return np.ascontiguousarray(lp, dtype=np.uint8)


def _check_spectral_args(flux, coeffs, ref_freq):
    fs, cs, rs = pl.shape_of(flux), pl.shape_of(coeffs), pl.shape_of(ref_freq)
    if len(fs) != 1 or len(cs) != 2 or len(rs) != 1 or not (fs[0] == cs[0] == rs[0]):
        raise ValueError("first dimensions of I, coeffs and ref_freq don't match.")
    return fs[0], cs[1]


def spectra(I, coeffs, log_poly, ref_freq, frequency):  # noqa: E741
    """WSClean polynomial spectral model, (source, chan) (spec_model.py:76-124): ordinary
    ``I + sum_c coeffs[c] (nu/ref - 1)^(c+1)`` or, where ``log_poly``, logarithmic
    ``I exp(sum_c coeffs[c] log(nu/ref)^(c+1))``.  Evaluated in float64; the result has the
    reference's dtype ``result_type(I, coeffs, ref_freq, frequency)``."""
    nsrc, ncoeffs = _check_spectral_args(I, coeffs, ref_freq)
    nchan = pl.shape_of(frequency)[0]
    out_dtype = np.result_type(*(pl.dtype_of(a) for a in (I, coeffs, ref_freq, frequency)))
    lp = _log_poly_array(log_poly, nsrc)
    device = pl.pick_device(I, coeffs, ref_freq, frequency)
    as_torch = pl.wants_torch(I, coeffs, ref_freq, frequency)
    f64 = np.float64
    with torch.cuda.device(device):
        d_i, d_c, d_r, d_f = (pl.to_device(a, f64, device) for a in (I, coeffs, ref_freq, frequency))
        d_lp = pl.to_device(lp, np.uint8, device)
        d_out = pl.empty_device((nsrc, nchan), f64, device)
        pl.call("afr_wsclean_spectra", device, pl.ptr(d_i), pl.ptr(d_c), pl.ptr(d_lp), pl.ptr(d_r),
                pl.ptr(d_f), nsrc, ncoeffs, nchan, pl.ptr(d_out), pl.stream_ptr(device))
        if out_dtype != f64:
            d_out = d_out.to(pl.torch_dtype(out_dtype))
        return d_out if as_torch else pl.to_host(d_out)


def wsclean_predict(uvw, lm, source_type, flux, coeffs, log_poly, ref_freq, gauss_shape, frequency):
    """Visibilities of a WSClean component list, (row, chan, 1) complex
    (wsclean_predict.py:86-116).  ``source_type`` (source,) of "POINT" / "GAUSSIAN";
    ``gauss_shape`` (source, 3) = (emaj, emin, angle) in radians; phase sign +2 pi/c (the CASA
    convention WSClean uses), n = sqrt(1 - l^2 - m^2) - 1 without clamp.

    POINT sources run on the FP64-pipe-bound phasor-stream kernel, GAUSSIAN ones on a
    per-term sincos + exp kernel.  All arithmetic is float64; the output dtype follows the
    reference, ``result_type(complex64, uvw, lm, flux, coeffs, ref_freq, frequency)`` (for
    all-float32 inputs the reference also evaluates the phase in float32; here it stays
    float64 and only the result is rounded).
    """
    ushape, lshape = pl.shape_of(uvw), pl.shape_of(lm)
    if len(ushape) != 2 or ushape[1] != 3 or len(lshape) != 2 or lshape[1] != 2:
        raise ValueError("wsclean_predict: expected uvw (row,3), lm (source,2)")
    nsrc, ncoeffs = _check_spectral_args(flux, coeffs, ref_freq)
    st = np.asarray(source_type.cpu().numpy() if pl.is_torch(source_type) else source_type)
    gshape = pl.shape_of(gauss_shape)
    if st.shape != (nsrc,) or lshape[0] != nsrc or tuple(gshape) != (nsrc, 3):
        raise ValueError("wsclean_predict: source arrays disagree on the number of sources")
    is_point, is_gauss = st == "POINT", st == "GAUSSIAN"
    if not np.all(is_point | is_gauss):
        raise ValueError("source_type must be POINT or GAUSSIAN")
    lp = _log_poly_array(log_poly, nsrc)
    nrow, nchan = ushape[0], pl.shape_of(frequency)[0]
    out_dtype = np.result_type(np.complex64, *(pl.dtype_of(a) for a in
                                               (uvw, lm, flux, coeffs, ref_freq, frequency)))
    arrays = (uvw, lm, flux, coeffs, ref_freq, gauss_shape, frequency)
    device = pl.pick_device(*arrays)
    as_torch = pl.wants_torch(*arrays)
    f64 = np.float64
    with torch.cuda.device(device):
        d_uvw, d_lm, d_i, d_c, d_r, d_g, d_f = (pl.to_device(a, f64, device) for a in
                                                (uvw, lm, flux, coeffs, ref_freq, gauss_shape, frequency))
        d_lp = pl.to_device(lp, np.uint8, device)
        d_isg = pl.to_device(is_gauss.astype(np.uint8), np.uint8, device)
        d_out = pl.empty_device((nrow, nchan, 1), np.complex128, device)
        pl.call("afr_wsclean_predict", device, pl.ptr(d_uvw), pl.ptr(d_lm), pl.ptr(d_isg), pl.ptr(d_g),
                pl.ptr(d_i), pl.ptr(d_c), pl.ptr(d_lp), pl.ptr(d_r), pl.ptr(d_f), nsrc, ncoeffs, nrow,
                nchan, int(is_gauss.sum()), pl.channel_mode(frequency), pl.ptr(d_out),
                pl.stream_ptr(device))
        if out_dtype != np.complex128:
            d_out = d_out.to(pl.torch_dtype(out_dtype))
        return d_out if as_torch else pl.to_host(d_out)
