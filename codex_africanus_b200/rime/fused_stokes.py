"""
Fused predict with the source brightness generated on the device from Stokes parameters
(SURVEY.md 8f-2).

The reference's predict example builds ``brightness (source, chan, 2, 2)`` on the host from the
catalogue columns -- ``spectral_model(stokes, spi, ref_freq, frequency)`` then
``convert(..., ["I","Q","U","V"], [["XX","XY"],["YX","YY"]])``
(africanus/rime/examples/predict.py:107-134, africanus/model/spectral/spec_model.py:160-211,
africanus/model/coherency/conversion.py:19-28) -- and feeds it to the phase (x) brightness ->
``predict_vis`` chain.  At 10^5 sources x 4096 channels that array is 26 GB.  Here the per-source
input is (stokes, spi, ref_freq): the brightness of one chunk of sources is produced by one kernel
in device memory and reduced into the visibilities straight away (``base_vis`` doubles as the
accumulator, the streaming idiom of africanus/rime/dask_predict.py:216-239), so neither the host
nor the device ever holds more than one chunk.  The DIEs are applied once, after the last chunk.
"""
import numpy as np
import torch

from .. import _plumbing as pl
from ..model.coherency import LINEAR, stokes_brightness
from .fused import fused_predict_vis
from .predict import apply_gains

# device bytes one chunk of brightness may occupy
_BRIGHTNESS_CHUNK_BYTES = 1 << 30


def fused_predict_vis_stokes(lm, uvw, frequency, stokes, spi, ref_freq, time_index, antenna1,
                             antenna2, dde1_jones=None, dde2_jones=None, die1_jones=None,
                             base_vis=None, die2_jones=None, convention="fourier", base=0,
                             stokes_schema=("I", "Q", "U", "V"), corr_schema=LINEAR,
                             implicit_stokes=False, dtype=None, source_chunk=None):
    """``fused_predict_vis(lm, uvw, frequency, brightness, ...)`` with
    ``brightness = convert(spectral_model(stokes, spi, ref_freq, frequency, base),
    stokes_schema, corr_schema, implicit_stokes)`` generated per chunk of sources on the device.

    stokes (source, pol), spi (source, spi-comps, pol), ref_freq (source,) real;
    ``corr_schema`` of 1, 2 or 2x2 correlations; the other arguments are those of
    ``fused_predict_vis``.  ``dtype`` (complex64/complex128) is the brightness dtype, default
    complex of the real inputs' precision; the output dtype is ``np.result_type`` of it and the
    Jones / base_vis inputs, as for ``fused_predict_vis``.  ``source_chunk`` = sources per chunk
    (default: what fits 1 GiB of brightness).  Returns (row, chan, corr...).
    """
    if (die1_jones is None) != (die2_jones is None):
        raise ValueError("Both die1_jones and die2_jones must be present or absent")
    if (dde1_jones is None) != (dde2_jones is None):
        raise ValueError("Both dde1_jones and dde2_jones must be present or absent")
    nsrc = pl.shape_of(lm)[0]
    if pl.shape_of(stokes)[:1] != (nsrc,):
        raise ValueError("fused_predict_vis_stokes: lm / stokes disagree on the number of sources")
    nchan = pl.shape_of(frequency)[0]
    ncorr = int(np.asarray(corr_schema, dtype=object).size)
    real = np.result_type(*(pl.dtype_of(a) for a in (stokes, spi, ref_freq, frequency)))
    b_dtype = np.dtype(dtype) if dtype is not None else np.result_type(real, np.complex64)
    cplx = [a for a in (dde1_jones, dde2_jones, die1_jones, base_vis, die2_jones) if a is not None]
    out_dtype = np.result_type(b_dtype, *(pl.dtype_of(a) for a in cplx))
    if source_chunk is None:
        source_chunk = max(1, _BRIGHTNESS_CHUNK_BYTES // max(1, nchan * ncorr * out_dtype.itemsize))
    source_chunk = int(max(1, min(source_chunk, max(nsrc, 1))))

    everything = (lm, uvw, frequency, stokes, spi, ref_freq, time_index, antenna1, antenna2,
                  dde1_jones, dde2_jones, die1_jones, base_vis, die2_jones)
    device = pl.pick_device(*everything)
    as_torch = pl.wants_torch(*everything)
    f64 = np.float64
    with torch.cuda.device(device):
        # small per-source / per-row inputs go up once; the public entry points then run
        # tensor -> tensor
        d_lm, d_uvw, d_f, d_st, d_spi, d_rf = (pl.to_device(a, f64, device) for a in
                                               (lm, uvw, frequency, stokes, spi, ref_freq))
        d_ti, d_a1, d_a2 = ((x if pl.is_torch(x) else torch.from_numpy(np.ascontiguousarray(x))).to(device)
                            for x in (time_index, antenna1, antenna2))
        acc = None if base_vis is None else pl.to_device(base_vis, out_dtype, device)
        same_dde = dde1_jones is dde2_jones
        for s0 in range(0, max(nsrc, 1), source_chunk):
            s1 = min(nsrc, s0 + source_chunk)
            d_b = stokes_brightness(d_st[s0:s1], d_spi[s0:s1], d_rf[s0:s1], d_f, base=base,
                                    stokes_schema=stokes_schema, corr_schema=corr_schema,
                                    implicit_stokes=implicit_stokes, dtype=out_dtype)
            d_e1 = d_e2 = None
            if dde1_jones is not None:  # only this chunk's source slice travels
                d_e1 = pl.to_device(dde1_jones[s0:s1], out_dtype, device)
                d_e2 = d_e1 if same_dde else pl.to_device(dde2_jones[s0:s1], out_dtype, device)
            acc = fused_predict_vis(d_lm[s0:s1], d_uvw, d_f, d_b, d_ti, d_a1, d_a2, d_e1, d_e2,
                                    None, acc, None, convention=convention, dtype=out_dtype)
            del d_b, d_e1, d_e2
        if die1_jones is not None:
            d_g1 = pl.to_device(die1_jones, out_dtype, device)
            d_g2 = d_g1 if die2_jones is die1_jones else pl.to_device(die2_jones, out_dtype, device)
            acc = apply_gains(d_ti, d_a1, d_a2, d_g1, acc, d_g2)
        return acc if as_torch else pl.to_host(acc)
