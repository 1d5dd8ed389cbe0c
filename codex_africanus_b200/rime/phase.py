"""``phase_delay`` on B200 -- africanus/rime/phase.py:11-63."""
import numpy as np
import torch

from .. import _plumbing as pl


def phase_delay(lm, uvw, frequency, convention="fourier"):
    """Complex phase delay K[s,r,f] = exp(-+2 pi i (l u + m v + (n-1) w) nu / c).

    lm (source, 2), uvw (row, 3), frequency (chan,) -> (source, row, chan) complex;
    dtype ``result_type(complex64, lm, uvw, frequency)`` (phase.py:26).  All-float32
    inputs are evaluated entirely in float32 like the reference (phase.py:23-25);
    with float32 ``lm`` the constant and ``n`` are float32-rounded.
    """
    sign = pl.convention_sign(convention)
    lshape, ushape, fshape = pl.shape_of(lm), pl.shape_of(uvw), pl.shape_of(frequency)
    if len(lshape) != 2 or lshape[1] != 2 or len(ushape) != 2 or ushape[1] != 3 or len(fshape) != 1:
        raise ValueError("phase_delay: expected lm (source,2), uvw (row,3), frequency (chan,)")
    out_dtype = np.result_type(np.complex64, *(pl.dtype_of(a) for a in (lm, uvw, frequency)))
    nsrc, nrow, nchan = lshape[0], ushape[0], fshape[0]
    device = pl.pick_device(lm, uvw, frequency)
    as_torch = pl.wants_torch(lm, uvw, frequency)
    with torch.cuda.device(device):
        d_out = pl.empty_device((nsrc, nrow, nchan), out_dtype, device)
        if out_dtype == np.complex64:
            f32 = np.float32
            d_lm, d_uvw, d_f = (pl.to_device(a, f32, device) for a in (lm, uvw, frequency))
            pl.call("afr_phase_delay_f32", device, pl.ptr(d_lm), pl.ptr(d_uvw), pl.ptr(d_f),
                    nsrc, nrow, nchan, sign, pl.ptr(d_out), pl.stream_ptr(device))
        else:
            f64 = np.float64
            d_lm, d_uvw, d_f = (pl.to_device(a, f64, device) for a in (lm, uvw, frequency))
            pl.call("afr_phase_delay_f64", device, pl.ptr(d_lm), pl.ptr(d_uvw), pl.ptr(d_f),
                    nsrc, nrow, nchan, sign, pl.f32_flags(lm, uvw, frequency),
                    pl.channel_mode(frequency), pl.ptr(d_out), pl.stream_ptr(device))
        return d_out if as_torch else pl.to_host(d_out)
