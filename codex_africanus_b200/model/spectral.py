"""``spectral_model`` on B200 -- africanus/model/spectral/spec_model.py:106-211."""
import ctypes

import numpy as np
import torch

from .. import _plumbing as pl

_BASES = {0: 0, "std": 0, 1: 1, "log": 1, 2: 2, "log10": 2}


def promote_bases(base, npol):
    """``base`` (int, str or list of them) -> npol codes; a list is padded with its last
    entry (spec_model.py:77-86).  The reference raises ValueError("Invalid base") from inside
    the loop and TypeError for other argument types (spec_model.py:139-140,210)."""
    bases = list(base) if isinstance(base, (list, tuple)) else [base]
    if len(bases) == 0:
        raise ValueError("Invalid base")
    bases = (bases + [bases[-1]] * npol)[:npol]
    for b in bases:
        if isinstance(b, (bool, np.bool_)) or not isinstance(b, (int, np.integer, str)):
            raise TypeError("base '%s' should be a string or integer" % (b,))
        if b not in _BASES:
            raise ValueError("Invalid base")
    return (ctypes.c_int * max(npol, 1))(*[_BASES[b] for b in bases])


def spectral_shapes(stokes, spi, ref_freq, frequency):
    """Argument checks of spec_model.py:157-171 -> (nsrc, nspi, npol, nchan, pol_shape)."""
    sshape, pshape = pl.shape_of(stokes), pl.shape_of(spi)
    if len(sshape) < 1 or len(pshape) < 2 or len(pshape) - 2 != len(sshape) - 1:
        raise ValueError("Dimensions on stokes and spi don't agree")
    npol = int(np.prod(sshape[1:], dtype=np.int64))
    if npol != int(np.prod(pshape[2:], dtype=np.int64)):
        raise ValueError("Correlations on stokes and spi don't agree")
    rshape, fshape = pl.shape_of(ref_freq), pl.shape_of(frequency)
    if len(rshape) != 1 or len(fshape) != 1 or not (sshape[0] == pshape[0] == rshape[0]):
        raise ValueError("stokes, spi and ref_freq disagree on the number of sources")
    return sshape[0], pshape[1], npol, fshape[0], tuple(sshape[1:])


def spectral_model(stokes, spi, ref_freq, frequency, base=0):
    """Spectral model per polarisation, ``(source, chan) + stokes.shape[1:]``.

    stokes (source,) or (source, pol); spi (source, spi-comps) or (source, spi-comps, pol);
    ref_freq (source,); frequency (chan,); base "std"/0: ``stokes prod_i (nu/ref)^spi_i``,
    "log"/1: ``stokes exp(sum_i spi_i ln(nu/ref)^(i+1))``, "log10"/2 likewise with base 10, or a
    per-polarisation list of those.  Evaluated in float64; the result has the reference's dtype
    ``result_type(stokes, spi, ref_freq, frequency)``.
    """
    nsrc, nspi, npol, nchan, pol_shape = spectral_shapes(stokes, spi, ref_freq, frequency)
    bases = promote_bases(base, npol)
    arrays = (stokes, spi, ref_freq, frequency)
    out_dtype = np.result_type(*(pl.dtype_of(a) for a in arrays))
    device = pl.pick_device(*arrays)
    as_torch = pl.wants_torch(*arrays)
    f64 = np.float64
    with torch.cuda.device(device):
        d_s, d_p, d_r, d_f = (pl.to_device(a, f64, device) for a in arrays)
        d_out = pl.empty_device((nsrc, nchan) + pol_shape, f64, device)
        if npol > 0:
            pl.call("afr_spectral_model", device, pl.ptr(d_s), pl.ptr(d_p), pl.ptr(d_r), pl.ptr(d_f), bases,
                    nsrc, nspi, npol, nchan, pl.ptr(d_out), pl.stream_ptr(device))
        if out_dtype != f64:
            d_out = d_out.to(pl.torch_dtype(out_dtype))
        return d_out if as_torch else pl.to_host(d_out)
