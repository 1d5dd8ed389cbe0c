"""
Fused ``phase_delay`` (x) brightness -> ``predict_vis``.

The reference composes this from three calls and two materialised intermediates
(africanus/rime/examples/predict.py:107-134,490,522-527; the recipe is asserted in
africanus/experimental/rime/fused/tests/test_rime.py:175-209):

    K   = phase_delay(lm, uvw, frequency)                       (source,row,chan)
    X   = einsum("srf,sfij->srfij", K, brightness)              (source,row,chan,2,2)
    vis = predict_vis(time_index, antenna1, antenna2, dde1, X, dde2, die1, base_vis, die2)

Here K and X never exist: one kernel forms the phasor per (source,row,chan), multiplies
the brightness and the DDE Jones in registers and reduces over sources.
"""
import numpy as np
import torch

from .. import _lib
from .. import _plumbing as pl
from .predict import normalise_indices, predict_checks


def fused_predict_vis(lm, uvw, frequency, brightness, time_index, antenna1, antenna2,
                      dde1_jones=None, dde2_jones=None, die1_jones=None, base_vis=None,
                      die2_jones=None, convention="fourier"):
    """V[r,f] = G1 (B[r,f] + sum_s E1 (K[s,r,f] brightness[s,f]) E2^H) G2^H.

    lm (source,2), uvw (row,3), frequency (chan,) real; brightness (source,chan,corr...)
    complex with corr... in {(1,), (2,), (2,2)}; the remaining arguments are exactly those
    of ``predict_vis`` (the brightness standing in for ``source_coh`` without its row axis).
    Output dtype is ``np.result_type`` of the complex inputs: complex128 runs the chain in
    FP64, complex64 in FP32 with an FP64 phase argument.
    """
    sign = pl.convention_sign(convention)
    bshape = pl.shape_of(brightness)
    if len(bshape) not in (3, 4):
        raise ValueError("brightness.ndim %d not in (3, 4)" % len(bshape))
    nsrc, nchan = bshape[:2]
    nrow = pl.shape_of(uvw)[0]

    class _AsCoh:  # brightness viewed as source_coh (extra row axis) for the ndim rules
        shape = (nsrc, nrow) + tuple(bshape[1:])

    mode, corr_shape = predict_checks(time_index, antenna1, antenna2, dde1_jones, _AsCoh,
                                      dde2_jones, die1_jones, base_vis, die2_jones)
    if pl.shape_of(lm) != (nsrc, 2) or pl.shape_of(frequency) != (nchan,):
        raise ValueError("fused_predict_vis: lm / frequency do not match the brightness shape")
    if pl.shape_of(uvw)[1:] != (3,) or pl.shape_of(time_index) != (nrow,):
        raise ValueError("fused_predict_vis: uvw / time_index rows mismatch")
    cplx = [a for a in (brightness, dde1_jones, dde2_jones, die1_jones, base_vis, die2_jones)
            if a is not None]
    out_dtype = np.result_type(np.complex64, *(pl.dtype_of(a) for a in cplx))
    ncorr = int(np.prod(corr_shape))
    ntime, nant = 1, 1
    if dde1_jones is not None:
        ntime, nant = pl.shape_of(dde1_jones)[1:3]
    elif die1_jones is not None:
        ntime, nant = pl.shape_of(die1_jones)[:2]

    everything = (lm, uvw, frequency, brightness, time_index, antenna1, antenna2, dde1_jones,
                  dde2_jones, die1_jones, base_vis, die2_jones)
    device = pl.pick_device(*everything)
    as_torch = pl.wants_torch(*everything)
    with torch.cuda.device(device):
        f64 = np.float64
        d_lm, d_uvw, d_f = (pl.to_device(a, f64, device) for a in (lm, uvw, frequency))
        d_b = pl.to_device(brightness, out_dtype, device)
        dj = [None if a is None else pl.to_device(a, out_dtype, device)
              for a in (dde1_jones, dde2_jones, die1_jones, base_vis, die2_jones)]
        ti, a1, a2 = normalise_indices(time_index, antenna1, antenna2, device)
        d_out = pl.empty_device((nrow, nchan) + tuple(corr_shape), out_dtype, device)
        pl.call("afr_predict_fused", device, pl.ptr(d_lm), pl.ptr(d_uvw), pl.ptr(d_f),
                pl.ptr(d_b), pl.ptr(ti), pl.ptr(a1), pl.ptr(a2), *(pl.ptr(x) for x in dj),
                nsrc, nrow, ntime, nant, nchan, ncorr, mode, sign, pl.channel_mode(frequency),
                int(out_dtype == np.complex64), pl.ptr(d_out), pl.stream_ptr(device))
        return d_out if as_torch else pl.to_host(d_out)
