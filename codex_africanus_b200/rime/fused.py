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

# output bytes above which the numpy path streams row blocks back while computing
_ROW_BLOCK_BYTES = 512 << 20


def fused_predict_vis(lm, uvw, frequency, brightness, time_index, antenna1, antenna2,
                      dde1_jones=None, dde2_jones=None, die1_jones=None, base_vis=None,
                      die2_jones=None, convention="fourier", dtype=None):
    """V[r,f] = G1 (B[r,f] + sum_s E1 (K[s,r,f] brightness[s,f]) E2^H) G2^H.

    lm (source,2), uvw (row,3), frequency (chan,) real; brightness (source,chan,corr...)
    complex with corr... in {(1,), (2,), (2,2)}; the remaining arguments are exactly those
    of ``predict_vis`` (the brightness standing in for ``source_coh`` without its row axis).
    Output dtype: that of the reference's un-fused chain -- ``phase_delay`` gives
    ``result_type(complex64, lm, uvw, frequency)`` (rime/phase.py:26), the einsum with the brightness
    and ``predict_vis`` promote further (rime/predict.py:542-544) -- so float64 coordinates give
    complex128 whatever the Jones precision.  ``dtype=np.complex64`` asks for the single-precision
    chain explicitly (FP32 Jones products and accumulators; the phase argument stays FP64).
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
    if dtype is None:
        out_dtype = np.result_type(np.complex64, *(pl.dtype_of(a) for a in (lm, uvw, frequency)),
                                   *(pl.dtype_of(a) for a in cplx))
    else:
        out_dtype = np.dtype(dtype)
        if out_dtype not in (np.complex64, np.complex128):
            raise TypeError("fused_predict_vis: dtype must be complex64 or complex128")
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
    chan_mode = pl.channel_mode(frequency)
    c64 = int(out_dtype == np.complex64)
    with torch.cuda.device(device):
        f64 = np.float64
        d_lm, d_uvw, d_f = (pl.to_device(a, f64, device) for a in (lm, uvw, frequency))
        d_b = pl.to_device(brightness, out_dtype, device)
        d_dde1, d_dde2, d_die1, d_die2 = (
            None if a is None else pl.to_device(a, out_dtype, device)
            for a in (dde1_jones, dde2_jones, die1_jones, die2_jones))
        if dde1_jones is dde2_jones:
            d_dde2 = d_dde1  # one upload; also lets the kernel stage a single copy per tile
        ti, a1, a2 = normalise_indices(time_index, antenna1, antenna2, device)
        out_shape = (nrow, nchan) + tuple(corr_shape)

        def launch(r0, r1, d_bvis, d_out):
            pl.call("afr_predict_fused", device, pl.ptr(d_lm), pl.ptr(d_uvw[r0:r1]), pl.ptr(d_f),
                    pl.ptr(d_b), pl.ptr(ti[r0:r1]), pl.ptr(a1[r0:r1]), pl.ptr(a2[r0:r1]),
                    pl.ptr(d_dde1), pl.ptr(d_dde2), pl.ptr(d_die1), pl.ptr(d_bvis), pl.ptr(d_die2),
                    nsrc, r1 - r0, ntime, nant, nchan, ncorr, mode, sign, chan_mode, c64,
                    pl.ptr(d_out), pl.stream_ptr(device))

        row_bytes = max(1, nchan * ncorr * out_dtype.itemsize)
        if as_torch or nrow * row_bytes <= _ROW_BLOCK_BYTES or nrow == 0:
            d_bvis = None if base_vis is None else pl.to_device(base_vis, out_dtype, device)
            d_out = pl.empty_device(out_shape, out_dtype, device)
            if nrow:
                launch(0, nrow, d_bvis, d_out)
            return d_out if as_torch else pl.to_host(d_out)

        # numpy path, large output: rows are independent, so stream row blocks -- the block's
        # base_vis goes up and its visibilities come back on a copy stream while the next
        # block computes (the reference's own answer to configs-3-sized outputs is row
        # chunking, africanus/rime/dask_predict.py:667-726)
        sink = pl.RowSink(out_shape, out_dtype, device)
        block = max(1024, _ROW_BLOCK_BYTES // row_bytes)
        if not sink.whole:
            block = min(block, sink.max_block_rows())
        compute = torch.cuda.current_stream(device)
        bufs = [pl.empty_device((min(block, nrow), nchan) + tuple(corr_shape), out_dtype, device)
                for _ in range(2)]
        done = [None, None]
        bv = None if base_vis is None else np.asarray(base_vis)
        for i, r0 in enumerate(range(0, nrow, block)):
            r1 = min(nrow, r0 + block)
            buf = bufs[i & 1][: r1 - r0]
            if done[i & 1] is not None:
                compute.wait_event(done[i & 1])  # the copy out of this buffer has finished
            d_bvis = None if bv is None else pl.to_device(bv[r0:r1], out_dtype, device)
            launch(r0, r1, d_bvis, buf)
            done[i & 1] = sink.push(r0, r1, buf, compute)
        out = sink.finish()
        compute.synchronize()
        return out
