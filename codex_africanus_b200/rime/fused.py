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


# (source, row, chan) terms from which DDE predicts in a layout without a fast kernel of its own are
# re-expressed as the complex128 2x2 time-ordered problem (below); smaller calls are launch-bound
_ADAPTER_MIN_TERMS = 1 << 24
# device bytes of re-expressed DDEs per source chunk
_ADAPTER_CHUNK_BYTES = 1 << 30


def _time_ordered(time_index):
    if pl.is_torch(time_index):
        return time_index.numel() < 2 or bool((time_index[1:] >= time_index[:-1]).all().item())
    ti = np.asarray(time_index)
    return ti.size < 2 or bool(np.all(ti[1:] >= ti[:-1]))


def _embed_2x2(x, corr_shape):
    """(..., c) diagonal Jones / coherency, c in (1, 2), or (..., 2, 2) -> complex128 (..., 2, 2)."""
    if tuple(corr_shape) == (2, 2):
        return x.to(torch.complex128)
    out = torch.zeros(tuple(x.shape[:-1]) + (2, 2), dtype=torch.complex128, device=x.device)
    out[..., 0, 0] = x[..., 0]
    if corr_shape[0] == 2:
        out[..., 1, 1] = x[..., 1]
    return out


def _predict_as_2x2_c128(lm, uvw, frequency, brightness, time_index, antenna1, antenna2, dde1_jones,
                         dde2_jones, die1_jones, base_vis, die2_jones, convention, corr_shape, out_dtype,
                         device):
    """DDE predicts whose layout has no fast kernel of its own -- diagonal Jones ((1,) / (2,)
    correlations, africanus/rime/predict.py:15-53,93-98), complex64 chains, rows that are not ordered
    by time -- through the kernels that have one: rows stably sorted by time, diagonal terms embedded
    as diagonal 2x2 matrices (the off-diagonal visibilities are exact zeros and are dropped), complex64
    promoted to complex128 (the result is rounded once, inside the 1e-5 gate by five orders), one chunk
    of sources at a time with the visibilities as the accumulator.  The GEMM kernel multiplies the
    embedded zeros too, and still runs at 4-8x the gather kernel these layouts used to get.  The DIEs
    are applied at the end in the caller's own layout and precision (``apply_gains``)."""
    from .predict import apply_gains
    f64 = np.float64
    nsrc = pl.shape_of(lm)[0]
    brightness, dde1_jones, dde2_jones = (a if pl.is_torch(a) or a is None else np.asarray(a)
                                          for a in (brightness, dde1_jones, dde2_jones))
    if dde2_jones is not dde1_jones and not pl.is_torch(dde1_jones) and dde2_jones.base is not None and \
            dde2_jones.base is dde1_jones.base and dde2_jones.shape == dde1_jones.shape and \
            dde2_jones.__array_interface__["data"] == dde1_jones.__array_interface__["data"]:
        dde2_jones = dde1_jones
    with torch.cuda.device(device):
        d_lm, d_uvw, d_f = (pl.to_device(a, f64, device) for a in (lm, uvw, frequency))
        ti, a1, a2 = normalise_indices(time_index, antenna1, antenna2, device)
        perm = None
        if not _time_ordered(time_index):
            perm = torch.argsort(ti, stable=True)
            d_uvw, ti, a1, a2 = d_uvw[perm].contiguous(), ti[perm].contiguous(), a1[perm].contiguous(), a2[perm].contiguous()
        acc = None
        if base_vis is not None:
            bv = pl.to_device(base_vis, pl.dtype_of(base_vis), device)
            acc = _embed_2x2(bv if perm is None else bv[perm], corr_shape)
        dshape = pl.shape_of(dde1_jones)
        per_source = max(1, int(np.prod(dshape[1:4])) * 64 * (1 if dde2_jones is dde1_jones else 2))
        chunk = int(max(1, min(nsrc, _ADAPTER_CHUNK_BYTES // per_source)))
        for s0 in range(0, nsrc, chunk):
            s1 = min(nsrc, s0 + chunk)
            e1 = _embed_2x2(pl.to_device(dde1_jones[s0:s1], pl.dtype_of(dde1_jones), device), corr_shape)
            e2 = e1 if dde2_jones is dde1_jones else \
                _embed_2x2(pl.to_device(dde2_jones[s0:s1], pl.dtype_of(dde2_jones), device), corr_shape)
            b = _embed_2x2(pl.to_device(brightness[s0:s1], pl.dtype_of(brightness), device), corr_shape)
            acc = fused_predict_vis(d_lm[s0:s1], d_uvw, d_f, b, ti, a1, a2, e1, e2, None, acc, None,
                                    convention=convention, dtype=np.complex128)
            del e1, e2, b
        if tuple(corr_shape) == (2, 2):
            res = acc
        elif corr_shape[0] == 2:
            res = torch.stack((acc[..., 0, 0], acc[..., 1, 1]), dim=-1)
        else:
            res = acc[..., 0, 0].unsqueeze(-1)
        res = res.to(pl.torch_dtype(out_dtype)).contiguous()
        if die1_jones is not None:
            g1 = pl.to_device(die1_jones, out_dtype, device)
            g2 = g1 if die2_jones is die1_jones else pl.to_device(die2_jones, out_dtype, device)
            res = apply_gains(ti, a1, a2, g1, res, g2)
        if perm is not None:
            out = torch.empty_like(res)
            out[perm] = res
            res = out
        return res


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
    if dde1_jones is not None and mode == _lib.AFR_JONES_DIAG and ncorr not in (1, 2, 4):
        # element-wise Jones with another correlation count (predict_vis takes any, rime/predict.py:15-53):
        # correlations are independent, so they run as blocks of 4 / 2 / 1 like the DFT kernels do
        parts, c = [], 0
        while c < ncorr:
            nc = 4 if ncorr - c >= 4 else (2 if ncorr - c >= 2 else 1)
            sl = [None if a is None else (a if pl.is_torch(a) else np.asarray(a))[..., c:c + nc]
                  for a in (brightness, dde1_jones, dde2_jones, die1_jones, base_vis, die2_jones)]
            if dde2_jones is dde1_jones:
                sl[2] = sl[1]
            if die2_jones is die1_jones:
                sl[5] = sl[3]
            parts.append(fused_predict_vis(lm, uvw, frequency, sl[0], time_index, antenna1, antenna2, sl[1], sl[2],
                                           sl[3], sl[4], sl[5], convention=convention, dtype=out_dtype))
            c += nc
        return torch.cat(parts, dim=-1) if pl.is_torch(parts[0]) else np.concatenate(parts, axis=-1)
    ntime, nant = 1, 1
    if dde1_jones is not None:
        ntime, nant = pl.shape_of(dde1_jones)[1:3]
    elif die1_jones is not None:
        ntime, nant = pl.shape_of(die1_jones)[:2]

    everything = (lm, uvw, frequency, brightness, time_index, antenna1, antenna2, dde1_jones,
                  dde2_jones, die1_jones, base_vis, die2_jones)
    device = pl.pick_device(*everything)
    as_torch = pl.wants_torch(*everything)
    if dde1_jones is not None and nsrc * nrow * nchan >= _ADAPTER_MIN_TERMS:
        diagonal = mode == _lib.AFR_JONES_DIAG and ncorr in (1, 2)
        if diagonal or out_dtype == np.complex64 or not _time_ordered(time_index):
            out = _predict_as_2x2_c128(lm, uvw, frequency, brightness, time_index, antenna1, antenna2,
                                       dde1_jones, dde2_jones, die1_jones, base_vis, die2_jones, convention,
                                       corr_shape, out_dtype, device)
            return out if as_torch else pl.to_host(out)
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
