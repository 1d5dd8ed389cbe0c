"""
Row-block streaming driver for predicts whose output does not fit anywhere at once.

SKA-Mid scale (BASELINE configs[3]): 19,306,000 rows x 4096 channels x 2x2 complex128 is
5.06 TB of visibilities from a 100k-source catalogue.  The reference's answer is dask row
chunking in whole-timestep blocks (africanus/rime/dask_predict.py:667-726) with the per-block
results written out as they appear.  This module is that loop for one GPU: rows are cut into
whole-timestep blocks, every block is predicted from the catalogue columns by
``fused_predict_vis_stokes`` (brightness generated per source chunk on the device, DDE / DIE
arrays sliced to the block's own timesteps), and finished blocks are copied to page-locked host
buffers on a copy stream while the next block computes.  Nothing of size (row, ...) beyond one
block ever exists on the device; nothing of size (source, chan, ...) beyond one source chunk.
"""
import numpy as np
import torch

from .. import _plumbing as pl


def timestep_row_blocks(time_index, rows_per_block):
    """[(row0, row1, t0, t1), ...]: contiguous row ranges cut at timestep boundaries, each as
    many whole timesteps as fit ``rows_per_block`` rows (at least one), with the half-open range
    of ``time_index - time_index.min()`` values they span -- the (time, ...) slice of the DDE /
    DIE arrays the block needs.  Rows must be ordered by time
    (africanus/rime/dask_predict.py:696-698)."""
    ti = np.asarray(time_index.cpu() if pl.is_torch(time_index) else time_index)
    nrow = ti.shape[0]
    if nrow == 0:
        return []
    if np.any(np.diff(ti) < 0):
        raise ValueError("timestep_row_blocks: time_index must be non-decreasing")
    rows_per_block = max(1, int(rows_per_block))
    tmin = int(ti[0])
    starts = np.flatnonzero(np.concatenate(([True], ti[1:] != ti[:-1])))
    bounds = np.concatenate((starts, [nrow]))
    blocks, i, nstep = [], 0, starts.shape[0]
    while i < nstep:
        j = i + 1
        while j < nstep and bounds[j + 1] - bounds[i] <= rows_per_block:
            j += 1
        r0, r1 = int(bounds[i]), int(bounds[j])
        blocks.append((r0, r1, int(ti[r0]) - tmin, int(ti[r1 - 1]) - tmin + 1))
        i = j
    return blocks


def _time_slice(a, axis, t0, t1):
    if a is None:
        return None
    idx = [slice(None)] * len(pl.shape_of(a))
    idx[axis] = slice(t0, t1)
    return a[tuple(idx)]


def _run_blocks(predict, blocks, to_host, device, nbuf):
    """Drive ``predict(r0, r1, t0, t1)`` over the row blocks.  ``to_host``: copy each CUDA block to
    one of ``nbuf`` rotating page-locked buffers on a copy stream while the next block computes and
    yield numpy views; otherwise yield what ``predict`` returns."""
    if not to_host:
        for r0, r1, t0, t1 in blocks:
            yield (r0, r1), predict(r0, r1, t0, t1)
        return
    nbuf = max(2, int(nbuf))
    with torch.cuda.device(device):
        compute = torch.cuda.current_stream(device)
        copier = pl.side_stream(device)
        bufs = [None] * nbuf
        pending = None  # ((r0, r1), host view, copy-done event)
        most = max(b1 - b0 for b0, b1, _, _ in blocks) if blocks else 0
        for i, (r0, r1, t0, t1) in enumerate(blocks):
            vis = predict(r0, r1, t0, t1)  # CUDA tensor: the per-source inputs are CUDA tensors
            k = i % nbuf
            if bufs[k] is None or bufs[k].numel() < vis.numel() or bufs[k].dtype != vis.dtype:
                bufs[k] = pl.empty_pinned((most * int(np.prod(vis.shape[1:])),), pl.dtype_of(vis))
            host = bufs[k][: vis.numel()].view(vis.shape)
            ready = torch.cuda.Event()
            ready.record(compute)
            copier.wait_event(ready)
            with torch.cuda.stream(copier):
                host.copy_(vis, non_blocking=True)
                vis.record_stream(copier)
                done = torch.cuda.Event()
                done.record(copier)
            if pending is not None:
                pending[2].synchronize()
                yield pending[0], pending[1].numpy()
            pending = ((r0, r1), host, done)
        if pending is not None:
            pending[2].synchronize()
            yield pending[0], pending[1].numpy()


def _stream(local_fn, cuda_path, per_source, f64_mask, uvw, time_index, antenna1, antenna2, dde1_jones,
            dde2_jones, die1_jones, base_vis, die2_jones, rows_per_block, block_bytes, ncorr, nchan,
            nbuf, kwargs, e_axis=1):
    """Shared body of the streaming generators: ``per_source`` are the arrays indexed by source /
    channel only (uploaded once on the CUDA path; ``f64_mask`` says which are real float64).  The
    two ``dde`` slots are sliced along ``e_axis`` to the block's timesteps, the DIEs along axis 0."""
    nrow = pl.shape_of(uvw)[0]
    if pl.shape_of(time_index) != (nrow,):
        raise ValueError("stream predict: uvw / time_index rows mismatch")
    if rows_per_block is None:
        rows_per_block = max(1, int(block_bytes) // max(1, nchan * ncorr * 16))
    blocks = timestep_row_blocks(time_index, rows_per_block)
    same_dde = dde1_jones is dde2_jones
    same_die = die1_jones is die2_jones
    every = tuple(per_source) + (uvw, time_index, antenna1, antenna2, dde1_jones, dde2_jones,
                                 die1_jones, base_vis, die2_jones)
    to_host = cuda_path and not pl.wants_torch(*every)
    device = None
    if cuda_path:  # per-source inputs go to the device once, not once per block
        device = pl.pick_device(*every)
        with torch.cuda.device(device):
            per_source = tuple(pl.to_device(a, np.float64 if is_f64 else pl.dtype_of(a), device)
                               for a, is_f64 in zip(per_source, f64_mask))

    def predict(r0, r1, t0, t1):
        e1 = _time_slice(dde1_jones, e_axis, t0, t1)
        e2 = e1 if same_dde else _time_slice(dde2_jones, e_axis, t0, t1)
        g1 = _time_slice(die1_jones, 0, t0, t1)
        g2 = g1 if same_die else _time_slice(die2_jones, 0, t0, t1)
        return local_fn(per_source, uvw[r0:r1], time_index[r0:r1], antenna1[r0:r1], antenna2[r0:r1],
                        e1, e2, g1, None if base_vis is None else base_vis[r0:r1], g2, kwargs)

    return _run_blocks(predict, blocks, to_host, device, nbuf)


def stream_predict_vis_stokes(lm, uvw, frequency, stokes, spi, ref_freq, time_index, antenna1,
                              antenna2, dde1_jones=None, dde2_jones=None, die1_jones=None,
                              base_vis=None, die2_jones=None, rows_per_block=None,
                              block_bytes=1 << 30, nbuf=3, local_fn=None, **kwargs):
    """Generator over whole-timestep row blocks: yields ``((row0, row1), vis_block)`` with
    ``vis_block = fused_predict_vis_stokes(...)`` of those rows, in row order.

    Arguments as ``fused_predict_vis_stokes`` (``kwargs``: convention, base, corr_schema, dtype,
    source_chunk, ...).  ``rows_per_block`` defaults to what fits ``block_bytes`` of output.
    With numpy inputs the blocks are numpy views of ``nbuf`` rotating page-locked buffers: a
    yielded block stays valid until ``nbuf - 1`` further blocks have been requested (consume or
    copy it before that); the device -> host copy of block i overlaps the computation of block
    i + 1.  With CUDA-tensor inputs the blocks are fresh CUDA tensors.  ``local_fn`` replaces the
    per-block predict (CPU tests substitute the oracle; the default is always the CUDA path).
    """
    cuda_path = local_fn is None
    if cuda_path:
        from .fused_stokes import fused_predict_vis_stokes as local_fn
    fn = local_fn

    def block_fn(ps, uvw_b, ti_b, a1_b, a2_b, e1, e2, g1, bv, g2, kw):
        d_lm, d_f, d_st, d_spi, d_rf = ps
        return fn(d_lm, uvw_b, d_f, d_st, d_spi, d_rf, ti_b, a1_b, a2_b, e1, e2, g1, bv, g2, **kw)

    ncorr = int(np.asarray(kwargs.get("corr_schema", [[0, 0], [0, 0]]), dtype=object).size)
    return _stream(block_fn, cuda_path, (lm, frequency, stokes, spi, ref_freq), (True,) * 5, uvw, time_index,
                   antenna1, antenna2, dde1_jones, dde2_jones, die1_jones, base_vis, die2_jones,
                   rows_per_block, block_bytes, ncorr, pl.shape_of(frequency)[0], nbuf, kwargs)


def stream_fused_predict_vis(lm, uvw, frequency, brightness, time_index, antenna1, antenna2,
                             dde1_jones=None, dde2_jones=None, die1_jones=None, base_vis=None,
                             die2_jones=None, convention="fourier", rows_per_block=None,
                             block_bytes=1 << 30, nbuf=3, local_fn=None):
    """The same streaming loop for ``fused_predict_vis`` (brightness given as a
    (source, chan, corr...) array, uploaded once): yields ``((row0, row1), vis_block)``."""
    cuda_path = local_fn is None
    if cuda_path:
        from .fused import fused_predict_vis as local_fn
    fn = local_fn

    def block_fn(ps, uvw_b, ti_b, a1_b, a2_b, e1, e2, g1, bv, g2, kw):
        d_lm, d_f, d_b = ps
        return fn(d_lm, uvw_b, d_f, d_b, ti_b, a1_b, a2_b, e1, e2, g1, bv, g2, **kw)

    ncorr = int(np.prod(pl.shape_of(brightness)[2:], dtype=np.int64))
    return _stream(block_fn, cuda_path, (lm, frequency, brightness), (True, True, False), uvw, time_index,
                   antenna1, antenna2, dde1_jones, dde2_jones, die1_jones, base_vis, die2_jones,
                   rows_per_block, block_bytes, ncorr, pl.shape_of(frequency)[0], nbuf,
                   {"convention": convention})


def stream_fused_predict_vis_beam(lm, uvw, frequency, brightness, time_index, antenna1, antenna2,
                                  beam, beam_lm_extents, beam_freq_map, parallactic_angles,
                                  point_errors, antenna_scaling, die1_jones=None, base_vis=None,
                                  die2_jones=None, rows_per_block=None, block_bytes=1 << 30, nbuf=3,
                                  local_fn=None, **kwargs):
    """The streaming loop for ``fused_predict_vis_beam`` (DDEs interpolated from the beam cube per
    source chunk): the (time, ant, ...) inputs of the beam -- parallactic angles and pointing
    errors -- and the DIEs are sliced to each block's timesteps; beam, brightness and the other
    per-source / per-antenna inputs go to the device once.  ``kwargs``: convention, source_chunk,
    feed_type.  Yields ``((row0, row1), vis_block)``."""
    cuda_path = local_fn is None
    if cuda_path:
        from .fused_beam import fused_predict_vis_beam as local_fn
    fn = local_fn
    pa, pe = parallactic_angles, point_errors

    def block_fn(ps, uvw_b, ti_b, a1_b, a2_b, pa_b, pe_b, g1, bv, g2, kw):
        d_lm, d_f, d_b, d_beam, d_ext, d_bfm, d_as = ps
        return fn(d_lm, uvw_b, d_f, d_b, ti_b, a1_b, a2_b, d_beam, d_ext, d_bfm, pa_b, pe_b, d_as, g1, bv, g2,
                  **kw)

    ncorr = int(np.prod(pl.shape_of(brightness)[2:], dtype=np.int64))
    return _stream(block_fn, cuda_path,
                   (lm, frequency, brightness, beam, beam_lm_extents, beam_freq_map, antenna_scaling),
                   (True, True, False, False, True, True, True), uvw, time_index, antenna1, antenna2,
                   pa, pe, die1_jones, base_vis, die2_jones, rows_per_block, block_bytes, ncorr,
                   pl.shape_of(frequency)[0], nbuf, kwargs, e_axis=0)
