"""
Dask-facing ``phase_delay``, ``beam_cube_dde`` and ``predict_vis(streams=)`` -- the graphs of
africanus/rime/dask.py:34-52,166-213 and africanus/rime/dask_predict.py:311-593 with the B200 kernels
as block functions: same signatures, chunk checks, ``blockwise`` index strings (``row`` substituted
for ``time`` in the Jones arrays, paired by block position) and ``streams`` semantics
(``False``: one partial visibility array per source chunk, tree-summed; ``True``: the source chunks
of a (row, chan) block are folded one after the other with ``base_vis`` as the accumulator, the
``LinearReduction`` of dask_predict.py:64-178).  Inputs: ``dask.array.Array`` (lazy; needs dask) or
``_chunked.ChunkedArray`` (eager; no dask needed).  One GPU per worker thread.
"""
import itertools

import numpy as np

from .. import _chunked as ck
from .. import _plumbing as pl
from ..dft.dask import _blockwise_for
from .fast_beam_cubes import beam_cube_dde as np_beam_cube_dde
from .phase import phase_delay as np_phase_delay
from .predict import predict_checks
from .predict import predict_vis as np_predict_vis


@ck.on_worker_device
def _phase_delay_wrap(lm, uvw, frequency, convention):
    return np_phase_delay(lm[0], uvw[0], frequency, convention=convention)


def phase_delay(lm, uvw, frequency, convention="fourier"):
    """Dask wrapper for phase_delay (africanus/rime/dask.py:38-52)."""
    return _blockwise_for(lm, uvw, frequency)(
        _phase_delay_wrap, ("source", "row", "chan"),
        lm, ("source", "(l,m)"),
        uvw, ("row", "(u,v,w)"),
        frequency, ("chan",),
        convention=convention,
        dtype=np.result_type(np.complex64, lm.dtype, uvw.dtype, frequency.dtype))


@ck.on_worker_device
def _beam_cube_dde_wrapper(beam, beam_lm_extents, beam_freq_map, lm, parallactic_angles, point_errors,
                           antenna_scaling, frequencies):
    return np_beam_cube_dde(beam[0][0][0], beam_lm_extents[0][0], beam_freq_map[0], lm[0], parallactic_angles,
                            point_errors[0], antenna_scaling[0], frequencies)


def beam_cube_dde(beam, beam_lm_extents, beam_freq_map, lm, parallactic_angles, point_errors, antenna_scaling,
                  frequencies):
    """Dask wrapper for beam_cube_dde (africanus/rime/dask.py:166-213)."""
    if not all(len(c) == 1 for c in beam.chunks):
        raise ValueError("Beam chunking unsupported")
    if not all(len(c) == 1 for c in beam_freq_map.chunks):
        raise ValueError("Beam frequency map chunking unsupported")
    if not all(len(c) == 1 for c in beam_lm_extents.chunks):
        raise ValueError("Chunking of beam_lm_extents unsupported")
    corr_dims = tuple("corr-%d" % i for i in range(len(beam.shape[3:])))
    return _blockwise_for(beam, beam_lm_extents, beam_freq_map, lm, parallactic_angles, point_errors,
                          antenna_scaling, frequencies)(
        _beam_cube_dde_wrapper, ("source", "time", "ant", "chan") + corr_dims,
        beam, ("beam-lw", "beam-mh", "beam-nud") + corr_dims,
        beam_lm_extents, ("beam-lm", "beam-ext"),
        beam_freq_map, ("beam-nud",),
        lm, ("source", "source-comp"),
        parallactic_angles, ("time", "ant"),
        point_errors, ("time", "ant", "chan", "pt-comp"),
        antenna_scaling, ("ant", "chan", "scale-comp"),
        frequencies, ("chan",),
        dtype=beam.dtype)


# ------------------------------------------------------------------------------------------ predict_vis
@ck.on_worker_device
def _predict_coh_wrapper(time_index, antenna1, antenna2, dde1_jones, source_coh, dde2_jones, base_vis,
                         reduce_single_source=False):
    # dask_predict.py:257-290
    if reduce_single_source:
        dde1_jones = dde1_jones[0] if dde1_jones else None
        source_coh = source_coh[0] if source_coh else None
        dde2_jones = dde2_jones[0] if dde2_jones else None
    # the DDE arrays contract over a single 'ant' chunk
    vis = np_predict_vis(time_index, antenna1, antenna2, dde1_jones[0] if dde1_jones is not None else None,
                         source_coh, dde2_jones[0] if dde2_jones is not None else None, None, base_vis, None)
    return vis if reduce_single_source else vis[None, ...]


@ck.on_worker_device
def _predict_dies_wrapper(time_index, antenna1, antenna2, die1_jones, base_vis, die2_jones):
    # dask_predict.py:293-308: the DIE arrays lose their single 'ant' chunk
    return np_predict_vis(time_index, antenna1, antenna2, None, None, None,
                          die1_jones[0] if die1_jones is not None else None, base_vis,
                          die2_jones[0] if die2_jones is not None else None)


def _cdims(a, first):
    return tuple("corr-%d" % i for i in range(len(a.shape[first:])))


def _parallel_reduction(time_index, antenna1, antenna2, dde1_jones, source_coh, dde2_jones, out_dtype):
    """One (1, row, chan, corr...) partial per source chunk, summed over the source-chunk axis
    (dask_predict.py:311-369)."""
    cdims = _cdims(dde1_jones, 4) if dde1_jones is not None else _cdims(source_coh, 3)
    ajones_dims = ("src", "row", "ant", "chan") + cdims
    src_coh_dims = ("src", "row", "chan") + cdims
    nsrc_blocks = len((dde1_jones if dde1_jones is not None else source_coh).chunks[0])
    coherencies = _blockwise_for(time_index, antenna1, antenna2, dde1_jones, source_coh, dde2_jones)(
        _predict_coh_wrapper, src_coh_dims,
        time_index, ("row",), antenna1, ("row",), antenna2, ("row",),
        dde1_jones, None if dde1_jones is None else ajones_dims,
        source_coh, None if source_coh is None else src_coh_dims,
        dde2_jones, None if dde2_jones is None else ajones_dims,
        None, None,
        # time + row chunks are equivalent but differently sized: pair blocks by position and give
        # the output the row chunking
        align_arrays=False,
        adjust_chunks={"row": time_index.chunks[0], "src": (1,) * nsrc_blocks},
        meta=np.empty((0,) * len(src_coh_dims), dtype=out_dtype), dtype=out_dtype)
    return coherencies.sum(axis=0)


def _linear_reduction_eager(time_index, antenna1, antenna2, dde1_jones, source_coh, dde2_jones, out_dtype):
    """streams=True over ChunkedArrays: per (row, chan) block the source chunks are folded in
    order, the running sum travelling as ``base_vis`` (dask_predict.py:216-239, feed_index = 7)."""
    lead = dde1_jones if dde1_jones is not None else source_coh
    first = 4 if dde1_jones is not None else 3
    chan_axis = 3 if dde1_jones is not None else 2
    row_chunks, chan_chunks = time_index.chunks[0], lead.chunks[chan_axis]
    corr_shape = tuple(lead.shape[first:])
    out = np.empty((sum(row_chunks), sum(chan_chunks)) + corr_shape, out_dtype)
    zc = (0,) * len(corr_shape)
    for r, f in itertools.product(range(len(row_chunks)), range(len(chan_chunks))):
        acc = None
        for s in range(len(lead.chunks[0])):
            e1 = None if dde1_jones is None else [dde1_jones.block((s, r, 0, f) + zc)]
            e2 = None if dde2_jones is None else [dde2_jones.block((s, r, 0, f) + zc)]
            coh = None if source_coh is None else source_coh.block((s, r, f) + zc)
            acc = _predict_coh_wrapper(time_index.block((r,)), antenna1.block((r,)), antenna2.block((r,)),
                                       e1, coh, e2, acc)[0]
        r0, f0 = sum(row_chunks[:r]), sum(chan_chunks[:f])
        out[r0:r0 + row_chunks[r], f0:f0 + chan_chunks[f]] = acc
    return ck.ChunkedArray(out, (row_chunks, chan_chunks) + tuple((c,) for c in corr_shape))


def _linear_reduction_dask(time_index, antenna1, antenna2, dde1_jones, source_coh, dde2_jones, out_dtype):
    """streams=True as a dask graph: a materialised layer whose task for source chunk s of a
    (row, chan) block takes the task of chunk s - 1 as its ``base_vis``."""
    from dask.base import tokenize
    from dask.highlevelgraph import HighLevelGraph

    da = ck.da
    lead = dde1_jones if dde1_jones is not None else source_coh
    first = 4 if dde1_jones is not None else 3
    chan_axis = 3 if dde1_jones is not None else 2
    row_chunks, chan_chunks = time_index.chunks[0], lead.chunks[chan_axis]
    corr_shape = tuple(lead.shape[first:])
    zc = (0,) * len(corr_shape)
    nsrc = len(lead.chunks[0])
    token = tokenize(time_index, antenna1, antenna2, dde1_jones, source_coh, dde2_jones)
    name = "predict-vis-stream-" + token
    layer = {}
    for r, f in itertools.product(range(len(row_chunks)), range(len(chan_chunks))):
        prev = None
        for s in range(nsrc):
            key = (name, r, f) + zc if s == nsrc - 1 else (name + "-partial-%d" % s, r, f) + zc
            e1 = None if dde1_jones is None else [(dde1_jones.name, s, r, 0, f) + zc]
            e2 = None if dde2_jones is None else [(dde2_jones.name, s, r, 0, f) + zc]
            coh = None if source_coh is None else (source_coh.name, s, r, f) + zc
            layer[key] = (_predict_coh_fold, (time_index.name, r), (antenna1.name, r), (antenna2.name, r),
                          e1, coh, e2, prev)
            prev = key
    deps = [a for a in (time_index, antenna1, antenna2, dde1_jones, source_coh, dde2_jones) if a is not None]
    graph = HighLevelGraph.from_collections(name, layer, dependencies=deps)
    chunks = (row_chunks, chan_chunks) + tuple((c,) for c in corr_shape)
    return da.Array(graph, name, chunks, dtype=out_dtype)


def _predict_coh_fold(time_index, antenna1, antenna2, dde1, coh, dde2, base_vis):
    return _predict_coh_wrapper(time_index, antenna1, antenna2, dde1, coh, dde2, base_vis)[0]


def _apply_dies(time_index, antenna1, antenna2, die1_jones, base_vis, die2_jones, out_dtype):
    # dask_predict.py:372-440
    cdims = _cdims(die1_jones, 3) if die1_jones is not None else _cdims(base_vis, 2)
    gjones_dims = ("row", "ant", "chan") + cdims
    vis_dims = ("row", "chan") + cdims
    return _blockwise_for(time_index, antenna1, antenna2, die1_jones, base_vis, die2_jones)(
        _predict_dies_wrapper, vis_dims,
        time_index, ("row",), antenna1, ("row",), antenna2, ("row",),
        die1_jones, None if die1_jones is None else gjones_dims,
        base_vis, None if base_vis is None else vis_dims,
        die2_jones, None if die2_jones is None else gjones_dims,
        align_arrays=False, adjust_chunks={"row": time_index.chunks[0]},
        meta=np.empty((0,) * len(vis_dims), dtype=out_dtype), dtype=out_dtype)


def predict_vis(time_index, antenna1, antenna2, dde1_jones=None, source_coh=None, dde2_jones=None,
                die1_jones=None, base_vis=None, die2_jones=None, streams=None):
    """Dask wrapper for predict_vis (africanus/rime/dask_predict.py:442-593)."""
    predict_checks(time_index, antenna1, antenna2, dde1_jones, source_coh, dde2_jones, die1_jones, base_vis,
                   die2_jones)
    have_ddes = dde1_jones is not None and dde2_jones is not None
    have_dies = die1_jones is not None and die2_jones is not None
    have_coh, have_bvis = source_coh is not None, base_vis is not None
    if have_ddes:
        for a in (dde1_jones, dde2_jones):
            if a.shape[2] != a.chunks[2][0]:
                raise ValueError("Subdivision of antenna dimension into multiple chunks is not supported.")
        if dde1_jones.chunks != dde2_jones.chunks:
            raise ValueError("dde1_jones.chunks != dde2_jones.chunks")
        if len(dde1_jones.chunks[1]) != len(time_index.chunks[0]):
            raise ValueError("Number of row chunks (%s) does not equal number of time chunks (%s)."
                             % (time_index.chunks[0], dde1_jones.chunks[1]))
    if have_dies:
        for a in (die1_jones, die2_jones):
            if a.shape[1] != a.chunks[1][0]:
                raise ValueError("Subdivision of antenna dimension into multiple chunks is not supported.")
        if die1_jones.chunks != die2_jones.chunks:
            raise ValueError("die1_jones.chunks != die2_jones.chunks")
        if len(die1_jones.chunks[0]) != len(time_index.chunks[0]):
            raise ValueError("Number of row chunks (%s) does not equal number of time chunks (%s)."
                             % (time_index.chunks[0], die1_jones.chunks[1]))
    # the dask wrapper leaves base_vis out of the dtype inference (dask_predict.py:527-530)
    out_dtype = np.result_type(*(np.dtype(a.dtype.name) for a in
                                 (dde1_jones, source_coh, dde2_jones, die1_jones, die2_jones) if a is not None))
    sum_coherencies = None
    if have_coh or have_ddes:
        if streams is True:
            red = (_linear_reduction_dask if any(ck.is_dask(a) for a in (time_index, dde1_jones, source_coh))
                   else _linear_reduction_eager)
            sum_coherencies = red(time_index, antenna1, antenna2, dde1_jones, source_coh, dde2_jones, out_dtype)
        else:
            sum_coherencies = _parallel_reduction(time_index, antenna1, antenna2, dde1_jones, source_coh,
                                                  dde2_jones, out_dtype)
    else:
        assert have_dies or have_bvis
    if not have_dies and not have_bvis:
        return sum_coherencies
    if sum_coherencies is not None:
        base_vis = sum_coherencies if not have_bvis else base_vis + sum_coherencies
    return _apply_dies(time_index, antenna1, antenna2, die1_jones, base_vis, die2_jones, out_dtype)
