"""
Fused predict with the DDE Jones interpolated from a beam cube, in source chunks.

The reference's predict example (africanus/rime/examples/predict.py:390-401,469-472,
522-527) materialises ``beam_cube_dde(...)`` for every source -- a
``(source, time, ant, chan, 2, 2)`` array, 16.8 GB per timestep at 1000 sources x 64
antennas x 4096 channels -- and hands it to ``predict_vis``.  Here the beam is sampled
for one chunk of sources at a time and that chunk is reduced into the visibilities
straight away (``base_vis`` doubles as the accumulator, the streaming idiom of
africanus/rime/dask_predict.py:216-239), so the DDE array never exists in full: device
memory holds one chunk, whatever the number of sources or timesteps.  The DIEs are
applied once, after the last chunk.
"""
import numpy as np
import torch

from .. import _plumbing as pl
from .fast_beam_cubes import beam_cube_dde_rotated
from .feeds import feed_rotation
from .fused import fused_predict_vis
from .predict import apply_gains

# device bytes one chunk of interpolated DDEs may occupy: at most 16 GiB and a quarter of the free device
# memory (a MeerKAT chunk of 4 GiB is 64 sources, where the GEMM kernel's pipeline fill and the
# accumulator pass per chunk still cost 10 %; 256 sources per chunk: 216 -> 2xx Gterms/s on configs[2])
_DDE_CHUNK_BYTES = 16 << 30


def _chunk_budget(device):
    try:
        free, _ = torch.cuda.mem_get_info(device)
    except Exception:  # pragma: no cover
        free = 4 * _DDE_CHUNK_BYTES
    return int(max(256 << 20, min(_DDE_CHUNK_BYTES, free // 4)))


# device bytes of plane-reduced beam per launch of the in-kernel sampling route
_PLANES_CHUNK_BYTES = 4 << 30


def _predict_sampling_in_kernel(d_lm, d_uvw, d_f, d_b, d_ti, d_a1, d_a2, d_beam, d_ext, d_bfm, d_pa, d_pe, d_as,
                                die1_jones, acc, die2_jones, convention, device, d_rot=None):
    """The in-kernel sampling route of ``fused_predict_vis_beam`` on device tensors; None when it does
    not apply (nothing was computed)."""
    import ctypes

    from .predict import normalise_indices
    nsrc = d_lm.shape[0]
    nrow = d_uvw.shape[0]
    ntime, nant = d_pa.shape
    nchan = d_f.shape[0]
    lw, mh, nud = d_beam.shape[:3]
    sign = pl.convention_sign(convention)
    ti, a1, a2 = normalise_indices(d_ti, d_a1, d_a2, device)
    c128 = np.complex128
    g1 = None if die1_jones is None else pl.to_device(die1_jones, c128, device)
    g2 = g1 if die2_jones is die1_jones else (None if die2_jones is None else pl.to_device(die2_jones, c128, device))
    per_source = max(1, ntime * nant * nud * 96)
    chunk = int(max(1, min(nsrc, _PLANES_CHUNK_BYTES // per_source)))
    fd = pl.empty_device((nchan, 3), np.float64, device)
    ok = torch.zeros(1, dtype=torch.int32, device=device)
    used = ctypes.c_int(0)
    out = None
    for s0 in range(0, nsrc, chunk):
        s1 = min(nsrc, s0 + chunk)
        planes = pl.empty_device((s1 - s0, ntime, nant, nud, 12), np.float64, device)
        pl.call("afr_beam_plane_reduce", device, pl.ptr(d_beam), pl.ptr(d_ext), pl.ptr(d_bfm), pl.ptr(d_lm[s0:s1]),
                pl.ptr(d_pa), pl.ptr(d_pe), pl.ptr(d_as), pl.ptr(d_f), lw, mh, nud, s1 - s0, ntime, nant, nchan,
                pl.ptr(planes), pl.ptr(fd), pl.ptr(ok), pl.stream_ptr(device))
        if s0 == 0 and int(ok.item()) == 0:
            return None  # some channel / row has its own grid position: the element kernel's business
        last = s1 == nsrc
        nxt = pl.empty_device((nrow, nchan, 2, 2), c128, device)
        pl.call("afr_predict_fused_planes", device, pl.ptr(d_lm[s0:s1]), pl.ptr(d_uvw), pl.ptr(d_f),
                pl.ptr(d_b[s0:s1]), pl.ptr(ti), pl.ptr(a1), pl.ptr(a2), pl.ptr(planes), pl.ptr(fd), nud, pl.ptr(d_rot),
                pl.ptr(g1 if last else None), pl.ptr(acc if out is None else out), pl.ptr(g2 if last else None),
                s1 - s0, nrow, ntime, nant, nchan, sign, ctypes.byref(used), pl.ptr(nxt), pl.stream_ptr(device))
        if not used.value:
            return None  # (decided on the first chunk: the admission test does not depend on the sources' count)
        out = nxt
        del planes
    return out


def fused_predict_vis_beam(lm, uvw, frequency, brightness, time_index, antenna1, antenna2,
                           beam, beam_lm_extents, beam_freq_map, parallactic_angles,
                           point_errors, antenna_scaling, die1_jones=None, base_vis=None,
                           die2_jones=None, convention="fourier", source_chunk=None, feed_type=None,
                           in_kernel=False):
    """``fused_predict_vis(..., dde1 = dde2 = beam_cube_dde(beam, ..., lm, ...))`` without the
    full DDE array.  Arguments: those of ``fused_predict_vis`` with the two DDE arrays replaced
    by the arguments of ``beam_cube_dde``; ``source_chunk`` (sources per chunk) defaults to
    what fits ``_DDE_CHUNK_BYTES``.  ``feed_type`` "linear" / "circular" additionally multiplies
    the beam by the feed rotation of the same parallactic angles, ``dde = beam_dde . L[t,a]``
    (africanus/rime/examples/predict.py:469-472, rime/feeds.py:13-48), inside the interpolation
    kernel.  ``in_kernel=True`` (2x2 complex128) samples the beam INSIDE the predict
    kernel (SURVEY 8f-1 proper; the reference's fused RIME does the same,
    experimental/rime/fused/terms/cube_dde.py:96-313): a pre-pass reduces the four spatial corners of
    every frequency plane per (source, time, antenna) -- 43x smaller than the Jones array at 4096
    channels -- and the kernel's producers combine the two planes of each channel; no
    (source, time, ant, chan, 2, 2) array exists, not even per chunk.  It applies when pointing errors and
    antenna scaling are constant along the channel axis, every channel lies inside the cube's frequency
    range and the baseline uvw are differences of antenna coordinates; otherwise (and by default: the
    chunked route through the GEMM kernel is 8 % faster) the beam is sampled per source chunk.
    Returns (row, chan, corr...) like ``predict_vis``."""
    if (die1_jones is None) != (die2_jones is None):
        raise ValueError("Both die1_jones and die2_jones must be present or absent")
    bshape = pl.shape_of(beam)
    if len(bshape) < 3:
        raise ValueError("beam must have at least 3 dimensions")
    nsrc = pl.shape_of(lm)[0]
    if pl.shape_of(brightness)[0] != nsrc:
        raise ValueError("fused_predict_vis_beam: lm / brightness disagree on the number of sources")
    ntime, nant = pl.shape_of(parallactic_angles)
    nchan = pl.shape_of(frequency)[0]
    ncorr = int(np.prod(bshape[3:])) if len(bshape) > 3 else 1
    bdt = pl.dtype_of(beam)
    per_source = max(1, ntime * nant * nchan * ncorr * bdt.itemsize)
    everything = (lm, uvw, frequency, brightness, time_index, antenna1, antenna2, beam,
                  beam_lm_extents, beam_freq_map, parallactic_angles, point_errors, antenna_scaling,
                  die1_jones, base_vis, die2_jones)
    device = pl.pick_device(*everything)
    if source_chunk is None:
        source_chunk = max(1, _chunk_budget(device) // per_source)
    source_chunk = int(max(1, min(source_chunk, max(nsrc, 1))))
    as_torch = pl.wants_torch(*everything)
    f64 = np.float64
    cplx = [a for a in (brightness, beam, die1_jones, base_vis, die2_jones) if a is not None]
    out_dtype = np.result_type(np.complex64, *(pl.dtype_of(a) for a in cplx))
    with torch.cuda.device(device):
        # everything to the device once; the public entry points then run tensor -> tensor
        d_lm, d_uvw, d_f = (pl.to_device(a, f64, device) for a in (lm, uvw, frequency))
        d_b = pl.to_device(brightness, out_dtype, device)
        d_beam = pl.to_device(beam, bdt, device)
        d_ext, d_bfm, d_pa, d_pe, d_as = (pl.to_device(a, f64, device) for a in (
            beam_lm_extents, beam_freq_map, parallactic_angles, point_errors, antenna_scaling))
        d_ti = time_index if pl.is_torch(time_index) else torch.from_numpy(np.ascontiguousarray(time_index))
        d_a1 = antenna1 if pl.is_torch(antenna1) else torch.from_numpy(np.ascontiguousarray(antenna1))
        d_a2 = antenna2 if pl.is_torch(antenna2) else torch.from_numpy(np.ascontiguousarray(antenna2))
        d_ti, d_a1, d_a2 = (x.to(device) for x in (d_ti, d_a1, d_a2))
        acc = None if base_vis is None else pl.to_device(base_vis, out_dtype, device)
        d_rot = None
        if feed_type is not None:
            d_rot = feed_rotation(d_pa, feed_type)
            if pl.dtype_of(d_rot) != bdt:
                d_rot = d_rot.to(pl.torch_dtype(bdt))
        if in_kernel and tuple(bshape[3:]) == (2, 2) and bdt == np.complex128 and \
                out_dtype == np.complex128 and nsrc > 0 and pl.shape_of(uvw)[0] > 0:
            res = _predict_sampling_in_kernel(d_lm, d_uvw, d_f, d_b, d_ti, d_a1, d_a2, d_beam, d_ext, d_bfm, d_pa,
                                              d_pe, d_as, die1_jones, acc, die2_jones, convention, device,
                                              None if d_rot is None else d_rot.contiguous())
            if res is not None:
                return res if as_torch else pl.to_host(res)
        for s0 in range(0, nsrc, source_chunk):
            s1 = min(nsrc, s0 + source_chunk)
            dde = beam_cube_dde_rotated(d_beam, d_ext, d_bfm, d_lm[s0:s1], d_pa, d_pe, d_as, d_f, d_rot)
            if dde.dtype != d_b.dtype:
                dde = dde.to(d_b.dtype)
            acc = fused_predict_vis(d_lm[s0:s1], d_uvw, d_f, d_b[s0:s1], d_ti, d_a1, d_a2, dde, dde,
                                    None, acc, None, convention=convention, dtype=out_dtype)
            del dde
        if acc is None:  # no sources and no base_vis
            acc = fused_predict_vis(d_lm, d_uvw, d_f, d_b, d_ti, d_a1, d_a2, convention=convention, dtype=out_dtype)
        if die1_jones is not None:
            d_g1 = pl.to_device(die1_jones, out_dtype, device)
            d_g2 = d_g1 if die2_jones is die1_jones else pl.to_device(die2_jones, out_dtype, device)
            acc = apply_gains(d_ti, d_a1, d_a2, d_g1, acc, d_g2)
        return acc if as_torch else pl.to_host(acc)
