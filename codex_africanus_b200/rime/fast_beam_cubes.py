"""``beam_cube_dde`` / ``freq_grid_interp`` on B200 --
africanus/rime/fast_beam_cubes.py:57-240 and :10-54."""
import numpy as np
import torch

from .. import _plumbing as pl


def freq_grid_interp(frequency, beam_freq_map):
    """(chan, 3) table of (scale, lower weight, lower grid index), fast_beam_cubes.py:10-54."""
    out_dtype = pl.dtype_of(frequency)
    nchan, nud = pl.shape_of(frequency)[0], pl.shape_of(beam_freq_map)[0]
    device = pl.pick_device(frequency, beam_freq_map)
    as_torch = pl.wants_torch(frequency, beam_freq_map)
    with torch.cuda.device(device):
        d_f = pl.to_device(frequency, np.float64, device)
        d_m = pl.to_device(beam_freq_map, np.float64, device)
        d_out = pl.empty_device((nchan, 3), np.float64, device)
        pl.call("afr_freq_grid_interp", device, pl.ptr(d_f), pl.ptr(d_m), nchan, nud,
                pl.ptr(d_out), pl.stream_ptr(device))
        d_out = d_out.to(pl.torch_dtype(out_dtype)) if out_dtype.kind == "f" else d_out
        return d_out if as_torch else pl.to_host(d_out)


def beam_cube_dde(beam, beam_lm_extents, beam_freq_map, lm, parallactic_angles, point_errors,
                  antenna_scaling, frequency):
    """Direction-dependent Jones from a beam cube, fast_beam_cubes.py:57-240.

    beam (lw, mh, nud, corr...) complex -> (source, time, ant, chan, corr...) in
    ``beam.dtype``.  Coordinate arithmetic is float64.
    """
    return beam_cube_dde_rotated(beam, beam_lm_extents, beam_freq_map, lm, parallactic_angles,
                                 point_errors, antenna_scaling, frequency, None)


def beam_cube_dde_rotated(beam, beam_lm_extents, beam_freq_map, lm, parallactic_angles, point_errors,
                          antenna_scaling, frequency, feed_rotation):
    """``einsum("stafij,tajk->stafik", beam_cube_dde(...), feed_rotation)`` -- the DDE term of
    africanus/rime/examples/predict.py:469-472 -- with the product done in the interpolation
    kernel's epilogue (no second pass over the DDE array).  ``feed_rotation`` (time, ant, 2, 2)
    complex (see ``feed_rotation``) or None for the plain ``beam_cube_dde``; needs a
    (lw, mh, nud, 2, 2) beam.
    """
    bshape = pl.shape_of(beam)
    if len(bshape) < 3:
        raise ValueError("beam must have at least 3 dimensions")
    lw, mh, nud = bshape[:3]
    corrs = tuple(bshape[3:])
    if lw < 2 or mh < 2 or nud < 2:
        raise ValueError("beam_lw, beam_mh and beam_nud must be >= 2")
    bdt = pl.dtype_of(beam)
    if bdt not in (np.complex64, np.complex128):
        raise TypeError("beam_cube_dde: beam must be complex64/complex128")
    ncorr = int(np.prod(corrs)) if corrs else 1
    nsrc = pl.shape_of(lm)[0]
    ntime, nant = pl.shape_of(parallactic_angles)
    nchan = pl.shape_of(frequency)[0]
    args = (beam, beam_lm_extents, beam_freq_map, lm, parallactic_angles, point_errors,
            antenna_scaling, frequency)
    if feed_rotation is not None:
        if corrs != (2, 2):
            raise ValueError("a feed rotation needs a beam of shape (lw, mh, nud, 2, 2)")
        if tuple(pl.shape_of(feed_rotation)) != (ntime, nant, 2, 2):
            raise ValueError("feed_rotation must have shape (time, ant, 2, 2)")
    device = pl.pick_device(*args, feed_rotation)
    as_torch = pl.wants_torch(*args, feed_rotation)
    f64 = np.float64
    with torch.cuda.device(device):
        d_beam = pl.to_device(beam, bdt, device)
        d = [pl.to_device(a, f64, device) for a in args[1:]]
        d_rot = None if feed_rotation is None else pl.to_device(feed_rotation, bdt, device)
        d_out = pl.empty_device((nsrc, ntime, nant, nchan) + corrs, bdt, device)
        pl.call("afr_beam_cube_dde_rot", device, pl.ptr(d_beam), *(pl.ptr(x) for x in d), pl.ptr(d_rot),
                lw, mh, nud, ncorr, nsrc, ntime, nant, nchan, int(bdt == np.complex64),
                pl.ptr(d_out), pl.stream_ptr(device))
        return d_out if as_torch else pl.to_host(d_out)
