"""``feed_rotation`` on B200 -- africanus/rime/feeds.py:13-71."""
import numpy as np
import torch

from .. import _lib
from .. import _plumbing as pl


def feed_rotation(parallactic_angles, feed_type="linear"):
    """2x2 feed rotation matrices, ``parallactic_angles.shape + (2, 2)``: linear
    [[cos pa, sin pa], [-sin pa, cos pa]], circular diag(exp(-i pa), exp(+i pa)).
    complex64 for float32 angles, complex128 for float64 (feeds.py:58-66)."""
    if feed_type == "linear":
        poltype = _lib.AFR_FEED_LINEAR
    elif feed_type == "circular":
        poltype = _lib.AFR_FEED_CIRCULAR
    else:
        raise ValueError("Invalid feed_type '%s'" % feed_type)
    pdt = pl.dtype_of(parallactic_angles)
    if pdt == np.float32:
        dtype = np.complex64
    elif pdt == np.float64:
        dtype = np.complex128
    else:
        raise ValueError("parallactic_angles has none-floating point type %s" % pdt)
    shape = tuple(pl.shape_of(parallactic_angles))
    n = int(np.prod(shape, dtype=np.int64))
    device = pl.pick_device(parallactic_angles)
    as_torch = pl.wants_torch(parallactic_angles)
    with torch.cuda.device(device):
        d_pa = pl.to_device(parallactic_angles, np.float64, device)
        d_out = pl.empty_device(shape + (2, 2), dtype, device)
        pl.call("afr_feed_rotation", device, pl.ptr(d_pa), n, poltype, int(dtype == np.complex64),
                pl.ptr(d_out), pl.stream_ptr(device))
        return d_out if as_torch else pl.to_host(d_out)
