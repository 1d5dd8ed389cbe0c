"""
Dask-facing ``im_to_vis`` / ``vis_to_im`` -- africanus/dft/dask.py:26-90 with the B200 kernels as
block functions.  Same signatures, chunk checks, ``blockwise`` index strings and defaults
(``dtype=complex128`` / ``float64``) as the reference.  Inputs are ``dask.array.Array`` (lazy graph;
needs dask) or ``codex_africanus_b200._chunked.ChunkedArray`` (evaluated eagerly, no dask needed).
Every block call runs on the calling worker thread's GPU (``_chunked.worker_device``).
"""
import numpy as np

from .. import _chunked as ck
from .kernels import im_to_vis as np_im_to_vis
from .kernels import vis_to_im as np_vis_to_im


def _blockwise_for(*arrays):
    arrays = [a for a in arrays if a is not None]
    if any(ck.is_dask(a) for a in arrays):
        return ck.da.core.blockwise
    if all(isinstance(a, ck.ChunkedArray) for a in arrays):
        return ck.blockwise
    if not ck.have_dask():
        raise ImportError("codex_africanus_b200.dft.dask needs dask.array.Array inputs (dask is not "
                          "installed) or codex_africanus_b200._chunked.ChunkedArray inputs")
    raise TypeError("expected dask.array.Array (or ChunkedArray) inputs")


@ck.on_worker_device
def _im_to_vis_wrapper(image, uvw, lm, frequency, convention, dtype_):
    # image contracts over one 'source' chunk, uvw over '(u,v,w)', lm over both of its indices
    return np_im_to_vis(image[0], uvw[0], lm[0][0], frequency, convention=convention, dtype=dtype_)


def im_to_vis(image, uvw, lm, frequency, convention="fourier", dtype=np.complex128):
    """Dask wrapper for im_to_vis (africanus/dft/dask.py:26-51)."""
    if lm.chunks[0][0] != lm.shape[0]:
        raise ValueError("lm chunks must match lm shape on first axis")
    if image.chunks[0][0] != image.shape[0]:
        raise ValueError("Image chunks must match image shape on first axis")
    if image.chunks[0][0] != lm.chunks[0][0]:
        raise ValueError("Image chunks and lm chunks must match on first axis")
    if image.chunks[1] != frequency.chunks[0]:
        raise ValueError("Image chunks must match frequency chunks on second axis")
    return _blockwise_for(image, uvw, lm, frequency)(
        _im_to_vis_wrapper, ("row", "chan", "corr"),
        image, ("source", "chan", "corr"),
        uvw, ("row", "(u,v,w)"),
        lm, ("source", "(l,m)"),
        frequency, ("chan",),
        convention=convention, dtype=dtype, dtype_=dtype)


@ck.on_worker_device
def _vis_to_im_wrapper(vis, uvw, lm, frequency, flags, convention, dtype_):
    return np_vis_to_im(vis, uvw[0], lm[0], frequency, flags, convention=convention, dtype=dtype_)[None, :]


def vis_to_im(vis, uvw, lm, frequency, flags, convention="fourier", dtype=np.float64):
    """Dask wrapper for vis_to_im (africanus/dft/dask.py:60-90): one partial image per row chunk,
    summed over the row-chunk axis (on several GPUs that sum is the NCCL all_reduce of
    ``distributed.sharded_vis_to_im``)."""
    if vis.chunks[0] != uvw.chunks[0]:
        raise ValueError("Vis chunks and uvw chunks must match on first axis")
    if vis.chunks[1] != frequency.chunks[0]:
        raise ValueError("Vis chunks must match frequency chunks on second axis")
    if vis.chunks != flags.chunks:
        raise ValueError("Vis chunks must match flags chunks on all axes")
    ims = _blockwise_for(vis, uvw, lm, frequency, flags)(
        _vis_to_im_wrapper, ("row", "source", "chan", "corr"),
        vis, ("row", "chan", "corr"),
        uvw, ("row", "(u,v,w)"),
        lm, ("source", "(l,m)"),
        frequency, ("chan",),
        flags, ("row", "chan", "corr"),
        adjust_chunks={"row": 1},
        convention=convention, dtype=dtype, dtype_=dtype)
    return ims.sum(axis=0)
