"""
Chunked-array plumbing of the dask-facing wrappers (``dft.dask``, ``rime.dask``).

The reference's dask wrappers (africanus/dft/dask.py:26-90, africanus/rime/dask.py:34-52,
africanus/rime/dask_predict.py:311-593) are ``dask.array.blockwise`` graphs whose block function is
the numpy-API kernel.  The wrappers here build the same graphs -- same index strings, same chunk
checks, same ``streams=`` semantics -- over either backend:

* real ``dask.array.Array`` inputs, when dask is importable (``requires dask``): lazy graphs,
  executed by dask's threaded scheduler; every worker thread is pinned to one GPU (round robin
  over the visible devices, ``worker_device``) and calls the C ABI with the GIL released;
* ``ChunkedArray`` inputs (this module): a numpy array plus a chunk layout, evaluated eagerly by
  ``blockwise`` below, which reproduces ``dask.array.blockwise``'s block-function calling
  convention (contracted indices arrive as nested lists of blocks; ``adjust_chunks``;
  ``align_arrays=False`` pairing blocks by position).  It exists so that the graphs can be
  exercised -- and used -- where dask is not installed (it is absent from this image), and is what
  ``tests/test_dask_wrappers*.py`` run the reference's ``test_dask_*`` parametrisations through.
"""
import itertools
import threading

import numpy as np

try:  # pragma: no cover - dask is not installed in the build image
    import dask.array as da
except ImportError:  # noqa: F401
    da = None


def have_dask():
    return da is not None


def is_dask(x):
    return da is not None and isinstance(x, da.Array)


# ------------------------------------------------------------------------------------------
# one GPU per worker thread
# ------------------------------------------------------------------------------------------
_tls = threading.local()
_next_device = itertools.count()


def worker_device():
    """The CUDA device of the calling (dask worker) thread: threads are assigned round robin over
    the visible GPUs on their first call; the library entry points then run on that device
    (``afr_set_device`` is thread-local state of the C ABI)."""
    import torch

    dev = getattr(_tls, "device", None)
    if dev is None:
        n = max(1, torch.cuda.device_count())
        dev = next(_next_device) % n
        _tls.device = dev
    return dev


def on_worker_device(fn):
    """Run ``fn`` with the calling thread's GPU as torch's current device."""
    import torch

    def wrapped(*args, **kwargs):
        if not torch.cuda.is_available():
            return fn(*args, **kwargs)  # the entry point raises AfricanusB200Error (no CPU fallback)
        with torch.cuda.device(worker_device()):
            return fn(*args, **kwargs)

    wrapped.__name__ = getattr(fn, "__name__", "wrapped")
    return wrapped


# ------------------------------------------------------------------------------------------
# eager chunked arrays
# ------------------------------------------------------------------------------------------
def _normalise_chunks(chunks, shape):
    if not isinstance(chunks, (tuple, list)):
        chunks = (chunks,) * len(shape)
    out = []
    for c, n in zip(chunks, shape):
        if isinstance(c, (tuple, list)):
            c = tuple(int(v) for v in c)
            if sum(c) != n:
                raise ValueError("chunks %s do not add up to the axis length %d" % (c, n))
        else:
            c = int(c)
            c = tuple([c] * (n // c) + ([n % c] if n % c else [])) if n else (0,)
        out.append(c)
    return tuple(out)


class ChunkedArray:
    """A numpy array with a dask-style chunk layout (``chunks``: one tuple of block lengths per axis)."""

    def __init__(self, data, chunks):
        self.data = np.asarray(data)
        self.chunks = _normalise_chunks(chunks, self.data.shape)

    shape = property(lambda self: self.data.shape)
    dtype = property(lambda self: self.data.dtype)
    ndim = property(lambda self: self.data.ndim)
    numblocks = property(lambda self: tuple(len(c) for c in self.chunks))

    def block(self, index):
        sl = []
        for c, i in zip(self.chunks, index):
            start = sum(c[:i])
            sl.append(slice(start, start + c[i]))
        return self.data[tuple(sl)]

    def sum(self, axis=0):
        chunks = tuple(c for i, c in enumerate(self.chunks) if i != axis)
        return ChunkedArray(self.data.sum(axis=axis), chunks)

    def compute(self):
        return self.data

    def __add__(self, other):
        return ChunkedArray(self.data + (other.data if isinstance(other, ChunkedArray) else other), self.chunks)

    __iadd__ = __add__


def from_array(x, chunks):
    return ChunkedArray(x, chunks)


def blockwise(func, out_ind, *args, adjust_chunks=None, align_arrays=True, dtype=None, meta=None, **kwargs):
    """``dask.array.blockwise`` evaluated eagerly over ChunkedArrays (``concatenate=None``: an index
    of an argument that is not in ``out_ind`` is contracted and the block function receives a list
    -- nested in the order of the argument's contracted indices -- of ALL blocks along it).
    ``args`` alternate array, index-tuple; a ``None`` index passes the value through unchanged."""
    del meta
    pairs = list(zip(args[0::2], args[1::2]))
    nblocks, sizes = {}, {}
    for arr, ind in pairs:
        if ind is None:
            continue
        if len(ind) != arr.ndim:
            raise ValueError("index %s does not match a %d-dimensional array" % (ind, arr.ndim))
        for d, c in zip(ind, arr.chunks):
            if d in nblocks:
                if nblocks[d] != len(c):
                    raise ValueError("dimension %r has %d and %d blocks" % (d, nblocks[d], len(c)))
                if align_arrays and sizes[d] != c:
                    raise ValueError("dimension %r has chunks %s and %s" % (d, sizes[d], c))
            else:
                nblocks[d], sizes[d] = len(c), c
    for d in out_ind:
        if d not in nblocks:
            raise ValueError("output index %r appears in no input" % (d,))

    def gather(arr, ind, fixed):
        free = [d for d in ind if d not in fixed]

        def rec(k, chosen):
            if k == len(free):
                return arr.block(tuple(chosen[d] for d in ind))
            return [rec(k + 1, dict(chosen, **{free[k]: b})) for b in range(nblocks[free[k]])]

        return rec(0, dict(fixed))

    blocks = {}
    for idx in itertools.product(*(range(nblocks[d]) for d in out_ind)):
        fixed = dict(zip(out_ind, idx))
        call = [a if ind is None else gather(a, ind, {d: fixed[d] for d in ind if d in fixed}) for a, ind in pairs]
        blocks[idx] = np.asarray(func(*call, **kwargs))
    out_chunks = []
    for k, d in enumerate(out_ind):
        if adjust_chunks and d in adjust_chunks:
            ac = adjust_chunks[d]
            c = tuple(ac) if isinstance(ac, (tuple, list)) else tuple([int(ac)] * nblocks[d])
        else:
            c = sizes[d]
        out_chunks.append(c)
    out_chunks = tuple(out_chunks)
    first = next(iter(blocks.values()))
    out = np.empty(tuple(sum(c) for c in out_chunks), dtype=dtype if dtype is not None else first.dtype)
    for idx, blk in blocks.items():
        sl = tuple(slice(sum(c[:i]), sum(c[:i]) + c[i]) for c, i in zip(out_chunks, idx))
        if blk.shape != tuple(s.stop - s.start for s in sl):
            raise ValueError("block %s has shape %s, expected %s" % (idx, blk.shape,
                                                                      tuple(s.stop - s.start for s in sl)))
        out[sl] = blk
    return ChunkedArray(out, out_chunks)
