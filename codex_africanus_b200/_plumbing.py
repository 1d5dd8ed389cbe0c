"""
Host-side plumbing shared by the entry points: dtype inference, input
normalisation (numpy or torch, any strides), host<->device staging through pinned
memory, CUDA streams.  PyTorch is used for device memory and streams only; all
arithmetic happens in libafricanus_b200.so.
"""
import ctypes
import weakref
import threading

import numpy as np
import torch

from . import _lib
from ._lib import AfricanusB200Error

_T2N = {
    torch.float32: np.dtype(np.float32), torch.float64: np.dtype(np.float64),
    torch.complex64: np.dtype(np.complex64), torch.complex128: np.dtype(np.complex128),
    torch.int8: np.dtype(np.int8), torch.int16: np.dtype(np.int16),
    torch.int32: np.dtype(np.int32), torch.int64: np.dtype(np.int64),
    torch.uint8: np.dtype(np.uint8), torch.bool: np.dtype(np.bool_),
    torch.float16: np.dtype(np.float16),
}
_N2T = {v: k for k, v in _T2N.items()}

_tls = threading.local()


def is_torch(a):
    return isinstance(a, torch.Tensor)


def dtype_of(a):
    """numpy dtype of a numpy array / torch tensor / array-like."""
    if is_torch(a):
        return _T2N[a.dtype]
    return np.asarray(a).dtype if not isinstance(a, np.ndarray) else a.dtype


def shape_of(a):
    return tuple(a.shape) if hasattr(a, "shape") else np.asarray(a).shape


def torch_dtype(np_dtype):
    return _N2T[np.dtype(np_dtype)]


def require_cuda():
    if not torch.cuda.is_available():
        raise AfricanusB200Error(
            "no CUDA device available: codex_africanus_b200 runs on B200 (sm_100a) only "
            "and has no CPU fallback")
    _lib.lib()  # fail loudly if the extension is missing


def pick_device(*arrays):
    """The CUDA device the call runs on: that of the first CUDA tensor among the
    inputs, else torch's current device."""
    require_cuda()
    for a in arrays:
        if is_torch(a) and a.is_cuda:
            return a.device
    return torch.device("cuda", torch.cuda.current_device())


def wants_torch(*arrays):
    """Outputs are torch CUDA tensors iff any input is a CUDA tensor."""
    return any(is_torch(a) and a.is_cuda for a in arrays)


def to_device(a, np_dtype, device):
    """Dense C-contiguous tensor of `np_dtype` on `device` holding the values of `a`
    (numpy array of any dtype/strides, array-like, or torch tensor anywhere)."""
    tdt = torch_dtype(np_dtype)
    if is_torch(a):
        return a.to(device=device, dtype=tdt, non_blocking=True).contiguous()
    arr = np.asarray(a)
    if arr.dtype != np.dtype(np_dtype) or not arr.flags.c_contiguous:
        arr = np.ascontiguousarray(arr, dtype=np_dtype)
    if arr.size == 0:
        return torch.empty(arr.shape, dtype=tdt, device=device)
    if not arr.flags.writeable:
        arr = arr.copy()
    host = torch.from_numpy(arr)
    if host.is_pinned():  # caller-provided page-locked memory: DMA straight from it
        return host.to(device, non_blocking=True)
    if arr.nbytes >= (1 << 20):
        # stage through (cached) pinned memory so the copy is a real async DMA
        pinned = torch.empty(host.shape, dtype=tdt, pin_memory=True)
        pinned.copy_(host)
        return pinned.to(device, non_blocking=True)
    return host.to(device)


def empty_device(shape, np_dtype, device):
    return torch.empty(tuple(int(s) for s in shape), dtype=torch_dtype(np_dtype), device=device)


def empty_pinned(shape, np_dtype):
    return torch.empty(tuple(int(s) for s in shape), dtype=torch_dtype(np_dtype), pin_memory=True)


def to_host(t):
    """Device tensor -> numpy array (through pinned memory, synchronised)."""
    if t.numel() == 0:
        return np.empty(tuple(t.shape), dtype=_T2N[t.dtype])
    if t.dim() >= 1 and t.numel() * t.element_size() > PIN_WHOLE_MAX and t.is_contiguous():
        sink = RowSink(tuple(t.shape), _T2N[t.dtype], t.device)  # bounded page-locked staging
        compute = torch.cuda.current_stream(t.device)
        step = sink.max_block_rows()
        for r0 in range(0, t.shape[0], step):
            r1 = min(t.shape[0], r0 + step)
            sink.push(r0, r1, t[r0:r1], compute)
        return sink.finish()
    pinned = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    pinned.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return pinned.numpy()


# A result this large is not pinned whole: PyTorch's caching host allocator rounds page-locked
# blocks up to a power of two and never returns them to the OS, so an 8 GB visibility array would
# lock 16 GB for as long as the caller keeps it (and SKA-size outputs could not be returned at all).
PIN_WHOLE_MAX = 2 << 30
_STAGE_BYTES = 512 << 20


_pool = None


def _copy_pool():
    global _pool
    if _pool is None:
        from concurrent.futures import ThreadPoolExecutor

        _pool = ThreadPoolExecutor(8, thread_name_prefix="afr-host-copy")
    return _pool


def _advise_hugepages(arr):
    """madvise(MADV_HUGEPAGE) on a freshly allocated result array (best effort, Linux only): the
    host threads that fill it then fault 2 MiB pages instead of 4 KiB ones."""
    try:
        libc = ctypes.CDLL(None, use_errno=True)
        addr = arr.ctypes.data
        lo = (addr + 4095) & ~4095
        hi = (addr + arr.nbytes) & ~4095
        if hi > lo:
            libc.madvise(ctypes.c_void_p(lo), ctypes.c_size_t(hi - lo), 14)  # MADV_HUGEPAGE
    except Exception:
        pass


class RowSink:
    """Host destination of a (row, ...) result produced in row blocks on the device.

    Up to PIN_WHOLE_MAX bytes the result is one page-locked array and every block is copied
    straight into place.  Above that the result is an ordinary numpy array: blocks travel through
    two rotating page-locked staging buffers (<= 512 MiB each) on the copy stream and are moved into
    place by the host while the device computes / copies the following blocks."""

    def __init__(self, shape, np_dtype, device):
        self.shape = tuple(int(v) for v in shape)
        self.dtype = np.dtype(np_dtype)
        self.device = device
        self.copier = side_stream(device)
        nbytes = int(np.prod(self.shape, dtype=np.int64)) * self.dtype.itemsize
        self.whole = nbytes <= PIN_WHOLE_MAX
        if self.whole:
            self._pinned = empty_pinned(self.shape, self.dtype)
            self.out = self._pinned.numpy()
        else:
            self.out = np.empty(self.shape, self.dtype)
            _advise_hugepages(self.out)  # 8 GB of first-touch 4 KiB page faults cost ~0.1 s per call
            self._bufs = [None, None]
            self._pending = [None, None]
            self._k = 0

    def max_block_rows(self):
        row_bytes = max(1, int(np.prod(self.shape[1:], dtype=np.int64)) * self.dtype.itemsize)
        return max(1, _STAGE_BYTES // row_bytes)

    def _drain(self, b):
        pend = self._pending[b]
        if pend is not None:
            r0, r1, done = pend
            done.synchronize()
            src = self._bufs[b][: r1 - r0].numpy()
            # numpy releases the GIL while copying: eight host threads move (and first-touch) a
            # 512 MiB block in ~15 ms (one thread: ~70 ms, which the last block's tail would add)
            parts = np.array_split(np.arange(r1 - r0), 8)
            jobs = [_copy_pool().submit(np.copyto, self.out[r0 + q[0]: r0 + q[-1] + 1], src[q[0]: q[-1] + 1])
                    for q in parts if q.size]
            for j in jobs:
                j.result()
            self._pending[b] = None

    def push(self, r0, r1, d_block, compute):
        """Queue the copy of ``d_block`` (rows r0:r1, complete once the work queued on ``compute`` so
        far has run); returns the event that marks the end of the device-side read."""
        ev = torch.cuda.Event()
        ev.record(compute)
        self.copier.wait_event(ev)
        done = torch.cuda.Event()
        if self.whole:
            with torch.cuda.stream(self.copier):
                self._pinned[r0:r1].copy_(d_block, non_blocking=True)
                done.record(self.copier)
            return done
        b = self._k & 1
        self._k += 1
        self._drain(b)
        n = r1 - r0
        if self._bufs[b] is None or self._bufs[b].shape[0] < n:
            self._bufs[b] = empty_pinned((max(n, min(self.max_block_rows(), self.shape[0])),) + self.shape[1:],
                                         self.dtype)
        with torch.cuda.stream(self.copier):
            self._bufs[b][:n].copy_(d_block, non_blocking=True)
            done.record(self.copier)
        self._pending[b] = (r0, r1, done)
        return done

    def finish(self):
        if not self.whole:
            first = self._k & 1  # the older of the two pending blocks
            self._drain(first)
            self._drain(first ^ 1)
        self.copier.synchronize()
        return self.out


def ptr(t):
    return None if t is None else ctypes.c_void_p(t.data_ptr())


def stream_ptr(device):
    return ctypes.c_void_p(torch.cuda.current_stream(device).cuda_stream)


def side_stream(device):
    """Per-thread, per-device copy stream (overlaps D2H with compute)."""
    key = "side_%d" % device.index
    s = getattr(_tls, key, None)
    if s is None:
        s = torch.cuda.Stream(device=device)
        setattr(_tls, key, s)
    return s


def call(name, device, *args):
    """Invoke a C-ABI entry point on `device` (GIL released by ctypes)."""
    lib = _lib.lib()
    _lib.check(lib.afr_set_device(int(device.index)))
    _lib.check(getattr(lib, name)(*args))


def convention_sign(convention):
    if convention == "fourier":
        return _lib.AFR_FOURIER
    elif convention == "casa":
        return _lib.AFR_CASA
    raise ValueError("convention not in ('fourier', 'casa')")


def host_frequency(frequency):
    """float64 host copy of the (tiny) frequency array."""
    if is_torch(frequency):
        return frequency.detach().to("cpu", torch.float64).contiguous().numpy()
    return np.ascontiguousarray(frequency, dtype=np.float64)


# Channels count as equispaced when every frequency is within this relative distance
# of the ideal grid; the induced phase error is |p| * rtol, i.e. a few double ulps.
UNIFORM_RTOL = 4e-16


_chan_mode_cache = {}  # id(tensor) -> (weakref, tensor._version, mode)


def channel_mode(frequency):
    """AFR_CHAN_UNIFORM when the channels are equispaced (recurrence kernels), else AFR_CHAN_EXACT.
    The check needs the values on the host; for a CUDA tensor that is a blocking copy, which a
    millisecond-scale call (BASELINE configs[0]) cannot afford on every invocation, so the answer is
    remembered per tensor OBJECT (weak reference) and version counter -- never per address."""
    if is_torch(frequency) and frequency.is_cuda:
        key = id(frequency)
        hit = _chan_mode_cache.get(key)
        if hit is not None and hit[0]() is frequency and hit[1] == frequency._version:
            return hit[2]
        mode = _channel_mode_host(host_frequency(frequency))
        try:
            ref = weakref.ref(frequency, lambda _r, k=key: _chan_mode_cache.pop(k, None))
            _chan_mode_cache[key] = (ref, frequency._version, mode)
        except TypeError:
            pass
        return mode
    return _channel_mode_host(host_frequency(frequency))


def _channel_mode_host(f):
    if f.ndim != 1:
        raise ValueError("frequency must be 1-dimensional")
    if f.shape[0] <= 1:
        return _lib.AFR_CHAN_UNIFORM
    if not np.all(np.isfinite(f)):
        return _lib.AFR_CHAN_EXACT
    ok = _lib.lib().afr_freq_is_uniform(f.ctypes.data_as(ctypes.c_void_p), f.shape[0], UNIFORM_RTOL)
    return _lib.AFR_CHAN_UNIFORM if ok else _lib.AFR_CHAN_EXACT


def f32_flags(lm=None, uvw=None, frequency=None):
    flags = 0
    if lm is not None and dtype_of(lm) == np.float32:
        flags |= _lib.AFR_F32_LM
    if uvw is not None and dtype_of(uvw) == np.float32:
        flags |= _lib.AFR_F32_UVW
    if frequency is not None and dtype_of(frequency) == np.float32:
        flags |= _lib.AFR_F32_FREQ
    return flags
