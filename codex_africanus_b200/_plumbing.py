"""
Host-side plumbing shared by the entry points: dtype inference, input
normalisation (numpy or torch, any strides), host<->device staging through pinned
memory, CUDA streams.  PyTorch is used for device memory and streams only; all
arithmetic happens in libafricanus_b200.so.
"""
import ctypes
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
    pinned = torch.empty(t.shape, dtype=t.dtype, pin_memory=True)
    pinned.copy_(t, non_blocking=True)
    torch.cuda.current_stream(t.device).synchronize()
    return pinned.numpy()


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


def channel_mode(frequency):
    f = host_frequency(frequency)
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
