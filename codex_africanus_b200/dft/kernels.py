"""
``im_to_vis`` / ``vis_to_im`` on B200 -- same signatures, dtype promotion, layouts
and errors as africanus/dft/kernels.py:14-69 and :72-148.

Inputs may be numpy arrays (any real/complex dtype, any strides; staged to the GPU
through pinned memory, result returned as a fresh numpy array) or torch CUDA tensors
(used in place; result returned as a torch CUDA tensor).
"""
import numpy as np
import torch

from .. import _lib
from .. import _plumbing as pl

# rows per launch when the output is streamed back to the host while later
# row blocks are still being computed
_ROW_BLOCK_BYTES = 512 << 20


def _check_convention(convention):
    return pl.convention_sign(convention)


def im_to_vis(image, uvw, lm, frequency, convention="fourier", dtype=None):
    """Image -> visibilities, africanus/dft/kernels.py:14-69.

    image (source, chan, corr) real or complex; uvw (row, 3); lm (source, 2);
    frequency (chan,) -> (row, chan, corr) complex.  Output dtype is
    ``result_type(complex64, inputs)`` unless ``dtype`` is given (kernels.py:26-31).
    """
    sign = _check_convention(convention)
    if dtype is None:
        out_dtype = np.result_type(np.complex64, *(pl.dtype_of(a) for a in (image, uvw, lm, frequency)))
    else:
        out_dtype = np.dtype(dtype)
    if out_dtype not in (np.complex64, np.complex128):
        raise TypeError("im_to_vis: output dtype %s is not complex64/complex128" % out_dtype)

    ishape, ushape, lshape = pl.shape_of(image), pl.shape_of(uvw), pl.shape_of(lm)
    if len(ishape) != 3 or len(ushape) != 2 or ushape[1] != 3 or len(lshape) != 2 or lshape[1] != 2:
        raise ValueError("im_to_vis: expected image (source,chan,corr), uvw (row,3), lm (source,2)")
    nsrc, nchan, ncorr = ishape
    nrow = ushape[0]
    if lshape[0] != nsrc or pl.shape_of(frequency) != (nchan,):
        raise ValueError("im_to_vis: lm / frequency do not match the image shape")

    device = pl.pick_device(image, uvw, lm, frequency)
    as_torch = pl.wants_torch(image, uvw, lm, frequency)
    flags = pl.f32_flags(lm=lm, uvw=uvw)
    chan_mode = pl.channel_mode(frequency)
    img_complex = pl.dtype_of(image).kind == "c"
    out_c64 = int(out_dtype == np.complex64)

    with torch.cuda.device(device):
        if out_c64:  # FP32 accumulator variant takes the image in single precision
            d_img = pl.to_device(image, np.complex64 if img_complex else np.float32, device)
        else:
            d_img = pl.to_device(image, np.complex128 if img_complex else np.float64, device)
        d_uvw = pl.to_device(uvw, np.float64, device)
        d_lm = pl.to_device(lm, np.float64, device)
        d_freq = pl.to_device(frequency, np.float64, device)
        d_out = pl.empty_device((nrow, nchan, ncorr), out_dtype, device)

        def launch(r0, r1):
            pl.call("afr_im_to_vis", device, pl.ptr(d_img), int(img_complex),
                    pl.ptr(d_uvw[r0:r1]) if r1 > r0 else None, pl.ptr(d_lm), pl.ptr(d_freq),
                    nsrc, r1 - r0, nchan, ncorr, sign, flags, chan_mode, out_c64,
                    pl.ptr(d_out[r0:r1]) if r1 > r0 else None, pl.stream_ptr(device))

        if as_torch:
            launch(0, nrow)
            return d_out

        # numpy path: rows are independent, so compute in row blocks and overlap the
        # device->host copy of block k with the kernel of block k+1
        row_bytes = max(1, nchan * ncorr * out_dtype.itemsize)
        block = max(4096, _ROW_BLOCK_BYTES // row_bytes)
        if nrow == 0 or nchan == 0 or ncorr == 0:
            launch(0, nrow)
            return np.zeros((nrow, nchan, ncorr), out_dtype)
        sink = pl.RowSink((nrow, nchan, ncorr), out_dtype, device)
        block = min(block, sink.max_block_rows()) if not sink.whole else block
        compute = torch.cuda.current_stream(device)
        for r0 in range(0, nrow, block):
            r1 = min(nrow, r0 + block)
            launch(r0, r1)
            sink.push(r0, r1, d_out[r0:r1], compute)
        out = sink.finish()
        compute.synchronize()
        return out


def vis_to_im(vis, uvw, lm, frequency, flags, convention="fourier", dtype=None):
    """Visibilities -> image (adjoint of im_to_vis), africanus/dft/kernels.py:72-148.

    vis (row, chan, corr) real or complex; flags bool of the same shape ->
    (source, chan, corr) real.  A (row, chan) sample is dropped if ANY of its
    correlations is flagged (kernels.py:136-137).
    """
    sign = _check_convention(convention)
    vdt = pl.dtype_of(vis)
    if dtype is None:
        vreal = np.dtype(np.float32) if vdt == np.complex64 else (
            np.dtype(np.float64) if vdt == np.complex128 else vdt)
        out_dtype = np.result_type(vreal, *(pl.dtype_of(a) for a in (uvw, lm, frequency)))
    else:
        out_dtype = np.dtype(dtype)
        if out_dtype.kind == "c":
            raise TypeError("dtype must be complex")  # sic: africanus/dft/kernels.py:97-98
    if out_dtype not in (np.float32, np.float64):
        raise TypeError("vis_to_im: output dtype %s is not float32/float64" % out_dtype)

    vshape = pl.shape_of(vis)
    assert vshape == pl.shape_of(flags)  # africanus/dft/kernels.py:102
    ushape, lshape = pl.shape_of(uvw), pl.shape_of(lm)
    if len(vshape) != 3 or len(ushape) != 2 or ushape[1] != 3 or len(lshape) != 2 or lshape[1] != 2:
        raise ValueError("vis_to_im: expected vis (row,chan,corr), uvw (row,3), lm (source,2)")
    nrow, nchan, ncorr = vshape
    nsrc = lshape[0]
    if ushape[0] != nrow or pl.shape_of(frequency) != (nchan,):
        raise ValueError("vis_to_im: uvw / frequency do not match the visibility shape")

    device = pl.pick_device(vis, uvw, lm, frequency, flags)
    as_torch = pl.wants_torch(vis, uvw, lm, frequency, flags)
    fflags = pl.f32_flags(lm=lm, uvw=uvw)
    chan_mode = pl.channel_mode(frequency)
    vis_complex = vdt.kind == "c"

    with torch.cuda.device(device):
        if out_dtype == np.float32:  # FP32 accumulator variant takes vis in single precision
            d_vis = pl.to_device(vis, np.complex64 if vis_complex else np.float32, device)
        else:
            d_vis = pl.to_device(vis, np.complex128 if vis_complex else np.float64, device)
        d_uvw = pl.to_device(uvw, np.float64, device)
        d_lm = pl.to_device(lm, np.float64, device)
        d_freq = pl.to_device(frequency, np.float64, device)
        # nothing flagged (the common case for simulated data): skip the flag path entirely
        if pl.is_torch(flags):
            any_flag = bool((flags != 0).any().item()) if flags.numel() else False
            d_flags = (flags != 0).to(device=device, dtype=torch.uint8).contiguous() if any_flag else None
        else:
            fl = np.asarray(flags) != 0
            d_flags = pl.to_device(fl, np.uint8, device) if fl.any() else None
        d_out = pl.empty_device((nsrc, nchan, ncorr), out_dtype, device)
        pl.call("afr_vis_to_im", device, pl.ptr(d_vis), int(vis_complex), pl.ptr(d_uvw),
                pl.ptr(d_lm), pl.ptr(d_freq), pl.ptr(d_flags), nsrc, nrow, nchan, ncorr, sign,
                fflags, chan_mode, int(out_dtype == np.float32), pl.ptr(d_out),
                pl.stream_ptr(device))
        if as_torch:
            return d_out
        return pl.to_host(d_out)
