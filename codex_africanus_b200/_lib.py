"""ctypes binding of libafricanus_b200.so (include/africanus_b200.h)."""
import ctypes
import os
import threading

_HERE = os.path.dirname(os.path.abspath(__file__))
LIB_PATH = os.path.join(_HERE, "libafricanus_b200.so")

AFR_FOURIER = 1
AFR_CASA = -1
AFR_F32_LM = 1
AFR_F32_UVW = 2
AFR_F32_FREQ = 4
AFR_CHAN_EXACT = 0
AFR_CHAN_UNIFORM = 1
AFR_JONES_DIAG = 0
AFR_JONES_2X2 = 1
AFR_FEED_LINEAR = 0
AFR_FEED_CIRCULAR = 1

_vp = ctypes.c_void_p
_i64 = ctypes.c_int64
_int = ctypes.c_int
_dbl = ctypes.c_double
_pint = ctypes.POINTER(ctypes.c_int)

# name -> (restype, argtypes); every symbol include/africanus_b200.h declares
PROTOTYPES = {
    "afr_version": (_int, []),
    "afr_last_error": (ctypes.c_char_p, []),
    "afr_device_count": (_int, []),
    "afr_kernel_launches": (ctypes.c_ulonglong, []),
    "afr_last_fused_path": (_int, []),
    "afr_last_dft_path": (_int, []),
    "afr_trim_scratch": (_int, [_int]),
    "afr_set_device": (_int, [_int]),
    "afr_device_info": (_int, [_int, _pint, _pint, _pint, _pint]),
    "afr_freq_is_uniform": (_int, [_vp, _i64, _dbl]),
    "afr_measure_fma_peak": (_int, [_int, _int, ctypes.POINTER(_dbl), _vp]),
    "afr_im_to_vis": (_int, [_vp, _int, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int, _int,
                             _int, _vp, _vp]),
    "afr_vis_to_im": (_int, [_vp, _int, _vp, _vp, _vp, _vp, _i64, _i64, _i64, _i64, _int, _int,
                             _int, _int, _vp, _vp]),
    "afr_phase_delay_f64": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _int, _int, _vp, _vp]),
    "afr_phase_delay_f32": (_int, [_vp, _vp, _vp, _i64, _i64, _i64, _int, _vp, _vp]),
    "afr_predict_vis": (_int, [_vp] * 9 + [_i64] * 6 + [_int, _int, _vp, _vp]),
    "afr_predict_fused": (_int, [_vp] * 12 + [_i64] * 6 + [_int, _int, _int, _int, _vp, _vp]),
    "afr_beam_cube_dde": (_int, [_vp] * 8 + [_i64] * 8 + [_int, _vp, _vp]),
    "afr_beam_cube_dde_rot": (_int, [_vp] * 9 + [_i64] * 8 + [_int, _vp, _vp]),
    "afr_beam_plane_reduce": (_int, [_vp] * 8 + [_i64] * 7 + [_vp] * 4),
    "afr_predict_fused_planes": (_int, [_vp] * 9 + [_i64] + [_vp] * 4 + [_i64] * 5 + [_int] + [_vp] * 3),
    "afr_feed_rotation": (_int, [_vp, _i64, _int, _int, _vp, _vp]),
    "afr_freq_grid_interp": (_int, [_vp, _vp, _i64, _i64, _vp, _vp]),
    "afr_wsclean_spectra": (_int, [_vp] * 5 + [_i64] * 3 + [_vp, _vp]),
    "afr_wsclean_predict": (_int, [_vp] * 9 + [_i64] * 5 + [_int, _vp, _vp]),
    "afr_spectral_model": (_int, [_vp] * 5 + [_i64] * 4 + [_vp, _vp]),
    "afr_convert": (_int, [_vp, _int, _i64, _i64, _vp, _vp, _vp, _i64, _vp, _vp]),
    "afr_stokes_brightness": (_int, [_vp] * 5 + [_i64] * 4 + [_vp] * 3 + [_i64, _int, _vp, _vp]),
}

_lock = threading.Lock()
_lib = None


class AfricanusB200Error(RuntimeError):
    pass


def lib():
    """Load the CUDA library.  Fails loudly -- there is no CPU fallback."""
    global _lib
    if _lib is not None:
        return _lib
    with _lock:
        if _lib is not None:
            return _lib
        if not os.path.exists(LIB_PATH):
            raise AfricanusB200Error(
                "%s not found: build it with `python -m codex_africanus_b200.build` "
                "(nvcc, sm_100a). codex_africanus_b200 has no CPU fallback." % LIB_PATH)
        handle = ctypes.CDLL(LIB_PATH)
        for name, (restype, argtypes) in PROTOTYPES.items():
            fn = getattr(handle, name)
            fn.restype = restype
            fn.argtypes = argtypes
        _lib = handle
    return _lib


def describe_dft_path(code=None):
    """afr_last_dft_path() as text."""
    code = lib().afr_last_dft_path() if code is None else code
    if code == 0:
        return "none"
    return "%s%s, %s, %s W tile, %s accumulators, %d channel runs per CTA, %d slice(s) of the streamed axis" % (
        ("warp-specialised (16 consumer + %d producer warps)" % (8 if code & 16 else 4)) if code & 1
        else "single-role (16 warps)", ", DMMA consumers" if code & 32 else "",
        "one sincos per term" if code & 2 else "three-term / rotation recurrence",
        "TMA bulk-copied" if code & 4 else "cp.async / staged", "FP32" if code & 8 else "FP64",
        (code >> 8) & 0xFF, (code >> 16) & 0xFFFF)


def last_error():
    msg = lib().afr_last_error()
    return msg.decode("utf-8", "replace") if msg else ""


def check(rc):
    if rc != 0:
        raise AfricanusB200Error("libafricanus_b200: %s" % last_error())
