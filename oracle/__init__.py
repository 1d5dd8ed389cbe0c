"""
CPU oracle for the codex-africanus RIME/DFT hot path -- TEST INFRASTRUCTURE ONLY.

numpy-facing wrappers (reference signatures) around ``libafr_oracle.so``, the
plain-C restatement in ``afr_oracle.c``.  Only ``tests/``,
``__graft_entry__.smoke()`` and the CPU-baseline legs of ``bench.py`` may
import this package; ``codex_africanus_b200`` never does.

Parity is pinned by ``tests/test_oracle.py`` against golden vectors generated
from the reference's own numba implementation (``oracle/gen_golden.py`` ->
``tests/golden/*.npz``) and against the reference's known-answer tests.

Reference entry points restated (paths relative to /root/reference/):
  phase_delay      africanus/rime/phase.py:11-63
  predict_vis      africanus/rime/predict.py:466-619
  apply_gains      africanus/rime/predict.py:622-649
  beam_cube_dde    africanus/rime/fast_beam_cubes.py:57-240
  freq_grid_interp africanus/rime/fast_beam_cubes.py:10-54
  im_to_vis        africanus/dft/kernels.py:14-69
  vis_to_im        africanus/dft/kernels.py:72-148
  fused_predict    composition in africanus/rime/examples/predict.py:107-134,490,522-527
  wsclean_spectra  africanus/model/wsclean/spec_model.py:76-124
  wsclean_predict  africanus/rime/wsclean_predict.py:11-116
"""
import ctypes
import os
import subprocess

import numpy as np

_HERE = os.path.dirname(os.path.abspath(__file__))
_LIB_PATH = os.path.join(_HERE, "libafr_oracle.so")
_lib = None

_i64 = ctypes.c_int64
_int = ctypes.c_int
_vp = ctypes.c_void_p


def build(force=False):
    """Compile libafr_oracle.so with the committed Makefile (gcc)."""
    src = os.path.join(_HERE, "afr_oracle.c")
    if (
        force
        or not os.path.exists(_LIB_PATH)
        or os.path.getmtime(_LIB_PATH) < os.path.getmtime(src)
    ):
        subprocess.check_call(["make", "-s", "-C", _HERE, "-B", "libafr_oracle.so"])
    return _LIB_PATH


def lib():
    global _lib
    if _lib is None:
        if not os.path.exists(_LIB_PATH):
            build()
        _lib = ctypes.CDLL(_LIB_PATH)
    return _lib


def max_threads():
    return int(lib().orc_max_threads())


def set_threads(n):
    """Thread count used by the OpenMP loops (bench CPU baseline only)."""
    os.environ["OMP_NUM_THREADS"] = str(int(n))
    try:
        omp = ctypes.CDLL("libgomp.so.1")
        omp.omp_set_num_threads(int(n))
    except OSError:
        pass


def _p(a):
    return None if a is None else a.ctypes.data_as(_vp)


def _as(a, dtype):
    return np.ascontiguousarray(a, dtype=dtype)


def _is_f32(a):
    return int(np.asarray(a).dtype == np.float32)


def _sign(convention):
    if convention == "fourier":
        return 1
    elif convention == "casa":
        return -1
    raise ValueError("convention not in ('fourier', 'casa')")


# ---------------------------------------------------------------------------
def phase_delay(lm, uvw, frequency, convention="fourier"):
    sign = _sign(convention)
    lm, uvw, frequency = (np.asarray(a) for a in (lm, uvw, frequency))
    out_dtype = np.result_type(np.complex64, lm.dtype, uvw.dtype, frequency.dtype)
    ns, nr, nf = lm.shape[0], uvw.shape[0], frequency.shape[0]
    if out_dtype == np.complex64:
        out = np.zeros((ns, nr, nf), np.complex64)
        rc = lib().orc_phase_delay_f32(
            _p(_as(lm, np.float32)), _p(_as(uvw, np.float32)), _p(_as(frequency, np.float32)),
            _i64(ns), _i64(nr), _i64(nf), _int(sign), _p(out))
    else:
        out = np.zeros((ns, nr, nf), np.complex128)
        a, b, c = _as(lm, np.float64), _as(uvw, np.float64), _as(frequency, np.float64)
        rc = lib().orc_phase_delay_f64(
            _p(a), _p(b), _p(c), _i64(ns), _i64(nr), _i64(nf), _int(sign),
            _int(_is_f32(lm)), _int(_is_f32(uvw)), _int(_is_f32(frequency)), _p(out))
    assert rc == 0, rc
    return out


# ---------------------------------------------------------------------------
def im_to_vis(image, uvw, lm, frequency, convention="fourier", dtype=None):
    sign = _sign(convention)
    image, uvw, lm, frequency = (np.asarray(a) for a in (image, uvw, lm, frequency))
    if dtype is None:
        out_dtype = np.result_type(np.complex64, image.dtype, uvw.dtype, lm.dtype, frequency.dtype)
    else:
        out_dtype = np.dtype(dtype)
    ns, nf, nc = image.shape
    nr = uvw.shape[0]
    cplx = int(np.iscomplexobj(image))
    img = _as(image, np.complex128 if cplx else np.float64)
    out = np.zeros((nr, nf, nc), np.complex128)
    rc = lib().orc_im_to_vis(
        _p(img), _int(cplx), _p(_as(uvw, np.float64)), _p(_as(lm, np.float64)),
        _p(_as(frequency, np.float64)), _i64(ns), _i64(nr), _i64(nf), _i64(nc), _int(sign),
        _int(_is_f32(lm)), _int(_is_f32(uvw)), _int(out_dtype == np.complex64), _p(out))
    assert rc == 0, rc
    return out.astype(out_dtype)


def vis_to_im(vis, uvw, lm, frequency, flags, convention="fourier", dtype=None):
    sign = _sign(convention)
    vis, uvw, lm, frequency, flags = (np.asarray(a) for a in (vis, uvw, lm, frequency, flags))
    cplx = int(np.iscomplexobj(vis))
    if dtype is None:
        vreal = vis.real.dtype if cplx else vis.dtype
        out_dtype = np.result_type(vreal, uvw.dtype, lm.dtype, frequency.dtype)
    else:
        out_dtype = np.dtype(dtype)
        if out_dtype.kind == "c":
            raise TypeError("dtype must be complex")  # sic, kernels.py:98
    assert vis.shape == flags.shape
    nr, nf, nc = vis.shape
    ns = lm.shape[0]
    v = _as(vis, np.complex128 if cplx else np.float64)
    fl = _as(flags, np.uint8)
    out = np.zeros((ns, nf, nc), np.float64)
    rc = lib().orc_vis_to_im(
        _p(v), _int(cplx), _p(_as(uvw, np.float64)), _p(_as(lm, np.float64)),
        _p(_as(frequency, np.float64)), _p(fl), _i64(ns), _i64(nr), _i64(nf), _i64(nc),
        _int(sign), _int(_is_f32(lm)), _int(_is_f32(uvw)), _int(out_dtype == np.float32), _p(out))
    assert rc == 0, rc
    return out.astype(out_dtype)


# ---------------------------------------------------------------------------
def _predict_layout(dde1, coh, dde2, die1, bvis, die2):
    """ndim / presence rules of predict_checks (predict.py:380-463) and
    _get_jones_types (predict.py:15-53, 546-563).  Returns (mode, corr_shape)."""
    if (dde1 is None) ^ (dde2 is None):
        raise ValueError("Both dde1_jones and dde2_jones must be present or absent")
    if (die1 is None) ^ (die2 is None):
        raise ValueError("Both die1_jones and die2_jones must be present or absent")
    spec = [("dde1_jones", dde1, 5, 6), ("source_coh", coh, 4, 5), ("dde2_jones", dde2, 5, 6),
            ("die1_jones", die1, 4, 5), ("base_vis", bvis, 3, 4), ("die2_jones", die2, 4, 5)]
    modes = []
    corr_shape = None
    for name, a, d1, d2 in spec:
        if a is None:
            continue
        if a.ndim == d1:
            modes.append(0)
        elif a.ndim == d2:
            modes.append(1)
        else:
            raise ValueError("%s.ndim %d not in (%d, %d)" % (name, a.ndim, d1, d2))
        if corr_shape is None:
            corr_shape = a.shape[d1 - 1:]
    if not modes:
        raise ValueError("No Jones Matrices were supplied")
    if any(m != modes[0] for m in modes):
        raise ValueError("Jones Matrix Correlations were mismatched")
    return modes[0], corr_shape


def predict_vis(time_index, antenna1, antenna2, dde1_jones=None, source_coh=None,
                dde2_jones=None, die1_jones=None, base_vis=None, die2_jones=None):
    arrs = [dde1_jones, source_coh, dde2_jones, die1_jones, base_vis, die2_jones]
    arrs = [None if a is None else np.asarray(a) for a in arrs]
    dde1, coh, dde2, die1, bvis, die2 = arrs
    mode, corr_shape = _predict_layout(*arrs)
    out_dtype = np.result_type(*(a.dtype for a in arrs if a is not None))
    ncorr = int(np.prod(corr_shape))
    nrow = np.asarray(time_index).shape[0]
    if dde1 is not None:
        nsrc, ntime, nant, nchan = dde1.shape[:4]
    else:
        nsrc = coh.shape[0] if coh is not None else 0
        ntime, nant = (die1.shape[:2] if die1 is not None else (1, 1))
        nchan = coh.shape[2] if coh is not None else (die1.shape[2] if die1 is not None else bvis.shape[1])
    if out_dtype == np.complex64:
        T, fn = np.complex64, lib().orc_predict_vis_c64
    else:
        T, fn = np.complex128, lib().orc_predict_vis_c128
    cv = [None if a is None else _as(a, T) for a in arrs]
    ti = _as(time_index, np.int64)
    a1 = _as(antenna1, np.int64)
    a2 = _as(antenna2, np.int64)
    out = np.zeros((nrow, nchan) + tuple(corr_shape), T)
    rc = fn(_p(ti), _p(a1), _p(a2), _p(cv[0]), _p(cv[1]), _p(cv[2]), _p(cv[3]), _p(cv[4]),
            _p(cv[5]), _i64(nsrc), _i64(nrow), _i64(ntime), _i64(nant), _i64(nchan),
            _i64(ncorr), _int(mode), _p(out))
    assert rc == 0, rc
    return out.astype(out_dtype, copy=False)


def apply_gains(time_index, antenna1, antenna2, die1_jones, corrupted_vis, die2_jones):
    return predict_vis(time_index, antenna1, antenna2, die1_jones=die1_jones,
                       base_vis=corrupted_vis, die2_jones=die2_jones)


# ---------------------------------------------------------------------------
def freq_grid_interp(frequency, beam_freq_map):
    frequency = np.asarray(frequency)
    f = _as(frequency, np.float64)
    m = _as(beam_freq_map, np.float64)
    out = np.empty((f.shape[0], 3), np.float64)
    rc = lib().orc_freq_grid_interp(_p(f), _p(m), _i64(f.shape[0]), _i64(m.shape[0]), _p(out))
    assert rc == 0, rc
    return out.astype(frequency.dtype, copy=False)


def beam_cube_dde(beam, beam_lm_extents, beam_freq_map, lm, parallactic_angles,
                  point_errors, antenna_scaling, frequency):
    beam = np.asarray(beam)
    lw, mh, nud = beam.shape[:3]
    corrs = beam.shape[3:]
    if lw < 2 or mh < 2 or nud < 2:
        raise ValueError("beam_lw, beam_mh and beam_nud must be >= 2")
    ncorr = int(np.prod(corrs)) if corrs else 1
    nsrc = np.asarray(lm).shape[0]
    ntime, nant = np.asarray(parallactic_angles).shape
    nchan = np.asarray(frequency).shape[0]
    if beam.dtype == np.complex64:
        T, fn = np.complex64, lib().orc_beam_cube_dde_c64
    else:
        T, fn = np.complex128, lib().orc_beam_cube_dde_c128
    b = _as(beam, T)
    out = np.empty((nsrc, ntime, nant, nchan) + tuple(corrs), T)
    f64 = np.float64
    rc = fn(_p(b), _p(_as(beam_lm_extents, f64)), _p(_as(beam_freq_map, f64)), _p(_as(lm, f64)),
            _p(_as(parallactic_angles, f64)), _p(_as(point_errors, f64)),
            _p(_as(antenna_scaling, f64)), _p(_as(frequency, f64)),
            _i64(lw), _i64(mh), _i64(nud), _i64(ncorr), _i64(nsrc), _i64(ntime), _i64(nant),
            _i64(nchan), _p(out))
    assert rc == 0, rc
    return out


# ---------------------------------------------------------------------------
def fused_predict(lm, uvw, frequency, brightness, time_index, antenna1, antenna2,
                  dde1_jones=None, dde2_jones=None, die1_jones=None, base_vis=None,
                  die2_jones=None, convention="fourier"):
    """phase_delay (x) brightness -> predict_vis without materialising (s,r,f,...)."""
    sign = _sign(convention)
    brightness = np.asarray(brightness)
    arrs = [dde1_jones, None, dde2_jones, die1_jones, base_vis, die2_jones]
    arrs = [None if a is None else np.asarray(a) for a in arrs]
    # brightness (source, chan, corr...) plays the role of source_coh minus the row axis
    corr_shape = brightness.shape[2:]
    mode = 1 if len(corr_shape) == 2 else 0
    ncorr = int(np.prod(corr_shape))
    nsrc, nchan = brightness.shape[:2]
    nrow = np.asarray(uvw).shape[0]
    if arrs[0] is not None:
        ntime, nant = arrs[0].shape[1:3]
    elif arrs[3] is not None:
        ntime, nant = arrs[3].shape[:2]
    else:
        ntime, nant = 1, 1
    c128, f64 = np.complex128, np.float64
    cv = [None if a is None else _as(a, c128) for a in arrs]
    out = np.zeros((nrow, nchan) + tuple(corr_shape), c128)
    rc = lib().orc_fused_predict_c128(
        _p(_as(lm, f64)), _p(_as(uvw, f64)), _p(_as(frequency, f64)), _p(_as(brightness, c128)),
        _p(_as(time_index, np.int64)), _p(_as(antenna1, np.int64)), _p(_as(antenna2, np.int64)),
        _p(cv[0]), _p(cv[2]), _p(cv[3]), _p(cv[4]), _p(cv[5]),
        _i64(nsrc), _i64(nrow), _i64(ntime), _i64(nant), _i64(nchan), _i64(ncorr),
        _int(mode), _int(sign), _p(out))
    assert rc == 0, rc
    return out


# ---------------------------------------------------------------------------
def wsclean_spectra(flux, coeffs, log_poly, ref_freq, frequency):
    flux, coeffs, ref_freq, frequency = (_as(a, np.float64) for a in (flux, coeffs, ref_freq, frequency))
    ns, nf = flux.shape[0], frequency.shape[0]
    lp = np.ascontiguousarray(np.broadcast_to(np.asarray(log_poly, dtype=np.uint8), (ns,)))
    out = np.zeros((ns, nf), np.float64)
    rc = lib().orc_wsclean_spectra(_p(flux), _p(coeffs), _p(lp), _p(ref_freq), _p(frequency),
                                   _i64(ns), _i64(coeffs.shape[1]), _i64(nf), _p(out))
    assert rc == 0, rc
    return out


def wsclean_predict(uvw, lm, source_type, flux, coeffs, log_poly, ref_freq, gauss_shape, frequency):
    spectrum = wsclean_spectra(flux, coeffs, log_poly, ref_freq, frequency)
    st = np.asarray(source_type)
    if not np.all((st == "POINT") | (st == "GAUSSIAN")):
        raise ValueError("source_type must be POINT or GAUSSIAN")
    isg = np.ascontiguousarray((st == "GAUSSIAN").astype(np.uint8))
    uvw, lm, gs, fr = (_as(a, np.float64) for a in (uvw, lm, gauss_shape, frequency))
    out = np.zeros((uvw.shape[0], fr.shape[0], 1), np.complex128)
    rc = lib().orc_wsclean_predict(_p(uvw), _p(lm), _p(isg), _p(gs), _p(fr), _p(spectrum),
                                   _i64(lm.shape[0]), _i64(uvw.shape[0]), _i64(fr.shape[0]), _p(out))
    assert rc == 0, rc
    return out


# ---------------------------------------------------------------------------
# Brightness from Stokes parameters
_BASES = {0: 0, "std": 0, 1: 1, "log": 1, 2: 2, "log10": 2}


def spectral_model(stokes, spi, ref_freq, frequency, base=0):
    """africanus/model/spectral/spec_model.py:106-211."""
    stokes, spi = np.asarray(stokes), np.asarray(spi)
    dtype = np.result_type(stokes, spi, np.asarray(ref_freq), np.asarray(frequency))
    if spi.ndim - 2 != stokes.ndim - 1:
        raise ValueError("Dimensions on stokes and spi don't agree")
    out_shape = (stokes.shape[0], np.asarray(frequency).shape[0]) + stokes.shape[1:]
    npol = int(np.prod(stokes.shape[1:], dtype=np.int64))
    if npol != int(np.prod(spi.shape[2:], dtype=np.int64)):
        raise ValueError("Correlations on stokes and spi don't agree")
    bases = list(base) if isinstance(base, (list, tuple)) else [base]
    bases = (bases + [bases[-1]] * npol)[:npol]
    if any(b not in _BASES for b in bases):
        raise ValueError("Invalid base")
    b = np.array([_BASES[x] for x in bases], np.int32)
    ns, nspi, nf = stokes.shape[0], spi.shape[1], out_shape[1]
    st = _as(stokes.reshape(ns, npol), np.float64)
    sp = _as(spi.reshape(ns, nspi, npol), np.float64)
    rf, fr = _as(ref_freq, np.float64), _as(frequency, np.float64)
    out = np.zeros((ns, nf, npol), np.float64)
    rc = lib().orc_spectral_model(_p(st), _p(sp), _p(rf), _p(fr), _p(b), _i64(ns), _i64(nspi),
                                  _i64(npol), _i64(nf), _p(out))
    assert rc == 0, rc
    return out.reshape(out_shape).astype(dtype, copy=False)


# output element -> ((input one, input two), op code of orc_convert); conversion.py:17-48
_CONV = {
    "RR": [(("I", "V"), 0)], "RL": [(("Q", "U"), 2)], "LR": [(("Q", "U"), 3)], "LL": [(("I", "V"), 1)],
    "XX": [(("I", "Q"), 0)], "XY": [(("U", "V"), 2)], "YX": [(("U", "V"), 3)], "YY": [(("I", "Q"), 1)],
    "I": [(("XX", "YY"), 4), (("RR", "LL"), 4)], "Q": [(("XX", "YY"), 5), (("RL", "LR"), 4)],
    "U": [(("XY", "YX"), 4), (("RL", "LR"), 6)], "V": [(("XY", "YX"), 6), (("RR", "LL"), 5)],
}
_STOKES_NAMES = ["Undefined", "I", "Q", "U", "V", "RR", "RL", "LR", "LL", "XX", "XY", "YX", "YY"]


def _schema(schema):
    """name -> flat index, shape (conversion.py:93-142)."""
    arr = np.asarray(schema, dtype=object)
    if arr.ndim == 0:
        arr = arr.reshape(1)
    names = {}
    for flat, e in enumerate(arr.reshape(-1)):
        if not isinstance(e, str):
            e = _STOKES_NAMES[int(e)]
        if e in names:
            raise ValueError("'%s' defined multiple times" % e)
        names[e] = flat
    return names, arr.shape


def convert(input, input_schema, output_schema, implicit_stokes=False):
    """africanus/model/coherency/conversion.py:145-243."""
    input = np.asarray(input)
    in_idx, in_shape = _schema(input_schema)
    out_idx, out_shape = _schema(output_schema)
    if input.shape[input.ndim - len(in_shape):] != in_shape:
        raise ValueError("Last dimension of input doesn't match input schema")
    nout = len(out_idx)
    s1, s2, op = (np.zeros(nout, np.int32) for _ in range(3))
    cplx = False
    for okey, o in out_idx.items():
        if okey not in _CONV:
            raise ValueError("Unknown output %s" % okey)
        defaults_ok = implicit_stokes and okey in ("RR", "RL", "LR", "LL", "XX", "XY", "YX", "YY")
        best = None
        for (c1, c2), code in _CONV[okey]:
            have = (c1 in in_idx) + (c2 in in_idx)
            if have < 2 and not defaults_ok:
                continue
            if best is None or have > best[0]:
                best = (have, in_idx.get(c1, -1), in_idx.get(c2, -1), code)
        if best is None:
            raise ValueError("None of the supplied inputs can produce output '%s'" % okey)
        s1[o], s2[o], op[o] = best[1:]
        cplx = cplx or best[3] not in (4, 5)
    zero = np.zeros((), input.dtype)
    dtype = np.result_type((zero + zero + 0j).dtype if cplx else (zero / 2).dtype, (zero / 2).dtype)
    lead = input.shape[:input.ndim - len(in_shape)]
    n, nin = int(np.prod(lead, dtype=np.int64)), len(in_idx)
    flat = _as(input.reshape(n, nin), np.complex128)
    out = np.zeros((n, nout), np.complex128)
    rc = lib().orc_convert(_p(flat), _i64(n), _i64(nin), _p(s1), _p(s2), _p(op), _i64(nout), _p(out))
    assert rc == 0, rc
    out = out.reshape(lead + out_shape)
    return (out if np.issubdtype(dtype, np.complexfloating) else out.real).astype(dtype)


# ---------------------------------------------------------------------------
def feed_rotation(parallactic_angles, feed_type="linear"):
    """africanus/rime/feeds.py:13-71."""
    if feed_type not in ("linear", "circular"):
        raise ValueError("Invalid feed_type '%s'" % feed_type)
    pa = np.asarray(parallactic_angles)
    if pa.dtype not in (np.float32, np.float64):
        raise ValueError("parallactic_angles has none-floating point type %s" % pa.dtype)
    f32 = pa.dtype == np.float32
    flat = _as(pa.reshape(-1), pa.dtype)
    out = np.zeros((flat.shape[0], 2, 2), np.complex64 if f32 else np.complex128)
    fn = lib().orc_feed_rotation_f32 if f32 else lib().orc_feed_rotation_f64
    rc = fn(_p(flat), _i64(flat.shape[0]), _int(feed_type == "circular"), _p(out))
    assert rc == 0, rc
    return out.reshape(pa.shape + (2, 2))
