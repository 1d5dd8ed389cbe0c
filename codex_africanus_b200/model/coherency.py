"""``convert`` (Stokes <-> correlations) on B200 -- africanus/model/coherency/conversion.py --
and the composed Stokes -> brightness used by the fused predict."""
import ctypes

import numpy as np
import torch

from .. import _plumbing as pl
from .spectral import promote_bases, spectral_shapes

# CASA Stokes ids of the elements handled here (africanus/util/casa_types.py:4-17)
_ID_TO_NAME = {1: "I", 2: "Q", 3: "U", 4: "V", 5: "RR", 6: "RL", 7: "LR", 8: "LL",
               9: "XX", 10: "XY", 11: "YX", 12: "YY"}
_ALL_IDS = 33  # casa_types.py lists ids 0..32

# op codes of afr_convert (include/africanus_b200.h)
_ADD, _SUB, _ADD_I, _SUB_I, _HALF_ADD, _HALF_SUB, _HALF_SUB_OVER_I = range(7)

# output element -> candidate ((input one, input two), op), in the reference's order
# (conversion.py:19-48)
_RULES = {
    "RR": ((("I", "V"), _ADD),), "RL": ((("Q", "U"), _ADD_I),),
    "LR": ((("Q", "U"), _SUB_I),), "LL": ((("I", "V"), _SUB),),
    "XX": ((("I", "Q"), _ADD),), "XY": ((("U", "V"), _ADD_I),),
    "YX": ((("U", "V"), _SUB_I),), "YY": ((("I", "Q"), _SUB),),
    "I": ((("XX", "YY"), _HALF_ADD), (("RR", "LL"), _HALF_ADD)),
    "Q": ((("XX", "YY"), _HALF_SUB), (("RL", "LR"), _HALF_ADD)),
    "U": ((("XY", "YX"), _HALF_ADD), (("RL", "LR"), _HALF_SUB_OVER_I)),
    "V": ((("XY", "YX"), _HALF_SUB_OVER_I), (("RR", "LL"), _HALF_SUB)),
}
_CORRELATIONS = ("RR", "RL", "LR", "LL", "XX", "XY", "YX", "YY")


class DimensionMismatch(Exception):
    pass


class MissingConversionInputs(Exception):
    pass


def schema_elements(schema):
    """Nested schema -> ({name: flat C-order index}, shape) (conversion.py:93-142)."""
    if not isinstance(schema, (tuple, list)):
        schema = [schema]
    names, shape = {}, []

    def walk(node, depth):
        if len(shape) <= depth:
            shape.append(len(node))
        elif shape[depth] != len(node):
            raise DimensionMismatch("Dimension mismatch %d != %d at depth %d"
                                    % (shape[depth], len(node), depth))
        for e in node:
            if isinstance(e, (tuple, list)):
                walk(e, depth + 1)
                continue
            if isinstance(e, str):
                name = e
            elif isinstance(e, (int, np.integer)) and not isinstance(e, (bool, np.bool_)):
                if not 0 <= int(e) < _ALL_IDS:
                    raise ValueError("Invalid id '%s'" % (e,))
                name = _ID_TO_NAME.get(int(e), "id%d" % int(e))
            else:
                raise TypeError("Invalid type '%s' for element '%s'" % (type(e), e))
            if name in names:
                raise ValueError("'%s' defined multiple times" % name)
            names[name] = len(names)

    walk(schema, 0)
    if len(names) != int(np.prod(shape)):
        raise DimensionMismatch("schema is ragged")
    return names, tuple(shape)


def resolve(in_names, out_names, implicit_stokes):
    """One (input one, input two, op) per output element (conversion.py:156-209): of the rules
    that can be satisfied, the one with most real (non-default) inputs, first listed on ties."""
    nout = len(out_names)
    s1, s2, op = ((ctypes.c_int * nout)() for _ in range(3))
    for okey, o in out_names.items():
        if okey not in _RULES:
            raise ValueError("Unknown output %s. Known outputs: %s" % (okey, list(_RULES)))
        defaults_ok = implicit_stokes and okey in _CORRELATIONS
        best = None
        for (c1, c2), code in _RULES[okey]:
            have = (c1 in in_names) + (c2 in in_names)
            if have < 2 and not defaults_ok:
                continue
            if best is None or have > best[0]:
                best = (have, in_names.get(c1, -1), in_names.get(c2, -1), code)
        if best is None:
            raise MissingConversionInputs(
                "None of the supplied inputs '%s' can produce output '%s'. It can be produced by "
                "the following combinations '%s'."
                % (list(in_names), okey, [r[0] for r in _RULES[okey]]))
        s1[o], s2[o], op[o] = best[1:]
    return s1, s2, op


def _output_dtype(in_dtype, op):
    """result_type of the per-output lambda results (conversion.py:194,212)."""
    zero = np.zeros((), in_dtype)
    kinds = [(zero / 2).dtype if code in (_HALF_ADD, _HALF_SUB) else (zero + zero + 0j).dtype
             for code in op]
    return np.result_type(*kinds)


def convert(input, input_schema, output_schema, implicit_stokes=False):
    """Convert between Stokes parameters and linear / circular correlations.

    ``input`` (..., icorr_1..icorr_m) real or complex with its last dimensions described by
    ``input_schema`` (nested lists of "I","Q","U","V","XX",... or CASA Stokes ids);
    returns (..., ocorr_1..ocorr_n) as ``output_schema``.  ``implicit_stokes`` lets missing Stokes
    inputs count as zero when producing correlations.  Output dtype as the reference: complex of
    the input's precision unless every output is a plain half-sum / half-difference.
    """
    in_names, in_shape = schema_elements(input_schema)
    out_names, out_shape = schema_elements(output_schema)
    shape = pl.shape_of(input)
    if tuple(shape[len(shape) - len(in_shape):]) != in_shape or len(shape) < len(in_shape):
        raise ValueError("Last dimension of input doesn't match input schema")
    s1, s2, op = resolve(in_names, out_names, implicit_stokes)
    in_dtype = pl.dtype_of(input)
    out_dtype = _output_dtype(in_dtype, op)
    lead = tuple(shape[:len(shape) - len(in_shape)])
    n, nin, nout = int(np.prod(lead, dtype=np.int64)), len(in_names), len(out_names)
    in_complex = int(np.issubdtype(in_dtype, np.complexfloating))
    device = pl.pick_device(input)
    as_torch = pl.wants_torch(input)
    with torch.cuda.device(device):
        d_in = pl.to_device(input, np.complex128 if in_complex else np.float64, device)
        d_out = pl.empty_device(lead + out_shape, np.complex128, device)
        pl.call("afr_convert", device, pl.ptr(d_in), in_complex, n, nin, s1, s2, op, nout, pl.ptr(d_out),
                pl.stream_ptr(device))
        if not np.issubdtype(out_dtype, np.complexfloating):
            d_out = d_out.real
        if pl.dtype_of(d_out) != out_dtype:
            d_out = d_out.to(pl.torch_dtype(out_dtype))
        d_out = d_out.contiguous()
        return d_out if as_torch else pl.to_host(d_out)


LINEAR = [["XX", "XY"], ["YX", "YY"]]
CIRCULAR = [["RR", "RL"], ["LR", "LL"]]


def stokes_brightness(stokes, spi, ref_freq, frequency, base=0, stokes_schema=("I", "Q", "U", "V"),
                      corr_schema=LINEAR, implicit_stokes=False, dtype=None, device_out=False):
    """``convert(spectral_model(stokes, spi, ref_freq, frequency, base), stokes_schema,
    corr_schema)`` in one kernel: the (source, chan, corr...) brightness of the predict
    (africanus/rime/examples/predict.py:107-134) straight from the catalogue columns.

    stokes (source, pol), spi (source, spi-comps, pol) with pol = len(stokes_schema) <= 4.
    ``dtype`` complex64 / complex128 (default: complex of the inputs' precision).
    ``device_out`` returns a CUDA tensor whatever the inputs are.
    """
    nsrc, nspi, npol, nchan, pol_shape = spectral_shapes(stokes, spi, ref_freq, frequency)
    in_names, in_shape = schema_elements(list(stokes_schema))
    out_names, out_shape = schema_elements(corr_schema)
    if len(pol_shape) != 1 or pol_shape != in_shape:
        raise ValueError("Last dimension of input doesn't match input schema")
    if npol > 4:
        raise ValueError("stokes_brightness: at most 4 Stokes parameters")
    for okey in out_names:
        if okey not in _CORRELATIONS:
            raise ValueError("stokes_brightness: '%s' is not a correlation" % okey)
    bases = promote_bases(base, npol)
    s1, s2, op = resolve(in_names, out_names, implicit_stokes)
    arrays = (stokes, spi, ref_freq, frequency)
    real = np.result_type(*(pl.dtype_of(a) for a in arrays))
    out_dtype = np.dtype(dtype) if dtype is not None else np.result_type(real, np.complex64)
    if out_dtype not in (np.complex64, np.complex128):
        raise ValueError("stokes_brightness: dtype must be complex64 or complex128")
    device = pl.pick_device(*arrays)
    as_torch = device_out or pl.wants_torch(*arrays)
    f64 = np.float64
    with torch.cuda.device(device):
        d_s, d_p, d_r, d_f = (pl.to_device(a, f64, device) for a in arrays)
        d_out = pl.empty_device((nsrc, nchan) + out_shape, out_dtype, device)
        pl.call("afr_stokes_brightness", device, pl.ptr(d_s), pl.ptr(d_p), pl.ptr(d_r), pl.ptr(d_f), bases,
                nsrc, nspi, npol, nchan, s1, s2, op, len(out_names), int(out_dtype == np.complex64),
                pl.ptr(d_out), pl.stream_ptr(device))
        return d_out if as_torch else pl.to_host(d_out)
