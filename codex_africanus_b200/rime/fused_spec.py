"""
The reference's fused-RIME specification front end on the B200 kernels.

``rime("(Lp, Ep, Kpq, Bpq, Eq, Lq): [I,Q,U,V] -> [XX,XY,YX,YY]", dataset, convention=...,
spi_base=...)`` of africanus/experimental/rime/fused/core.py:227-241 (specification grammar:
experimental/rime/fused/specification.py:78-115,166-185,440-453).  The reference compiles the
term list into one numba loop nest; here the same string selects among the fused CUDA entry
points, whose terms are exactly the reference's:

    Kpq  phase delay              terms/phase.py            -> the phasor-stream / GEMM kernels
    Bpq  brightness from Stokes   terms/brightness.py       -> ``stokes_brightness`` on the device
    Ep/Eq  beam cube DDE          terms/cube_dde.py:96-313  -> ``fused_predict_vis_beam``
    Lp/Lq  feed rotation          terms/feed_rotation.py    -> DIE (outside E) or the beam kernel's
                                                               rotation epilogue (inside E)

Inputs are looked up by the reference's names (time, antenna1, antenna2, feed1, feed2, radec +
phase_dir or lm, uvw, chan_freq, stokes, spi, ref_freq, beam, beam_lm_extents, beam_freq_map,
beam_parangle, feed_parangle) in the mappings / keyword arguments given after the specification.
The parallactic-angle arrays are taken in the layouts the reference's transformer produces
(transformers/parangle.py:80-117: ``feed_parangle (time, feed, ant, 2, 2)`` and ``beam_parangle
(time, feed, ant, 2)`` holding sin / cos) or as plain ``parallactic_angles (time, ant)``; they are
not derived from antenna positions here (the reference calls casacore for that).  One feed per
antenna; the Gaussian shape term (Cpq) is not built.  ``in_kernel=True`` asks for the beam to be sampled
inside the predict kernel as the reference's fused loop does (``fused_predict_vis_beam(in_kernel=True)``).
"""
import re
from collections.abc import Mapping

import numpy as np
import torch

from .. import _plumbing as pl

REQUIRED_ARGS = ("time", "antenna1", "antenna2", "feed1", "feed2")
DEFAULT_SPEC = "(Kpq, Bpq): [I, Q, U, V] -> [XX, XY, YX, YY]"
_TERM = re.compile(r"^([A-Z])(pq|p|q)$")
_NAME = re.compile(r"^[A-Za-z][A-Za-z0-9]*$")
_LINEAR, _CIRCULAR = {"XX", "XY", "YX", "YY"}, {"RR", "RL", "LR", "LL"}
_BASES = {"standard": "std", "std": "std", "log": "log", "log10": "log10"}


class RimeParseError(ValueError):
    pass


class RimeSpecificationError(ValueError):
    pass


def _parse_list(text, what, brackets):
    text = text.strip()
    if len(text) < 2 or text[0] not in brackets or text[-1] != brackets[text[0]]:
        raise RimeParseError("%s must be of the form %s. Got %s." % (what[0], what[1], text))
    names = [t.strip() for t in text[1:-1].split(",")]
    if names and names[-1] == "":
        names.pop()  # a trailing comma, as in a Python tuple
    if not names or not all(_NAME.match(n) for n in names):
        raise RimeParseError("%s must be of the form %s. Got %s." % (what[0], what[1], text))
    return names


def parse_rime(rime_spec):
    """``"(Kpq, Bpq): [I,Q,U,V] -> [XX,XY,YX,YY]"`` -> (terms, stokes, corrs), three lists of names
    (specification.py:78-115)."""
    bits = [s.strip() for s in str(rime_spec).split(":")]
    if len(bits) != 2:
        raise RimeParseError("RIME must be of the form [Gp, (Kpq, Bpq), Gq]: [I,Q,U,V] -> [XX,XY,YX,YY]. "
                             "Got %s." % (rime_spec,))
    pol = [s.strip() for s in bits[1].split("->")]
    if len(pol) != 2:
        raise RimeParseError("Polarisation specification must be of the form [I,Q,U,V] -> [XX,XY,YX,YY]. "
                             "Got %s." % bits[1])
    stokes = _parse_list(pol[0], ("Stokes specification", "[I,Q,U,V]"), {"[": "]"})
    corrs = _parse_list(pol[1], ("Correlation specification", "[XX,XY,YX,YY]"), {"[": "]"})
    terms = _parse_list(bits[0], ("RIME", "a tuple/list of Terms (Kpq, Bpq)"), {"(": ")", "[": "]"})
    return terms, [s.upper() for s in stokes], [c.upper() for c in corrs]


def feed_type_of(corrs):
    """"linear" / "circular" from the correlation names (specification.py:440-453)."""
    sc = set(corrs)
    if sc.issubset(_LINEAR):
        return "linear"
    if sc.issubset(_CIRCULAR):
        return "circular"
    raise RimeSpecificationError("Correlations must be purely linear or circular. Got %s" % (corrs,))


def split_terms(terms):
    """(left letters, middle letters, right letters) of a term list: antenna-p terms, then baseline
    (pq) terms, then antenna-q terms mirroring the left ones, e.g. (Lp, Ep, Kpq, Bpq, Eq, Lq)."""
    parsed = []
    for t in terms:
        m = _TERM.match(t)
        if not m:
            raise RimeSpecificationError("%s does not match %s" % (t, _TERM.pattern))
        parsed.append(m.groups())
    order = {"p": 0, "pq": 1, "q": 2}
    ranks = [order[i] for _, i in parsed]
    if ranks != sorted(ranks):
        raise RimeSpecificationError("terms must run from antenna p over baseline pq to antenna q: %s" % (terms,))
    left = [x for x, i in parsed if i == "p"]
    mid = [x for x, i in parsed if i == "pq"]
    right = [x for x, i in parsed if i == "q"]
    if right != left[::-1]:
        raise RimeSpecificationError("the antenna-q terms must mirror the antenna-p terms: %s" % (terms,))
    for x in left + mid:
        if x == "C":
            raise NotImplementedError("the Gaussian shape term (C) is not built on the B200 path")
        if x not in ("K", "B", "E", "L"):
            raise RimeSpecificationError("Unknown term %s" % x)
    if any(x in ("K", "B") for x in left) or any(x in ("E", "L") for x in mid):
        raise RimeSpecificationError("K and B are baseline (pq) terms, E and L antenna (p / q) terms: %s" % (terms,))
    if len(set(left)) != len(left) or len(set(mid)) != len(mid):
        raise RimeSpecificationError("a term may appear once per side: %s" % (terms,))
    if "B" not in mid:
        raise NotImplementedError("a RIME without a brightness term (Bpq) is not built")
    return left, mid, right


def consolidate_args(args, kw):
    """Mappings (a dataset is a mapping of lower-case names to arrays), then positional arrays in
    the order of ``REQUIRED_ARGS``, then keywords (core.py:208-224)."""
    mapping, positional = {}, []
    for element in args:
        if isinstance(element, Mapping):
            mapping.update((str(k).lower(), v) for k, v in element.items())
        else:
            positional.append(element)
    mapping.update(zip(REQUIRED_ARGS, positional))
    mapping.update(kw)
    return mapping


def radec_to_lm(radec, phase_dir):
    """transformers/lm.py:22-40."""
    radec, phase_dir = np.asarray(radec, np.float64), np.asarray(phase_dir, np.float64)
    da = radec[:, 0] - phase_dir[0]
    lm = np.empty_like(radec)
    lm[:, 0] = np.cos(radec[:, 1]) * np.sin(da)
    lm[:, 1] = np.sin(radec[:, 1]) * np.cos(phase_dir[1]) - np.cos(radec[:, 1]) * np.sin(phase_dir[1]) * np.cos(da)
    return lm


def _host(a):
    return a.detach().cpu().numpy() if pl.is_torch(a) else np.asarray(a)


def _feed_matrices(m, corrs, ntime, nant):
    """(time, ant, 2, 2) feed rotation from ``feed_parangle`` (time, feed, ant, 2, 2) sin / cos of the
    two receptor angles (terms/feed_rotation.py:44-65) or from ``parallactic_angles`` (time, ant)."""
    linear = feed_type_of(corrs) == "linear"
    if m.get("feed_parangle") is not None:
        fp = _host(m["feed_parangle"])
        if fp.ndim != 5 or fp.shape[1] != 1 or fp.shape[3:] != (2, 2):
            raise NotImplementedError("feed_parangle must be (time, 1 feed, ant, 2, 2); got %s" % (fp.shape,))
        sa, ca, sb, cb = fp[:, 0, :, 0, 0], fp[:, 0, :, 0, 1], fp[:, 0, :, 1, 0], fp[:, 0, :, 1, 1]
    elif m.get("parallactic_angles") is not None:
        pa = _host(m["parallactic_angles"])
        sa = sb = np.sin(pa)
        ca = cb = np.cos(pa)
    else:
        raise ValueError("feed rotation (L) needs feed_parangle or parallactic_angles")
    if sa.shape != (ntime, nant):
        raise ValueError("feed parallactic angles have shape %s, expected (%d, %d)" % (sa.shape, ntime, nant))
    L = np.empty((ntime, nant, 2, 2), np.complex128)
    if linear:
        L[..., 0, 0], L[..., 0, 1], L[..., 1, 0], L[..., 1, 1] = ca, sa, -sb, cb
    else:
        L[..., 0, 0] = 0.5 * ((ca + cb) - (sa + sb) * 1j)
        L[..., 0, 1] = 0.5 * ((ca - cb) + (sa - sb) * 1j)
        L[..., 1, 0] = 0.5 * ((ca - cb) - (sa - sb) * 1j)
        L[..., 1, 1] = 0.5 * ((ca + cb) + (sa + sb) * 1j)
    return L


def _beam_angles(m, ntime, nant):
    if m.get("beam_parangle") is not None:
        bp = _host(m["beam_parangle"])
        if bp.ndim != 4 or bp.shape[1] != 1 or bp.shape[3] != 2:
            raise NotImplementedError("beam_parangle must be (time, 1 feed, ant, 2); got %s" % (bp.shape,))
        pa = np.arctan2(bp[:, 0, :, 0], bp[:, 0, :, 1])
    elif m.get("parallactic_angles") is not None:
        pa = _host(m["parallactic_angles"]).astype(np.float64)
    else:
        raise ValueError("the beam term (E) needs beam_parangle or parallactic_angles")
    if pa.shape != (ntime, nant):
        raise ValueError("beam parallactic angles have shape %s, expected (%d, %d)" % (pa.shape, ntime, nant))
    return pa


def rime(rime_spec, *args, **kw):
    """Evaluate the RIME given by ``rime_spec`` on the inputs in ``*args`` (mappings / datasets, or
    the five required arrays) and ``**kw``; returns (row, chan, corr) complex128 visibilities like
    africanus.experimental.rime.fused.core.rime.  numpy inputs give a numpy result, any torch CUDA
    input a CUDA tensor."""
    from ..model import stokes_brightness
    from .fused_beam import fused_predict_vis_beam
    from .fused_stokes import fused_predict_vis_stokes

    m = consolidate_args(args, kw)
    terms, stokes_schema, corrs = parse_rime(rime_spec)
    left, mid, _ = split_terms(terms)
    for name in REQUIRED_ARGS + ("uvw", "chan_freq", "stokes", "spi", "ref_freq"):
        if m.get(name) is None:
            raise ValueError("rime: missing input '%s'" % name)
    if len(corrs) not in (1, 2, 4):
        raise RimeSpecificationError("1, 2 or 4 correlations are supported. Got %s" % (corrs,))
    if left and len(corrs) != 4:
        raise RimeSpecificationError("Four correlations required for %s terms but %s were specified"
                                     % ("/".join(left), corrs))
    for f in ("feed1", "feed2"):
        if np.unique(_host(m[f])).size > 1:
            raise NotImplementedError("one feed per antenna is supported")
    convention = m.get("convention", "fourier")
    spi_base = m.get("spi_base", "standard")
    if spi_base not in _BASES:
        raise ValueError("Invalid base")
    base = _BASES[spi_base]

    # unique times / antennas index the (time, ...) and (ant, ...) axes (core.py term state)
    utime, time_index = np.unique(_host(m["time"]), return_inverse=True)
    a1h, a2h = _host(m["antenna1"]), _host(m["antenna2"])
    uant, inv = np.unique(np.concatenate((a1h, a2h)), return_inverse=True)
    ant1, ant2 = inv[: a1h.size].astype(np.int32), inv[a1h.size:].astype(np.int32)
    ntime, nant = utime.size, uant.size
    time_index = time_index.astype(np.int32)

    lm = m["lm"] if m.get("lm") is not None else radec_to_lm(_host(m["radec"]), _host(m["phase_dir"]))
    uvw, freq = m["uvw"], m["chan_freq"]
    if "K" not in mid:  # no phase term: every phasor is one
        uvw = np.zeros(pl.shape_of(uvw))
    nchan = pl.shape_of(freq)[0]
    corr_schema = [corrs[:2], corrs[2:]] if len(corrs) == 4 else list(corrs)
    stokes_args = (m["stokes"], m["spi"], m["ref_freq"])
    everything = (lm, uvw, freq) + stokes_args + tuple(m.get(k) for k in ("beam", "feed_parangle", "beam_parangle"))
    as_torch = pl.wants_torch(*everything)

    die = None
    if "L" in left and (left[0] == "L" or "E" not in left):
        # feed rotation outermost: a direction-independent 2x2 term, the same for every channel
        L = _feed_matrices(m, corrs, ntime, nant)
        die = np.ascontiguousarray(np.broadcast_to(L[:, :, None], (ntime, nant, nchan, 2, 2)))
    if "E" not in left:
        out = fused_predict_vis_stokes(lm, uvw, freq, *stokes_args, time_index, ant1, ant2, die1_jones=die,
                                       die2_jones=die, convention=convention, base=base,
                                       stokes_schema=tuple(stokes_schema), corr_schema=corr_schema,
                                       dtype=np.complex128)
    else:
        for name in ("beam", "beam_lm_extents", "beam_freq_map"):
            if m.get(name) is None:
                raise ValueError("rime: the beam term (E) needs '%s'" % name)
        pa = _beam_angles(m, ntime, nant)
        feed_type = None
        if "L" in left and left[0] == "E":
            # feed rotation between the beam and the sky: dde = beam . L in the beam kernel's epilogue;
            # it rotates by the beam's own parallactic angles, so the two must agree
            L = _feed_matrices(m, corrs, ntime, nant)
            ref = _feed_matrices({"parallactic_angles": pa}, corrs, ntime, nant)
            if not np.allclose(L, ref, rtol=0, atol=1e-12):
                raise NotImplementedError("(Ep, Lp, ...) with receptor angles: feed and beam angles must agree")
            feed_type = feed_type_of(corrs)
        beam = m["beam"]
        bshape = pl.shape_of(beam)
        if len(bshape) != 4 or bshape[3] != 4:
            raise ValueError("beam %s should be a (beam_lw, beam_mh, beam_nud, 4 corr) array" % (bshape,))
        beam = beam.reshape(bshape[:3] + (2, 2))
        bright = stokes_brightness(*stokes_args, freq, base=base, stokes_schema=tuple(stokes_schema),
                                   corr_schema=corr_schema, dtype=np.complex128, device_out=True)
        # the reference's cube term applies neither pointing errors nor antenna scaling (cube_dde.py:207-218)
        perr = np.zeros((ntime, nant, nchan, 2))
        ascale = np.ones((nant, nchan, 2))
        out = fused_predict_vis_beam(lm, uvw, freq, bright, time_index, ant1, ant2, beam, m["beam_lm_extents"],
                                     m["beam_freq_map"], pa, perr, ascale, die, None, die, convention=convention,
                                     feed_type=feed_type, in_kernel=bool(m.get("in_kernel", False)))
        if not as_torch and pl.is_torch(out):
            out = pl.to_host(out)
    nrow = pl.shape_of(out)[0]
    return out.reshape(nrow, nchan, len(corrs))
