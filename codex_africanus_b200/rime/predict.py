"""``predict_vis`` / ``apply_gains`` on B200 -- africanus/rime/predict.py:466-649."""
import numpy as np
import torch

from .. import _lib
from .. import _plumbing as pl


def predict_checks(time_index, antenna1, antenna2, dde1_jones, source_coh, dde2_jones,
                   die1_jones, base_vis, die2_jones):
    """Presence / ndim rules of africanus/rime/predict.py:380-463 plus the correlation-mode
    rules of :15-53,546-563.  Returns (jones_mode, corr_shape)."""
    for idx in (time_index, antenna1, antenna2):
        assert len(pl.shape_of(idx)) == 1
    have = [a is not None for a in (dde1_jones, source_coh, dde2_jones, die1_jones, base_vis,
                                    die2_jones)]
    if have[0] ^ have[2]:
        raise ValueError("Both dde1_jones and dde2_jones must be present or absent")
    if have[3] ^ have[5]:
        raise ValueError("Both die1_jones and die2_jones must be present or absent")
    spec = [("dde1_jones", dde1_jones, 5, 6), ("source_coh", source_coh, 4, 5),
            ("dde2_jones", dde2_jones, 5, 6), ("die1_jones", die1_jones, 4, 5),
            ("base_vis", base_vis, 3, 4), ("die2_jones", die2_jones, 4, 5)]
    ndims = {name: len(pl.shape_of(a)) for name, a, _, _ in spec if a is not None}
    for name, a, d1, d2 in spec:
        if a is not None and ndims[name] not in (d1, d2):
            raise ValueError("%s.ndim %d not in (%d, %d)" % (name, ndims[name], d1, d2))
    if have[0] and ndims["dde1_jones"] != ndims["dde2_jones"]:
        raise ValueError("dde1_jones.ndim != dde2_jones.ndim")
    if have[3] and ndims["die1_jones"] != ndims["die2_jones"]:
        raise ValueError("die1_jones.ndim != die2_jones.ndim")
    modes = [ndims[name] - d1 for name, a, d1, _ in spec if a is not None]
    if not modes:
        raise ValueError("No Jones Matrices were supplied")
    if any(m != modes[0] for m in modes):
        # the reference reports inconsistent ndims either as the pre-condition error
        # (predict.py:454-461) or as mismatched correlations (:557-558)
        raise ValueError("Jones Matrix Correlations were mismatched: one of the following "
                         "pre-conditions is broken (missing values are ignored):\n"
                         "dde_jones{1,2}.ndim == source_coh.ndim + 1\n"
                         "dde_jones{1,2}.ndim == base_vis.ndim + 2\n"
                         "dde_jones{1,2}.ndim == die_jones{1,2}.ndim + 1")
    corr_shape = None
    for name, a, d1, _ in spec:
        if a is not None:
            cs = tuple(pl.shape_of(a)[d1 - 1:])
            if corr_shape is None:
                corr_shape = cs
            elif cs != corr_shape:
                raise ValueError("Jones Matrix Correlations were mismatched")
    if modes[0] == 1 and corr_shape != (2, 2):
        raise ValueError("Jones Matrix Correlations were mismatched: matrix mode needs (2, 2)")
    return (_lib.AFR_JONES_2X2 if modes[0] == 1 else _lib.AFR_JONES_DIAG), corr_shape


def normalise_indices(time_index, antenna1, antenna2, device):
    """int32 device copies; time_index minus its minimum (predict.py:597)."""
    ti = pl.to_device(time_index, np.int64, device) if not pl.is_torch(time_index) else \
        time_index.to(device=device, dtype=torch.int64)
    if ti.numel():
        ti = ti - ti.min()
    ti = ti.to(torch.int32).contiguous()
    a1 = pl.to_device(antenna1, np.int32, device)
    a2 = pl.to_device(antenna2, np.int32, device)
    return ti, a1, a2


def predict_vis(time_index, antenna1, antenna2, dde1_jones=None, source_coh=None,
                dde2_jones=None, die1_jones=None, base_vis=None, die2_jones=None):
    """V = G1 (B + sum_s E1 X E2^H) G2^H, africanus/rime/predict.py:466-619.

    Same argument meaning, presence rules, correlation layouts ((1,), (2,) element-wise;
    (2,2) matrix) and output dtype (``np.result_type`` of the present arrays) as the
    reference.  Index arrays may have any integer dtype; inputs may be strided.
    """
    arrs = (dde1_jones, source_coh, dde2_jones, die1_jones, base_vis, die2_jones)
    mode, corr_shape = predict_checks(time_index, antenna1, antenna2, *arrs)
    out_dtype = np.result_type(*(pl.dtype_of(a) for a in arrs if a is not None))
    if out_dtype not in (np.complex64, np.complex128):
        raise TypeError("predict_vis: Jones terms must be complex (got %s)" % out_dtype)
    nrow = pl.shape_of(time_index)[0]
    ncorr = int(np.prod(corr_shape))
    nsrc, ntime, nant = 0, 1, 1
    if dde1_jones is not None:
        nsrc, ntime, nant, nchan = pl.shape_of(dde1_jones)[:4]
    elif source_coh is not None:
        nsrc, _, nchan = pl.shape_of(source_coh)[:3]
    elif die1_jones is not None:
        nchan = pl.shape_of(die1_jones)[2]
    else:
        nchan = pl.shape_of(base_vis)[1]
    if die1_jones is not None:
        ntime, nant = pl.shape_of(die1_jones)[:2]

    device = pl.pick_device(*arrs, time_index)
    as_torch = pl.wants_torch(*arrs, time_index, antenna1, antenna2)
    with torch.cuda.device(device):
        dev = [None if a is None else pl.to_device(a, out_dtype, device) for a in arrs]
        ti, a1, a2 = normalise_indices(time_index, antenna1, antenna2, device)
        d_out = pl.empty_device((nrow, nchan) + tuple(corr_shape), out_dtype, device)
        pl.call("afr_predict_vis", device, pl.ptr(ti), pl.ptr(a1), pl.ptr(a2),
                *(pl.ptr(d) for d in dev), nsrc, nrow, ntime, nant, nchan, ncorr, mode,
                int(out_dtype == np.complex64), pl.ptr(d_out), pl.stream_ptr(device))
        return d_out if as_torch else pl.to_host(d_out)


def apply_gains(time_index, antenna1, antenna2, die1_jones, corrupted_vis, die2_jones):
    """africanus/rime/predict.py:622-649."""
    return predict_vis(time_index, antenna1, antenna2, die1_jones=die1_jones,
                       base_vis=corrupted_vis, die2_jones=die2_jones)
