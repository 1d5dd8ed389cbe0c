#!/usr/bin/env python
"""Quick device-resident timing of the main kernels (development aid, not the bench)."""
import ctypes
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import _lib, dft, rime  # noqa: E402


def timed(fn, reps=3):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best


def main():
    lib = _lib.lib()
    lib.afr_set_device(0)
    peak = ctypes.c_double()
    for fp64 in (1, 0):
        _lib.check(lib.afr_measure_fma_peak(fp64, 20000, ctypes.byref(peak), None))
        print("FMA peak %s: %.2f TFLOP/s" % ("fp64" if fp64 else "fp32", peak.value / 1e12), flush=True)
    fp64_peak = None
    _lib.check(lib.afr_measure_fma_peak(1, 20000, ctypes.byref(peak), None))
    fp64_peak = peak.value
    rng = np.random.default_rng(0)
    dev = torch.device("cuda:0")

    def T(a):
        return torch.from_numpy(a).to(dev)

    cases = [("cfg2-slice", 4000, 201600, 256, 1), ("few-chan", 4000, 400000, 64, 1),
             ("ncorr4", 2000, 100000, 256, 4)]
    for name, nsrc, nrow, nchan, ncorr in cases:
        lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2)))
        uvw = T(rng.standard_normal((nrow, 3)) * 3000.0)
        freq = T(np.linspace(0.856e9, 1.712e9, nchan))
        image = T(rng.standard_normal((nsrc, nchan, ncorr)))
        terms = nsrc * nrow * nchan
        t = timed(lambda: dft.im_to_vis(image, uvw, lm, freq))
        fl = 7 + 4 * ncorr
        print("%-10s im_to_vis c128 : %.3f s  %.3f Tterm/s  %.1f%% of fp64 FMA peak (%d flop/term)"
              % (name, t, terms / t / 1e12, 100 * terms * fl / t / fp64_peak, fl), flush=True)
        t = timed(lambda: dft.im_to_vis(image, uvw, lm, freq, dtype=np.complex64))
        print("%-10s im_to_vis c64  : %.3f s  %.3f Tterm/s" % (name, t, terms / t / 1e12), flush=True)
        vis = dft.im_to_vis(image, uvw, lm, freq)
        flags = torch.zeros(vis.shape, dtype=torch.bool, device=dev)
        t = timed(lambda: dft.vis_to_im(vis, uvw, lm, freq, flags))
        print("%-10s vis_to_im f64  : %.3f s  %.3f Tterm/s  %.1f%% of fp64 FMA peak"
              % (name, t, terms / t / 1e12, 100 * terms * fl / t / fp64_peak), flush=True)
        t = timed(lambda: dft.vis_to_im(vis, uvw, lm, freq, flags, dtype=np.float32))
        print("%-10s vis_to_im f32  : %.3f s  %.3f Tterm/s" % (name, t, terms / t / 1e12), flush=True)
        del vis, flags

    # fused point predict, config-1 shape
    na, ntime, nchan, nsrc = 64, 100, 64, 100
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    tidx = np.repeat(np.arange(ntime), a1.size)
    nrow = ant1.size
    uvw = T(rng.standard_normal((nrow, 3)) * 3000.0)
    lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2)))
    freq = T(np.linspace(0.856e9, 1.712e9, nchan))
    bright = T(rng.standard_normal((nsrc, nchan, 2, 2)) + 1j * rng.standard_normal((nsrc, nchan, 2, 2)))
    ti, a1t, a2t = T(tidx), T(ant1), T(ant2)
    terms = nsrc * nrow * nchan
    t = timed(lambda: rime.fused_predict_vis(lm, uvw, freq, bright, ti, a1t, a2t))
    print("cfg1 fused point predict c128: %.4f s  %.3f Tterm/s  %.1f%% of fp64 FMA peak (39 flop/term)"
          % (t, terms / t / 1e12, 100 * terms * 39 / t / fp64_peak), flush=True)
    # fused DDE predict, small config-3 slice: 64 ant, 1 time, 1024 chan, 64 src
    ntime, nchan, nsrc = 1, 1024, 64
    ant1, ant2 = a1, a2
    tidx = np.zeros(a1.size, np.int64)
    nrow = a1.size
    uvw = T(rng.standard_normal((nrow, 3)) * 3000.0)
    lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2)))
    freq = T(np.linspace(0.856e9, 1.712e9, nchan))
    bright = T(rng.standard_normal((nsrc, nchan, 2, 2)) + 1j * rng.standard_normal((nsrc, nchan, 2, 2)))
    dde = T(1 + 0.1 * (rng.standard_normal((nsrc, ntime, na, nchan, 2, 2)) + 1j * rng.standard_normal((nsrc, ntime, na, nchan, 2, 2))))
    die = T(1 + 0.1 * (rng.standard_normal((ntime, na, nchan, 2, 2)) + 1j * rng.standard_normal((ntime, na, nchan, 2, 2))))
    ti, a1t, a2t = T(tidx), T(ant1), T(ant2)
    terms = nsrc * nrow * nchan
    t = timed(lambda: rime.fused_predict_vis(lm, uvw, freq, bright, ti, a1t, a2t, dde, dde, die, None, die))
    print("cfg3-slice fused DDE predict c128: %.4f s  %.4f Tterm/s  %.1f%% of fp64 FMA peak (95 flop/term)"
          % (t, terms / t / 1e12, 100 * terms * 95 / t / fp64_peak), flush=True)


if __name__ == "__main__":
    main()
