"""Accuracy of the FP32 (complex64 / float32) phasor-stream variants against the oracle's complex64
path and against the FP64 kernel: relative L2 error (the project gate is 1e-5), incl. the regime the
FP32 three-term sub-runs are most sensitive to (tiny step angles: short baselines, sources near
the phase centre)."""
import os, sys, numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
import oracle
from codex_africanus_b200 import dft
oracle.build()
rng = np.random.default_rng(5)
def rel(a, b): return float(np.linalg.norm((a.astype(np.complex128) - b).ravel()) / np.linalg.norm(b.ravel()))
for name, uvw_scale, field, nchan, ncorr in (("MeerKAT-like", 3000.0, 0.02, 256, 1), ("tiny angles", 30.0, 0.002, 256, 1),
                                             ("mixed", 800.0, 0.01, 96, 2), ("long baselines", 60000.0, 0.05, 128, 4),
                                             ("zero baselines", 0.0, 0.02, 64, 1)):
    nsrc, nrow = 300, 500
    lm = rng.uniform(-field, field, (nsrc, 2)); uvw = rng.standard_normal((nrow, 3)) * uvw_scale
    freq = np.linspace(0.856e9, 1.712e9, nchan); image = rng.standard_normal((nsrc, nchan, ncorr))
    ref = oracle.im_to_vis(image, uvw, lm, freq)
    got = dft.im_to_vis(image, uvw, lm, freq, dtype=np.complex64)
    vis = ref + 0.1 * (rng.standard_normal(ref.shape) + 1j * rng.standard_normal(ref.shape))
    flags = rng.random(vis.shape) < 0.05
    refi = oracle.vis_to_im(vis, uvw, lm, freq, flags)
    goti = dft.vis_to_im(vis, uvw, lm, freq, flags, dtype=np.float32)
    print("%-15s nchan %3d ncorr %d: im_to_vis c64 rel L2 %.2e   vis_to_im f32 rel L2 %.2e" % (
        name, nchan, ncorr, rel(got, ref), rel(goti, refi)), flush=True)
