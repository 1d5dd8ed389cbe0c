"""End-to-end timing of dft.im_to_vis with host buffers on the configs[1] shape (or BENCH_NTIME)."""
import os, sys, time, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth
from codex_africanus_b200 import dft
rng = np.random.default_rng(2)
ntime = int(os.environ.get("BENCH_NTIME", 1000))
uvw, tidx, a1, a2 = synth.uvw_tracks(64, ntime, rng)
lm = synth.sky_lm(10000, rng); freq = synth.frequencies(256); image = synth.stokes_image(10000, 256, 1, rng, freq)
dft.im_to_vis(image, uvw, lm, freq)
for _ in range(3):
    t0 = time.perf_counter(); out = dft.im_to_vis(image, uvw, lm, freq); t = time.perf_counter() - t0
    print("e2e %.1f ms  %.0f Gterms/s" % (t * 1e3, 1e4 * uvw.shape[0] * 256 / t / 1e9), flush=True); del out
