#!/usr/bin/env python
"""beam_cube_dde on the configs[2] geometry (257x257x64 cube, 64 antennas, 4096 channels, 96 sources):
best of 10 device-timed calls, TB/s of output, and the worst deviation from the element kernel."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth
from codex_africanus_b200 import rime
rng = np.random.default_rng(3); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
na, nchan, nsrc = 64, 4096, 96
freq = synth.frequencies(nchan); lm = synth.sky_lm(nsrc, rng)
beam, ext, bfreq = synth.beam_cube(257, 64, rng)
pa = rng.uniform(-0.3, 0.3, (1, na)); perr = np.zeros((1, na, nchan, 2)); ascale = np.ones((na, nchan, 2))
args = (T(beam), ext, bfreq, T(lm), T(pa), T(perr), T(ascale), T(freq))
for _ in range(3):
    out = rime.beam_cube_dde(*args)
torch.cuda.synchronize()
best = 1e9
for _ in range(10):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = rime.beam_cube_dde(*args); e1.record(); torch.cuda.synchronize()
    best = min(best, e0.elapsed_time(e1))
print("beam_cube_dde: %.3f ms, %.2f TB/s of output" % (best, out.numel() * 16 / best / 1e9))
os.environ["AFR_BEAM_PLANES"] = "0"
ref = rime.beam_cube_dde(*args)
print("planes kernel vs element kernel: max |diff| / max |ref| = %.2e" % ((out - ref).abs().max() / ref.abs().max()).item())
