"""ncu target: one beam_cube_dde call on the configs[2] geometry (257x257x64 cube, 64 antennas,
4096 channels, 96 sources)."""
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
for _ in range(2):
    out = rime.beam_cube_dde(*args)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record(); out = rime.beam_cube_dde(*args); e1.record(); torch.cuda.synchronize()
print("beam_cube_dde: %.3f ms, %.1f GB/s of output" % (e0.elapsed_time(e1), out.numel() * 16 / e0.elapsed_time(e1) / 1e6))
