import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth
from codex_africanus_b200 import rime
rng = np.random.default_rng(3); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
na, ntime, nchan, nsrc = 64, 1, int(sys.argv[1]) if len(sys.argv) > 1 else 1024, int(sys.argv[2]) if len(sys.argv) > 2 else 64
uvw, tidx, a1, a2 = synth.uvw_tracks(na, ntime, rng, ntime_total=100)
freq = synth.frequencies(nchan); lm = synth.sky_lm(nsrc, rng)
bright = T(synth.brightness_2x2(nsrc, nchan, rng, freq))
g = torch.Generator(device=dev).manual_seed(1)
dde = torch.randn((nsrc, ntime, na, nchan, 2, 2), dtype=torch.complex128, device=dev, generator=g) * 0.1
for _ in range(2):
    out = rime.fused_predict_vis(T(lm), T(uvw), T(freq), bright, T(tidx), T(a1), T(a2), dde, dde)
torch.cuda.synchronize(); print("done")
