import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth
from codex_africanus_b200 import rime
rng = np.random.default_rng(3); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
na, ntime, nchan, nsrc = 64, int(sys.argv[1]) if len(sys.argv) > 1 else 2, 4096, int(sys.argv[2]) if len(sys.argv) > 2 else 400
uvw, tidx, a1, a2 = synth.uvw_tracks(na, ntime, rng, ntime_total=100)
freq = synth.frequencies(nchan); lm = synth.sky_lm(nsrc, rng)
bright = T(synth.brightness_2x2(nsrc, nchan, rng, freq)); die = T(synth.gains(ntime, na, nchan, rng))
g = torch.Generator(device=dev).manual_seed(1)
dde = torch.randn((nsrc, ntime, na, nchan, 2, 2), dtype=torch.complex128, device=dev, generator=g) * 0.1
dde[..., 0, 0] += 1; dde[..., 1, 1] += 1
d_uvw, d_lm, d_f, d_t, d_a1, d_a2 = T(uvw), T(lm), T(freq), T(tidx), T(a1), T(a2)
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)
terms = float(nsrc) * uvw.shape[0] * nchan
t = timed(lambda: rime.fused_predict_vis(d_lm, d_uvw, d_f, bright, d_t, d_a1, d_a2, dde, dde, die, None, die))
from codex_africanus_b200 import _lib
print("path", _lib.lib().afr_last_fused_path(), end=" ")
print("fused DDE predict: %d src x %d rows x %d chan, DDE %.1f GB: %.3f s  %.1f Gterms/s  DDE read-once %.0f GB/s" % (nsrc, uvw.shape[0], nchan, dde.numel() * 16 / 1e9, t, terms / t / 1e9, dde.numel() * 16 / t / 1e9))
