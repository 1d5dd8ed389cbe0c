import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth
from codex_africanus_b200 import rime
rng = np.random.default_rng(3); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e-3); del out
    return best
uvw, tidx, a1, a2 = synth.uvw_tracks(64, 50, rng, ntime_total=1000)
nsw, nchan = 1000, 256
arcsec = np.pi / 180.0 / 3600.0
gshape = np.stack([rng.uniform(10, 90, nsw) * arcsec, rng.uniform(3, 10, nsw) * arcsec, rng.uniform(0, np.pi, nsw)], axis=1)
args = (T(uvw), T(synth.sky_lm(nsw, rng)), None, T(np.abs(rng.standard_normal(nsw)) + 0.1), T(rng.standard_normal((nsw, 2)) * 0.2),
        rng.random(nsw) < 0.5, T(np.full(nsw, 1.284e9)), T(gshape), T(synth.frequencies(nchan)))
for kind in ("GAUSSIAN", "POINT"):
    a = list(args); a[2] = np.full(nsw, kind)
    t = timed(lambda: rime.wsclean_predict(*a))
    print("%s only: %.1f Gterms/s" % (kind, nsw * uvw.shape[0] * nchan / t / 1e9))
