import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import dft
rng = np.random.default_rng(0); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(a).to(dev)
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best
nsrc, nrow, nchan = 3000, 120000, 256
lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2))); uvw = T(rng.standard_normal((nrow, 3)) * 3000.0)
freq = T(np.linspace(0.856e9, 1.712e9, nchan))
tag = "WS=%s" % os.environ.get("AFR_WS", "default")
for ncorr in (1, 2, 4):
    n = nrow // ncorr
    terms = nsrc * n * nchan
    vis = T(rng.standard_normal((n, nchan, ncorr)) + 1j * rng.standard_normal((n, nchan, ncorr)))
    flags = T(np.repeat(rng.random((n, nchan, 1)) < 0.05, ncorr, axis=2))
    t = timed(lambda: dft.vis_to_im(vis, uvw[:n], lm, freq, flags)); print(tag, "adj cplx flagged ncorr%d: %.1f Gterm/s" % (ncorr, terms / t / 1e9))
