import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import dft
rng = np.random.default_rng(0); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(a).to(dev)
nsrc, nrow, nchan = 4000, 201600, 256
lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2))); uvw = T(rng.standard_normal((nrow, 3)) * 3000.0)
freq = T(np.linspace(0.856e9, 1.712e9, nchan)); image = T(rng.standard_normal((nsrc, nchan, 1)))
def timed(fn, reps=3):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e-3)
    return best
t = timed(lambda: dft.im_to_vis(image, uvw, lm, freq))
print("AFR_DEBUG=%s im_to_vis c128: %.4f s %.3f Tterm/s" % (os.environ.get("AFR_DEBUG", "0"), t, nsrc * nrow * nchan / t / 1e12))
