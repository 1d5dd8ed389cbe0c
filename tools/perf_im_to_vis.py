import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import dft
rng = np.random.default_rng(0); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(a).to(dev)
nsrc, nrow, nchan = 4000, 201600, 256
lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2))); uvw = T(rng.standard_normal((nrow, 3)) * 3000.0)
freq = T(np.linspace(0.856e9, 1.712e9, nchan)); image = T(rng.standard_normal((nsrc, nchan, 1)))
def timed(fn, reps=6):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts), sorted(ts)[len(ts)//2]
for var in sys.argv[1:] or [""]:
    for kv in var.split(","):
        if "=" in kv:
            k, v = kv.split("="); os.environ[k] = v
    tmin, tmed = timed(lambda: dft.im_to_vis(image, uvw, lm, freq))
    print("%-24s im_to_vis c128: min %.4f s (%.3f Tterm/s)  median %.4f s" % (var, tmin, nsrc * nrow * nchan / tmin / 1e12, tmed), flush=True)
