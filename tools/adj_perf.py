import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import dft
rng = np.random.default_rng(0); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(a).to(dev)
nsrc, nrow, nchan = 10000, 100800, 256
lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2))); uvw = T(rng.standard_normal((nrow, 3)) * 3000.0)
freq = T(np.linspace(0.856e9, 1.712e9, nchan))
vis = T(rng.standard_normal((nrow, nchan, 1)) + 1j * rng.standard_normal((nrow, nchan, 1)))
flags = torch.rand(vis.shape, device=dev) < 0.05
noflags = torch.zeros(vis.shape, dtype=torch.bool, device=dev)
def timed(fn, reps=4):
    fn(); torch.cuda.synchronize(); ts = []
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); fn(); e1.record(); torch.cuda.synchronize(); ts.append(e0.elapsed_time(e1) * 1e-3)
    return min(ts)
for var in sys.argv[1:] or [""]:
    for k in ("AFR_WS", "AFR_ADJ16"): os.environ.pop(k, None)
    for kv in var.split(","):
        if "=" in kv:
            k, v = kv.split("="); os.environ[k] = v
    t = timed(lambda: dft.vis_to_im(vis, uvw, lm, freq, flags))
    print("%-28s vis_to_im f64 5%% flags: %.4f s (%.3f Tterm/s)" % (var, t, nsrc * nrow * nchan / t / 1e12), flush=True)
    t = timed(lambda: dft.vis_to_im(vis, uvw, lm, freq, noflags))
    print("%-28s vis_to_im f64 no flags: %.4f s (%.3f Tterm/s)" % (var, t, nsrc * nrow * nchan / t / 1e12), flush=True)
