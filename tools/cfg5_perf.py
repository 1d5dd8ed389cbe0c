import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth
from codex_africanus_b200 import dft, rime
rng = np.random.default_rng(3); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
def timed(fn, reps=2):
    fn(); torch.cuda.synchronize(); best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e-3); del out
    return best
na, ntime = 64, 2
uvw, tidx, a1, a2 = synth.uvw_tracks(na, ntime, rng, ntime_total=1000)
npix = 1024; cell = 4.0 / 3600.0 * np.pi / 180.0
gl = (np.arange(npix) - npix // 2) * cell
lm5 = T(np.stack(np.meshgrid(gl, gl, indexing="ij"), axis=-1).reshape(-1, 2))
for nchan in (128, 64):
    freq5 = T(synth.frequencies(nchan))
    vis5 = torch.randn((uvw.shape[0], nchan, 1), dtype=torch.complex128, device=dev)
    flags5 = (torch.rand(vis5.shape, device=dev) < 0.05)
    for ws in ("0", "1", "default"):
        os.environ["AFR_WS"] = ws
        if ws == "default": os.environ.pop("AFR_WS")
        t = timed(lambda: dft.vis_to_im(vis5, T(uvw), lm5, freq5, flags5))
        print("nchan %d AFR_WS=%s vis_to_im 1024^2: %.1f Gterms/s" % (nchan, ws, npix * npix * uvw.shape[0] * nchan / t / 1e9), flush=True)
    img = torch.randn((4000, nchan, 1), dtype=torch.float64, device=dev)
    lmi = T(synth.sky_lm(4000, rng)); uv = T(rng.standard_normal((201600, 3)) * 3000)
    for ws in ("0", "1", "default"):
        os.environ["AFR_WS"] = ws
        if ws == "default": os.environ.pop("AFR_WS")
        t = timed(lambda: dft.im_to_vis(img, uv, lmi, freq5))
        print("nchan %d AFR_WS=%s im_to_vis: %.1f Gterms/s" % (nchan, ws, 4000 * 201600 * nchan / t / 1e9), flush=True)
os.environ.pop("AFR_WS", None)
# beam kernel
nchan3, nsrc3 = 4096, 96
freq3 = synth.frequencies(nchan3); lm3 = synth.sky_lm(nsrc3, rng)
beam, ext, bfreq = synth.beam_cube(257, 64, rng)
args = (T(beam), ext, bfreq, T(lm3), T(rng.uniform(-0.3, 0.3, (1, na))), T(np.zeros((1, na, nchan3, 2))), T(np.ones((na, nchan3, 2))), T(freq3))
t = timed(lambda: rime.beam_cube_dde(*args))
print("beam_cube_dde: %.3f ms, %.0f GB/s of output" % (t * 1e3, nsrc3 * na * nchan3 * 64 / t / 1e9))
