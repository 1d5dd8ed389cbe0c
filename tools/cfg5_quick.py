"""vis_to_im on the configs[4] geometry (1024^2 pixels x 64 channels, 2 timesteps of rows, 5 % flags)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth
from codex_africanus_b200 import dft, _lib
rng = np.random.default_rng(3); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
uvw, tidx, a1, a2 = synth.uvw_tracks(64, 2, rng, ntime_total=1000)
npix = 1024; cell = 4.0 / 3600.0 * np.pi / 180.0
gl = (np.arange(npix) - npix // 2) * cell
lm5 = T(np.stack(np.meshgrid(gl, gl, indexing="ij"), axis=-1).reshape(-1, 2))
nchan = 64
freq5 = T(synth.frequencies(nchan))
vis5 = torch.randn((uvw.shape[0], nchan, 1), dtype=torch.complex128, device=dev)
flags5 = (torch.rand(vis5.shape, device=dev) < 0.05)
d_uvw = T(uvw)
fn = lambda: dft.vis_to_im(vis5, d_uvw, lm5, freq5, flags5)
fn(); torch.cuda.synchronize(); best = 1e30
for _ in range(3):
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record(); out = fn(); e1.record(); torch.cuda.synchronize(); best = min(best, e0.elapsed_time(e1) * 1e-3)
print("vis_to_im 1024^2 x 64 chan: %.1f Gterms/s  (%s)" % (npix * npix * uvw.shape[0] * nchan / best / 1e9, _lib.describe_dft_path()))
