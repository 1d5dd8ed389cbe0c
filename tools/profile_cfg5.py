"""ncu target: vis_to_im on the configs[4] geometry (1024^2 pixels x 64 channels, 2 timesteps of rows)."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth
from codex_africanus_b200 import dft
rng = np.random.default_rng(3); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
uvw, tidx, a1, a2 = synth.uvw_tracks(64, 1, rng, ntime_total=1000)
npix = 1024; cell = 4.0 / 3600.0 * np.pi / 180.0
gl = (np.arange(npix) - npix // 2) * cell
lm5 = T(np.stack(np.meshgrid(gl, gl, indexing="ij"), axis=-1).reshape(-1, 2))
nchan = 64
freq5 = T(synth.frequencies(nchan))
vis5 = torch.randn((uvw.shape[0], nchan, 1), dtype=torch.complex128, device=dev)
flags5 = (torch.rand(vis5.shape, device=dev) < 0.05)
d_uvw = T(uvw)
for _ in range(2):
    out = dft.vis_to_im(vis5, d_uvw, lm5, freq5, flags5)
torch.cuda.synchronize()
print("done")
