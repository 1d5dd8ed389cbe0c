"""ncu target: the FP32 (complex64 / float32) phasor-stream variants on a 256-channel problem."""
import os, sys, numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import dft
rng = np.random.default_rng(0); dev = torch.device("cuda:0")
T = lambda a: torch.from_numpy(a).to(dev)
nsrc, nrow, nchan = 2048, 148 * 32 * 8, 256
lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2))); uvw = T(rng.standard_normal((nrow, 3)) * 3000.0)
freq = T(np.linspace(0.856e9, 1.712e9, nchan)); image = T(rng.standard_normal((nsrc, nchan, 1)))
for _ in range(2):
    vis = dft.im_to_vis(image, uvw, lm, freq, dtype=np.complex64)
v128 = dft.im_to_vis(image, uvw[:4736], lm, freq)
flags = torch.zeros(v128.shape, dtype=torch.bool, device=dev)
for _ in range(2):
    img = dft.vis_to_im(v128, uvw[:4736], lm, freq, flags, dtype=np.float32)
torch.cuda.synchronize(); print("done")
