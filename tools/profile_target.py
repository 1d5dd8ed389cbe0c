#!/usr/bin/env python
"""Small device-resident workload for ncu captures: one launch of each main kernel."""
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import dft, rime  # noqa: E402

which = sys.argv[1] if len(sys.argv) > 1 else "all"
rng = np.random.default_rng(0)
dev = torch.device("cuda:0")


def T(a):
    return torch.from_numpy(a).to(dev)


nsrc, nrow, nchan = 1024, 148 * 32 * 8, 256
lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2)))
uvw = T(rng.standard_normal((nrow, 3)) * 3000.0)
freq = T(np.linspace(0.856e9, 1.712e9, nchan))
image = T(rng.standard_normal((nsrc, nchan, 1)))
if which in ("all", "i2v"):
    vis = dft.im_to_vis(image, uvw, lm, freq)
    vis = dft.im_to_vis(image, uvw, lm, freq)
if which in ("all", "v2i"):
    nsrc2 = 148 * 32 * 2
    lm2 = T(rng.uniform(-0.02, 0.02, (nsrc2, 2)))
    nrow2 = 4096
    uvw2 = T(rng.standard_normal((nrow2, 3)) * 3000.0)
    vis2 = T(rng.standard_normal((nrow2, nchan, 1)) + 1j * rng.standard_normal((nrow2, nchan, 1)))
    flags = torch.zeros(vis2.shape, dtype=torch.bool, device=dev)
    im = dft.vis_to_im(vis2, uvw2, lm2, freq, flags)
    im = dft.vis_to_im(vis2, uvw2, lm2, freq, flags)
if which in ("all", "fused"):
    na, ntime, nchan3, nsrc3 = 64, 1, 1024, 64
    a1, a2 = np.triu_indices(na, 1)
    tidx = np.zeros(a1.size, np.int64)
    uvw3 = T(rng.standard_normal((a1.size, 3)) * 3000.0)
    lm3 = T(rng.uniform(-0.02, 0.02, (nsrc3, 2)))
    freq3 = T(np.linspace(0.856e9, 1.712e9, nchan3))
    shp = (nsrc3, nchan3, 2, 2)
    bright = T(rng.standard_normal(shp) + 1j * rng.standard_normal(shp))
    shp = (nsrc3, ntime, na, nchan3, 2, 2)
    dde = T(1 + 0.1 * (rng.standard_normal(shp) + 1j * rng.standard_normal(shp)))
    out = rime.fused_predict_vis(lm3, uvw3, freq3, bright, T(tidx), T(a1), T(a2), dde, dde)
    out = rime.fused_predict_vis(lm3, uvw3, freq3, bright, T(tidx), T(a1), T(a2), dde, dde)
torch.cuda.synchronize()
print("done")
