#!/usr/bin/env python
"""Fused point predict, ncorr = 4 complex brightness (the configs[0] / configs[3] kernel):
one SKA-Mid timestep (197 antennas, 19306 rows) x 1024 channels x nsrc sources, device resident.
With no argument it prints Gterms/s (best of 3); with `once` it launches twice (ncu target)."""
import os
import sys
import time

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import rime  # noqa: E402

once = len(sys.argv) > 1 and sys.argv[1] == "once"
nsrc = int(os.environ.get("NSRC", "256" if once else "2000"))
nchan = int(os.environ.get("NCHAN", "1024"))
rng = np.random.default_rng(0)
dev = torch.device("cuda:0")
na = 197
a1, a2 = np.triu_indices(na, 1)
pos = rng.standard_normal((na, 3)) * 30e3
uvw = torch.from_numpy(pos[a1] - pos[a2]).to(dev)
lm = torch.from_numpy(rng.uniform(-0.02, 0.02, (nsrc, 2))).to(dev)
freq = torch.from_numpy(np.linspace(0.856e9, 1.712e9, nchan)).to(dev)
shp = (nsrc, nchan, 2, 2)
bright = torch.from_numpy(rng.standard_normal(shp) + 1j * rng.standard_normal(shp)).to(dev)
tidx = torch.zeros(a1.size, dtype=torch.int32, device=dev)
ta1 = torch.from_numpy(a1.astype(np.int32)).to(dev)
ta2 = torch.from_numpy(a2.astype(np.int32)).to(dev)


def call():
    return rime.fused_predict_vis(lm, uvw, freq, bright, tidx, ta1, ta2)


call()
torch.cuda.synchronize()
if once:
    call()
    torch.cuda.synchronize()
    print("done")
else:
    best = 1e9
    for _ in range(3):
        t0 = time.perf_counter()
        call()
        torch.cuda.synchronize()
        best = min(best, time.perf_counter() - t0)
    terms = a1.size * nchan * nsrc
    print("fused point predict ncorr=4: %.1f Gterms/s (%.2f ms)" % (terms / best / 1e9, best * 1e3))
