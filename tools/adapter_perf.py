#!/usr/bin/env python
"""DDE predict in layouts without a fast kernel of their own (diagonal Jones, complex64, unordered
rows): the re-expression as the complex128 2x2 problem (rime/fused.py) against the gather kernel."""
import os, sys, time
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import rime, _lib
from codex_africanus_b200.rime import fused as fused_mod
rng = np.random.default_rng(0)
dev = torch.device("cuda:0")
na, ntime, nsrc, nchan = 64, 2, 96, 1024
a1, a2 = np.triu_indices(na, 1)
ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
ti = np.repeat(np.arange(ntime), a1.size)
pos = rng.standard_normal((ntime, na, 3)) * 1500.0
uvw = (pos[:, a1] - pos[:, a2]).reshape(-1, 3)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
rc = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2))); freq = T(np.linspace(0.856e9, 1.712e9, nchan))
d_uvw, d_ti, d_a1, d_a2 = T(uvw), T(ti.astype(np.int32)), T(ant1.astype(np.int32)), T(ant2.astype(np.int32))
terms = ant1.size * nchan * nsrc
for name, corr, cdt in (("diagonal (2,) c128", (2,), np.complex128), ("2x2 c64", (2, 2), np.complex64)):
    b = T(rc((nsrc, nchan) + corr).astype(cdt)); dde = T((1 + 0.2 * rc((nsrc, ntime, na, nchan) + corr)).astype(cdt))
    for thr, tag in ((1 << 62, "own kernel"), (0, "as 2x2 c128")):
        fused_mod._ADAPTER_MIN_TERMS = thr
        kw = dict(dtype=cdt)
        out = rime.fused_predict_vis(lm, d_uvw, freq, b, d_ti, d_a1, d_a2, dde, dde, **kw)
        torch.cuda.synchronize()
        best = 1e9
        for _ in range(3):
            t0 = time.perf_counter()
            out = rime.fused_predict_vis(lm, d_uvw, freq, b, d_ti, d_a1, d_a2, dde, dde, **kw)
            torch.cuda.synchronize()
            best = min(best, time.perf_counter() - t0)
        print("%-20s %-12s path %d: %.1f Gterms/s" % (name, tag, _lib.lib().afr_last_fused_path(), terms / best / 1e9))
