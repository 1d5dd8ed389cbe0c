#!/usr/bin/env python
"""configs[2] full RIME (64 antennas x 4 timesteps x 4096 channels, beam 257x257x64, DIE + DDE): the beam
sampled per source chunk (GEMM kernel, default) against the beam sampled INSIDE the predict kernel from
the plane-reduced beam (SURVEY 8f-1 proper).  Device-resident, best of 3; peak device memory of each."""
import os, sys
import numpy as np
import torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
import synth
from codex_africanus_b200 import rime, _lib
rng = np.random.default_rng(5); dev = torch.device("cuda:0")
na, ntime, nchan, nsrc = 64, 4, 4096, int(os.environ.get("NSRC", "1000"))
uvw, ti, a1, a2 = synth.uvw_tracks(na, ntime, rng, ntime_total=1000)
T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).to(dev)
freq = synth.frequencies(nchan); lm = synth.sky_lm(nsrc, rng)
beam, ext, bfreq = synth.beam_cube(257, 64, rng)
rc = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
pa = rng.uniform(-0.3, 0.3, (ntime, na)); perr = np.zeros((ntime, na, nchan, 2)); ascale = np.ones((na, nchan, 2))
bright = rc((nsrc, nchan, 2, 2)); die = 1 + 0.1 * rc((ntime, na, nchan, 2, 2))
args = [T(x) for x in (lm, uvw, freq, bright, ti.astype(np.int32), a1.astype(np.int32), a2.astype(np.int32), beam)] + \
       [ext, bfreq] + [T(x) for x in (pa, perr, ascale, die)] + [None, None]
args[-1] = args[-3]
terms = float(ti.size) * nchan * nsrc
outs = {}
for name, kw in (("beam sampled per source chunk (GEMM kernel)", {}), ("beam sampled inside the predict kernel", {"in_kernel": True})):
    out = rime.fused_predict_vis_beam(*args, **kw); torch.cuda.synchronize()
    torch.cuda.reset_peak_memory_stats(); best = 1e9
    for _ in range(3):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record(); out = rime.fused_predict_vis_beam(*args, **kw); e1.record(); torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1))
    outs[name] = out
    print("%-46s path %d: %.1f ms, %.1f Gterms/s, torch peak %.1f GB" % (name, _lib.lib().afr_last_fused_path(), best,
          terms / best / 1e6, torch.cuda.max_memory_allocated() / 1e9))
    del out
a, b = outs.values()
print("max |diff| / max |ref| between the two: %.2e" % ((a - b).abs().max() / a.abs().max()).item())
