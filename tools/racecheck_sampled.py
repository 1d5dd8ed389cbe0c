"""Small fused_predict_vis_beam(in_kernel=True) calls (beam sampled inside the warp-specialised DDE kernel:
cp.async of the frequency planes + mbarrier pipeline) for compute-sanitizer memcheck / racecheck (AFR_SANITIZE=1)."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import rime, _lib
rng = np.random.default_rng(3)
rc = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
for na, ntime, nsrc, nchan in ((9, 2, 7, 5), (37, 1, 4, 2)):
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size)
    pos = rng.standard_normal((ntime, na, 3)) * 1500.0
    uvw = pos[ti, ant1] - pos[ti, ant2]
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2)); freq = np.linspace(0.9e9, 1.7e9, nchan)
    beam = rc((9, 9, 5, 2, 2)); ext = np.array([[-0.03, 0.03], [-0.03, 0.03]]); bfm = np.linspace(0.8e9, 1.8e9, 5)
    pa = rng.uniform(-1, 1, (ntime, na)); pe = np.zeros((ntime, na, nchan, 2)); asc = np.ones((na, nchan, 2))
    rime.fused_predict_vis_beam(lm, uvw, freq, rc((nsrc, nchan, 2, 2)), ti, ant1, ant2, beam, ext, bfm, pa, pe, asc,
                                in_kernel=True)
    assert _lib.lib().afr_last_fused_path() == 7
print("sampled target done")
