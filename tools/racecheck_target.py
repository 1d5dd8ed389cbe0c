"""Small calls of the round-2 kernels that synchronise through shared memory (antenna-mode GEMM
kernel: mbarrier pipeline + cp.async; plane-interpolated beam kernel: __syncthreads) for
`AFR_SANITIZE=1 compute-sanitizer --tool racecheck`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import rime, _lib
rng = np.random.default_rng(1)
rc = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
for na, nsrc, nchan, same in ((9, 11, 2, True), (70, 6, 1, False)):
    a1, a2 = np.triu_indices(na, 1)
    pos = rng.standard_normal((1, na, 3)) * 1500.0
    uvw = pos[0, a1] - pos[0, a2]
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2)); freq = np.linspace(1e9, 1.1e9, nchan) if nchan > 1 else np.array([1e9])
    d1 = rc((nsrc, 1, na, nchan, 2, 2)); d2 = d1 if same else rc((nsrc, 1, na, nchan, 2, 2))
    rime.fused_predict_vis(lm, uvw, freq, rc((nsrc, nchan, 2, 2)), np.zeros(a1.size, int), a1, a2, d1, d2)
    assert _lib.lib().afr_last_fused_path() == 6
beam = rc((9, 9, 5, 2, 2)); f80 = np.linspace(0.75e9, 1.85e9, 80)
rime.beam_cube_dde(beam, np.array([[-0.03, 0.03], [-0.03, 0.03]]), np.linspace(0.8e9, 1.8e9, 5),
                   rng.uniform(-0.02, 0.02, (5, 2)), rng.uniform(-1, 1, (2, 3)), np.zeros((2, 3, 80, 2)),
                   np.ones((3, 80, 2)), f80)
print("racecheck target done")
