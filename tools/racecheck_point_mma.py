"""Small fused point predicts on the DMMA-consumer schedule of the phasor-stream kernel (mbarrier
pipeline, TMA bulk-copied W tile, padded anchor pitches) for
`compute-sanitizer --tool memcheck` and `AFR_SANITIZE=1 compute-sanitizer --tool racecheck`."""
import os, sys
import numpy as np
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from codex_africanus_b200 import rime, _lib
rng = np.random.default_rng(2)
rc = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
for nrow, nsrc, nchan in ((37, 11, 6), (70, 19, 136), (20, 3, 300)):
    uvw = rng.standard_normal((nrow, 3)) * 1500.0
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    freq = np.linspace(1e9, 1.1e9, nchan)
    z = np.zeros(nrow, int)
    rime.fused_predict_vis(lm, uvw, freq, rc((nsrc, nchan, 2, 2)), z, z, z + 1)
    assert _lib.lib().afr_last_dft_path() & 32
print("point-mma target done")
