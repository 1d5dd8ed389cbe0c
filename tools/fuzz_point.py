"""Randomised parity sweep of the fused POINT predict (2x2 complex brightness: DMMA consumers of the
phasor-stream kernel; diagonal / scalar brightness and non-equispaced channels: scalar consumers) and
of the DDE layout adapter (rime/fused.py::_predict_as_2x2_c128) against the CPU oracle.
usage: fuzz_point.py [ncases] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from codex_africanus_b200 import rime, _lib
from codex_africanus_b200.rime import fused as fused_mod
from conftest import assert_c128_close, rel_l2

n = int(sys.argv[1]) if len(sys.argv) > 1 else 60
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
rc = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
bad = 0; mma = 0; adapted = 0
for case in range(n):
    nrow = int(rng.choice([1, 5, 15, 16, 17, 33, 100, 257, 700]))
    nsrc = int(rng.choice([1, 2, 3, 7, 8, 9, 16, 31, 100]))
    nchan = int(rng.choice([1, 2, 7, 8, 9, 16, 63, 64, 65, 127, 128, 129, 200, 300]))
    scale = float(rng.choice([30.0, 3e3, 1.5e5]))
    corr = [(2, 2), (2, 2), (2, 2), (2,), (1,)][int(rng.integers(0, 5))]
    uvw = rng.standard_normal((nrow, 3)) * scale
    lm = rng.uniform(-0.03, 0.03, (nsrc, 2))
    freq = np.linspace(0.856e9, 1.712e9, nchan) if nchan > 1 else np.array([1.0e9])
    if nchan > 2 and rng.random() < 0.15: freq = np.sort(rng.uniform(0.856e9, 1.712e9, nchan))
    conv = "casa" if rng.random() < 0.3 else "fourier"
    ntime, na = int(rng.choice([1, 3])), int(rng.choice([2, 5]))
    ti = np.sort(rng.integers(0, ntime, nrow)); a1 = rng.integers(0, na, nrow); a2 = rng.integers(0, na, nrow)
    bright = rc((nsrc, nchan) + corr)
    die = None if rng.random() < 0.5 else 1 + 0.1 * rc((ntime, na, nchan) + corr)
    bvis = None if rng.random() < 0.5 else rc((nrow, nchan) + corr)
    tag = "case %d: nrow %d nsrc %d nchan %d scale %g corr %s %s" % (case, nrow, nsrc, nchan, scale, corr, conv)
    try:
        got = rime.fused_predict_vis(lm, uvw, freq, bright, ti, a1, a2, None, None, die, bvis, die, convention=conv)
        mma += bool(_lib.lib().afr_last_dft_path() & 32)
        assert_c128_close(got, oracle.fused_predict(lm, uvw, freq, bright, ti, a1, a2, None, None, die, bvis, die,
                                                    convention=conv))
        if rng.random() < 0.4 and nrow * nsrc * nchan < 400000:  # the adapter on the same shapes, with DDEs
            dde = 1 + 0.2 * rc((nsrc, ntime, na, nchan) + corr)
            dde_b = dde if rng.random() < 0.5 else 1 + 0.2 * rc((nsrc, ntime, na, nchan) + corr)
            perm = rng.permutation(nrow) if rng.random() < 0.5 else np.arange(nrow)
            ref = oracle.fused_predict(lm, uvw, freq, bright, ti, a1, a2, dde, dde_b, die, bvis, die, convention=conv)
            fused_mod._ADAPTER_MIN_TERMS, fused_mod._ADAPTER_CHUNK_BYTES = 0, 1 << 16
            got = rime.fused_predict_vis(lm, uvw[perm], freq, bright, ti[perm], a1[perm], a2[perm], dde, dde_b, die,
                                         None if bvis is None else bvis[perm], die, convention=conv)
            fused_mod._ADAPTER_MIN_TERMS = 1 << 24
            adapted += 1
            assert_c128_close(got, ref[perm])
    except AssertionError as e:
        bad += 1
        print("FAIL", tag, str(e).split("\n")[0][:200], flush=True)
print("fuzz_point: %d cases (%d on DMMA consumers, %d through the layout adapter), %d failures" % (n, mma, adapted, bad))
sys.exit(1 if bad else 0)
