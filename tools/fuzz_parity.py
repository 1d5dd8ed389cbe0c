"""Randomised parity sweep of the phasor-stream entry points against the CPU oracle
(shapes around the tiling edges: channel tails of TMA-staged tiles, row tails, y-splits,
flag groups, ncorr blocks).  usage: fuzz_parity.py [ncases] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from codex_africanus_b200 import dft, rime
from conftest import assert_c128_close

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
chans = [1, 2, 15, 16, 17, 31, 32, 48, 64, 80, 96, 128, 144, 255, 256, 257, 272, 300, 512, 528]
bad = 0
for case in range(n):
    nchan = int(rng.choice(chans)); ncorr = int(rng.choice([1, 1, 2, 3, 4]))
    nsrc = int(rng.choice([1, 3, 8, 9, 17, 40, 65])); nrow = int(rng.choice([1, 5, 31, 32, 33, 100, 257]))
    if nsrc * nrow * nchan * ncorr > 6e6:
        nrow = max(1, int(6e6 / (nsrc * nchan * ncorr)))
    lm = rng.uniform(-0.03, 0.03, (nsrc, 2)); uvw = rng.standard_normal((nrow, 3)) * 3000.0
    freq = np.linspace(0.856e9, 1.712e9, nchan) if nchan > 1 else np.array([1.2e9])
    if rng.random() < 0.2 and nchan > 2:
        freq = np.sort(rng.uniform(0.856e9, 1.712e9, nchan))
    cplx = rng.random() < 0.5
    image = rng.standard_normal((nsrc, nchan, ncorr)) + (1j * rng.standard_normal((nsrc, nchan, ncorr)) if cplx else 0)
    if not cplx: image = image.real
    vis = rng.standard_normal((nrow, nchan, ncorr)) + 1j * rng.standard_normal((nrow, nchan, ncorr))
    flags = rng.random((nrow, nchan, ncorr)) < rng.choice([0.0, 0.05, 0.5])
    tag = "case %d: nsrc %d nrow %d nchan %d ncorr %d cplx %d uniform %d" % (
        case, nsrc, nrow, nchan, ncorr, cplx, int(np.allclose(np.diff(freq), np.diff(freq)[0]) if nchan > 2 else 1))
    try:
        assert_c128_close(dft.im_to_vis(image, uvw, lm, freq), oracle.im_to_vis(image, uvw, lm, freq))
        assert_c128_close(dft.vis_to_im(vis, uvw, lm, freq, flags), oracle.vis_to_im(vis, uvw, lm, freq, flags))
        # owners = sources: many sources, few rows -> y-split path of the adjoint
        assert_c128_close(dft.vis_to_im(vis.real.copy(), uvw, lm, freq, flags),
                          oracle.vis_to_im(vis.real.copy(), uvw, lm, freq, flags))
    except AssertionError as e:
        bad += 1
        print("FAIL", tag, str(e).split("\n")[0][:200], flush=True)
print("fuzz: %d cases, %d failures" % (n, bad))
sys.exit(1 if bad else 0)
