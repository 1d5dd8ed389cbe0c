"""Randomised parity sweep of fused_predict_vis (DDE kernels: antenna-mode DMMA GEMM incl. multi-pass
panels, swapped antenna pairs and duplicate baselines, scalar antenna-phasor / per-row / tiled /
gather paths) against the CPU oracle.  usage: fuzz_fused.py [ncases] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from codex_africanus_b200 import rime, _lib
from conftest import assert_c128_close

n = int(sys.argv[1]) if len(sys.argv) > 1 else 30
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
bad = 0; paths = {}
for case in range(n):
    na = int(rng.choice([2, 3, 7, 12, 33, 40, 64, 70, 130], p=[.14, .14, .14, .14, .14, .12, .06, .06, .06]))
    ntime = int(rng.choice([1, 2, 5])) if na < 64 else int(rng.choice([1, 2]))
    nchan = int(rng.choice([1, 3, 4, 5, 16, 37])); nsrc = int(rng.choice([1, 2, 7, 19]))
    a1, a2 = np.triu_indices(na, 1)
    keep = [np.sort(rng.choice(a1.size, size=max(1, a1.size - int(rng.integers(0, 3))), replace=False)) for _ in range(ntime)]
    ant1 = np.concatenate([a1[k] for k in keep]); ant2 = np.concatenate([a2[k] for k in keep])
    ti = np.concatenate([np.full(k.size, t) for t, k in enumerate(keep)])
    if rng.random() < 0.3:  # swap some antenna pairs (a1 > a2 rows)
        sw = rng.random(ant1.size) < 0.3
        ant1, ant2 = np.where(sw, ant2, ant1), np.where(sw, ant1, ant2)
    if rng.random() < 0.1 and ant1.size > 3:  # a baseline listed twice in one timestep: no (t, a1, a2) -> row map
        dup = rng.integers(0, ant1.size, 2)
        order = np.sort(np.concatenate([np.arange(ant1.size), dup]))
        ant1, ant2, ti = ant1[order], ant2[order], ti[order]
    antpos = rng.standard_normal((ntime, na, 3)) * 1500.0
    consistent = rng.random() < 0.6
    uvw = antpos[ti, ant1] - antpos[ti, ant2] if consistent else rng.standard_normal((ti.size, 3)) * 1500.0
    if rng.random() < 0.15:  # rows not ordered by time -> gather kernel
        perm = rng.permutation(ti.size); ant1, ant2, ti, uvw = ant1[perm], ant2[perm], ti[perm], uvw[perm]
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    freq = np.linspace(0.856e9, 1.712e9, nchan) if nchan > 1 else np.array([1.0e9])
    if nchan > 2 and rng.random() < 0.25: freq = np.sort(rng.uniform(0.856e9, 1.712e9, nchan))
    rc = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)
    bright = rc((nsrc, nchan, 2, 2)); dde = 1 + 0.2 * rc((nsrc, ntime, na, nchan, 2, 2))
    dde_b = dde if rng.random() < 0.5 else 1 + 0.2 * rc((nsrc, ntime, na, nchan, 2, 2))
    die = None if rng.random() < 0.5 else 1 + 0.1 * rc((ntime, na, nchan, 2, 2))
    bvis = None if rng.random() < 0.5 else rc((ti.size, nchan, 2, 2))
    tag = "case %d: na %d ntime %d nchan %d nsrc %d consistent %d" % (case, na, ntime, nchan, nsrc, consistent)
    try:
        got = rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde_b, die, bvis, die)
        path = _lib.lib().afr_last_fused_path(); paths[path] = paths.get(path, 0) + 1
        assert_c128_close(got, oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde_b, die, bvis, die))
    except AssertionError as e:
        bad += 1
        print("FAIL", tag, "path", _lib.lib().afr_last_fused_path(), str(e).split("\n")[0][:200], flush=True)
print("fuzz_fused: %d cases, %d failures, paths used %s" % (n, bad, paths))
sys.exit(1 if bad else 0)
