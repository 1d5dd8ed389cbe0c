"""Randomised parity sweep of the Stokes -> brightness rows (spectral_model, convert,
stokes_brightness, fused_predict_vis_stokes, beam_cube_dde_rotated) against the CPU oracle.
usage: fuzz_brightness.py [ncases] [seed]"""
import os, sys
import numpy as np
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT); sys.path.insert(0, os.path.join(ROOT, "tests"))
import oracle
from codex_africanus_b200 import model, rime
from conftest import assert_c128_close

n = int(sys.argv[1]) if len(sys.argv) > 1 else 40
rng = np.random.default_rng(int(sys.argv[2]) if len(sys.argv) > 2 else 0)
STOKES = ["I", "Q", "U", "V"]
LIN, CIRC = ["XX", "XY", "YX", "YY"], ["RR", "RL", "LR", "LL"]
bad = 0
rc = lambda shape: rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def check(tag, got, ref):
    global bad
    try:
        assert_c128_close(got, ref)
    except AssertionError as e:
        bad += 1
        print("FAIL", tag, str(e).split("\n")[0][:200], flush=True)


for case in range(n):
    nsrc = int(rng.choice([1, 2, 9, 40])); nchan = int(rng.choice([1, 3, 16, 33])); nspi = int(rng.choice([1, 2, 3, 5]))
    freq = np.linspace(0.856e9, 1.712e9, nchan) if nchan > 1 else np.array([1.1e9])
    rf = rng.uniform(0.9e9, 1.6e9, nsrc)
    bases = [rng.choice(["std", "log", "log10"]) for _ in range(int(rng.integers(1, 5)))]
    base = bases if rng.random() < 0.5 else bases[0]
    if rng.random() < 0.3:
        base = [{"std": 0, "log": 1, "log10": 2}[b] for b in bases] if isinstance(base, list) else {"std": 0, "log": 1, "log10": 2}[base]
    # --- spectral_model with an arbitrary polarisation count
    npol = int(rng.choice([1, 2, 4, 7]))
    st = rng.standard_normal((nsrc, npol)); spi = rng.standard_normal((nsrc, nspi, npol)) * 0.4
    tag = "case %d: nsrc %d nchan %d nspi %d npol %d base %s" % (case, nsrc, nchan, nspi, npol, base)
    check(tag + " spectral_model", model.spectral_model(st, spi, rf, freq, base=base),
          oracle.spectral_model(st, spi, rf, freq, base=base))
    # --- convert: random subset / order of a schema, both directions
    sub = list(rng.permutation(4)[: int(rng.integers(1, 5))])
    corr = LIN if rng.random() < 0.5 else CIRC
    out_schema = [corr[i] for i in sub]
    x = rng.standard_normal((nsrc, nchan, 4))
    check(tag + " convert s->c %s" % out_schema, model.convert(x, STOKES, out_schema), oracle.convert(x, STOKES, out_schema))
    v = rc((nsrc, nchan, 2, 2))
    out_schema = [STOKES[i] for i in sub]
    nested = [corr[:2], corr[2:]]
    check(tag + " convert c->s %s" % out_schema, model.convert(v, nested, out_schema), oracle.convert(v, nested, out_schema))
    # --- stokes_brightness with missing Stokes parameters (implicit zeros)
    have = sorted(rng.permutation(4)[: int(rng.integers(1, 5))])
    sch = [STOKES[i] for i in have]
    st4 = rng.standard_normal((nsrc, len(have))); spi4 = rng.standard_normal((nsrc, nspi, len(have))) * 0.4
    nested = [corr[:2], corr[2:]]
    ref_b = oracle.convert(oracle.spectral_model(st4, spi4, rf, freq, base=base), sch, nested, implicit_stokes=True)
    check(tag + " stokes_brightness %s" % sch,
          model.stokes_brightness(st4, spi4, rf, freq, base=base, stokes_schema=sch, corr_schema=nested,
                                  implicit_stokes=True), ref_b)
    # --- predict through it, random chunking, optional DDE / DIE / base_vis
    na = int(rng.choice([3, 5, 9])); ntime = int(rng.choice([1, 2, 4]))
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime); ti = np.repeat(np.arange(ntime), a1.size) + int(rng.integers(0, 5))
    pos = rng.standard_normal((ntime, na, 3)) * 1200.0
    uvw = pos[ti - ti.min(), ant1] - pos[ti - ti.min(), ant2]
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    dde = None if rng.random() < 0.5 else 1 + 0.2 * rc((nsrc, ntime, na, nchan, 2, 2))
    die = None if rng.random() < 0.5 else 1 + 0.1 * rc((ntime, na, nchan, 2, 2))
    bvis = None if rng.random() < 0.5 else rc((ti.size, nchan, 2, 2))
    chunk = int(rng.integers(1, nsrc + 2))
    got = rime.fused_predict_vis_stokes(lm, uvw, freq, st4, spi4, rf, ti, ant1, ant2, dde, dde, die, bvis, die,
                                        base=base, stokes_schema=sch, corr_schema=nested, implicit_stokes=True,
                                        source_chunk=chunk)
    check(tag + " fused_predict_vis_stokes chunk %d" % chunk, got,
          oracle.fused_predict(lm, uvw, freq, ref_b, ti, ant1, ant2, dde, dde, die, bvis, die))
    # --- streaming driver over the same rows
    parts = [b.copy() for _, b in rime.stream_predict_vis_stokes(
        lm, uvw, freq, st4, spi4, rf, ti, ant1, ant2, dde, dde, die, bvis, die,
        rows_per_block=int(rng.integers(1, ti.size + 1)), base=base, stokes_schema=sch, corr_schema=nested,
        implicit_stokes=True, source_chunk=chunk)]
    check(tag + " stream", np.concatenate(parts), got)
    # --- rotated beam DDE
    beam = rc((5, 6, 4, 2, 2)); ext = np.array([[-0.03, 0.03], [-0.03, 0.03]]); bfm = np.linspace(0.8e9, 1.8e9, 4)
    pa = rng.uniform(-3, 3, (ntime, na)); pe = rng.uniform(-1e-3, 1e-3, (ntime, na, nchan, 2))
    asc = rng.uniform(0.9, 1.1, (na, nchan, 2)); ft = "linear" if rng.random() < 0.5 else "circular"
    ref = np.einsum("stafij,tajk->stafik", oracle.beam_cube_dde(beam, ext, bfm, lm, pa, pe, asc, freq),
                    oracle.feed_rotation(pa, ft))
    check(tag + " beam_cube_dde_rotated " + ft,
          rime.beam_cube_dde_rotated(beam, ext, bfm, lm, pa, pe, asc, freq, rime.feed_rotation(pa, ft)), ref)
print("fuzz_brightness: %d cases, %d failures" % (n, bad))
sys.exit(1 if bad else 0)
