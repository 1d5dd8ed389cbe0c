"""
The dask-facing wrappers (codex_africanus_b200.dft.dask / .rime.dask) on CPU: the graphs -- index
strings, chunk checks, block-function calling convention, ``streams=`` semantics -- are exercised
through the eager ChunkedArray backend with the ORACLE standing in for the CUDA block functions
(tests may use it), on the reference's own ``test_dask_*`` parametrisations
(africanus/rime/tests/test_predict.py:20-50,129-212, africanus/dft/tests/test_dft.py:219-294).
With dask installed the same cases also run through real dask graphs.
"""
import itertools

import numpy as np
import pytest

from codex_africanus_b200 import _chunked as ck


def rc(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


@pytest.fixture
def cpu_blocks(oracle, monkeypatch):
    """Swap the CUDA block functions for the oracle's."""
    import codex_africanus_b200.dft.dask as dd
    import codex_africanus_b200.rime.dask as rd

    monkeypatch.setattr(dd, "np_im_to_vis", oracle.im_to_vis)
    monkeypatch.setattr(dd, "np_vis_to_im", oracle.vis_to_im)
    monkeypatch.setattr(rd, "np_predict_vis", oracle.predict_vis)
    monkeypatch.setattr(rd, "np_phase_delay", oracle.phase_delay)
    monkeypatch.setattr(rd, "np_beam_cube_dde", oracle.beam_cube_dde)
    return dd, rd


BACKENDS = ["chunked"] + (["dask"] if ck.have_dask() else [])


def _from_array(backend, x, chunks):
    if backend == "dask":
        return ck.da.from_array(x, chunks=chunks)
    return ck.from_array(x, chunks)


def _compute(x):
    return x.compute()


def test_blockwise_emulation_matches_dask_calling_convention():
    a = ck.from_array(np.arange(24.0).reshape(4, 6), ((1, 3), (2, 4)))
    b = ck.from_array(np.arange(6.0), ((2, 4),))
    assert a.numblocks == (2, 2) and a.block((1, 1)).shape == (3, 4)
    seen = []

    def f(x, y, scale=1.0):
        # "j" is contracted: x arrives as the list of its blocks along j, y as the list of its blocks
        assert isinstance(x, list) and isinstance(y, list) and len(x) == 2 and len(y) == 2
        seen.append(tuple(blk.shape for blk in x))
        return scale * sum(blk @ v for blk, v in zip(x, y))

    out = ck.blockwise(f, ("i",), a, ("i", "j"), b, ("j",), scale=2.0)
    np.testing.assert_allclose(out.compute(), 2.0 * a.data @ b.data)
    assert out.chunks == ((1, 3),) and seen == [((1, 2), (1, 4)), ((3, 2), (3, 4))]
    # adjust_chunks + a new leading block axis, then the sum over it (the vis_to_im pattern)
    out = ck.blockwise(lambda x: x.sum(axis=0)[None, :], ("i", "j"), a, ("i", "j"), adjust_chunks={"i": 1})
    assert out.shape == (2, 6) and out.chunks == ((1, 1), (2, 4))
    np.testing.assert_allclose(out.sum(axis=0).compute(), a.data.sum(axis=0))
    # align_arrays=False pairs blocks by position although the chunk sizes differ
    t = ck.from_array(np.arange(3.0), ((2, 1),))
    r = ck.from_array(np.arange(5.0), ((3, 2),))
    out = ck.blockwise(lambda rr, tt: rr + tt.sum(), ("row",), r, ("row",), t, ("row",), align_arrays=False,
                       adjust_chunks={"row": r.chunks[0]})
    np.testing.assert_allclose(out.compute(), [1, 2, 3, 5, 6])
    with pytest.raises(ValueError):
        ck.blockwise(lambda rr, tt: rr, ("row",), r, ("row",), t, ("row",))
    with pytest.raises(ValueError):
        ck.from_array(np.zeros(5), ((2, 2),))


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("corr_shape", [(1,), (2,), (2, 2)])
@pytest.mark.parametrize("a1j,blj,a2j", [(True, True, True), (True, False, True), (False, True, False)])
@pytest.mark.parametrize("g1j,bvis,g2j", [(True, True, True), (True, False, True), (False, True, False)])
def test_dask_predict_vis(cpu_blocks, oracle, backend, corr_shape, a1j, blj, a2j, g1j, bvis, g2j):
    """africanus/rime/tests/test_predict.py:129-212: streams=True and streams=False against the
    single-call result, the reference's chunk layout (9 source chunks, time chunks (2,1,1) paired
    with row chunks (4,4,2))."""
    _, rd = cpu_blocks
    rng = np.random.default_rng(42)
    sc, tc, rrc, ac, cc = (2, 3, 4, 2, 2, 2, 2, 2, 2), (2, 1, 1), (4, 4, 2), (4,), (3, 2)
    s, t, a, c, r = sum(sc), sum(tc), sum(ac), sum(cc), sum(rrc)
    a1_jones, a2_jones = rc(rng, (s, t, a, c) + corr_shape), rc(rng, (s, t, a, c) + corr_shape)
    bl_jones = rc(rng, (s, r, c) + corr_shape)
    g1_jones, g2_jones = rc(rng, (t, a, c) + corr_shape), rc(rng, (t, a, c) + corr_shape)
    base_vis = rc(rng, (r, c) + corr_shape)
    time_idx = np.asarray([0, 0, 1, 1, 2, 2, 2, 2, 3, 3])
    ant1 = np.asarray([0, 0, 0, 0, 1, 1, 1, 2, 2, 3])
    ant2 = np.asarray([0, 1, 2, 3, 1, 2, 3, 2, 3, 3])
    ref = oracle.predict_vis(time_idx, ant1, ant2, a1_jones if a1j else None, bl_jones if blj else None,
                             a2_jones if a2j else None, g1_jones if g1j else None, base_vis if bvis else None,
                             g2_jones if g2j else None)
    F = lambda x, ch: _from_array(backend, x, ch)  # noqa: E731
    args = (F(time_idx, (rrc,)), F(ant1, (rrc,)), F(ant2, (rrc,)),
            F(a1_jones, (sc, tc, ac, cc) + corr_shape) if a1j else None,
            F(bl_jones, (sc, rrc, cc) + corr_shape) if blj else None,
            F(a2_jones, (sc, tc, ac, cc) + corr_shape) if a2j else None,
            F(g1_jones, (tc, ac, cc) + corr_shape) if g1j else None,
            F(base_vis, (rrc, cc) + corr_shape) if bvis else None,
            F(g2_jones, (tc, ac, cc) + corr_shape) if g2j else None)
    fan = _compute(rd.predict_vis(*args, streams=False))
    stream = _compute(rd.predict_vis(*args, streams=True))
    np.testing.assert_allclose(fan, ref, rtol=1e-12, atol=1e-12)
    np.testing.assert_allclose(stream, ref, rtol=1e-12, atol=1e-12)


@pytest.mark.parametrize("backend", BACKENDS)
def test_dask_predict_vis_chunk_errors(cpu_blocks, backend):
    _, rd = cpu_blocks
    rng = np.random.default_rng(1)
    F = lambda x, ch: _from_array(backend, x, ch)  # noqa: E731
    ti, a1, a2 = (F(np.zeros(4, np.int64), ((2, 2),)) for _ in range(3))
    dde = rc(rng, (3, 2, 4, 5, 2, 2))
    with pytest.raises(ValueError, match="antenna dimension"):
        rd.predict_vis(ti, a1, a2, F(dde, ((3,), (1, 1), (2, 2), (5,), 2, 2)), None,
                       F(dde, ((3,), (1, 1), (2, 2), (5,), 2, 2)))
    with pytest.raises(ValueError, match="row chunks"):
        rd.predict_vis(ti, a1, a2, F(dde, ((3,), (2,), (4,), (5,), 2, 2)), None, F(dde, ((3,), (2,), (4,), (5,), 2, 2)))
    with pytest.raises(ValueError, match="dde1_jones.chunks != dde2_jones.chunks"):
        rd.predict_vis(ti, a1, a2, F(dde, ((3,), (1, 1), (4,), (5,), 2, 2)), None,
                       F(dde, ((1, 2), (1, 1), (4,), (5,), 2, 2)))
    with pytest.raises(ValueError, match="present or absent"):
        rd.predict_vis(ti, a1, a2, F(dde, ((3,), (1, 1), (4,), (5,), 2, 2)), None, None)


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("convention", ["fourier", "casa"])
def test_dask_dft(cpu_blocks, oracle, backend, convention):
    """africanus/dft/tests/test_dft.py:219-294 (smaller extents, same chunk structure: rows in 8
    chunks, channels in chunks of nchan // 2, one source chunk)."""
    dd, _ = cpu_blocks
    rng = np.random.default_rng(3)
    nrow, nsource, nchan, ncorr = 160, 37, 11, 4
    uvw = 100 * rng.random((nrow, 3))
    lm = 0.01 * rng.standard_normal((nsource, 2))
    frequency = np.linspace(1.0, 2.0, nchan) * 2.99792458e8
    image = rng.standard_normal((nsource, nchan, ncorr))
    F = lambda x, ch: _from_array(backend, x, ch)  # noqa: E731
    uvw_c, lm_c, f_c = F(uvw, (nrow // 8, 3)), F(lm, (nsource, 2)), F(frequency, (nchan // 2,))
    vis = oracle.im_to_vis(image, uvw, lm, frequency, convention=convention)
    got = _compute(dd.im_to_vis(F(image, (nsource, nchan // 2, ncorr)), uvw_c, lm_c, f_c, convention=convention))
    assert got.dtype == np.complex128
    np.testing.assert_allclose(got, vis, rtol=1e-13, atol=1e-13 * np.abs(vis).max())
    visr = rng.standard_normal((nrow, nchan, ncorr))
    flags = rng.random((nrow, nchan, ncorr)) < 0.55
    img = oracle.vis_to_im(visr, uvw, lm, frequency, flags, convention=convention)
    got = _compute(dd.vis_to_im(F(visr, (nrow // 8, nchan // 2, ncorr)), uvw_c, lm_c, f_c,
                                F(flags, (nrow // 8, nchan // 2, ncorr)), convention=convention))
    assert got.dtype == np.float64 and got.shape == img.shape
    np.testing.assert_allclose(got, img, rtol=1e-12, atol=1e-12 * np.abs(img).max())
    with pytest.raises(ValueError, match="lm chunks"):
        dd.im_to_vis(F(image, (nsource, nchan // 2, ncorr)), uvw_c, F(lm, (10, 2)), f_c)
    with pytest.raises(ValueError, match="frequency chunks"):
        dd.im_to_vis(F(image, (nsource, nchan, ncorr)), uvw_c, lm_c, f_c)
    with pytest.raises(ValueError, match="flags chunks"):
        dd.vis_to_im(F(visr, (nrow // 8, nchan // 2, ncorr)), uvw_c, lm_c, f_c, F(flags, (nrow // 4, nchan // 2, ncorr)))


@pytest.mark.parametrize("backend", BACKENDS)
def test_dask_phase_delay_and_beam(cpu_blocks, oracle, backend):
    _, rd = cpu_blocks
    rng = np.random.default_rng(5)
    F = lambda x, ch: _from_array(backend, x, ch)  # noqa: E731
    lm = rng.uniform(-0.1, 0.1, (9, 2))
    uvw = rng.standard_normal((20, 3)) * 500
    freq = np.linspace(1e9, 1.5e9, 7)
    got = _compute(rd.phase_delay(F(lm, ((4, 5), 2)), F(uvw, ((7, 7, 6), 3)), F(freq, ((3, 4),)), convention="casa"))
    np.testing.assert_array_equal(got, oracle.phase_delay(lm, uvw, freq, convention="casa"))
    lw, mh, nud, ntime, nant = 11, 9, 5, 4, 3
    beam = rc(rng, (lw, mh, nud, 2, 2))
    ext = np.array([[-0.1, 0.1], [-0.1, 0.1]])
    bfm = np.linspace(1e9, 1.5e9, nud)
    pa = rng.uniform(-1, 1, (ntime, nant))
    perr = rng.uniform(-0.01, 0.01, (ntime, nant, 7, 2))
    asc = rng.uniform(0.9, 1.1, (nant, 7, 2))
    got = _compute(rd.beam_cube_dde(F(beam, beam.shape), F(ext, ext.shape), F(bfm, bfm.shape), F(lm, ((4, 5), 2)),
                                    F(pa, ((1, 3), nant)), F(perr, ((1, 3), nant, (3, 4), 2)),
                                    F(asc, (nant, (3, 4), 2)), F(freq, ((3, 4),))))
    np.testing.assert_array_equal(got, oracle.beam_cube_dde(beam, ext, bfm, lm, pa, perr, asc, freq))
    with pytest.raises(ValueError, match="Beam chunking"):
        rd.beam_cube_dde(F(beam, ((5, 6), mh, nud, 2, 2)), F(ext, ext.shape), F(bfm, bfm.shape), F(lm, ((4, 5), 2)),
                         F(pa, ((1, 3), nant)), F(perr, ((1, 3), nant, (3, 4), 2)), F(asc, (nant, (3, 4), 2)),
                         F(freq, ((3, 4),)))


def test_plain_numpy_inputs_are_refused_without_dask():
    import codex_africanus_b200.dft.dask as dd

    class Fake:  # has .chunks / .shape but is neither a dask array nor a ChunkedArray
        def __init__(self, x):
            self.shape, self.chunks, self.dtype, self.ndim = x.shape, tuple((n,) for n in x.shape), x.dtype, x.ndim

    x = [Fake(np.zeros(s)) for s in ((3, 4, 1), (5, 3), (3, 2), (4,))]
    with pytest.raises((ImportError, TypeError)):
        dd.im_to_vis(*x)
