"""
The dask-facing wrappers with the CUDA kernels as block functions (ChunkedArray backend; real dask
graphs too when dask is installed): chunked evaluation must reproduce the single-call result and
the oracle.  Chunk layouts of the reference's test_dask_* (see tests/test_dask_wrappers.py).
"""
import numpy as np
import pytest

from conftest import assert_c128_close

from codex_africanus_b200 import _chunked as ck

pytestmark = pytest.mark.gpu

BACKENDS = ["chunked"] + (["dask"] if ck.have_dask() else [])


def rc(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def _F(backend):
    if backend == "dask":
        return lambda x, ch: ck.da.from_array(x, chunks=ch)
    return ck.from_array


@pytest.mark.parametrize("backend", BACKENDS)
@pytest.mark.parametrize("corr_shape", [(1,), (2, 2)])
def test_dask_predict_vis_cuda(oracle, backend, corr_shape):
    import codex_africanus_b200.rime as rime
    import codex_africanus_b200.rime.dask as rd

    rng = np.random.default_rng(42)
    sc, tc, rrc, ac, cc = (2, 3, 4, 2, 2, 2, 2, 2, 2), (2, 1, 1), (4, 4, 2), (4,), (3, 2)
    s, t, a, c, r = sum(sc), sum(tc), sum(ac), sum(cc), sum(rrc)
    dde1, dde2 = rc(rng, (s, t, a, c) + corr_shape), rc(rng, (s, t, a, c) + corr_shape)
    coh = rc(rng, (s, r, c) + corr_shape)
    g1, g2 = rc(rng, (t, a, c) + corr_shape), rc(rng, (t, a, c) + corr_shape)
    bvis = rc(rng, (r, c) + corr_shape)
    ti = np.asarray([0, 0, 1, 1, 2, 2, 2, 2, 3, 3]) + 7
    a1 = np.asarray([0, 0, 0, 0, 1, 1, 1, 2, 2, 3])
    a2 = np.asarray([0, 1, 2, 3, 1, 2, 3, 2, 3, 3])
    F = _F(backend)
    for present in ((1, 1, 1, 1, 1, 1), (1, 0, 1, 0, 0, 0), (0, 1, 0, 1, 0, 1), (0, 0, 0, 1, 1, 1)):
        full = [x if p else None for x, p in zip((dde1, coh, dde2, g1, bvis, g2), present)]
        chunks = ((sc, tc, ac, cc), (sc, rrc, cc), (sc, tc, ac, cc), (tc, ac, cc), (rrc, cc), (tc, ac, cc))
        ch = [None if x is None else F(x, c_ + corr_shape) for x, c_ in zip(full, chunks)]
        one = rime.predict_vis(ti, a1, a2, *full)
        ref = oracle.predict_vis(ti, a1, a2, *full)
        idx = (F(ti, (rrc,)), F(a1, (rrc,)), F(a2, (rrc,)))
        for streams in (False, True):
            got = rd.predict_vis(*idx, *ch, streams=streams).compute()
            assert_c128_close(got, one, rtol=1e-12)
            assert_c128_close(got, ref)


@pytest.mark.parametrize("backend", BACKENDS)
def test_dask_dft_and_phase_cuda(oracle, backend):
    import codex_africanus_b200.dft as dft
    import codex_africanus_b200.dft.dask as dd
    import codex_africanus_b200.rime.dask as rd

    rng = np.random.default_rng(3)
    nrow, nsource, nchan, ncorr = 800, 81, 11, 4
    uvw = 100 * rng.random((nrow, 3))
    lm = 0.01 * rng.standard_normal((nsource, 2))
    frequency = np.linspace(1.0, 2.0, nchan) * 2.99792458e8
    image = rng.standard_normal((nsource, nchan, ncorr))
    F = _F(backend)
    uvw_c, lm_c, f_c = F(uvw, (nrow // 8, 3)), F(lm, (nsource, 2)), F(frequency, (nchan // 2,))
    got = dd.im_to_vis(F(image, (nsource, nchan // 2, ncorr)), uvw_c, lm_c, f_c).compute()
    assert_c128_close(got, dft.im_to_vis(image, uvw, lm, frequency), rtol=1e-11)
    assert_c128_close(got, oracle.im_to_vis(image, uvw, lm, frequency))
    vis = rc(rng, (nrow, nchan, ncorr))
    flags = rng.random((nrow, nchan, ncorr)) < 0.55
    got = dd.vis_to_im(F(vis, (nrow // 8, nchan // 2, ncorr)), uvw_c, lm_c, f_c,
                       F(flags, (nrow // 8, nchan // 2, ncorr))).compute()
    assert_c128_close(got, oracle.vis_to_im(vis, uvw, lm, frequency, flags))
    got = rd.phase_delay(F(lm, ((40, 41), 2)), F(uvw, (nrow // 4, 3)), f_c).compute()
    assert_c128_close(got, oracle.phase_delay(lm, uvw, frequency))
