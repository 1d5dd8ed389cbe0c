"""
GPU parity tests (B200): every entry point of codex_africanus_b200, called through the
ctypes C-ABI, against (i) the committed golden vectors produced by the reference's numba
implementation and (ii) the CPU oracle on fresh seeded inputs.

Gates:  complex128 / float64 -> allclose(rtol=1e-10, atol=1e-10*max|ref|);
complex64 / float32 -> ||got - ref||_2 <= 1e-5 ||ref||_2.
"""
import itertools

import numpy as np
import pytest

from conftest import assert_c128_close, assert_c64_close, rel_l2

pytestmark = pytest.mark.gpu

C = 2.99792458e8


@pytest.fixture(scope="module")
def b200():
    import codex_africanus_b200.dft as dft
    import codex_africanus_b200.model as model
    import codex_africanus_b200.rime as rime

    class NS:
        pass

    ns = NS()
    ns.dft, ns.rime, ns.model = dft, rime, model
    return ns


# ----------------------------------------------------------------------------- phase_delay
def test_phase_delay_golden(b200, golden):
    g = golden("phase_delay")
    lm, uvw, freq = g["lm"], g["uvw"], g["freq"]
    for conv in ("fourier", "casa"):
        assert_c128_close(b200.rime.phase_delay(lm, uvw, freq, convention=conv), g["f64_" + conv])
    assert_c128_close(b200.rime.phase_delay(lm, uvw, g["freq_nu"]), g["f64_nonuniform"])
    f32 = np.float32
    lm_s, uvw_s = g["lm_s"], g["uvw_s"]
    got = b200.rime.phase_delay(lm_s.astype(f32), uvw_s.astype(f32), freq.astype(f32))
    assert_c64_close(got, g["f32_all"])
    assert np.max(np.abs(got - g["f32_all"])) < 2e-6
    assert_c128_close(b200.rime.phase_delay(lm_s.astype(f32), uvw_s, freq), g["mix_lm32"])
    assert_c128_close(b200.rime.phase_delay(lm_s.astype(f32), uvw_s.astype(f32), freq), g["mix_lm32_uvw32"])
    assert_c128_close(b200.rime.phase_delay(lm_s, uvw_s.astype(f32), freq), g["mix_uvw32"])
    assert_c128_close(b200.rime.phase_delay(lm_s, uvw_s, freq.astype(f32)), g["mix_freq32"])


@pytest.mark.parametrize("convention, sign", [("fourier", 1), ("casa", -1)])
def test_phase_delay_known_answer(b200, convention, sign):
    # rime/tests/test_rime.py:19-47; the rotation recurrence is not bit-exact, so the
    # reference's == becomes the project gate
    rng = np.random.default_rng(0)
    uvw = rng.random((100, 3))
    lm = rng.random((10, 2))
    frequency = np.linspace(0.856e9, 0.856e9 * 2, 64, endpoint=True)
    uvw[2] = [1, 2, 3]
    lm[3] = [0.1, 0.2]
    cp = b200.rime.phase_delay(lm, uvw, frequency, convention=convention)
    n = np.sqrt(1.0 - 0.1**2 - 0.2**2) - 1.0
    phase = sign * (-2 * np.pi / C) * (1 * 0.1 + 2 * 0.2 + 3 * n) * frequency[5]
    assert abs(np.exp(1j * phase) - cp[3, 2, 5]) < 1e-10


def test_phase_delay_vs_oracle_ragged(b200, oracle):
    rng = np.random.default_rng(11)
    for nsrc, nrow, nchan in [(1, 1, 1), (3, 5, 7), (2, 129, 33), (5, 64, 256)]:
        lm = rng.uniform(-0.1, 0.1, (nsrc, 2))
        uvw = rng.standard_normal((nrow, 3)) * 3000
        freq = np.linspace(0.856e9, 1.712e9, nchan) if nchan > 1 else np.array([1.0e9])
        assert_c128_close(b200.rime.phase_delay(lm, uvw, freq), oracle.phase_delay(lm, uvw, freq))
    out = b200.rime.phase_delay(np.zeros((0, 2)), np.zeros((4, 3)), np.ones(3))
    assert out.shape == (0, 4, 3) and out.dtype == np.complex128


# ----------------------------------------------------------------------------- dft
def test_dft_golden(b200, golden):
    g = golden("dft")
    lm, uvw, freq = g["lm"], g["uvw"], g["freq"]
    f32 = np.float32
    for nc in (1, 2, 4):
        assert_c128_close(b200.dft.im_to_vis(g["image_r%d" % nc], uvw, lm, freq), g["i2v_r%d" % nc])
        got = b200.dft.vis_to_im(g["vis_c%d" % nc], uvw, lm, freq, g["flags_%d" % nc])
        assert_c128_close(got, g["v2i_c%d" % nc])
    assert_c128_close(b200.dft.im_to_vis(g["image_c2"], uvw, lm, freq), g["i2v_c2"])
    assert_c128_close(b200.dft.im_to_vis(g["image_c2"], uvw, lm, freq, convention="casa"), g["i2v_c2_casa"])
    assert_c128_close(b200.dft.im_to_vis(g["image_r1"], uvw, lm, g["freq_nu"]), g["i2v_r1_nonuniform"])
    assert_c64_close(b200.dft.im_to_vis(g["image_r1"], uvw, lm, freq, dtype=np.complex64), g["i2v_r1_c64"])
    got = b200.dft.im_to_vis(g["image_r1"].astype(f32), uvw.astype(f32), lm.astype(f32), freq.astype(f32))
    assert_c64_close(got, g["i2v_r1_in32"], tol=2e-5)
    assert_c128_close(b200.dft.im_to_vis(g["image_r1"], uvw, lm.astype(f32), freq), g["i2v_r1_lm32"])
    assert_c128_close(b200.dft.vis_to_im(g["vis_r2"], uvw, lm, freq, g["flags_2"]), g["v2i_r2"])
    assert_c128_close(
        b200.dft.vis_to_im(g["vis_c1"], uvw, lm, freq, g["flags_1"], convention="casa"), g["v2i_c1_casa"])
    assert_c64_close(
        b200.dft.vis_to_im(g["vis_c1"], uvw, lm, freq, g["flags_1"], dtype=np.float32), g["v2i_c1_f32"])
    assert_c128_close(
        b200.dft.vis_to_im(g["vis_c1"], uvw, lm, g["freq_nu"], g["flags_1"]), g["v2i_c1_nonuniform"])
    got = b200.dft.vis_to_im(g["vis_c1"].astype(np.complex64), uvw.astype(f32), lm.astype(f32),
                             freq.astype(f32), g["flags_1"])
    assert_c64_close(got, g["v2i_c1_in32"], tol=2e-5)


def test_dft_padded_grid_golden(b200, golden):
    """A zero-padded image grid reaching past l^2 + m^2 = 1 (n = NaN, dft/kernels.py:54) stays
    finite because zero pixels are skipped (kernels.py:64); a bright pixel outside the disc gives
    NaN for every row at exactly the (chan, corr) entries where it is non-zero."""
    import torch

    g = golden("dft_padded")
    lm, uvw, freq = g["lm"], g["uvw"], g["freq"]
    got = b200.dft.im_to_vis(g["image"], uvw, lm, freq)
    assert np.all(np.isfinite(got))
    assert_c128_close(got, g["i2v"])
    assert_c64_close(b200.dft.im_to_vis(g["image"], uvw, lm, freq, dtype=np.complex64), g["i2v_c64"])
    t = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    bad = b200.dft.im_to_vis(t(g["image_bad"]), t(uvw), t(lm), t(freq)).cpu().numpy()
    np.testing.assert_array_equal(np.isnan(bad), np.isnan(g["i2v_bad"]))
    assert_c128_close(np.nan_to_num(bad), np.nan_to_num(g["i2v_bad"]))
    # complex image, non-equispaced channels (one sincos per term)
    imgc = g["image_bad"] * (1.0 + 0.5j)
    fq = np.sort(freq * (1 + 0.01 * np.cos(np.arange(freq.size))))
    got = b200.dft.im_to_vis(imgc, uvw, lm, fq)
    pat = np.isnan(g["i2v_bad"])
    np.testing.assert_array_equal(np.isnan(got), pat)


@pytest.mark.parametrize("nsrc,nrow,nchan,ncorr", [
    (1, 1, 1, 1), (7, 33, 5, 1), (40, 300, 64, 1), (19, 77, 256, 1), (9, 41, 300, 2),
    (11, 65, 70, 4), (6, 35, 17, 3), (3000, 40, 16, 1), (33, 257, 48, 2), (50, 531, 128, 4),
    (21, 1203, 32, 1),
])
def test_dft_vs_oracle_shapes(b200, oracle, nsrc, nrow, nchan, ncorr):
    """ragged / non-multiple-of-tile shapes, channel tails, y-split path (many sources, few rows)"""
    rng = np.random.default_rng(nsrc * 1000 + nrow)
    lm = rng.uniform(-0.03, 0.03, (nsrc, 2))
    uvw = rng.standard_normal((nrow, 3)) * 4000.0
    freq = np.linspace(0.856e9, 1.712e9, nchan) if nchan > 1 else np.array([1.4e9])
    image = rng.standard_normal((nsrc, nchan, ncorr))
    assert_c128_close(b200.dft.im_to_vis(image, uvw, lm, freq), oracle.im_to_vis(image, uvw, lm, freq))
    imagec = image + 1j * rng.standard_normal((nsrc, nchan, ncorr))
    assert_c128_close(b200.dft.im_to_vis(imagec, uvw, lm, freq), oracle.im_to_vis(imagec, uvw, lm, freq))
    vis = rng.standard_normal((nrow, nchan, ncorr)) + 1j * rng.standard_normal((nrow, nchan, ncorr))
    flags = rng.random((nrow, nchan, ncorr)) < 0.05
    assert_c128_close(b200.dft.vis_to_im(vis, uvw, lm, freq, flags),
                      oracle.vis_to_im(vis, uvw, lm, freq, flags))
    assert_c128_close(b200.dft.vis_to_im(vis.real.copy(), uvw, lm, freq, flags),
                      oracle.vis_to_im(vis.real.copy(), uvw, lm, freq, flags))


def test_dft_c64_vs_oracle(b200, oracle):
    rng = np.random.default_rng(77)
    nsrc, nrow, nchan = 500, 200, 64
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    uvw = rng.standard_normal((nrow, 3)) * 3000.0
    freq = np.linspace(0.856e9, 1.712e9, nchan)
    image = np.abs(rng.standard_normal((nsrc, nchan, 1)))
    got = b200.dft.im_to_vis(image, uvw, lm, freq, dtype=np.complex64)
    assert_c64_close(got, oracle.im_to_vis(image, uvw, lm, freq, dtype=np.complex64))
    # the float64 oracle is the tighter yardstick for the FP32 kernel's own error
    assert rel_l2(got.astype(np.complex128), oracle.im_to_vis(image, uvw, lm, freq)) < 1e-5
    vis = rng.standard_normal((nrow, nchan, 1)) + 1j * rng.standard_normal((nrow, nchan, 1))
    flags = rng.random((nrow, nchan, 1)) < 0.05
    got = b200.dft.vis_to_im(vis, uvw, lm, freq, flags, dtype=np.float32)
    assert_c64_close(got, oracle.vis_to_im(vis, uvw, lm, freq, flags, dtype=np.float32))


def test_im_to_vis_phase_centre(b200):
    # dft/tests/test_dft.py:12-42
    nrow, npix, nchan, ncorr = 100, 35, 11, 2
    uvw = np.random.default_rng(1).random((nrow, 3))
    x = np.linspace(-0.1, 0.1, npix)
    ll, mm = np.meshgrid(x, x)
    lm = np.vstack((ll.flatten(), mm.flatten())).T
    frequency = np.linspace(1.0, 2.0, nchan, endpoint=True)
    image = np.zeros((npix, npix, nchan, ncorr))
    Inu = (frequency / frequency[nchan // 2]) ** (-0.7)
    for corr in range(ncorr):
        image[npix // 2, npix // 2, :, corr] = Inu
    vis = b200.dft.im_to_vis(image.reshape(npix**2, nchan, ncorr), uvw, lm, frequency)
    tmp = vis - Inu[None, :, None]
    assert np.all(np.abs(tmp.real) < 1e-13) and np.all(np.abs(tmp.imag) < 1e-13)


@pytest.mark.parametrize("convention", ["fourier", "casa"])
def test_im_to_vis_fft(b200, convention):
    # dft/tests/test_dft.py:86-133
    np.random.seed(123)
    Fs, iFs = np.fft.fftshift, np.fft.ifftshift
    npix, ncorr, nsource = 29, 1, 25
    image = np.zeros((npix, npix, ncorr))
    Ix = np.random.randint(5, npix - 5, nsource)
    Iy = np.random.randint(5, npix - 5, nsource)
    image[Ix, Iy, 0] = np.random.randn(nsource)
    fft_image = Fs(np.fft.fft2(iFs(image[:, :, 0]))).reshape(npix**2, 1, 1)
    deltal = 0.001
    l_coord = np.arange(-(npix // 2), npix // 2 + 1) * deltal
    ll, mm = np.meshgrid(l_coord, l_coord)
    lm = np.vstack((ll.flatten(), mm.flatten())).T
    u = Fs(np.fft.fftfreq(npix, d=deltal))
    uu, vv = np.meshgrid(u, u)
    uvw = np.zeros((npix**2, 3))
    uvw[:, 0], uvw[:, 1] = uu.flatten(), vv.flatten()
    vis = b200.dft.im_to_vis(image.reshape(npix**2, 1, ncorr), uvw, lm, np.ones(1) * C,
                             convention=convention)
    fft_image = np.conj(fft_image) if convention == "casa" else fft_image
    np.testing.assert_array_almost_equal(vis, fft_image, decimal=12)


def test_adjointness_flags_symmetry(b200):
    # dft/tests/test_dft.py:136-177, :180-215, :297-331
    np.random.seed(123)
    nsource, nrow, nchan, ncorr = 21, 31, 3, 4
    uvw = 100 * np.random.random(size=(nrow, 3))
    lm = np.vstack((0.01 * np.random.randn(nsource), 0.01 * np.random.randn(nsource))).T
    frequency = np.arange(1, nchan + 1) * C
    gamma_im = np.random.randn(nsource, nchan, ncorr)
    gamma_vis = np.random.randn(nrow, nchan, ncorr)
    flag = np.zeros((nrow, nchan, ncorr), dtype=bool)
    LHS = np.vdot(gamma_vis.ravel(), b200.dft.im_to_vis(gamma_im, uvw, lm, frequency).ravel()).real
    RHS = np.dot(b200.dft.vis_to_im(gamma_vis, uvw, lm, frequency, flag).ravel(), gamma_im.ravel())
    assert np.abs(LHS - RHS) < 1e-11 * max(1.0, abs(LHS))
    uvw[0, :] = 0.0
    vis = np.random.randn(nrow, nchan, ncorr) + 1.0j * np.random.randn(nrow, nchan, ncorr)
    vis[0, :, :] = 1.0
    flags = np.ones((nrow, nchan, ncorr), dtype=bool)
    flags[0, :, :] = 0
    im = b200.dft.vis_to_im(vis, uvw, lm, np.ones(nchan) * C, flags)
    np.testing.assert_array_almost_equal(im, np.ones((nsource, nchan, ncorr)), decimal=13)
    # R^H R symmetric
    nsource = 25
    lm = np.random.uniform(-0.05, 0.05, (nsource, 2))
    uvw = np.random.randn(1000, 3) * 1000
    uvw[:, 2] = 0.0
    freq = np.array([1.0e9])
    flags = np.zeros((1000, 1, 1), dtype=bool)
    psf = np.zeros((nsource, nsource))
    for s in range(nsource):
        Ki = b200.dft.im_to_vis(np.ones((1, 1, 1)), uvw, lm[s].reshape(1, 2), freq)
        psf[:, s] = b200.dft.vis_to_im(Ki, uvw, lm, freq, flags).squeeze()
    np.testing.assert_array_almost_equal(psf, psf.T, decimal=9)


def test_dft_torch_inputs_and_strides(b200, oracle):
    import torch

    rng = np.random.default_rng(5)
    nsrc, nrow, nchan, ncorr = 17, 50, 32, 2
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    uvw = rng.standard_normal((nrow, 3)) * 1000.0
    freq = np.linspace(1e9, 2e9, nchan)
    image = rng.standard_normal((nsrc, nchan, ncorr))
    ref = oracle.im_to_vis(image, uvw, lm, freq)
    got = b200.dft.im_to_vis(torch.from_numpy(image).cuda(), torch.from_numpy(uvw).cuda(),
                             torch.from_numpy(lm).cuda(), torch.from_numpy(freq).cuda())
    assert isinstance(got, torch.Tensor) and got.is_cuda
    assert_c128_close(got.cpu().numpy(), ref)
    # non-contiguous (transposed) numpy inputs, as calibration/utils/tests/test_utils.py:57-70
    image_t = np.ascontiguousarray(image.transpose(2, 0, 1)).transpose(1, 2, 0)
    uvw_t = np.ascontiguousarray(uvw.T).T
    assert not image_t.flags.c_contiguous
    assert_c128_close(b200.dft.im_to_vis(image_t, uvw_t, lm, freq), ref)


def test_dft_large_properties(b200):
    """Size-independent checks at a size the oracle cannot reach: linearity and adjointness."""
    rng = np.random.default_rng(2)
    nsrc, nrow, nchan = 2000, 20000, 256
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    uvw = rng.standard_normal((nrow, 3)) * 3000.0
    freq = np.linspace(0.856e9, 1.712e9, nchan)
    a = rng.standard_normal((nsrc, nchan, 1))
    b = rng.standard_normal((nsrc, nchan, 1))
    va = b200.dft.im_to_vis(a, uvw, lm, freq)
    vb = b200.dft.im_to_vis(b, uvw, lm, freq)
    vab = b200.dft.im_to_vis(2.0 * a - 3.0 * b, uvw, lm, freq)
    assert rel_l2(vab, 2.0 * va - 3.0 * vb) < 1e-12
    y = rng.standard_normal((nrow, nchan, 1)) + 1j * rng.standard_normal((nrow, nchan, 1))
    flags = np.zeros(y.shape, bool)
    lhs = np.vdot(y.ravel(), va.ravel()).real
    rhs = np.dot(b200.dft.vis_to_im(y, uvw, lm, freq, flags).ravel(), a.ravel())
    assert abs(lhs - rhs) < 1e-10 * max(abs(lhs), np.linalg.norm(y) * np.linalg.norm(va))


# ----------------------------------------------------------------------------- predict_vis
PRESENCE = [(True, True, True), (True, False, True), (False, True, False)]


@pytest.mark.parametrize("cname", ["c1", "c2", "c22"])
def test_predict_vis_golden(b200, golden, oracle, cname):
    g = golden("predict_vis")
    ti, a1, a2 = g["time_idx"], g["ant1"], g["ant2"]
    arrs = {k: g["%s_%s" % (cname, k)] for k in ("a1j", "blj", "a2j", "g1j", "bvis", "g2j")}
    for (d1, bl, d2), (g1, bv, g2) in itertools.product(PRESENCE, PRESENCE):
        key = "%s_out_%d%d%d_%d%d%d" % (cname, d1, bl, d2, g1, bv, g2)
        got = b200.rime.predict_vis(
            ti, a1, a2, arrs["a1j"] if d1 else None, arrs["blj"] if bl else None,
            arrs["a2j"] if d2 else None, arrs["g1j"] if g1 else None,
            arrs["bvis"] if bv else None, arrs["g2j"] if g2 else None)
        assert_c128_close(got, g[key])
    a64 = {k: v.astype(np.complex64) for k, v in arrs.items()}
    got = b200.rime.predict_vis(ti.astype(np.int16), a1.astype(np.int16), a2.astype(np.int16),
                                a64["a1j"], a64["blj"], a64["a2j"], a64["g1j"], a64["bvis"], a64["g2j"])
    assert_c64_close(got, g["%s_out_c64" % cname])
    got = b200.rime.apply_gains(ti, a1, a2, arrs["g1j"], arrs["bvis"], arrs["g2j"])
    assert_c128_close(got, oracle.apply_gains(ti, a1, a2, arrs["g1j"], arrs["bvis"], arrs["g2j"]))


def test_predict_vis_vs_oracle_larger(b200, oracle):
    rng = np.random.default_rng(21)
    na, ntime, nchan, nsrc = 7, 5, 19, 13
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size) + 100
    nrow = ant1.size

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    for corr in ((2, 2), (2,), (1,), (3,)):
        dde = rc((nsrc, ntime, na, nchan) + corr)
        coh = rc((nsrc, nrow, nchan) + corr)
        die = rc((ntime, na, nchan) + corr)
        bvis = rc((nrow, nchan) + corr)
        got = b200.rime.predict_vis(ti, ant1, ant2, dde, coh, dde, die, bvis, die)
        assert_c128_close(got, oracle.predict_vis(ti, ant1, ant2, dde, coh, dde, die, bvis, die))


# ----------------------------------------------------------------------------- beam
def test_beam_cube_dde_golden(b200, golden):
    g = golden("beam_cube_dde")
    args = (g["ext"], g["beam_freq_map"], g["lm"], g["pa"], g["perr"], g["ascale"], g["freq"])
    for cname in ("c22", "c4", "c2", "c1"):
        assert_c128_close(b200.rime.beam_cube_dde(g["beam_" + cname], *args), g["dde_" + cname])
    got = b200.rime.beam_cube_dde(g["beam_c22"].astype(np.complex64), *args)
    assert_c64_close(got, g["dde_c22_c64"])
    ka = b200.rime.beam_cube_dde(
        g["ka_beam"], np.asarray([[-1.0, 1.0], [-1.0, 1.0]]), np.asarray([0.0, 1.0]),
        np.asarray([[0.1, 0.1]]), np.zeros((1, 1)), np.zeros((1, 1, 1, 2)), np.ones((1, 1, 2)),
        np.asarray([0.3]))
    np.testing.assert_array_almost_equal([[[[[0.470255 + 0.4786j]]]]], ka)  # test_fast_beams.py:126
    assert_c128_close(ka, g["ka_dde"])
    np.testing.assert_array_equal(b200.rime.freq_grid_interp(g["freq"], g["beam_freq_map"]), g["freq_data"])


def test_beam_cube_dde_vs_oracle(b200, oracle):
    rng = np.random.default_rng(31)
    lw, mh, nud = 33, 31, 7
    nsrc, ntime, nant, nchan = 23, 3, 5, 37
    bfm = np.linspace(0.856e9, 1.712e9, nud)
    freq = np.linspace(0.8e9, 1.8e9, nchan)
    ext = np.array([[-0.05, 0.05], [-0.04, 0.06]])
    lm = rng.uniform(-0.06, 0.06, (nsrc, 2))
    pa = rng.uniform(-np.pi, np.pi, (ntime, nant))
    perr = rng.uniform(-0.005, 0.005, (ntime, nant, nchan, 2))
    ascale = rng.uniform(0.9, 1.1, (nant, nchan, 2))
    for corr in ((2, 2), (3,)):
        beam = rng.standard_normal((lw, mh, nud) + corr) + 1j * rng.standard_normal((lw, mh, nud) + corr)
        got = b200.rime.beam_cube_dde(beam, ext, bfm, lm, pa, perr, ascale, freq)
        assert_c128_close(got, oracle.beam_cube_dde(beam, ext, bfm, lm, pa, perr, ascale, freq))


def test_beam_cube_dde_planes_kernel(b200, oracle, monkeypatch):
    """The plane-interpolated kernel (four spatial corners reduced per frequency plane once per
    (source, time, antenna) row, then two planes per channel) re-associates the reference's sum over
    the 8 corners: it must agree with the element-per-thread kernel to a few ulp and with the oracle
    at the project gate.  Cases: channel-constant pointing errors (plane path), per-channel pointing
    errors / scaling (rows fall back to the element path: bit-identical), out-of-band channels (lm
    scaled per channel: element path inside a plane row), complex64 beams, 1 / 2 / 3 / 4
    correlations, and the feed-rotation epilogue."""
    rng = np.random.default_rng(77)
    lw, mh, nud = 41, 37, 9
    nsrc, ntime, nant = 13, 2, 5
    bfm = np.linspace(0.856e9, 1.712e9, nud)
    ext = np.array([[-0.05, 0.05], [-0.04, 0.06]])
    lm = rng.uniform(-0.045, 0.045, (nsrc, 2))
    pa = rng.uniform(-np.pi, np.pi, (ntime, nant))

    def both(fn, *args):
        got = fn(*args)
        monkeypatch.setenv("AFR_BEAM_PLANES", "0")
        elem = fn(*args)
        monkeypatch.delenv("AFR_BEAM_PLANES")
        return got, elem

    for nchan, f_lo, f_hi in ((256, 0.856e9, 1.712e9), (77, 0.80e9, 1.80e9)):
        freq = np.linspace(f_lo, f_hi, nchan)
        perr_c = np.ascontiguousarray(
            np.broadcast_to(rng.uniform(-0.004, 0.004, (ntime, nant, 1, 2)), (ntime, nant, nchan, 2)))
        perr_v = rng.uniform(-0.004, 0.004, (ntime, nant, nchan, 2))
        perr_v[1] = perr_c[1]  # timestep 1 keeps constant errors: plane rows and element rows in one call
        asc_c = np.ones((nant, nchan, 2)) * rng.uniform(0.95, 1.05, (nant, 1, 2))
        asc_v = rng.uniform(0.95, 1.05, (nant, nchan, 2))
        for corr, dt in (((2, 2), np.complex128), ((2, 2), np.complex64), ((3,), np.complex128), ((1,), np.complex128)):
            beam = (rng.standard_normal((lw, mh, nud) + corr) + 1j * rng.standard_normal((lw, mh, nud) + corr)).astype(dt)
            for perr, asc in ((perr_c, asc_c), (perr_v, asc_c), (perr_c, asc_v)):
                args = (beam, ext, bfm, lm, pa, perr, asc, freq)
                got, elem = both(b200.rime.beam_cube_dde, *args)
                if asc is asc_v:
                    np.testing.assert_array_equal(got, elem)  # no plane rows at all
                elif dt == np.complex128:
                    assert_c128_close(got, elem, rtol=1e-12)
                    assert_c128_close(got, oracle.beam_cube_dde(*args))
                else:
                    assert_c64_close(got, elem, tol=2e-6)
        beam = rng.standard_normal((lw, mh, nud, 2, 2)) + 1j * rng.standard_normal((lw, mh, nud, 2, 2))
        for feed in ("linear", "circular"):
            args = (beam, ext, bfm, lm, pa, perr_c, asc_c, freq, b200.rime.feed_rotation(pa, feed))
            got, elem = both(b200.rime.beam_cube_dde_rotated, *args)
            assert_c128_close(got, elem, rtol=1e-12)


# ----------------------------------------------------------------------------- fused
def test_fused_predict_golden(b200, golden):
    g = golden("fused_predict")
    ti, a1, a2 = g["time_idx"], g["ant1"], g["ant2"]
    lm, uvw, freq = g["lm"], g["uvw"], g["freq"]
    for conv in ("fourier", "casa"):
        for cname in ("c22", "c2", "c1"):
            key = "%s_%s" % (conv, cname)
            br, die, bvis = g["bright_" + key], g["die_" + key], g["bvis_" + key]
            got = b200.rime.fused_predict_vis(lm, uvw, freq, br, ti, a1, a2, convention=conv)
            assert_c128_close(got, g["point_" + key])
            got = b200.rime.fused_predict_vis(lm, uvw, freq, br, ti, a1, a2, die1_jones=die,
                                              base_vis=bvis, die2_jones=die, convention=conv)
            assert_c128_close(got, g["point_die_" + key])
            dde = g["dde_beam_c22"] if cname == "c22" else g["dde_" + key]
            got = b200.rime.fused_predict_vis(lm, uvw, freq, br, ti, a1, a2, dde, dde, die, bvis,
                                              die, convention=conv)
            assert_c128_close(got, g["full_" + key])
    # beam_cube_dde -> fused predict end to end on the GPU (no host round trip for the DDEs)
    import torch

    nchan = freq.shape[0]
    ntime, na = g["pa"].shape
    dde = b200.rime.beam_cube_dde(
        torch.from_numpy(g["beam"]).cuda(), g["ext"], g["bfm"], lm, g["pa"],
        np.zeros((ntime, na, nchan, 2)), np.ones((na, nchan, 2)), freq)
    key = "fourier_c22"
    got = b200.rime.fused_predict_vis(lm, uvw, freq, g["bright_" + key], ti, a1, a2, dde, dde,
                                      g["die_" + key], g["bvis_" + key], g["die_" + key])
    assert_c128_close(got.cpu().numpy(), g["full_" + key])


def test_fused_predict_vs_oracle(b200, oracle):
    """MeerKAT-like slice: 12 antennas, 300 channels (recurrence across several runs/segments),
    non-uniform channels (exact path), complex64 chain."""
    rng = np.random.default_rng(41)
    na, ntime, nsrc = 12, 3, 29
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size) + 7
    nrow = ant1.size
    uvw = rng.standard_normal((nrow, 3)) * 2500.0
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    for nchan, uniform in ((300, True), (64, True), (40, False)):
        freq = np.linspace(0.856e9, 1.712e9, nchan)
        if not uniform:
            freq = np.sort(rng.uniform(0.856e9, 1.712e9, nchan))
        for corr in ((2, 2), (2,)):
            bright = rc((nsrc, nchan) + corr)
            dde = 1.0 + 0.2 * rc((nsrc, ntime, na, nchan) + corr)
            die = 1.0 + 0.1 * rc((ntime, na, nchan) + corr)
            bvis = rc((nrow, nchan) + corr)
            ref = oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2)
            assert_c128_close(b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2), ref)
            ref = oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die)
            got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die)
            assert_c128_close(got, ref)
            c64 = np.complex64
            args64 = (bright.astype(c64), ti, ant1, ant2, dde.astype(c64), dde.astype(c64), die.astype(c64),
                      bvis.astype(c64), die.astype(c64))
            got = b200.rime.fused_predict_vis(lm, uvw, freq, *args64, dtype=c64)
            assert got.dtype == c64
            assert rel_l2(got.astype(np.complex128), ref) < 1e-5
            # dtype promotion of the reference chain: float64 coordinates make K complex128
            # (rime/phase.py:26) and with it the visibilities; all-float32 coordinates keep complex64
            assert b200.rime.fused_predict_vis(lm, uvw, freq, *args64).dtype == np.complex128
            f32 = np.float32
            assert b200.rime.fused_predict_vis(lm.astype(f32), (uvw * 1e-3).astype(f32), freq.astype(f32),
                                               *args64).dtype == c64


def test_fused_point_predict_tensor_pipe(b200, oracle, monkeypatch):
    """2x2 complex brightness without DDEs on equispaced channels runs the phasor-stream kernel with
    DMMA consumers (afr_last_dft_path bit 5).  Ragged shapes: rows that do not fill the 16-row tile,
    odd source counts (the k-step holds two sources), source counts that are not a multiple of the
    8-item tile, channel counts from one partial run to several CTAs of 128, long baselines; each
    against the oracle, and the scalar schedule (AFR_POINT_MMA=0) against the same oracle values."""
    from codex_africanus_b200 import _lib
    rng = np.random.default_rng(77)

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    cases = [(37, 1, 6, 3e3), (100, 7, 8, 3e3), (53, 19, 41, 3e3), (210, 33, 137, 3e3), (64, 64, 300, 1.5e5), (9, 5, 1, 3e3),
             (17, 250, 2, 3e3)]
    for nrow, nsrc, nchan, scale in cases:
        uvw = rng.standard_normal((nrow, 3)) * scale
        lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
        freq = np.linspace(0.856e9, 1.712e9, nchan)
        bright = rc((nsrc, nchan, 2, 2))
        ti = np.zeros(nrow, np.int32)
        a1 = np.zeros(nrow, np.int32)
        a2 = np.ones(nrow, np.int32)
        die = 1.0 + 0.1 * rc((1, 2, nchan, 2, 2))
        bvis = rc((nrow, nchan, 2, 2))
        ref = oracle.fused_predict(lm, uvw, freq, bright, ti, a1, a2)
        got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, a1, a2)
        assert _lib.lib().afr_last_dft_path() & 32, (nrow, nsrc, nchan)
        assert_c128_close(got, ref)
        ref2 = oracle.fused_predict(lm, uvw, freq, bright, ti, a1, a2, None, None, die, bvis, die)
        assert_c128_close(b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, a1, a2, None, None, die, bvis, die),
                          ref2)
        monkeypatch.setenv("AFR_POINT_MMA", "0")
        got0 = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, a1, a2)
        assert not (_lib.lib().afr_last_dft_path() & 32)
        monkeypatch.delenv("AFR_POINT_MMA")
        assert_c128_close(got0, ref)
        # the schedules anchor their recurrences at different channels (runs of 8 vs 4, 128 vs 64 channels
        # per CTA); at 150 km the anchor phases themselves carry ulp(1e5 rad) = 1.5e-11
        assert np.abs(got - got0).max() <= 1e-10 * np.abs(ref).max()
    # non-equispaced channels and diagonal brightness keep the scalar consumers
    freq = np.sort(rng.uniform(0.856e9, 1.712e9, 24))
    uvw = rng.standard_normal((40, 3)) * 3e3
    lm = rng.uniform(-0.02, 0.02, (9, 2))
    ti = np.zeros(40, np.int32)
    bright = rc((9, 24, 2, 2))
    got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ti, ti + 1)
    assert not (_lib.lib().afr_last_dft_path() & 32)
    assert_c128_close(got, oracle.fused_predict(lm, uvw, freq, bright, ti, ti, ti + 1))


def test_fused_dde_diagonal_any_ncorr(b200, oracle):
    """Element-wise Jones with 3, 5 and 7 correlations and DDEs (predict_vis accepts any count; the
    fused kernels have 1 / 2 / 4): blocks of correlations, concatenated."""
    rng = np.random.default_rng(53)
    na, ntime, nsrc, nchan = 5, 2, 6, 9
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size)
    uvw = rng.standard_normal((ti.size, 3)) * 900.0
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    freq = np.linspace(0.9e9, 1.1e9, nchan)

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    for ncorr in (3, 5, 7):
        bright = rc((nsrc, nchan, ncorr))
        dde = 1.0 + 0.2 * rc((nsrc, ntime, na, nchan, ncorr))
        die = 1.0 + 0.1 * rc((ntime, na, nchan, ncorr))
        bvis = rc((ti.size, nchan, ncorr))
        ref = oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die)
        got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die)
        assert got.shape == ref.shape
        assert_c128_close(got, ref)


def test_fused_dde_layout_adapter(b200, oracle, monkeypatch):
    """Diagonal Jones ((2,), (1,)), complex64 chains and rows that are not ordered by time have no fast
    DDE kernel of their own: above a size threshold they are re-expressed as the time-ordered complex128
    2x2 problem (rime/fused.py) and must land on the GEMM / warp-specialised kernels with the oracle's
    values; below it the gather kernel still serves them (covered by the tests above)."""
    from codex_africanus_b200 import _lib
    from codex_africanus_b200.rime import fused as fused_mod
    monkeypatch.setattr(fused_mod, "_ADAPTER_MIN_TERMS", 0)
    monkeypatch.setattr(fused_mod, "_ADAPTER_CHUNK_BYTES", 1 << 20)  # several source chunks
    rng = np.random.default_rng(97)
    na, ntime, nsrc, nchan = 10, 3, 23, 20
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size) + 3
    pos = rng.standard_normal((ntime, na, 3)) * 1500.0
    uvw = (pos[:, a1] - pos[:, a2]).reshape(-1, 3)
    nrow = ant1.size
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    freq = np.linspace(0.856e9, 1.712e9, nchan)

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    import torch
    fast = (6, 2, 3)  # AFR_PATH_DDE_MMA_ANT, AFR_PATH_DDE_WS_ANT, AFR_PATH_DDE_WS_ROW
    for corr in ((2,), (1,), (2, 2)):
        bright = rc((nsrc, nchan) + corr)
        dde = 1.0 + 0.2 * rc((nsrc, ntime, na, nchan) + corr)
        dde_b = 1.0 + 0.2 * rc((nsrc, ntime, na, nchan) + corr)
        die = 1.0 + 0.1 * rc((ntime, na, nchan) + corr)
        bvis = rc((nrow, nchan) + corr)
        if corr != (2, 2):
            for d2 in (dde, dde_b):
                ref = oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, d2, die, bvis, die)
                got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, d2, die, bvis, die)
                assert _lib.lib().afr_last_fused_path() in fast
                assert got.shape == ref.shape
                assert_c128_close(got, ref)
            ref = oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde)
            assert_c128_close(b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde), ref)
        # complex64 chain
        c64 = np.complex64
        ref = oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die)
        got = b200.rime.fused_predict_vis(lm, uvw, freq, bright.astype(c64), ti, ant1, ant2, dde.astype(c64),
                                          dde.astype(c64), die.astype(c64), bvis.astype(c64), die.astype(c64),
                                          dtype=c64)
        assert got.dtype == c64 and _lib.lib().afr_last_fused_path() in fast
        assert rel_l2(got.astype(np.complex128), ref) < 1e-5
        # rows in a random order (torch inputs: the result stays on the device)
        perm = rng.permutation(nrow)
        got = b200.rime.fused_predict_vis(*(torch.from_numpy(np.ascontiguousarray(a)).cuda() for a in (
            lm, uvw[perm], freq, bright, ti[perm], ant1[perm], ant2[perm], dde, dde, die, bvis[perm], die)))
        assert got.is_cuda and _lib.lib().afr_last_fused_path() in fast
        assert_c128_close(got.cpu().numpy(), ref[perm])


def test_fused_dde_ws_modes_vs_oracle(b200, oracle, monkeypatch):
    """The DDE kernels: antenna-phasor mode (baseline uvw that are differences of per-antenna
    coordinates, as in a Measurement Set) as a DMMA GEMM (afr_rime_mma.cu) and as the scalar
    warp-specialised kernel (AFR_DDE_MMA=0), per-row phasor mode (forced, and
    chosen automatically for uvw that are not antenna-consistent), and the previous tiled
    kernel must all match the oracle.  Ragged timesteps, channel tail, E1 != E2."""
    rng = np.random.default_rng(2024)
    na, ntime, nsrc = 9, 4, 23
    a1, a2 = np.triu_indices(na, 1)
    # ragged timesteps; the baselines to antenna 0 (the reference of the decomposition) stay
    opt = np.flatnonzero(a1 != 0)
    keep = [np.sort(np.concatenate([np.flatnonzero(a1 == 0),
                                    rng.choice(opt, size=opt.size - 3 * t, replace=False)]))
            for t in range(ntime)]
    ant1 = np.concatenate([a1[k] for k in keep])
    ant2 = np.concatenate([a2[k] for k in keep])
    ti = np.concatenate([np.full(k.size, t + 3) for t, k in enumerate(keep)])
    antpos = rng.standard_normal((ntime, na, 3)) * 1500.0
    uvw_ant = antpos[ti - 3, ant1] - antpos[ti - 3, ant2]
    uvw_rnd = rng.standard_normal(uvw_ant.shape) * 1500.0
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    for nchan, uniform in ((37, True), (16, False)):
        freq = np.linspace(0.856e9, 1.712e9, nchan)
        if not uniform:
            freq = np.sort(rng.uniform(0.856e9, 1.712e9, nchan))
        bright = rc((nsrc, nchan, 2, 2))
        dde = 1.0 + 0.2 * rc((nsrc, ntime, na, nchan, 2, 2))
        dde_b = 1.0 + 0.2 * rc((nsrc, ntime, na, nchan, 2, 2))
        die = 1.0 + 0.1 * rc((ntime, na, nchan, 2, 2))
        for uvw in (uvw_ant, uvw_rnd):
            for d1, d2 in ((dde, dde), (dde, dde_b)):
                ref = oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, d1, d2, die, None, die)
                for env in ({}, {"AFR_DDE_MMA": "0"}, {"AFR_DDE_ANT": "0"}, {"AFR_DDE_WS": "0"}):
                    for k, v in env.items():
                        monkeypatch.setenv(k, v)
                    got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, d1, d2,
                                                      die, None, die)
                    for k in env:
                        monkeypatch.delenv(k)
                    assert_c128_close(got, ref)
                    # the path that ran: antenna mode only for antenna-consistent uvw
                    from codex_africanus_b200 import _lib
                    path = _lib.lib().afr_last_fused_path()
                    if "AFR_DDE_WS" in env:
                        assert path == 4
                    elif "AFR_DDE_ANT" in env or uvw is uvw_rnd:
                        assert path == 3
                    elif "AFR_DDE_MMA" in env:
                        assert path == 2
                    else:
                        assert path == 6  # source sum as a GEMM on the FP64 tensor pipe



# ----------------------------------------------------------------------------- wsclean_predict
def test_wsclean_predict_golden(b200, golden):
    """africanus.rime.wsclean_predict / model.wsclean.spectra against the reference's outputs:
    its own test inputs (seed 42) and a MeerKAT-like POINT/GAUSSIAN mix, scalar and per-source
    log_poly, non-uniform channels."""
    g = golden("wsclean")
    for pre in ("t_", "m_"):
        args = [g[pre + k] for k in ("flux", "coeffs", "log_poly", "ref_freq", "freq")]
        assert_c128_close(b200.rime.wsclean_spectra(*args), g[pre + "spectra"])
        got = b200.rime.wsclean_predict(g[pre + "uvw"], g[pre + "lm"], g[pre + "source_type"], g[pre + "flux"],
                                        g[pre + "coeffs"], g[pre + "log_poly"], g[pre + "ref_freq"],
                                        g[pre + "gauss_shape"], g[pre + "freq"])
        assert got.shape == g[pre + "vis"].shape and got.dtype == np.complex128
        assert_c128_close(got, g[pre + "vis"])
    margs = [g["m_" + k] for k in ("uvw", "lm", "source_type", "flux", "coeffs")]
    for lp, key in ((True, "m_vis_logpoly_true"), (False, "m_vis_logpoly_false")):
        got = b200.rime.wsclean_predict(*margs, lp, g["m_ref_freq"], g["m_gauss_shape"], g["m_freq"])
        assert_c128_close(got, g[key])
    got = b200.rime.wsclean_predict(*margs, g["m_log_poly"], g["m_ref_freq"], g["m_gauss_shape"],
                                    g["m_freq_nu"])
    assert_c128_close(got, g["m_vis_nu"])
    with pytest.raises(ValueError):
        b200.rime.wsclean_predict(*margs[:2], np.array(["DISK"] * g["m_lm"].shape[0]), *margs[3:],
                                  True, g["m_ref_freq"], g["m_gauss_shape"], g["m_freq"])


def test_wsclean_predict_vs_oracle(b200, oracle):
    """larger seeded case: 500 sources (30 % Gaussian), 700 rows, 96 channels; torch inputs;
    only POINT / only GAUSSIAN; complex64 result for float32 inputs."""
    import torch

    rng = np.random.default_rng(5)
    nsrc, nrow, nchan = 500, 700, 96
    st = np.where(rng.random(nsrc) < 0.7, "POINT", "GAUSSIAN")
    arcsec = np.pi / 180 / 3600
    gs = np.stack([rng.uniform(10, 90, nsrc) * arcsec, rng.uniform(3, 10, nsrc) * arcsec,
                   rng.uniform(0, np.pi, nsrc)], axis=1)
    uvw = rng.standard_normal((nrow, 3)) * 2000.0
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    flux = np.abs(rng.standard_normal(nsrc)) + 0.1
    coeffs = rng.standard_normal((nsrc, 2)) * 0.2
    lp = rng.random(nsrc) < 0.5
    freq = np.linspace(0.856e9, 1.712e9, nchan)
    rf = np.full(nsrc, 1.284e9)
    ref = oracle.wsclean_predict(uvw, lm, st, flux, coeffs, lp, rf, gs, freq)
    assert_c128_close(b200.rime.wsclean_predict(uvw, lm, st, flux, coeffs, lp, rf, gs, freq), ref)
    dev = torch.device("cuda:0")
    T = lambda a: torch.from_numpy(a).to(dev)  # noqa: E731
    got = b200.rime.wsclean_predict(T(uvw), T(lm), st, T(flux), T(coeffs), lp, T(rf), T(gs), T(freq))
    assert isinstance(got, torch.Tensor)
    assert_c128_close(got.cpu().numpy(), ref)
    for kind in ("POINT", "GAUSSIAN"):
        st1 = np.full(nsrc, kind)
        assert_c128_close(b200.rime.wsclean_predict(uvw, lm, st1, flux, coeffs, lp, rf, gs, freq),
                          oracle.wsclean_predict(uvw, lm, st1, flux, coeffs, lp, rf, gs, freq))
    # long baselines x big Gaussians: the taper underflows for part of the (row, source) pairs,
    # which switches those lanes from the recurrences to per-term exp (and decreasing channels)
    uvw_l = uvw * 60.0
    for fr in (freq, freq[::-1].copy()):
        assert_c128_close(b200.rime.wsclean_predict(uvw_l, lm, st, flux, coeffs, lp, rf, gs, fr),
                          oracle.wsclean_predict(uvw_l, lm, st, flux, coeffs, lp, rf, gs, fr))
    f32 = np.float32
    got = b200.rime.wsclean_predict(*(a.astype(f32) for a in (uvw, lm)), st, flux.astype(f32),
                                    coeffs.astype(f32), lp, rf.astype(f32), gs, freq.astype(f32))
    assert got.dtype == np.complex64



def test_fused_predict_vis_beam_chunks(b200, oracle):
    """Beam-interpolated DDEs reduced chunk by chunk (chunk granularity) equal the
    reference composition beam_cube_dde -> phase_delay -> einsum -> predict_vis with the whole
    DDE array materialised (oracle), for several chunk sizes incl. one that does not divide nsrc."""
    rng = np.random.default_rng(314)
    na, ntime, nsrc, nchan = 7, 3, 13, 24
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size)
    nrow = ti.size
    antpos = rng.standard_normal((ntime, na, 3)) * 1200.0
    uvw = antpos[ti, ant1] - antpos[ti, ant2]
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    freq = np.linspace(0.856e9, 1.712e9, nchan)

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    beam = rc((11, 11, 6, 2, 2))
    ext = np.array([[-0.03, 0.03], [-0.03, 0.03]])
    bfm = np.linspace(0.8e9, 1.8e9, 6)
    pa = rng.uniform(-1, 1, (ntime, na))
    pe = rng.uniform(-1e-3, 1e-3, (ntime, na, nchan, 2))
    asc = rng.uniform(0.9, 1.1, (na, nchan, 2))
    bright = rc((nsrc, nchan, 2, 2))
    die = 1.0 + 0.1 * rc((ntime, na, nchan, 2, 2))
    bvis = rc((nrow, nchan, 2, 2))
    dde = oracle.beam_cube_dde(beam, ext, bfm, lm, pa, pe, asc, freq)
    ref = oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die)
    for chunk in (None, 1, 5, 13):
        got = b200.rime.fused_predict_vis_beam(lm, uvw, freq, bright, ti, ant1, ant2, beam, ext, bfm, pa,
                                               pe, asc, die, bvis, die, source_chunk=chunk)
        assert_c128_close(got, ref)
    got = b200.rime.fused_predict_vis_beam(lm, uvw, freq, bright, ti, ant1, ant2, beam, ext, bfm, pa, pe,
                                           asc, source_chunk=4)
    assert_c128_close(got, oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde))



def test_fused_dde_mma_ragged_source_counts(b200, oracle):
    """Source counts that are not a multiple of the GEMM kernel's four sources per stage: the slots of
    the missing sources must contribute exact zeros whatever shared memory held before (found by
    tools/fuzz_fused.py: the brightness slot of a missing source is never copied, and 0 x stale NaN
    poisoned the sum).  Shared memory is first filled with NaN by a DFT over a NaN image."""
    from codex_africanus_b200 import _lib
    rng = np.random.default_rng(99)
    na, ntime, nchan = 12, 2, 5
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size)
    antpos = rng.standard_normal((ntime, na, 3)) * 1500.0
    uvw = antpos[ti, ant1] - antpos[ti, ant2]
    freq = np.linspace(0.856e9, 1.712e9, nchan)

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    nan_img = np.full((64, 256, 4), np.nan)
    for nsrc in (1, 2, 3, 5, 6, 7, 19):
        lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
        bright = rc((nsrc, nchan, 2, 2))
        dde = 1.0 + 0.2 * rc((nsrc, ntime, na, nchan, 2, 2))
        dde_b = 1.0 + 0.2 * rc((nsrc, ntime, na, nchan, 2, 2))
        for d2 in (dde, dde_b):
            b200.dft.im_to_vis(nan_img, rng.standard_normal((148 * 64, 3)), rng.uniform(-.01, .01, (64, 2)),
                               np.linspace(1e9, 2e9, 256))
            got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, d2)
            assert _lib.lib().afr_last_fused_path() == 6
            assert np.all(np.isfinite(got))
            assert_c128_close(got, oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, d2))


def test_fused_dde_ws_many_antennas(b200, oracle, monkeypatch):
    """140 antennas (9730 baselines): the three-stage antenna tile of the 512-row x 4-channel CTA
    does not fit in shared memory; antenna mode takes the 2048-row x 1-channel tile, random uvw the
    per-row mode with 512 rows x 2 channels per CTA (1 channel from 259 antennas on)."""
    from codex_africanus_b200 import _lib
    rng = np.random.default_rng(7)
    na, nsrc, nchan = 140, 4, 3
    ant1, ant2 = np.triu_indices(na, 1)
    ti = np.zeros(ant1.size, np.int64)
    antpos = rng.standard_normal((na, 3)) * 3000.0
    uvw = antpos[ant1] - antpos[ant2]
    lm = rng.uniform(-0.01, 0.01, (nsrc, 2))
    freq = np.linspace(0.95e9, 1.05e9, nchan)

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    bright = rc((nsrc, nchan, 2, 2))
    dde = 1.0 + 0.2 * rc((nsrc, 1, na, nchan, 2, 2))
    ref = oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde)
    got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde)
    assert _lib.lib().afr_last_fused_path() == 6  # GEMM path: 18 x 18 tiles in 3 x 3 panels of 6
    assert_c128_close(got, ref)
    monkeypatch.setenv("AFR_DDE_MMA", "0")
    got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde)
    monkeypatch.delenv("AFR_DDE_MMA")
    assert _lib.lib().afr_last_fused_path() == 2
    assert_c128_close(got, ref)
    uvw_r = rng.standard_normal(uvw.shape) * 3000.0
    ref = oracle.fused_predict(lm, uvw_r, freq, bright, ti, ant1, ant2, dde, dde)
    got = b200.rime.fused_predict_vis(lm, uvw_r, freq, bright, ti, ant1, ant2, dde, dde)
    assert _lib.lib().afr_last_fused_path() == 3
    assert_c128_close(got, ref)



# ----------------------------------------------------------------------------- cross-kernel
def test_fused_equals_unfused_composition_on_gpu(b200):
    """Size-independent property at a size the CPU oracle cannot reach: the fused kernel must
    equal the reference's own three-step composition (rime/examples/predict.py:107-134,490,
    522-527) built from the OTHER kernels of this library: phase_delay -> einsum -> predict_vis.
    64 antennas x 6 times (12096 rows), 128 channels, 40 sources, 2x2, DIE + DDE."""
    import torch

    rng = np.random.default_rng(123)
    na, ntime, nchan, nsrc = 64, 6, 128, 40
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size)
    nrow = ti.size
    dev = torch.device("cuda:0")

    def T(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def rc(shape, scale=1.0):
        return scale * (rng.standard_normal(shape) + 1j * rng.standard_normal(shape))

    uvw = T(rng.standard_normal((nrow, 3)) * 3000.0)
    lm = T(rng.uniform(-0.02, 0.02, (nsrc, 2)))
    freq = T(np.linspace(0.856e9, 1.712e9, nchan))
    bright = T(rc((nsrc, nchan, 2, 2)))
    dde = T(1.0 + rc((nsrc, ntime, na, nchan, 2, 2), 0.2))
    dde_b = T(1.0 + rc((nsrc, ntime, na, nchan, 2, 2), 0.2))
    die = T(1.0 + rc((ntime, na, nchan, 2, 2), 0.1))
    bvis = T(rc((nrow, nchan, 2, 2)))
    tiT, a1T, a2T = T(ti), T(ant1), T(ant2)
    K = b200.rime.phase_delay(lm, uvw, freq)
    coh = torch.einsum("srf,sfij->srfij", K, bright).contiguous()
    for d1, d2 in ((None, None), (dde, dde), (dde, dde_b)):
        ref = b200.rime.predict_vis(tiT, a1T, a2T, d1, coh, d2, die, bvis, die)
        got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, tiT, a1T, a2T, d1, d2, die, bvis, die)
        assert_c128_close(got.cpu().numpy(), ref.cpu().numpy())
    # Measurement-Set-like uvw (differences of per-antenna coordinates, 8 km array): the phasor
    # folds into the antenna Jones and the source sum runs as a GEMM per (time, channel): one
    # pass of 36 tiles for 64 antennas
    from codex_africanus_b200 import _lib
    antpos = rng.standard_normal((ntime, na, 3)) * 2500.0
    uvw_a = T(antpos[ti, ant1] - antpos[ti, ant2])
    Ka = b200.rime.phase_delay(lm, uvw_a, freq)
    coh_a = torch.einsum("srf,sfij->srfij", Ka, bright).contiguous()
    for d1, d2 in ((dde, dde), (dde, dde_b)):
        ref = b200.rime.predict_vis(tiT, a1T, a2T, d1, coh_a, d2, die, bvis, die)
        got = b200.rime.fused_predict_vis(lm, uvw_a, freq, bright, tiT, a1T, a2T, d1, d2, die, bvis, die)
        assert _lib.lib().afr_last_fused_path() == 6
        assert_c128_close(got.cpu().numpy(), ref.cpu().numpy())
    # rows not ordered by time take the gather kernel; same answer
    perm = torch.from_numpy(rng.permutation(nrow)).to(dev)
    ref = b200.rime.predict_vis(tiT, a1T, a2T, dde, coh, dde_b, die, bvis, die)
    got = b200.rime.fused_predict_vis(lm, uvw[perm], freq, bright, tiT[perm], a1T[perm], a2T[perm],
                                      dde, dde_b, die, bvis[perm], die)
    assert_c128_close(got.cpu().numpy(), ref[perm].cpu().numpy())


def test_phasor_stream_kernel_variants_agree(b200, monkeypatch):
    """The warp-specialised and the single-role phasor-stream kernels are two schedules of
    the same arithmetic: forcing either one (AFR_WS) must give the same visibilities."""
    rng = np.random.default_rng(77)
    nsrc, nrow, nchan = 300, 1000, 96
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    uvw = rng.standard_normal((nrow, 3)) * 3000.0
    freq = np.linspace(0.856e9, 1.712e9, nchan)
    for ncorr in (1, 2, 4):
        image = rng.standard_normal((nsrc, nchan, ncorr)) + 1j * rng.standard_normal((nsrc, nchan, ncorr))
        out = {}
        for ws in ("0", "1"):
            monkeypatch.setenv("AFR_WS", ws)
            out[ws] = b200.dft.im_to_vis(image, uvw, lm, freq)
        monkeypatch.delenv("AFR_WS")
        assert_c128_close(out["1"], out["0"], rtol=1e-12)
        # adjoint, flagged (TMA-staged tile edited by the producers) and unflagged
        vis = rng.standard_normal((nrow, nchan, ncorr)) + 1j * rng.standard_normal((nrow, nchan, ncorr))
        for flags in (rng.random((nrow, nchan, ncorr)) < 0.05, np.zeros((nrow, nchan, ncorr), bool)):
            out = {}
            for ws in ("0", "1"):
                monkeypatch.setenv("AFR_WS", ws)
                out[ws] = b200.dft.vis_to_im(vis, uvw, lm, freq, flags)
            monkeypatch.delenv("AFR_WS")
            assert_c128_close(out["1"], out["0"], rtol=1e-12)


def test_row_block_streaming_paths(b200, oracle, monkeypatch):
    """numpy inputs with a large output are computed in row blocks whose results stream back
    on a copy stream; force tiny blocks and compare with the oracle (rows are independent)."""
    import codex_africanus_b200.dft.kernels as dk
    import codex_africanus_b200.rime.fused as fu

    rng = np.random.default_rng(19)
    na, ntime, nchan, nsrc = 9, 11, 24, 7
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size) + 2
    nrow = ti.size
    uvw = rng.standard_normal((nrow, 3)) * 2000.0
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    freq = np.linspace(0.856e9, 1.712e9, nchan)

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    image = rng.standard_normal((nsrc, nchan, 2))
    bright = rc((nsrc, nchan, 2, 2))
    dde = 1.0 + 0.2 * rc((nsrc, ntime, na, nchan, 2, 2))
    die = 1.0 + 0.1 * rc((ntime, na, nchan, 2, 2))
    bvis = rc((nrow, nchan, 2, 2))
    monkeypatch.setattr(dk, "_ROW_BLOCK_BYTES", 1)   # -> minimum block of 4096 rows > nrow: 1 block
    assert_c128_close(b200.dft.im_to_vis(image, uvw, lm, freq), oracle.im_to_vis(image, uvw, lm, freq))
    monkeypatch.setattr(fu, "_ROW_BLOCK_BYTES", 100 * nchan * 4 * 16)  # 100-row blocks (min 1024)
    ref = oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die)
    got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die)
    assert_c128_close(got, ref)
    # genuinely multi-block: shrink the floor as well
    real_max = max
    monkeypatch.setattr(fu, "max", lambda a, b: real_max(37, b) if a == 1024 else real_max(a, b), raising=False)
    got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die)
    assert_c128_close(got, ref)
    got = b200.rime.fused_predict_vis(lm, uvw, freq, bright, ti, ant1, ant2)
    assert_c128_close(got, oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2))


# ----------------------------------------------------------------------------- brightness (8f-2)
_LIN = [["XX", "XY"], ["YX", "YY"]]
_CIRC = [["RR", "RL"], ["LR", "LL"]]
_IQUV = ["I", "Q", "U", "V"]


def test_spectral_model_golden(b200, golden):
    """africanus.model.spectral.spectral_model against the reference's outputs: every base, 1/2/6
    spectral-index components, per-polarisation base list, no polarisation axis, float32."""
    g = golden("brightness")
    st, rf, fr = g["stokes"], g["ref_freq"], g["freq"]
    sm = b200.model.spectral_model
    for n in (1, 2, 6):
        spi = g["spi%d" % n]
        for i, b in enumerate(("std", "log", "log10")):
            got = sm(st, spi, rf, fr, base=b)
            assert got.dtype == np.float64
            assert_c128_close(got, g["sm_%s_%d" % (b, n)])
            assert_c128_close(sm(st, spi, rf, fr, base=i), g["sm_int_%d" % n][i])
    assert_c128_close(sm(st, g["spi2"], rf, fr, base=["std", "log", "log10"]), g["sm_list"])
    assert_c128_close(sm(st[:, 0], g["spi2"][:, :, 0], rf, fr, base="log"), g["sm_nopol"])  # strided views
    f32 = np.float32
    got = sm(st.astype(f32), g["spi2"].astype(f32), rf.astype(f32), fr.astype(f32))
    assert got.dtype == np.float32 and got.shape == g["sm_f32"].shape
    assert rel_l2(got, g["sm_f32"]) < 1e-5
    assert sm(st[:0], g["spi2"][:0], rf[:0], fr).shape == (0, fr.shape[0], 4)


def test_convert_golden(b200, golden):
    """africanus.model.coherency.convert against the reference's outputs (dtype and shape incl.)
    and the reference test's known answers (model/coherency/tests/test_convert.py:69-110)."""
    import torch

    g = golden("brightness")
    sm, vis = g["sm_std_2"], g["vis"]
    cases = [
        ("b_linear", sm, _IQUV, _LIN, False), ("b_circular", sm, _IQUV, _CIRC, False),
        ("b_flat", sm, _IQUV, ["XX", "XY", "YX", "YY"], False),
        ("b_implicit", sm[..., :1], ["I"], ["XX", "XY", "YX", "YY"], True),
        ("b_diag", sm[..., :2], ["I", "Q"], ["XX", "YY"], False),
        ("s_linear", vis, _LIN, _IQUV, False), ("s_circular", vis, _CIRC, [["I", "Q"], ["U", "V"]], False),
        ("s_int", vis[..., 0, :], [9, 12], [1, 2], False),
        ("s_real", vis.real, _LIN, ["I", "Q"], False),
    ]
    for key, inp, isch, osch, implicit in cases:
        got = b200.model.convert(inp, isch, osch, implicit_stokes=implicit)
        assert_c128_close(got, g[key])
    got = b200.model.convert(sm.astype(np.float32), _IQUV, _LIN)
    assert_c64_close(got, g["b_f32"])
    got = b200.model.convert(torch.from_numpy(sm).cuda(), _IQUV, _LIN)
    assert isinstance(got, torch.Tensor) and got.is_cuda
    assert_c128_close(got.cpu().numpy(), g["b_linear"])
    I, Q, U, V = 1.0 + 1j, 2.0 + 2j, 3.0 + 3j, 4.0 + 4j  # noqa: E741
    inp = np.asarray([[I, Q, U, V]])
    assert np.all(b200.model.convert(inp, _IQUV, ["XX", "XY", "YX", "YY"])
                  == [[I + Q, U + V * 1j, U - V * 1j, I - Q]])
    assert np.all(b200.model.convert(inp, [1, 2, 3, 4], [5, 6, 7, 8])
                  == [[I + V, Q + U * 1j, Q - U * 1j, I - V]])
    # round trip Stokes -> correlations -> Stokes
    back = b200.model.convert(b200.model.convert(sm, _IQUV, _CIRC), _CIRC, _IQUV)
    assert_c128_close(back, sm.astype(np.complex128), rtol=1e-14)


def test_stokes_brightness_and_wsclean_powers(b200, golden):
    """The composed kernel equals convert(spectral_model(...)) of the reference for both feed
    types / every base; WSClean spectra with 6 coefficients (integer powers >= 4)."""
    g = golden("brightness")
    st, rf, fr = g["stokes"], g["ref_freq"], g["freq"]
    sb = b200.model.stokes_brightness
    assert_c128_close(sb(st, g["spi2"], rf, fr), g["b_linear"])
    assert_c128_close(sb(st, g["spi2"], rf, fr, corr_schema=_CIRC), g["b_circular"])
    assert_c128_close(sb(st, g["spi2"], rf, fr, corr_schema=["XX", "XY", "YX", "YY"]), g["b_flat"])
    assert_c128_close(sb(st[:, :1], g["spi2"][:, :, :1], rf, fr, stokes_schema=["I"],
                         corr_schema=["XX", "XY", "YX", "YY"], implicit_stokes=True), g["b_implicit"])
    got = sb(st, g["spi2"], rf, fr, dtype=np.complex64)
    assert_c64_close(got, g["b_linear"].astype(np.complex64))
    assert_c128_close(b200.rime.wsclean_spectra(st[:, 0], g["w_coeffs"], g["w_log_poly"], rf, fr),
                      g["w_spectra"])


def test_fused_predict_vis_stokes(b200, golden, oracle):
    """fused_predict_vis_stokes (brightness generated per source chunk on the device) equals the
    reference composition spectral_model -> convert -> phase_delay (x) brightness -> predict_vis
    (golden), for chunk sizes that do and do not divide nsrc; with DDEs and torch inputs vs the
    oracle; complex64 on request."""
    import torch

    g = golden("brightness")
    st, spi, rf, fr = g["stokes"], g["spi2"], g["ref_freq"], g["freq"]
    lm, uvw, ti, a1, a2 = g["p_lm"], g["p_uvw"], g["p_time_index"], g["p_ant1"], g["p_ant2"]
    die, bvis = g["p_die"], g["p_base_vis"]
    f = b200.rime.fused_predict_vis_stokes
    for feed, schema in (("linear", _LIN), ("circular", _CIRC)):
        for chunk in (None, 5, 23, 1):
            got = f(lm, uvw, fr, st, spi, rf, ti, a1, a2, None, None, die, bvis, die,
                    corr_schema=schema, source_chunk=chunk)
            assert_c128_close(got, g["p_" + feed])
    # DDEs (sliced per chunk), log-polynomial spectra, casa convention, int16 indices
    rng = np.random.default_rng(77)
    nsrc, nchan = lm.shape[0], fr.shape[0]
    ntime, na = die.shape[:2]
    dde = 1.0 + 0.2 * (rng.standard_normal((nsrc, ntime, na, nchan, 2, 2))
                       + 1j * rng.standard_normal((nsrc, ntime, na, nchan, 2, 2)))
    bright = oracle.convert(oracle.spectral_model(st, spi, rf, fr, base="log"), _IQUV, _LIN)
    ref = oracle.fused_predict(lm, uvw, fr, bright, ti, a1, a2, dde, dde, die, bvis, die, convention="casa")
    i16 = np.int16
    got = f(lm, uvw, fr, st, spi, rf, ti.astype(i16), a1.astype(i16), a2.astype(i16), dde, dde, die, bvis,
            die, convention="casa", base="log", source_chunk=7)
    assert_c128_close(got, ref)
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    got = f(T(lm), T(uvw), T(fr), T(st), T(spi), T(rf), T(ti), T(a1), T(a2), T(dde), T(dde), T(die),
            T(bvis), T(die), convention="casa", base="log", source_chunk=4)
    assert isinstance(got, torch.Tensor)
    assert_c128_close(got.cpu().numpy(), ref)
    # no Jones terms at all; complex64 brightness -> complex64 visibilities
    ref0 = oracle.fused_predict(lm, uvw, fr, g["b_linear"], ti, a1, a2)
    assert_c128_close(f(lm, uvw, fr, st, spi, rf, ti, a1, a2, source_chunk=6), ref0)
    got = f(lm, uvw, fr, st, spi, rf, ti, a1, a2, dtype=np.complex64)
    assert got.dtype == np.complex64
    assert rel_l2(got.astype(np.complex128), ref0) < 1e-5
    with pytest.raises(ValueError):
        f(lm, uvw, fr, st, spi, rf, ti, a1, a2, die1_jones=die)


# ----------------------------------------------------------------------------- feed rotation (8f-1)
def test_feed_rotation_and_rotated_beam_predict(b200, golden):
    """feed_rotation, the rotated DDE produced in the beam kernel's epilogue, and the chunked
    beam-interpolated predict with feed_type -- against the reference's outputs
    (rime/feeds.py:13-71, rime/examples/predict.py:469-472,522-527)."""
    import torch

    g = golden("feeds")
    beam_args = (g["beam"], g["ext"], g["bfm"], g["lm"], g["pa"], g["pe"], g["asc"], g["freq"])
    for ft in ("linear", "circular"):
        rot = b200.rime.feed_rotation(g["pa"], ft)
        assert_c128_close(rot, g["rot_" + ft], rtol=1e-15)
        rot32 = b200.rime.feed_rotation(g["pa"].astype(np.float32), ft)
        assert_c64_close(rot32, g["rot32_" + ft], tol=2e-7)
        dde = b200.rime.beam_cube_dde_rotated(*beam_args, rot)
        assert_c128_close(dde, g["dde_" + ft])
        got = b200.rime.fused_predict_vis_beam(g["lm"], g["uvw"], g["freq"], g["bright"], g["time_index"],
                                               g["ant1"], g["ant2"], *beam_args[:3], *beam_args[4:7],
                                               g["die"], None, g["die"], source_chunk=4, feed_type=ft)
        assert_c128_close(got, g["vis_" + ft])
    # complex64 beam: rotation applied in float32; torch inputs
    c64 = np.complex64
    ref = np.einsum("stafij,tajk->stafik",
                    b200.rime.beam_cube_dde(g["beam"].astype(c64), *beam_args[1:]).astype(np.complex128),
                    g["rot_linear"])
    T = lambda a: torch.from_numpy(np.ascontiguousarray(a)).cuda()  # noqa: E731
    got = b200.rime.beam_cube_dde_rotated(T(g["beam"].astype(c64)), *(T(a) for a in beam_args[1:]),
                                          T(g["rot_linear"].astype(c64)))
    assert got.dtype == torch.complex64
    assert rel_l2(got.cpu().numpy().astype(np.complex128), ref) < 1e-6
    with pytest.raises(ValueError):
        b200.rime.feed_rotation(g["pa"], "elliptical")
    with pytest.raises(ValueError):
        b200.rime.feed_rotation(g["pa"].astype(np.int32))
    with pytest.raises(ValueError):
        b200.rime.beam_cube_dde_rotated(g["beam"][..., 0], *beam_args[1:], g["rot_linear"])


@pytest.mark.parametrize("base", [0, 2, "log", ["log", "std", "std", "std"]])
@pytest.mark.parametrize("npol", [0, 1, 2, 4])
def test_spectral_model_multiple_spi(b200, oracle, base, npol):
    """The reference's own parametrisation (model/spectral/tests/test_spectral_model.py:41-76):
    6 spectral indices, 0/1/2/4 polarisations, broadcast (strided) stokes -- CUDA vs oracle."""
    rng = np.random.default_rng(7)
    nsrc, nchan, nspi = 10, 16, 6
    if isinstance(base, list):
        base = base[0] if npol == 0 else base[:npol]
    flux = rng.normal(size=nsrc)
    if npol > 0:
        stokes = np.broadcast_to(flux[:, None], (nsrc, npol))
        spi = 0.7 + rng.random((nsrc, nspi, npol)) * 0.2
    else:
        stokes = flux
        spi = 0.7 + rng.random((nsrc, nspi)) * 0.2
    ref_freq = np.full(nsrc, 3 * 0.856e9 / 2)
    freq = np.linspace(0.856e9, 2 * 0.856e9, nchan)
    got = b200.model.spectral_model(stokes, spi, ref_freq, freq, base=base)
    assert got.flags.c_contiguous
    assert_c128_close(got, oracle.spectral_model(stokes, spi, ref_freq, freq, base=base))


def _corrupt_vis_case(g, tag):
    """Inputs of calibration/utils/tests/test_utils.py:21-80 in predict_vis order."""
    model, jones = g[tag + "_model"], g[tag + "_jones"]
    if tag == "d22":  # DIAG Jones against 2x2 model: broadcast onto the diagonal (:47-53)
        tmp = np.zeros(jones.shape[:4] + (2, 2), np.complex128)
        tmp[..., 0, 0] = jones[..., 0]
        tmp[..., 1, 1] = jones[..., 1]
        jones = tmp
    if model.ndim == 5:
        jones, model = np.transpose(jones, [3, 0, 1, 2, 4, 5]), np.transpose(model, [2, 0, 1, 3, 4])
    else:
        jones, model = np.transpose(jones, [3, 0, 1, 2, 4]), np.transpose(model, [2, 0, 1, 3])
    time_index = np.unique(g["time"], return_inverse=True)[1]
    return time_index, g["antenna1"], g["antenna2"], jones, model


@pytest.mark.parametrize("tag", ["c1", "c2", "d22", "f22"])
def test_predict_vis_equals_reference_corrupt_vis(b200, golden, tag):
    """CUDA predict_vis against the reference's independent corrupt_vis on its own cross-check
    inputs (calibration/utils/tests/test_utils.py:21-80): int16 indices, antenna1 > antenna2,
    transposed views."""
    g = golden("corrupt_vis")
    ti, a1, a2, jones, model = _corrupt_vis_case(g, tag)
    got = b200.rime.predict_vis(ti, a1, a2, dde1_jones=jones, source_coh=model, dde2_jones=jones)
    assert_c128_close(got, g[tag + "_vis"])


def test_fused_spec_front_end_golden(b200, golden):
    """``rime(spec, dataset)`` against the reference's own fused-RIME front end
    (africanus/experimental/rime/fused/core.py:227-241; goldens generated by oracle/gen_golden.py from the
    unmodified reference): (Kpq, Bpq) for linear / circular / two-correlation schemas, the three spectral
    bases and both conventions; feed rotation outside and inside the beam term; the beam cube term."""
    from codex_africanus_b200.rime.fused_spec import rime
    g = golden("fused_spec")
    ds = {k: g[k] for k in ("time", "antenna1", "antenna2", "feed1", "feed2", "radec", "phase_dir", "uvw", "chan_freq",
                           "stokes", "spi", "ref_freq")}
    lin, circ = "[XX,XY,YX,YY]", "[RR,RL,LR,LL]"
    for tag, corrs, conv, base in (("kb_lin_casa_std", lin, "casa", "standard"), ("kb_circ_fourier_log", circ, "fourier", "log"),
                                   ("kb_lin_fourier_log10", lin, "fourier", "log10"), ("kb_rrll_fourier_std", "[RR,LL]", "fourier", "standard")):
        got = rime("(Kpq, Bpq): [I,Q,U,V] -> %s" % corrs, ds, convention=conv, spi_base=base)
        assert isinstance(got, np.ndarray)
        assert_c128_close(got, g[tag])
    for tag, corrs in (("lkbl_lin", lin), ("lkbl_circ", circ)):
        got = rime("(Lp, Kpq, Bpq, Lq): [I,Q,U,V] -> %s" % corrs, ds, feed_parangle=g["feed_parangle"],
                   convention="casa", spi_base="standard")
        assert_c128_close(got, g[tag])
        got = rime("(Lp, Kpq, Bpq, Lq): [I,Q,U,V] -> %s" % corrs, ds, parallactic_angles=g["parallactic_angles"],
                   convention="casa")
        assert_c128_close(got, g[tag])
    eds = {**ds, "beam": g["beam"], "beam_lm_extents": g["beam_lm_extents"], "beam_freq_map": g["beam_freq_map"],
           "beam_parangle": g["beam_parangle"]}
    assert_c128_close(rime("(Ep, Kpq, Bpq, Eq): [I,Q,U,V] -> %s" % lin, eds, convention="casa"), g["ekbe_lin"])
    assert_c128_close(rime("(Ep, Kpq, Bpq, Eq): [I,Q,U,V] -> %s" % lin, eds, convention="casa", in_kernel=True), g["ekbe_lin"])
    assert_c128_close(rime("(Lp, Ep, Kpq, Bpq, Eq, Lq): [I,Q,U,V] -> %s" % lin, eds, feed_parangle=g["feed_parangle"]),
                      g["lekbel_lin"])
    assert_c128_close(rime("(Ep, Lp, Kpq, Bpq, Lq, Eq): [I,Q,U,V] -> %s" % lin, eds, feed_parangle=g["feed_parangle"]),
                      g["elkble_lin"])
    # the five required arrays may be given positionally, the rest by keyword; CUDA tensors in -> CUDA tensor out
    import torch
    kw = {k: (torch.from_numpy(v).cuda() if k in ("uvw", "chan_freq", "stokes", "spi", "ref_freq") else v)
          for k, v in ds.items() if k not in fs_required()}
    got = rime("(Kpq, Bpq): [I,Q,U,V] -> %s" % lin, *(ds[k] for k in fs_required()), convention="casa", **kw)
    assert got.is_cuda
    assert_c128_close(got.cpu().numpy(), g["kb_lin_casa_std"])


def fs_required():
    from codex_africanus_b200.rime.fused_spec import REQUIRED_ARGS
    return REQUIRED_ARGS


def test_fused_predict_vis_beam_sampled_in_kernel(b200, oracle, monkeypatch):
    """SURVEY 8f-1 proper: ``fused_predict_vis_beam(in_kernel=True)`` forms the beam Jones inside the
    predict kernel from the plane-reduced beam (no (source,time,ant,chan,2,2) array, not even per chunk).
    Same values as the oracle's beam_cube_dde -> fused_predict chain and as the chunked route; the
    kernel that ran is asserted; inputs the route does not cover fall back to the chunked one."""
    from codex_africanus_b200 import _lib
    from codex_africanus_b200.rime import fused_beam
    rng = np.random.default_rng(1618)

    def rc(shape):
        return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)

    for na, ntime, nsrc, nchan in ((7, 3, 13, 24), (40, 2, 6, 5), (3, 1, 1, 1)):
        a1, a2 = np.triu_indices(na, 1)
        ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
        ti = np.repeat(np.arange(ntime), a1.size) + 4
        nrow = ti.size
        antpos = rng.standard_normal((ntime, na, 3)) * 1200.0
        uvw = antpos[ti - 4, ant1] - antpos[ti - 4, ant2]
        lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
        freq = np.linspace(0.9e9, 1.7e9, nchan) if nchan > 1 else np.array([1.2e9])
        beam = rc((11, 9, 6, 2, 2))
        ext = np.array([[-0.03, 0.03], [-0.03, 0.03]])
        bfm = np.linspace(0.8e9, 1.8e9, 6)
        pa = rng.uniform(-1, 1, (ntime, na))
        pe = np.broadcast_to(rng.uniform(-1e-3, 1e-3, (ntime, na, 1, 2)), (ntime, na, nchan, 2)).copy()
        asc = np.broadcast_to(rng.uniform(0.9, 1.1, (na, 1, 2)), (na, nchan, 2)).copy()
        bright = rc((nsrc, nchan, 2, 2))
        die = 1.0 + 0.1 * rc((ntime, na, nchan, 2, 2))
        bvis = rc((nrow, nchan, 2, 2))
        dde = oracle.beam_cube_dde(beam, ext, bfm, lm, pa, pe, asc, freq)
        args = (lm, uvw, freq, bright, ti, ant1, ant2, beam, ext, bfm, pa, pe, asc)
        for extra, ref in (((die, bvis, die), oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die)),
                           ((), oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde))):
            got = b200.rime.fused_predict_vis_beam(*args, *extra, in_kernel=True)
            assert _lib.lib().afr_last_fused_path() == 7, (na, ntime, nsrc, nchan)
            assert_c128_close(got, ref)
            assert_c128_close(b200.rime.fused_predict_vis_beam(*args, *extra), ref)
        # with the feed rotation on the right of the beam Jones (dde = beam . L), both feed types
        for ft in ("linear", "circular"):
            want = b200.rime.fused_predict_vis_beam(*args, die, bvis, die, feed_type=ft)
            got = b200.rime.fused_predict_vis_beam(*args, die, bvis, die, feed_type=ft, in_kernel=True)
            assert _lib.lib().afr_last_fused_path() == 7
            assert_c128_close(got, want)
        # several plane chunks (sources per launch bounded by _PLANES_CHUNK_BYTES)
        monkeypatch.setattr(fused_beam, "_PLANES_CHUNK_BYTES", 4 * ntime * na * 6 * 96)
        got = b200.rime.fused_predict_vis_beam(*args, die, bvis, die, in_kernel=True)
        monkeypatch.undo()
        assert_c128_close(got, oracle.fused_predict(lm, uvw, freq, bright, ti, ant1, ant2, dde, dde, die, bvis, die))
    # not covered: pointing errors that change along the channel axis; uvw that are not antenna differences;
    # a channel outside the cube's frequency range -> the chunked route, same values
    pe2 = rng.uniform(-1e-3, 1e-3, pe.shape)
    na, ntime, nsrc, nchan = 7, 3, 13, 24
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime), np.tile(a2, ntime)
    ti = np.repeat(np.arange(ntime), a1.size)
    antpos = rng.standard_normal((ntime, na, 3)) * 1200.0
    uvw = antpos[ti, ant1] - antpos[ti, ant2]
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    bright = rc((nsrc, nchan, 2, 2))
    pa = rng.uniform(-1, 1, (ntime, na))
    pe = np.zeros((ntime, na, nchan, 2))
    asc = np.ones((na, nchan, 2))
    for freq, uv, perr in ((np.linspace(0.9e9, 1.7e9, nchan), uvw, rng.uniform(-1e-3, 1e-3, pe.shape)),
                           (np.linspace(0.9e9, 1.7e9, nchan), rng.standard_normal(uvw.shape) * 900.0, pe),
                           (np.linspace(0.7e9, 1.7e9, nchan), uvw, pe)):
        dde = oracle.beam_cube_dde(beam, ext, bfm, lm, pa, perr, asc, freq)
        ref = oracle.fused_predict(lm, uv, freq, bright, ti, ant1, ant2, dde, dde)
        got = b200.rime.fused_predict_vis_beam(lm, uv, freq, bright, ti, ant1, ant2, beam, ext, bfm, pa, perr, asc,
                                               in_kernel=True)
        assert _lib.lib().afr_last_fused_path() != 7
        assert_c128_close(got, ref)
