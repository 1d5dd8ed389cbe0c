"""
Pin the CPU oracle (oracle/afr_oracle.c) against the reference.

* golden vectors produced by the reference's numba implementation
  (oracle/gen_golden.py -> tests/golden/*.npz), compared BIT-EXACTLY wherever the
  oracle restates the same operation order with the same libm, otherwise to 1e-13;
* the reference's own known-answer tests, re-stated against the oracle:
  rime/tests/test_rime.py:19-47, rime/tests/test_fast_beams.py:43-150,
  dft/tests/test_dft.py:12-215,297-331, rime/tests/test_predict.py:62-126.

CPU only (no gpu marker).
"""
import itertools

import numpy as np
import pytest
from numpy.testing import assert_array_almost_equal, assert_array_equal

from conftest import rel_l2

C = 2.99792458e8
MINUS_TWO_PI_OVER_C = -2 * np.pi / C


def _tight(got, ref, tol=1e-13):
    assert got.shape == ref.shape and got.dtype == ref.dtype
    scale = max(np.max(np.abs(ref)), 1e-300)
    assert np.max(np.abs(got - ref)) <= tol * scale


# ----------------------------------------------------------------------------- phase
def test_phase_delay_golden(oracle, golden):
    g = golden("phase_delay")
    lm, uvw, freq = g["lm"], g["uvw"], g["freq"]
    for conv in ("fourier", "casa"):
        assert_array_equal(oracle.phase_delay(lm, uvw, freq, convention=conv), g["f64_" + conv])
    assert_array_equal(oracle.phase_delay(lm, uvw, g["freq_nu"]), g["f64_nonuniform"])
    f32 = np.float32
    lm_s, uvw_s = g["lm_s"], g["uvw_s"]
    got = oracle.phase_delay(lm_s.astype(f32), uvw_s.astype(f32), freq.astype(f32))
    assert got.dtype == np.complex64
    # float32 libm (cosf/sinf) may differ in the last ulp between builds
    assert np.max(np.abs(got - g["f32_all"])) <= 3e-7
    assert_array_equal(oracle.phase_delay(lm_s.astype(f32), uvw_s, freq), g["mix_lm32"])
    assert_array_equal(oracle.phase_delay(lm_s.astype(f32), uvw_s.astype(f32), freq), g["mix_lm32_uvw32"])
    assert_array_equal(oracle.phase_delay(lm_s, uvw_s.astype(f32), freq), g["mix_uvw32"])
    assert_array_equal(oracle.phase_delay(lm_s, uvw_s, freq.astype(f32)), g["mix_freq32"])


@pytest.mark.parametrize("convention, sign", [("fourier", 1), ("casa", -1)])
def test_phase_delay_known_answer(oracle, convention, sign):
    # rime/tests/test_rime.py:19-47 (bit-exact)
    rng = np.random.default_rng(0)
    uvw = rng.random((100, 3))
    lm = rng.random((10, 2))
    frequency = np.linspace(0.856e9, 0.856e9 * 2, 64, endpoint=True)
    uvw[2] = [1, 2, 3]
    lm[3] = [0.1, 0.2]
    frequency[5] = 0.856e9
    cp = oracle.phase_delay(lm, uvw, frequency, convention=convention)
    n = np.sqrt(1.0 - 0.1**2 - 0.2**2) - 1.0
    phase = sign * MINUS_TWO_PI_OVER_C * (1 * 0.1 + 2 * 0.2 + 3 * n) * 0.856e9
    assert np.all(np.exp(1j * phase) == cp[3, 2, 5])


def test_phase_delay_bad_convention(oracle):
    with pytest.raises(ValueError, match="convention not in"):
        oracle.phase_delay(np.zeros((1, 2)), np.zeros((1, 3)), np.ones(1), convention="x")


# ----------------------------------------------------------------------------- dft
def test_dft_golden(oracle, golden):
    g = golden("dft")
    lm, uvw, freq = g["lm"], g["uvw"], g["freq"]
    f32 = np.float32
    for nc in (1, 2, 4):
        assert_array_equal(oracle.im_to_vis(g["image_r%d" % nc], uvw, lm, freq), g["i2v_r%d" % nc])
        assert_array_equal(
            oracle.vis_to_im(g["vis_c%d" % nc], uvw, lm, freq, g["flags_%d" % nc]), g["v2i_c%d" % nc])
    assert_array_equal(oracle.im_to_vis(g["image_c2"], uvw, lm, freq), g["i2v_c2"])
    assert_array_equal(oracle.im_to_vis(g["image_c2"], uvw, lm, freq, convention="casa"), g["i2v_c2_casa"])
    assert_array_equal(oracle.im_to_vis(g["image_r1"], uvw, lm, g["freq_nu"]), g["i2v_r1_nonuniform"])
    got = oracle.im_to_vis(g["image_r1"], uvw, lm, freq, dtype=np.complex64)
    assert got.dtype == np.complex64
    assert_array_equal(got, g["i2v_r1_c64"])
    got = oracle.im_to_vis(g["image_r1"].astype(f32), uvw.astype(f32), lm.astype(f32), freq.astype(f32))
    assert got.dtype == g["i2v_r1_in32"].dtype
    assert_array_equal(got, g["i2v_r1_in32"])
    assert_array_equal(oracle.im_to_vis(g["image_r1"], uvw, lm.astype(f32), freq), g["i2v_r1_lm32"])
    assert_array_equal(oracle.vis_to_im(g["vis_r2"], uvw, lm, freq, g["flags_2"]), g["v2i_r2"])
    assert_array_equal(
        oracle.vis_to_im(g["vis_c1"], uvw, lm, freq, g["flags_1"], convention="casa"), g["v2i_c1_casa"])
    got = oracle.vis_to_im(g["vis_c1"], uvw, lm, freq, g["flags_1"], dtype=np.float32)
    assert got.dtype == np.float32
    assert_array_equal(got, g["v2i_c1_f32"])
    assert_array_equal(
        oracle.vis_to_im(g["vis_c1"], uvw, lm, g["freq_nu"], g["flags_1"]), g["v2i_c1_nonuniform"])
    got = oracle.vis_to_im(g["vis_c1"].astype(np.complex64), uvw.astype(f32), lm.astype(f32),
                           freq.astype(f32), g["flags_1"])
    assert got.dtype == g["v2i_c1_in32"].dtype
    assert_array_equal(got, g["v2i_c1_in32"])


def test_dft_padded_grid_golden(oracle, golden):
    """zero-padded image grid past the unit disc: n = NaN there (kernels.py:54), zero pixels are
    skipped (kernels.py:64); a bright pixel out there gives NaN exactly where it is non-zero"""
    g = golden("dft_padded")
    lm, uvw, freq = g["lm"], g["uvw"], g["freq"]
    with np.errstate(invalid="ignore"):
        got = oracle.im_to_vis(g["image"], uvw, lm, freq)
        assert np.all(np.isfinite(got))
        assert_array_equal(got, g["i2v"])
        assert_array_equal(oracle.im_to_vis(g["image"], uvw, lm, freq, dtype=np.complex64), g["i2v_c64"])
        bad = oracle.im_to_vis(g["image_bad"], uvw, lm, freq)
    assert_array_equal(np.isnan(bad), np.isnan(g["i2v_bad"]))
    ok = ~np.isnan(bad)
    assert_array_equal(bad[ok], g["i2v_bad"][ok])


def test_im_to_vis_phase_centre(oracle):
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
    vis = oracle.im_to_vis(image.reshape(npix**2, nchan, ncorr), uvw, lm, frequency)
    tmp = vis - Inu[None, :, None]
    assert np.all(np.abs(tmp.real) < 1e-13) and np.all(np.abs(tmp.imag) < 1e-13)


@pytest.mark.parametrize("convention", ["fourier", "casa"])
def test_im_to_vis_fft(oracle, convention):
    # dft/tests/test_dft.py:86-133
    np.random.seed(123)
    Fs, iFs = np.fft.fftshift, np.fft.ifftshift
    npix, ncorr, nsource = 29, 1, 25
    image = np.zeros((npix, npix, ncorr))
    fft_image = np.zeros((npix, npix, ncorr), np.complex128)
    Ix = np.random.randint(5, npix - 5, nsource)
    Iy = np.random.randint(5, npix - 5, nsource)
    image[Ix, Iy, 0] = np.random.randn(nsource)
    fft_image[:, :, 0] = Fs(np.fft.fft2(iFs(image[:, :, 0])))
    deltal = 0.001
    l_coord = np.arange(-(npix // 2), npix // 2 + 1) * deltal
    ll, mm = np.meshgrid(l_coord, l_coord)
    lm = np.vstack((ll.flatten(), mm.flatten())).T
    u = Fs(np.fft.fftfreq(npix, d=deltal))
    uu, vv = np.meshgrid(u, u)
    uvw = np.zeros((npix**2, 3))
    uvw[:, 0], uvw[:, 1] = uu.flatten(), vv.flatten()
    frequency = np.ones(1) * C
    vis = oracle.im_to_vis(image.reshape(npix**2, 1, ncorr), uvw, lm, frequency, convention=convention)
    fft_image = fft_image.reshape(npix**2, 1, ncorr)
    fft_image = np.conj(fft_image) if convention == "casa" else fft_image
    assert_array_almost_equal(vis, fft_image, decimal=13)


def test_adjointness_and_flags(oracle):
    # dft/tests/test_dft.py:136-177 and :180-215
    np.random.seed(123)
    nsource, nrow, nchan, ncorr = 21, 31, 3, 4
    uvw = 100 * np.random.random(size=(nrow, 3))
    lm = np.vstack((0.01 * np.random.randn(nsource), 0.01 * np.random.randn(nsource))).T
    frequency = np.arange(1, nchan + 1) * C
    gamma_im = np.random.randn(nsource, nchan, ncorr)
    gamma_vis = np.random.randn(nrow, nchan, ncorr)
    flag = np.zeros((nrow, nchan, ncorr), dtype=bool)
    LHS = np.vdot(gamma_vis.ravel(), oracle.im_to_vis(gamma_im, uvw, lm, frequency).ravel()).real
    RHS = np.dot(oracle.vis_to_im(gamma_vis, uvw, lm, frequency, flag).ravel(), gamma_im.ravel())
    assert np.abs(LHS - RHS) < 1e-13 * max(1.0, abs(LHS))
    # flagged data: only the zero-uvw, all-ones row survives
    uvw[0, :] = 0.0
    vis = np.random.randn(nrow, nchan, ncorr) + 1.0j * np.random.randn(nrow, nchan, ncorr)
    vis[0, :, :] = 1.0
    flags = np.ones((nrow, nchan, ncorr), dtype=bool)
    flags[0, :, :] = 0
    im = oracle.vis_to_im(vis, uvw, lm, np.ones(nchan) * C, flags)
    assert_array_almost_equal(im, np.ones((nsource, nchan, ncorr)), decimal=13)


def test_vis_to_im_dtype_errors(oracle):
    z = np.zeros((2, 1, 1))
    with pytest.raises(TypeError):
        oracle.vis_to_im(z, np.zeros((2, 3)), np.zeros((1, 2)), np.ones(1), z > 0, dtype=np.complex128)
    with pytest.raises(AssertionError):
        oracle.vis_to_im(z, np.zeros((2, 3)), np.zeros((1, 2)), np.ones(1), np.zeros((2, 1, 2), bool))


# ----------------------------------------------------------------------------- predict
PRESENCE = [(True, True, True), (True, False, True), (False, True, False)]


@pytest.mark.parametrize("cname", ["c1", "c2", "c22"])
def test_predict_vis_golden(oracle, golden, cname):
    g = golden("predict_vis")
    ti, a1, a2 = g["time_idx"], g["ant1"], g["ant2"]
    arrs = {k: g["%s_%s" % (cname, k)] for k in ("a1j", "blj", "a2j", "g1j", "bvis", "g2j")}
    for (d1, bl, d2), (g1, bv, g2) in itertools.product(PRESENCE, PRESENCE):
        key = "%s_out_%d%d%d_%d%d%d" % (cname, d1, bl, d2, g1, bv, g2)
        got = oracle.predict_vis(
            ti, a1, a2, arrs["a1j"] if d1 else None, arrs["blj"] if bl else None,
            arrs["a2j"] if d2 else None, arrs["g1j"] if g1 else None,
            arrs["bvis"] if bv else None, arrs["g2j"] if g2 else None)
        assert_array_equal(got, g[key], err_msg=key)
    a64 = {k: v.astype(np.complex64) for k, v in arrs.items()}
    got = oracle.predict_vis(ti.astype(np.int16), a1.astype(np.int16), a2.astype(np.int16),
                             a64["a1j"], a64["blj"], a64["a2j"], a64["g1j"], a64["bvis"], a64["g2j"])
    assert got.dtype == np.complex64
    assert_array_equal(got, g["%s_out_c64" % cname])


@pytest.mark.parametrize("corr_shape, idm, sig1, sig2", [
    ((1,), (1,), "srci,srci,srci->rci", "rci,rci,rci->rci"),
    ((2,), (1, 1), "srci,srci,srci->rci", "rci,rci,rci->rci"),
    ((2, 2), ((1, 0), (0, 1)), "srcij,srcjk,srclk->rcil", "rcij,rcjk,rclk->rcil"),
])
def test_predict_vis_einsum(oracle, corr_shape, idm, sig1, sig2):
    # rime/tests/test_predict.py:62-126 (independent numpy oracle), all 9 presence combos
    rng = np.random.default_rng(5)
    s, t, a, c, r = 21, 4, 4, 5, 10

    def rcx(shape):
        return rng.random(shape) + 1j * rng.random(shape)

    time_idx = np.asarray([0, 0, 1, 1, 2, 2, 2, 2, 3, 3])
    ant1 = np.asarray([0, 0, 0, 0, 1, 1, 1, 2, 2, 3])
    ant2 = np.asarray([0, 1, 2, 3, 1, 2, 3, 2, 3, 3])
    A1, BL, A2 = rcx((s, t, a, c) + corr_shape), rcx((s, r, c) + corr_shape), rcx((s, t, a, c) + corr_shape)
    G1, BV, G2 = rcx((t, a, c) + corr_shape), rcx((r, c) + corr_shape), rcx((t, a, c) + corr_shape)
    for (d1, bl, d2), (g1, bv, g2) in itertools.product(PRESENCE, PRESENCE):
        got = oracle.predict_vis(time_idx, ant1, ant2, A1 if d1 else None, BL if bl else None,
                                 A2 if d2 else None, G1 if g1 else None, BV if bv else None,
                                 G2 if g2 else None)
        ident = lambda arr: np.broadcast_to(idm, arr.shape)  # noqa: E731
        e1 = A1[:, time_idx, ant1] if d1 else ident(BL)
        x = BL if bl else ident(BL)
        e2 = A2[:, time_idx, ant2].conj() if d2 else ident(BL)
        v = np.einsum(sig1, e1, x, e2)
        if bv:
            v = v + BV
        gg1 = G1[time_idx, ant1] if g1 else ident(v)
        gg2 = G2[time_idx, ant2].conj() if g2 else ident(v)
        v = np.einsum(sig2, gg1, v, gg2)
        assert_array_almost_equal(v, got)


def test_predict_vis_errors(oracle):
    ti = a1 = a2 = np.zeros(2, np.int32)
    dde = np.zeros((1, 1, 1, 1, 2, 2), np.complex128)
    coh = np.zeros((1, 2, 1, 2), np.complex128)
    with pytest.raises(ValueError, match="Both dde1_jones and dde2_jones"):
        oracle.predict_vis(ti, a1, a2, dde1_jones=dde)
    with pytest.raises(ValueError, match="Both die1_jones and die2_jones"):
        oracle.predict_vis(ti, a1, a2, die1_jones=dde[0])
    with pytest.raises(ValueError, match="mismatched"):
        oracle.predict_vis(ti, a1, a2, dde, coh, dde)
    with pytest.raises(ValueError, match="No Jones"):
        oracle.predict_vis(ti, a1, a2)
    with pytest.raises(ValueError, match="ndim"):
        oracle.predict_vis(ti, a1, a2, source_coh=np.zeros((1, 2, 1), np.complex128))


# ----------------------------------------------------------------------------- beam
def test_freq_grid_interp_known_answer(oracle, golden):
    # rime/tests/test_fast_beams.py:130-150
    freqs = np.array([0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1])
    bfm = np.array([0.5, 0.56, 0.7, 0.91, 1.0])
    fd = oracle.freq_grid_interp(freqs, bfm)
    assert_array_almost_equal(fd[:, 0], [0.8, 1.0, 1.0, 1.0, 1.0, 1.0, 1.0, 1.1])
    assert_array_equal(np.int32(fd[:, 2]), [0, 0, 1, 2, 2, 2, 3, 3])
    assert_array_almost_equal(fd[:, 1], [1.0, 1.0, 0.71428571, 1.0, 0.52380952, 0.04761905, 0.0, 0.0])
    assert_array_equal(fd, golden("beam_cube_dde")["freq_data"])


def test_beam_cube_dde_golden(oracle, golden):
    g = golden("beam_cube_dde")
    args = (g["ext"], g["beam_freq_map"], g["lm"], g["pa"], g["perr"], g["ascale"], g["freq"])
    for cname in ("c22", "c4", "c2", "c1"):
        got = oracle.beam_cube_dde(g["beam_" + cname], *args)
        _tight(got, g["dde_" + cname], 1e-15)
    got = oracle.beam_cube_dde(g["beam_c22"].astype(np.complex64), *args)
    assert got.dtype == np.complex64
    assert np.max(np.abs(got - g["dde_c22_c64"])) <= 5e-7 * np.max(np.abs(g["dde_c22_c64"]))
    # known-answer (rime/tests/test_fast_beams.py:43-127)
    ka = oracle.beam_cube_dde(
        g["ka_beam"], np.asarray([[-1.0, 1.0], [-1.0, 1.0]]), np.asarray([0.0, 1.0]),
        np.asarray([[0.1, 0.1]]), np.zeros((1, 1)), np.zeros((1, 1, 1, 2)), np.ones((1, 1, 2)),
        np.asarray([0.3]))
    assert_array_almost_equal([[[[[0.470255 + 0.4786j]]]]], ka)
    _tight(ka, g["ka_dde"], 1e-15)


def test_beam_cube_dde_errors(oracle):
    with pytest.raises(ValueError, match="must be >= 2"):
        oracle.beam_cube_dde(np.zeros((1, 2, 2, 1), np.complex128), np.zeros((2, 2)), np.zeros(2),
                             np.zeros((1, 2)), np.zeros((1, 1)), np.zeros((1, 1, 1, 2)),
                             np.ones((1, 1, 2)), np.ones(1))


# ----------------------------------------------------------------------------- fused
def test_fused_predict_golden(oracle, golden):
    g = golden("fused_predict")
    ti, a1, a2 = g["time_idx"], g["ant1"], g["ant2"]
    lm, uvw, freq = g["lm"], g["uvw"], g["freq"]
    for conv in ("fourier", "casa"):
        for cname in ("c22", "c2", "c1"):
            key = "%s_%s" % (conv, cname)
            br, die, bvis = g["bright_" + key], g["die_" + key], g["bvis_" + key]
            got = oracle.fused_predict(lm, uvw, freq, br, ti, a1, a2, convention=conv)
            assert_array_equal(got, g["point_" + key])
            got = oracle.fused_predict(lm, uvw, freq, br, ti, a1, a2, die1_jones=die,
                                       base_vis=bvis, die2_jones=die, convention=conv)
            assert_array_equal(got, g["point_die_" + key])
            dde = g["dde_beam_c22"] if cname == "c22" else g["dde_" + key]
            got = oracle.fused_predict(lm, uvw, freq, br, ti, a1, a2, dde, dde, die, bvis, die,
                                       convention=conv)
            assert_array_equal(got, g["full_" + key])


# ----------------------------------------------------------------------------- wsclean
def test_wsclean_spectra_and_predict_golden(golden, oracle):
    """africanus.model.wsclean.spectra + africanus.rime.wsclean_predict: the reference's own test
    inputs (rime/tests/test_wsclean_predict.py:26-61) and a MeerKAT-like mixed POINT/GAUSSIAN
    model, bit-exact."""
    g = golden("wsclean")
    for pre in ("t_", "m_"):
        args = [g[pre + k] for k in ("flux", "coeffs", "log_poly", "ref_freq", "freq")]
        assert np.array_equal(oracle.wsclean_spectra(*args), g[pre + "spectra"])
        got = oracle.wsclean_predict(g[pre + "uvw"], g[pre + "lm"], g[pre + "source_type"], g[pre + "flux"],
                                     g[pre + "coeffs"], g[pre + "log_poly"], g[pre + "ref_freq"],
                                     g[pre + "gauss_shape"], g[pre + "freq"])
        assert np.array_equal(got, g[pre + "vis"])
    for lp, key in ((True, "m_vis_logpoly_true"), (False, "m_vis_logpoly_false")):
        got = oracle.wsclean_predict(g["m_uvw"], g["m_lm"], g["m_source_type"], g["m_flux"], g["m_coeffs"],
                                     lp, g["m_ref_freq"], g["m_gauss_shape"], g["m_freq"])
        assert np.array_equal(got, g[key])


def test_wsclean_predict_known_answer(golden, oracle):
    """The reference test's independent composition: einsum(shape, phase_delay(casa), spectra)
    with shape = 1 for POINT sources (rime/tests/test_wsclean_predict.py:52-61)."""
    g = golden("wsclean")
    uvw, lm, freq, gs = g["t_uvw"], g["t_lm"], g["t_freq"], g["t_gauss_shape"]
    phase = oracle.phase_delay(lm, uvw, freq, convention="casa")
    spectrum = oracle.wsclean_spectra(g["t_flux"], g["t_coeffs"], g["t_log_poly"], g["t_ref_freq"], freq)
    # africanus/model/shape/gaussian_shape.py:31-63
    fwhminv = 1.0 / (2.0 * np.sqrt(2.0 * np.log(2.0)))
    scale = fwhminv * np.sqrt(2.0) * np.pi / 2.99792458e8
    emaj, emin, ang = gs[:, 0], gs[:, 1], gs[:, 2]
    el, em = emaj * np.sin(ang), emaj * np.cos(ang)
    er = emin / np.where(emaj == 0.0, 1.0, emaj)
    u, v = uvw[:, 0], uvw[:, 1]
    u1 = (u[None, :] * em[:, None] - v[None, :] * el[:, None]) * er[:, None]
    v1 = u[None, :] * el[:, None] + v[None, :] * em[:, None]
    sf = freq * scale
    shape = np.exp(-((u1[:, :, None] * sf) ** 2 + (v1[:, :, None] * sf) ** 2))
    shape[g["t_source_type"] == "POINT"] = 1.0
    ref = np.einsum("srf,srf,sf->rf", shape, phase, spectrum)[:, :, None]
    np.testing.assert_almost_equal(ref, g["t_vis"])
    got = oracle.wsclean_predict(uvw, lm, g["t_source_type"], g["t_flux"], g["t_coeffs"], g["t_log_poly"],
                                 g["t_ref_freq"], gs, freq)
    np.testing.assert_almost_equal(ref, got)


# ----------------------------------------------------------------------------- brightness (8f-2)
_LIN = [["XX", "XY"], ["YX", "YY"]]
_CIRC = [["RR", "RL"], ["LR", "LL"]]
_IQUV = ["I", "Q", "U", "V"]


def test_spectral_model_golden(golden, oracle):
    """africanus.model.spectral.spectral_model: every base (string and integer spelling), 1 / 2 /
    6 spectral-index components (integer powers >= 4 pin numba's square-and-multiply order),
    per-polarisation base list, no polarisation axis -- bit-exact; float32 inputs to 1e-6."""
    g = golden("brightness")
    st, rf, fr = g["stokes"], g["ref_freq"], g["freq"]
    for n in (1, 2, 6):
        spi = g["spi%d" % n]
        for i, b in enumerate(("std", "log", "log10")):
            ref = g["sm_%s_%d" % (b, n)]
            assert np.array_equal(oracle.spectral_model(st, spi, rf, fr, base=b), ref)
            assert np.array_equal(oracle.spectral_model(st, spi, rf, fr, base=i), g["sm_int_%d" % n][i])
            assert np.array_equal(ref, g["sm_int_%d" % n][i])
    assert np.array_equal(oracle.spectral_model(st, g["spi2"], rf, fr, base=["std", "log", "log10"]),
                          g["sm_list"])
    assert np.array_equal(oracle.spectral_model(st[:, 0], g["spi2"][:, :, 0], rf, fr, base="log"),
                          g["sm_nopol"])
    f32 = np.float32
    got = oracle.spectral_model(st.astype(f32), g["spi2"].astype(f32), rf.astype(f32), fr.astype(f32))
    assert got.dtype == np.float32 and got.shape == g["sm_f32"].shape
    assert rel_l2(got, g["sm_f32"]) < 1e-6
    with pytest.raises(ValueError):
        oracle.spectral_model(st, g["spi2"][:, :, 0], rf, fr)
    with pytest.raises(ValueError):
        oracle.spectral_model(st, g["spi2"], rf, fr, base="ln")


def test_spectral_model_known_answer(golden, oracle):
    """The reference test's independent numpy formulation
    (model/spectral/tests/test_spectral_model.py / spec_model.py:11-54)."""
    g = golden("brightness")
    st, rf, fr, spi = g["stokes"], g["ref_freq"], g["freq"], g["spi2"]
    ratio = fr[None, :] / rf[:, None]
    std = st[:, None, :] * np.prod(ratio[:, None, :, None] ** spi[:, :, None, :], axis=1)
    np.testing.assert_allclose(oracle.spectral_model(st, spi, rf, fr, base="std"), std, rtol=1e-13)
    exps = np.arange(1, spi.shape[1] + 1)
    lg = st[:, None, :] * np.exp(np.sum(spi[:, :, None, :] * np.log(ratio)[:, None, :, None]
                                        ** exps[None, :, None, None], axis=1))
    np.testing.assert_allclose(oracle.spectral_model(st, spi, rf, fr, base="log"), lg, rtol=1e-13)


def test_convert_golden(golden, oracle):
    """africanus.model.coherency.convert: Stokes -> linear / circular brightness (nested, flat,
    implicit Stokes, two-element), correlations -> Stokes (complex, real, integer ids) --
    bit-exact incl. dtype; and the reference test's known answers
    (model/coherency/tests/test_convert.py:69-110)."""
    g = golden("brightness")
    sm, vis = g["sm_std_2"], g["vis"]
    cases = [
        ("b_linear", sm, _IQUV, _LIN, False), ("b_circular", sm, _IQUV, _CIRC, False),
        ("b_flat", sm, _IQUV, ["XX", "XY", "YX", "YY"], False),
        ("b_implicit", sm[..., :1], ["I"], ["XX", "XY", "YX", "YY"], True),
        ("b_diag", sm[..., :2], ["I", "Q"], ["XX", "YY"], False),
        ("b_f32", sm.astype(np.float32), _IQUV, _LIN, False),
        ("s_linear", vis, _LIN, _IQUV, False), ("s_circular", vis, _CIRC, [["I", "Q"], ["U", "V"]], False),
        ("s_int", vis[..., 0, :], [9, 12], [1, 2], False),
        ("s_real", vis.real, _LIN, ["I", "Q"], False),
    ]
    for key, inp, isch, osch, implicit in cases:
        got = oracle.convert(inp, isch, osch, implicit_stokes=implicit)
        assert got.dtype == g[key].dtype and got.shape == g[key].shape, key
        assert np.array_equal(got, g[key]), key
    I, Q, U, V = 1.0 + 1j, 2.0 + 2j, 3.0 + 3j, 4.0 + 4j  # noqa: E741
    inp = np.asarray([[I, Q, U, V]])
    assert np.all(oracle.convert(inp, _IQUV, ["XX", "XY", "YX", "YY"])
                  == [[I + Q, U + V * 1j, U - V * 1j, I - Q]])
    assert np.all(oracle.convert(inp, _IQUV, ["RR", "RL", "LR", "LL"])
                  == [[I + V, Q + U * 1j, Q - U * 1j, I - V]])
    with pytest.raises(ValueError):
        oracle.convert(sm[..., :1], ["I"], ["XX", "XY", "YX", "YY"])  # no implicit Stokes


def test_wsclean_spectra_six_coefficients(golden, oracle):
    """model.wsclean.spectra with powers up to 6 (numba int_power ordering), bit-exact."""
    g = golden("brightness")
    got = oracle.wsclean_spectra(g["stokes"][:, 0], g["w_coeffs"], g["w_log_poly"], g["ref_freq"], g["freq"])
    assert np.array_equal(got, g["w_spectra"])


def test_stokes_predict_composition_golden(golden, oracle):
    """spectral_model -> convert -> phase_delay (x) brightness -> predict_vis with DIEs and base_vis
    (rime/examples/predict.py:107-134,490,522-527), both feed types, bit-exact."""
    g = golden("brightness")
    sm = oracle.spectral_model(g["stokes"], g["spi2"], g["ref_freq"], g["freq"])
    for feed, schema in (("linear", _LIN), ("circular", _CIRC)):
        b = oracle.convert(sm, _IQUV, schema)
        got = oracle.fused_predict(g["p_lm"], g["p_uvw"], g["freq"], b, g["p_time_index"], g["p_ant1"],
                                   g["p_ant2"], None, None, g["p_die"], g["p_base_vis"], g["p_die"])
        assert np.array_equal(got, g["p_" + feed])


# ----------------------------------------------------------------------------- feed rotation (8f-1)
def test_feed_rotation_and_rotated_dde_golden(golden, oracle):
    """africanus.rime.feed_rotation (feeds.py:13-71), bit-exact for float64 and float32 angles, and
    the DDE term of rime/examples/predict.py:469-472 -- einsum("stafij,tajk->stafik", beam_cube_dde,
    feed_rot) -- carried through the fused predict, bit-exact."""
    g = golden("feeds")
    for ft in ("linear", "circular"):
        rot = oracle.feed_rotation(g["pa"], ft)
        assert rot.dtype == np.complex128 and np.array_equal(rot, g["rot_" + ft])
        rot32 = oracle.feed_rotation(g["pa"].astype(np.float32), ft)
        assert rot32.dtype == np.complex64 and np.array_equal(rot32, g["rot32_" + ft])
        bd = oracle.beam_cube_dde(g["beam"], g["ext"], g["bfm"], g["lm"], g["pa"], g["pe"], g["asc"], g["freq"])
        dde = np.einsum("stafij,tajk->stafik", bd, rot)
        assert np.array_equal(dde, g["dde_" + ft])
        vis = oracle.fused_predict(g["lm"], g["uvw"], g["freq"], g["bright"], g["time_index"], g["ant1"],
                                   g["ant2"], dde, dde, g["die"], None, g["die"])
        assert np.array_equal(vis, g["vis_" + ft])
    # known answers (feeds.py:80-95): unitary; linear is a real rotation, circular is diagonal
    lin, circ = g["rot_linear"], g["rot_circular"]
    eye = np.broadcast_to(np.eye(2), lin.shape)
    np.testing.assert_allclose(np.einsum("taij,takj->taik", lin, lin.conj()), eye, atol=1e-15)
    np.testing.assert_allclose(np.einsum("taij,takj->taik", circ, circ.conj()), eye, atol=1e-15)
    np.testing.assert_allclose(circ[..., 0, 0], np.exp(-1j * g["pa"]), rtol=1e-15)
    assert np.all(circ[..., 0, 1] == 0) and np.all(lin.imag == 0)
    with pytest.raises(ValueError):
        oracle.feed_rotation(g["pa"], "elliptical")


def _numpy_spectral(stokes, spi, ref_freq, freq, bases):
    """Independent vectorised formulation of the three spectral models."""
    st = stokes.reshape(stokes.shape[0], -1)
    sp = spi.reshape(spi.shape[0], spi.shape[1], -1)
    ratio = freq[None, :] / ref_freq[:, None]
    out = np.empty((st.shape[0], freq.shape[0], st.shape[1]))
    exps = np.arange(1, sp.shape[1] + 1)
    for p, b in enumerate(bases):
        if b in ("std", 0):
            out[:, :, p] = st[:, None, p] * np.prod(ratio[:, None, :] ** sp[:, :, None, p], axis=1)
        else:
            lr = np.log(ratio) if b in ("log", 1) else np.log10(ratio)
            poly = np.sum(sp[:, :, None, p] * lr[:, None, :] ** exps[None, :, None], axis=1)
            out[:, :, p] = st[:, None, p] * (np.exp(poly) if b in ("log", 1) else 10.0 ** poly)
    return out.reshape((st.shape[0], freq.shape[0]) + stokes.shape[1:])


@pytest.mark.parametrize("base", [0, 1, 2, "std", "log", "log10", ["log", "std", "std", "std"]])
@pytest.mark.parametrize("npol", [0, 1, 2, 4])
def test_spectral_model_multiple_spi(oracle, base, npol):
    """The reference's own parametrisation (model/spectral/tests/test_spectral_model.py:41-76):
    6 spectral indices, 0/1/2/4 polarisations, every base spelling, broadcast (strided) stokes."""
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
    got = oracle.spectral_model(stokes, spi, ref_freq, freq, base=base)
    bases = list(base) if isinstance(base, list) else [base]
    bases = (bases + [bases[-1]] * max(npol, 1))[:max(npol, 1)]
    ref = _numpy_spectral(np.asarray(stokes), spi, ref_freq, freq, bases)
    assert got.shape == ref.shape and got.flags.c_contiguous
    np.testing.assert_allclose(got, ref, rtol=1e-12)


_SCHEMA_CASES = [
    ([["XX"], ["YY"]], ["I", "Q"]),
    (["XX", "YY"], ["I", "Q"]),
    (["XX", "XY", "YX", "YY"], ["I", "Q", "U", "V"]),
    ([["XX", "XY"], ["YX", "YY"]], [["I", "Q"], ["U", "V"]]),
    (["I", "Q", "U", "V"], ["XX", "XY", "YX", "YY"]),
    ([["I", "Q"], ["U", "V"]], [["XX", "XY"], ["YX", "YY"]]),
    ([["I", "Q"], ["U", "V"]], [["XX", "XY", "YX", "YY"]]),
    ([["I", "Q"], ["U", "V"]], [["RR", "RL", "LR", "LL"]]),
    (["I", "V"], ["RR", "LL"]),
    (["I", "Q"], ["XX", "YY"]),
    ([9, 12], [1, 2]),
]


@pytest.mark.parametrize("input_schema, output_schema", _SCHEMA_CASES)
@pytest.mark.parametrize("vis_shape", [(10, 5, 3), (6, 8), (15,)])
def test_conversion_schemas(oracle, input_schema, output_schema, vis_shape):
    """The reference's schema cases (model/coherency/tests/test_convert.py:12-66): output shape, and
    the same resolution by the product's host logic (codex_africanus_b200.model.coherency)."""
    from codex_africanus_b200.model import coherency as coh

    in_shape = np.asarray(input_schema).shape
    out_shape = np.asarray(output_schema).shape
    vis = np.arange(1.0, np.prod(vis_shape + in_shape) + 1.0).reshape(vis_shape + in_shape)
    got = oracle.convert(vis, input_schema, output_schema)
    assert got.shape == vis_shape + out_shape
    in_names, ishape = coh.schema_elements(input_schema)
    out_names, oshape = coh.schema_elements(output_schema)
    assert ishape == in_shape and oshape == out_shape
    s1, s2, op = coh.resolve(in_names, out_names, False)
    # apply the product's resolved mapping with numpy and compare with the oracle
    flat = vis.reshape(-1, len(in_names)).astype(np.complex128)
    fns = [lambda a, b: a + b, lambda a, b: a - b, lambda a, b: a + 1j * b, lambda a, b: a - 1j * b,
           lambda a, b: (a + b) / 2, lambda a, b: (a - b) / 2, lambda a, b: (a - b) / 2j]
    ref = np.stack([fns[op[o]](flat[:, s1[o]], flat[:, s2[o]]) for o in range(len(out_names))], axis=1)
    ref = ref.reshape(vis_shape + out_shape)
    assert got.dtype == coh._output_dtype(vis.dtype, list(op))
    np.testing.assert_array_equal(got, ref if np.iscomplexobj(got) else ref.real)


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
def test_predict_vis_equals_reference_corrupt_vis(golden, oracle, tag):
    """predict_vis against the reference's independent corrupt_vis (calibration/utils/
    corrupt_vis.py:58-103) on its own cross-check inputs -- int16 antenna columns, antenna1 >
    antenna2, transposed (non-contiguous) Jones and model views -- at the reference's decimal 10."""
    g = golden("corrupt_vis")
    ti, a1, a2, jones, model = _corrupt_vis_case(g, tag)
    got = oracle.predict_vis(ti, a1, a2, dde1_jones=jones, source_coh=model, dde2_jones=jones)
    assert_array_almost_equal(got, g[tag + "_vis"], decimal=10)
