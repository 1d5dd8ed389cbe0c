#!/usr/bin/env python
"""
Generate golden input/output vectors for the hot path by importing the
UNMODIFIED reference (codex-africanus, numba) from /root/reference.

Run in the build container only (the reference does not travel to the GPU box):

    NUMBA_CACHE_DIR=/tmp/numba_cache python oracle/gen_golden.py

Writes tests/golden/*.npz (inputs + reference outputs).  Every case is seeded;
re-running reproduces the files bit-for-bit given the same numpy/numba.
"""
import itertools
import os
import sys

os.environ.setdefault("NUMBA_CACHE_DIR", "/tmp/numba_cache")
sys.path.insert(0, "/root/reference")

import numpy as np  # noqa: E402

from africanus.dft import im_to_vis, vis_to_im  # noqa: E402
from africanus.rime import beam_cube_dde, phase_delay, predict_vis  # noqa: E402
from africanus.rime.fast_beam_cubes import freq_grid_interp  # noqa: E402

OUT = os.path.join(os.path.dirname(os.path.abspath(__file__)), "..", "tests", "golden")
os.makedirs(OUT, exist_ok=True)


def rc(rng, shape):
    return rng.standard_normal(shape) + 1j * rng.standard_normal(shape)


def save(name, **kw):
    path = os.path.join(OUT, name + ".npz")
    np.savez_compressed(path, **kw)
    print("wrote", path, os.path.getsize(path), "bytes")


# --------------------------------------------------------------------------
def gen_phase():
    rng = np.random.default_rng(101)
    lm = rng.uniform(-0.3, 0.3, (7, 2))
    lm[3] = [0.9, 0.9]  # outside the unit disc: n clamped (phase.py:42-43)
    uvw = rng.standard_normal((33, 3)) * 2000.0
    freq = np.linspace(0.856e9, 1.712e9, 24)
    freq_nu = np.sort(rng.uniform(0.856e9, 1.712e9, 11))  # non-uniform channels
    cases = {}
    for ci, conv in enumerate(("fourier", "casa")):
        cases["f64_%s" % conv] = phase_delay(lm, uvw, freq, convention=conv)
    cases["f64_nonuniform"] = phase_delay(lm, uvw, freq_nu)
    f32 = np.float32
    lm_s, uvw_s = lm * 0.1, uvw * 0.01  # keep |p| modest for the float32 paths
    cases["f32_all"] = phase_delay(lm_s.astype(f32), uvw_s.astype(f32), freq.astype(f32))
    cases["mix_lm32"] = phase_delay(lm_s.astype(f32), uvw_s, freq)
    cases["mix_lm32_uvw32"] = phase_delay(lm_s.astype(f32), uvw_s.astype(f32), freq)
    cases["mix_uvw32"] = phase_delay(lm_s, uvw_s.astype(f32), freq)
    cases["mix_freq32"] = phase_delay(lm_s, uvw_s, freq.astype(f32))
    save("phase_delay", lm=lm, uvw=uvw, freq=freq, freq_nu=freq_nu, lm_s=lm_s, uvw_s=uvw_s, **cases)


# --------------------------------------------------------------------------
def gen_dft():
    rng = np.random.default_rng(202)
    nsrc, nrow, nchan = 13, 37, 9
    lm = rng.uniform(-0.05, 0.05, (nsrc, 2))
    uvw = rng.standard_normal((nrow, 3)) * 1500.0
    freq = np.linspace(0.9e9, 1.3e9, nchan)
    freq_nu = np.sort(rng.uniform(0.9e9, 1.3e9, nchan))
    out = dict(lm=lm, uvw=uvw, freq=freq, freq_nu=freq_nu)
    for ncorr in (1, 2, 4):
        img = rng.standard_normal((nsrc, nchan, ncorr))
        img[2] = 0.0  # exactly-zero pixels are skipped (kernels.py:64)
        out["image_r%d" % ncorr] = img
        out["i2v_r%d" % ncorr] = im_to_vis(img, uvw, lm, freq)
        vis = rc(rng, (nrow, nchan, ncorr))
        flags = rng.random((nrow, nchan, ncorr)) < 0.1
        out["vis_c%d" % ncorr] = vis
        out["flags_%d" % ncorr] = flags
        out["v2i_c%d" % ncorr] = vis_to_im(vis, uvw, lm, freq, flags)
    imgc = rc(rng, (nsrc, nchan, 2))
    out["image_c2"] = imgc
    out["i2v_c2"] = im_to_vis(imgc, uvw, lm, freq)
    out["i2v_c2_casa"] = im_to_vis(imgc, uvw, lm, freq, convention="casa")
    out["i2v_r1_nonuniform"] = im_to_vis(out["image_r1"], uvw, lm, freq_nu)
    out["i2v_r1_c64"] = im_to_vis(out["image_r1"], uvw, lm, freq, dtype=np.complex64)
    f32 = np.float32
    out["i2v_r1_in32"] = im_to_vis(out["image_r1"].astype(f32), uvw.astype(f32),
                                   lm.astype(f32), freq.astype(f32))
    out["i2v_r1_lm32"] = im_to_vis(out["image_r1"], uvw, lm.astype(f32), freq)
    visr = rng.standard_normal((nrow, nchan, 2))
    out["vis_r2"] = visr
    out["v2i_r2"] = vis_to_im(visr, uvw, lm, freq, out["flags_2"])
    out["v2i_c1_casa"] = vis_to_im(out["vis_c1"], uvw, lm, freq, out["flags_1"], convention="casa")
    out["v2i_c1_f32"] = vis_to_im(out["vis_c1"], uvw, lm, freq, out["flags_1"], dtype=np.float32)
    out["v2i_c1_nonuniform"] = vis_to_im(out["vis_c1"], uvw, lm, freq_nu, out["flags_1"])
    out["v2i_c1_in32"] = vis_to_im(out["vis_c1"].astype(np.complex64), uvw.astype(f32),
                                   lm.astype(f32), freq.astype(f32), out["flags_1"])
    save("dft", **out)


def gen_dft_padded():
    """A zero-padded image grid reaching past l^2 + m^2 = 1: n = sqrt(negative) - 1 = NaN there
    (kernels.py:54 has no clamp), but exactly-zero pixels are skipped (kernels.py:64), so the
    reference stays finite as long as every pixel outside the unit disc is zero.  Second case: one
    bright pixel outside the disc poisons (NaN) exactly the channels / correlations where it is
    non-zero."""
    rng = np.random.default_rng(212)
    npix, nrow, nchan, ncorr = 9, 21, 8, 2
    x = np.linspace(-1.2, 1.2, npix)
    ll, mm = np.meshgrid(x, x, indexing="ij")
    lm = np.stack([ll.ravel(), mm.ravel()], axis=1)
    inside = (lm**2).sum(axis=1) < 1.0
    uvw = rng.standard_normal((nrow, 3)) * 40.0
    freq = np.linspace(0.9e9, 1.3e9, nchan)
    img = rng.standard_normal((npix * npix, nchan, ncorr))
    img[~inside] = 0.0
    out = dict(lm=lm, uvw=uvw, freq=freq, image=img)
    import warnings
    with warnings.catch_warnings():
        warnings.simplefilter("ignore")
        out["i2v"] = im_to_vis(img, uvw, lm, freq)
        out["i2v_c64"] = im_to_vis(img, uvw, lm, freq, dtype=np.complex64)
        assert np.all(np.isfinite(out["i2v"]))
        bad = img.copy()
        k = int(np.flatnonzero(~inside)[3])
        bad[k, 2:5, 1] = 1.5
        out["image_bad"] = bad
        out["i2v_bad"] = im_to_vis(bad, uvw, lm, freq)
        assert np.isnan(out["i2v_bad"][:, 2:5, 1]).all() and np.isfinite(out["i2v_bad"][:, :, 0]).all()
    save("dft_padded", **out)


# --------------------------------------------------------------------------
def gen_predict():
    """The reference's 27-case matrix (rime/tests/test_predict.py:33-126) with the
    cupy test's +10 time-index offset (rime/cuda/tests/test_cuda_predict.py:42-45)."""
    rng = np.random.default_rng(303)
    s, t, a, c, r = 6, 4, 4, 5, 10
    time_idx = np.asarray([0, 0, 1, 1, 2, 2, 2, 2, 3, 3]) + 10
    ant1 = np.asarray([0, 0, 0, 0, 1, 1, 1, 2, 2, 3])
    ant2 = np.asarray([0, 1, 2, 3, 1, 2, 3, 2, 3, 3])
    out = dict(time_idx=time_idx, ant1=ant1, ant2=ant2)
    presence = [(True, True, True), (True, False, True), (False, True, False)]
    for cname, corr in (("c1", (1,)), ("c2", (2,)), ("c22", (2, 2))):
        arrs = dict(
            a1j=rc(rng, (s, t, a, c) + corr), blj=rc(rng, (s, r, c) + corr),
            a2j=rc(rng, (s, t, a, c) + corr), g1j=rc(rng, (t, a, c) + corr),
            bvis=rc(rng, (r, c) + corr), g2j=rc(rng, (t, a, c) + corr))
        for k, v in arrs.items():
            out["%s_%s" % (cname, k)] = v
        for (d1, bl, d2), (g1, bv, g2) in itertools.product(presence, presence):
            key = "%s_out_%d%d%d_%d%d%d" % (cname, d1, bl, d2, g1, bv, g2)
            out[key] = predict_vis(
                time_idx, ant1, ant2,
                arrs["a1j"] if d1 else None, arrs["blj"] if bl else None,
                arrs["a2j"] if d2 else None, arrs["g1j"] if g1 else None,
                arrs["bvis"] if bv else None, arrs["g2j"] if g2 else None)
        # complex64 arithmetic, int16 indices
        a64 = {k: v.astype(np.complex64) for k, v in arrs.items()}
        out["%s_out_c64" % cname] = predict_vis(
            time_idx.astype(np.int16), ant1.astype(np.int16), ant2.astype(np.int16),
            a64["a1j"], a64["blj"], a64["a2j"], a64["g1j"], a64["bvis"], a64["g2j"])
    save("predict_vis", **out)


# --------------------------------------------------------------------------
def gen_beam():
    rng = np.random.default_rng(404)
    lw, mh, nud = 9, 8, 5
    nsrc, ntime, nant, nchan = 5, 3, 4, 8
    beam_freq_map = np.array([0.5, 0.56, 0.7, 0.91, 1.0])
    freq = np.array([0.4, 0.5, 0.6, 0.7, 0.8, 0.9, 1.0, 1.1])
    ext = np.array([[-0.9, 0.9], [-0.8, 1.0]])
    lm = rng.uniform(-1.0, 1.0, (nsrc, 2))  # some sources fall off the cube: clamping
    pa = rng.uniform(-np.pi, np.pi, (ntime, nant))
    perr = rng.uniform(-0.05, 0.05, (ntime, nant, nchan, 2))
    ascale = rng.uniform(0.9, 1.1, (nant, nchan, 2))
    out = dict(beam_freq_map=beam_freq_map, freq=freq, ext=ext, lm=lm, pa=pa, perr=perr,
               ascale=ascale)
    out["freq_data"] = freq_grid_interp(freq, beam_freq_map)
    for cname, corr in (("c22", (2, 2)), ("c4", (4,)), ("c2", (2,)), ("c1", (1,))):
        beam = rc(rng, (lw, mh, nud) + corr)
        out["beam_" + cname] = beam
        out["dde_" + cname] = beam_cube_dde(beam, ext, beam_freq_map, lm, pa, perr, ascale, freq)
    b64 = out["beam_c22"].astype(np.complex64)
    out["dde_c22_c64"] = beam_cube_dde(b64, ext, beam_freq_map, lm, pa, perr, ascale, freq)
    # the reference's own known-answer test (rime/tests/test_fast_beams.py:43-127)
    np.random.seed(42)
    kb = np.random.random((2, 2, 2, 1)) + 1j * np.random.random((2, 2, 2, 1))
    out["ka_beam"] = kb
    out["ka_dde"] = beam_cube_dde(
        kb, np.asarray([[-1.0, 1.0], [-1.0, 1.0]]), np.asarray([0.0, 1.0]),
        np.asarray([[0.1, 0.1]]), np.zeros((1, 1)), np.zeros((1, 1, 1, 2)),
        np.ones((1, 1, 2)), np.asarray([0.3]))
    save("beam_cube_dde", **out)


# --------------------------------------------------------------------------
def gen_fused():
    """Un-fused composition the new fused kernel must reproduce
    (rime/examples/predict.py:107-134,490,522-527; recipe asserted in
    experimental/rime/fused/tests/test_rime.py:175-209)."""
    rng = np.random.default_rng(505)
    na, ntime, nchan, nsrc = 5, 3, 12, 9
    a1, a2 = np.triu_indices(na, 1)
    nbl = a1.size
    ant1 = np.tile(a1, ntime)
    ant2 = np.tile(a2, ntime)
    time_idx = np.repeat(np.arange(ntime), nbl) + 3
    nrow = ant1.size
    uvw = rng.standard_normal((nrow, 3)) * 800.0
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    freq = np.linspace(0.856e9, 1.712e9, nchan)
    out = dict(ant1=ant1, ant2=ant2, time_idx=time_idx, uvw=uvw, lm=lm, freq=freq)
    # beam-cube DDEs
    lw = mh = 17
    nud = 6
    ext = np.array([[-0.03, 0.03], [-0.03, 0.03]])
    bfm = np.linspace(0.856e9, 1.712e9, nud)
    beam = rc(rng, (lw, mh, nud, 2, 2))
    pa = rng.uniform(-0.5, 0.5, (ntime, na))
    perr = np.zeros((ntime, na, nchan, 2))
    ascale = np.ones((na, nchan, 2))
    out.update(beam=beam, ext=ext, bfm=bfm, pa=pa)
    for conv in ("fourier", "casa"):
        K = phase_delay(lm, uvw, freq, convention=conv)
        for cname, corr in (("c22", (2, 2)), ("c2", (2,)), ("c1", (1,))):
            bright = rc(rng, (nsrc, nchan) + corr)
            die = 1.0 + 0.1 * rc(rng, (ntime, na, nchan) + corr)
            bvis = rc(rng, (nrow, nchan) + corr)
            sub = "fij" if len(corr) == 2 else "fi"
            coh = np.einsum("srf,s%s->sr%s" % (sub, sub), K, bright)
            key = "%s_%s" % (conv, cname)
            out["bright_" + key] = bright
            out["die_" + key] = die
            out["bvis_" + key] = bvis
            out["point_" + key] = predict_vis(time_idx, ant1, ant2, None, coh, None, None, None, None)
            out["point_die_" + key] = predict_vis(time_idx, ant1, ant2, None, coh, None, die, bvis, die)
            if corr == (2, 2):
                dde = beam_cube_dde(beam, ext, bfm, lm, pa, perr, ascale, freq)
            else:
                dde = rc(rng, (nsrc, ntime, na, nchan) + corr)
                out["dde_" + key] = dde
            out["full_" + key] = predict_vis(time_idx, ant1, ant2, dde, coh, dde, die, bvis, die)
    out["dde_beam_c22"] = beam_cube_dde(beam, ext, bfm, lm, pa, perr, ascale, freq)
    save("fused_predict", **out)


def gen_wsclean():
    """africanus.rime.wsclean_predict (rime/wsclean_predict.py) + model.wsclean.spectra:
    (i) the reference's own test case (rime/tests/test_wsclean_predict.py:26-61, seed 42);
    (ii) a MeerKAT-like case: arcminute Gaussians, kilometre baselines, mixed log/ordinary
    polynomials, boolean and per-source log_poly."""
    from africanus.model.wsclean.spec_model import spectra
    from africanus.rime.wsclean_predict import wsclean_predict

    out = {}
    # (i) reference test inputs
    rs = np.random.RandomState(42)
    row, src, chan = 10, 21, 5
    source_sel = rs.randint(0, 2, src).astype(np.bool_)
    source_type = np.where(source_sel, "POINT", "GAUSSIAN")
    gauss_shape = rs.normal(size=(src, 3))
    uvw = rs.normal(size=(row, 3))
    lm = rs.normal(size=(src, 2)) * 1e-5
    flux = rs.normal(size=src)
    coeffs = rs.normal(size=(src, 2))
    log_poly = rs.randint(0, 2, src, dtype=np.bool_)
    flux[log_poly] = np.abs(flux[log_poly])
    coeffs[log_poly] = np.abs(coeffs[log_poly])
    freq = np.linspace(0.856e9, 2 * 0.856e9, chan)
    ref_freq = np.full(src, freq[freq.shape[0] // 2])
    out.update(t_uvw=uvw, t_lm=lm, t_source_type=source_type, t_flux=flux, t_coeffs=coeffs,
               t_log_poly=log_poly, t_ref_freq=ref_freq, t_gauss_shape=gauss_shape, t_freq=freq,
               t_spectra=spectra(flux, coeffs, log_poly, ref_freq, freq),
               t_vis=wsclean_predict(uvw, lm, source_type, flux, coeffs, log_poly, ref_freq,
                                     gauss_shape, freq))
    # (ii) MeerKAT-like
    rng = np.random.default_rng(909)
    row, src, chan = 57, 31, 48
    source_type = np.where(rng.random(src) < 0.6, "POINT", "GAUSSIAN")
    arcsec = np.pi / 180.0 / 3600.0
    gauss_shape = np.stack([rng.uniform(20, 120, src) * arcsec, rng.uniform(5, 20, src) * arcsec,
                            rng.uniform(0, np.pi, src)], axis=1)
    gauss_shape[3, 0] = 0.0  # emaj == 0 branch (wsclean_predict.py:52)
    uvw = rng.standard_normal((row, 3)) * 1500.0
    lm = rng.uniform(-0.02, 0.02, (src, 2))
    flux = np.abs(rng.standard_normal(src)) + 0.1
    coeffs = rng.standard_normal((src, 3)) * 0.3
    log_poly = rng.random(src) < 0.5
    freq = np.linspace(0.856e9, 1.712e9, chan)
    ref_freq = rng.uniform(0.9e9, 1.6e9, src)
    out.update(m_uvw=uvw, m_lm=lm, m_source_type=source_type, m_flux=flux, m_coeffs=coeffs,
               m_log_poly=log_poly, m_ref_freq=ref_freq, m_gauss_shape=gauss_shape, m_freq=freq,
               m_spectra=spectra(flux, coeffs, log_poly, ref_freq, freq),
               m_vis=wsclean_predict(uvw, lm, source_type, flux, coeffs, log_poly, ref_freq,
                                     gauss_shape, freq),
               m_vis_logpoly_true=wsclean_predict(uvw, lm, source_type, flux, coeffs, True, ref_freq,
                                                  gauss_shape, freq),
               m_vis_logpoly_false=wsclean_predict(uvw, lm, source_type, flux, coeffs, False,
                                                   ref_freq, gauss_shape, freq))
    freq_nu = np.sort(rng.uniform(0.856e9, 1.712e9, 13))  # non-uniform channels
    out.update(m_freq_nu=freq_nu,
               m_vis_nu=wsclean_predict(uvw, lm, source_type, flux, coeffs, log_poly, ref_freq,
                                        gauss_shape, freq_nu))
    save("wsclean", **out)


def gen_brightness():
    """africanus.model.spectral.spectral_model, africanus.model.coherency.convert
    and their composition into the (source, chan, 2, 2) brightness the predict consumes
    (rime/examples/predict.py:107-134), plus the reference test's own known answers
    (model/coherency/tests/test_convert.py:69-110)."""
    from africanus.model.coherency.conversion import convert
    from africanus.model.spectral.spec_model import spectral_model
    from africanus.model.wsclean.spec_model import spectra

    out = {}
    rng = np.random.default_rng(2024)
    src, chan = 23, 17
    freq = np.linspace(0.856e9, 1.712e9, chan)
    ref_freq = rng.uniform(0.9e9, 1.6e9, src)
    stokes = rng.standard_normal((src, 4)) * 0.1
    stokes[:, 0] = np.abs(rng.standard_normal(src)) + 0.5
    out.update(freq=freq, ref_freq=ref_freq, stokes=stokes)
    for nspi in (1, 2, 6):
        spi = rng.standard_normal((src, nspi, 4)) * 0.4 - 0.3
        out["spi%d" % nspi] = spi
        for base in ("std", "log", "log10"):
            out["sm_%s_%d" % (base, nspi)] = spectral_model(stokes, spi, ref_freq, freq, base=base)
        out["sm_int_%d" % nspi] = np.stack([spectral_model(stokes, spi, ref_freq, freq, base=b)
                                            for b in (0, 1, 2)])
    # per-polarisation base list, padded with its last entry (spec_model.py:77-83)
    out["sm_list"] = spectral_model(stokes, out["spi2"], ref_freq, freq, base=["std", "log", "log10"])
    # no polarisation dimension
    out["sm_nopol"] = spectral_model(stokes[:, 0].copy(), out["spi2"][:, :, 0].copy(), ref_freq, freq,
                                     base="log")
    # float32 inputs
    f32 = np.float32
    out["sm_f32"] = spectral_model(stokes.astype(f32), out["spi2"].astype(f32), ref_freq.astype(f32),
                                   freq.astype(f32), base="std")
    # the composed brightness for both feed types
    sm = out["sm_std_2"]
    out["b_linear"] = convert(sm, ["I", "Q", "U", "V"], [["XX", "XY"], ["YX", "YY"]])
    out["b_circular"] = convert(sm, ["I", "Q", "U", "V"], [["RR", "RL"], ["LR", "LL"]])
    out["b_flat"] = convert(sm, ["I", "Q", "U", "V"], ["XX", "XY", "YX", "YY"])
    out["b_implicit"] = convert(sm[..., :1], ["I"], ["XX", "XY", "YX", "YY"], implicit_stokes=True)
    out["b_diag"] = convert(sm[..., :2], ["I", "Q"], ["XX", "YY"])
    out["b_f32"] = convert(sm.astype(f32), ["I", "Q", "U", "V"], [["XX", "XY"], ["YX", "YY"]])
    # back to Stokes from complex correlations (int schema too)
    vis = rc(rng, (9, 5, 2, 2))
    out["vis"] = vis
    out["s_linear"] = convert(vis, [["XX", "XY"], ["YX", "YY"]], ["I", "Q", "U", "V"])
    out["s_circular"] = convert(vis, [["RR", "RL"], ["LR", "LL"]], [["I", "Q"], ["U", "V"]])
    out["s_int"] = convert(vis[..., 0, :], [9, 12], [1, 2])
    out["s_real"] = convert(vis.real, [["XX", "XY"], ["YX", "YY"]], ["I", "Q"])
    # WSClean spectra with 6 coefficients: integer powers >= 4 (numba int_power ordering)
    coeffs = rng.standard_normal((src, 6)) * 0.3
    log_poly = rng.random(src) < 0.5
    out.update(w_coeffs=coeffs, w_log_poly=log_poly,
               w_spectra=spectra(stokes[:, 0], coeffs, log_poly, ref_freq, freq))
    # predict with that brightness: MeerKAT-like, DIE gains, base_vis
    na, ntime = 7, 3
    a1, a2 = np.triu_indices(na, 1)
    nbl = a1.size
    ant1, ant2 = np.tile(a1, ntime).astype(np.int32), np.tile(a2, ntime).astype(np.int32)
    time_index = np.repeat(np.arange(ntime), nbl).astype(np.int32)
    pos = rng.standard_normal((ntime, na, 3)) * 800.0
    uvw = (pos[:, a1] - pos[:, a2]).reshape(-1, 3)
    lm = rng.uniform(-0.02, 0.02, (src, 2))
    die = 1.0 + 0.1 * rc(rng, (ntime, na, chan, 2, 2))
    base_vis = rc(rng, (uvw.shape[0], chan, 2, 2))
    K = phase_delay(lm, uvw, freq)
    for feed in ("linear", "circular"):
        coh = np.einsum("srf,sfij->srfij", K, out["b_" + feed])
        out["p_" + feed] = predict_vis(time_index, ant1, ant2, None, coh, None, die, base_vis, die)
    out.update(p_lm=lm, p_uvw=uvw, p_time_index=time_index, p_ant1=ant1, p_ant2=ant2, p_die=die,
               p_base_vis=base_vis)
    save("brightness", **out)


def gen_feeds():
    """africanus.rime.feed_rotation (rime/feeds.py) and the DDE term of the predict example
    (rime/examples/predict.py:469-472): einsum("stafij,tajk->stafik", beam_cube_dde, feed_rot),
    carried through phase_delay (x) brightness -> predict_vis."""
    from africanus.rime.feeds import feed_rotation

    out = {}
    rng = np.random.default_rng(4711)
    na, ntime, nsrc, nchan = 6, 3, 9, 12
    pa = rng.uniform(-np.pi, np.pi, (ntime, na))
    out["pa"] = pa
    for ft in ("linear", "circular"):
        out["rot_" + ft] = feed_rotation(pa, ft)
        out["rot32_" + ft] = feed_rotation(pa.astype(np.float32), ft)
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime).astype(np.int32), np.tile(a2, ntime).astype(np.int32)
    time_index = np.repeat(np.arange(ntime), a1.size).astype(np.int32)
    pos = rng.standard_normal((ntime, na, 3)) * 900.0
    uvw = (pos[:, a1] - pos[:, a2]).reshape(-1, 3)
    lm = rng.uniform(-0.02, 0.02, (nsrc, 2))
    freq = np.linspace(0.856e9, 1.712e9, nchan)
    beam = rc(rng, (9, 9, 5, 2, 2))
    ext = np.array([[-0.03, 0.03], [-0.03, 0.03]])
    bfm = np.linspace(0.8e9, 1.8e9, 5)
    pe = rng.uniform(-1e-3, 1e-3, (ntime, na, nchan, 2))
    asc = rng.uniform(0.9, 1.1, (na, nchan, 2))
    bright = rc(rng, (nsrc, nchan, 2, 2))
    die = 1.0 + 0.1 * rc(rng, (ntime, na, nchan, 2, 2))
    K = phase_delay(lm, uvw, freq)
    coh = np.einsum("srf,sfij->srfij", K, bright)
    beam_dde = beam_cube_dde(beam, ext, bfm, lm, pa, pe, asc, freq)
    out.update(uvw=uvw, lm=lm, freq=freq, beam=beam, ext=ext, bfm=bfm, pe=pe, asc=asc, bright=bright,
               die=die, ant1=ant1, ant2=ant2, time_index=time_index)
    for ft in ("linear", "circular"):
        dde = np.einsum("stafij,tajk->stafik", beam_dde, out["rot_" + ft])
        out["dde_" + ft] = dde
        out["vis_" + ft] = predict_vis(time_index, ant1, ant2, dde, coh, dde, die, None, die)
    save("feeds", **out)


def gen_corrupt_vis():
    """The reference's INDEPENDENT implementation of the DDE chain,
    africanus.calibration.utils.corrupt_vis (calibration/utils/corrupt_vis.py:58-103), on the inputs
    of its own cross-check against predict_vis (calibration/utils/tests/test_utils.py:21-80,
    conftest.py:30-115): int16 antenna columns with antenna1 > antenna2, (time, ant, chan, dir, corr)
    Jones and (row, chan, dir, corr) model layouts that the test transposes into predict_vis order."""
    from africanus.averaging.support import unique_time
    from africanus.calibration.utils import corrupt_vis

    out = {}
    rs = np.random.RandomState(42)
    n_dir, n_time, n_chan, n_ant = 3, 8, 6, 7
    n_bl = n_ant * (n_ant - 1) // 2
    n_row = n_bl * n_time
    antenna1 = np.zeros(n_row, dtype=np.int16)
    antenna2 = np.zeros(n_row, dtype=np.int16)
    time = np.zeros(n_row, dtype=np.float64)
    time_values = np.linspace(0, 1, n_time)
    for i in range(n_time):
        row = 0
        for p in range(n_ant):
            for q in range(p):
                time[i * n_bl + row] = time_values[i]
                antenna1[i * n_bl + row] = p
                antenna2[i * n_bl + row] = q
                row += 1
    uvw = rs.randn(n_row, 3)
    freq = np.linspace(1e9, 2e9, n_chan)
    lm = 0.1 * rs.randn(n_dir, 2)
    _, time_bin_indices, _, time_bin_counts = unique_time(time)
    out.update(antenna1=antenna1, antenna2=antenna2, time=time, uvw=uvw, freq=freq, lm=lm)
    for tag, corr_shape, jones_shape in (("c1", (1,), (1,)), ("c2", (2,), (2,)), ("d22", (2, 2), (2,)),
                                         ("f22", (2, 2), (2, 2))):
        flux = np.abs(rs.normal(size=(n_dir, 1) + corr_shape)) * (freq / freq[n_chan // 2])[None, :, None].reshape(
            (1, n_chan) + (1,) * len(corr_shape)) ** -0.7
        model = np.zeros((n_row, n_chan, n_dir) + corr_shape, dtype=np.complex128)
        for d in range(n_dir):
            tmp = im_to_vis(flux[d].reshape(1, n_chan, -1), uvw, lm[d].reshape(1, 2), freq)
            model[:, :, d] = tmp.reshape((n_row, n_chan) + corr_shape)
        jones = np.ones((n_time, n_ant, n_chan, n_dir) + jones_shape, dtype=np.complex128)
        jones += rs.normal(0.0, 0.05, jones.shape) + 1.0j * rs.normal(0.0, 0.05, jones.shape)
        vis = corrupt_vis(time_bin_indices.copy(), time_bin_counts, antenna1, antenna2, jones, model)
        out.update({tag + "_model": model, tag + "_jones": jones, tag + "_vis": vis})
    save("corrupt_vis", **out)


def gen_fused_spec():
    """The reference's fused-RIME spec front end (africanus/experimental/rime/fused/core.py:227-241,
    tests: experimental/rime/fused/tests/test_rime.py:225-297,97-219): (Kpq, Bpq) for both feed schemas, a
    two-correlation schema, the three spectral bases and both conventions; feed rotation (Lp .. Lq) and
    the beam cube (Ep .. Eq) with the parallactic-angle arrays supplied directly (the transformer that
    derives them needs casacore, absent here)."""
    from africanus.experimental.rime.fused.core import rime

    rng = np.random.default_rng(4242)
    nsrc, nspi, nchan, na, ntime = 7, 2, 6, 4, 3
    a1, a2 = np.triu_indices(na, 1)
    ant1, ant2 = np.tile(a1, ntime).astype(np.int32), np.tile(a2, ntime).astype(np.int32)
    nrow = ant1.size
    time = np.repeat(np.linspace(0.1, 1.0, ntime), a1.size)
    feed = np.zeros(nrow, np.int32)
    radec = rng.random((nsrc, 2)) * 2e-2
    phase_dir = rng.random(2) * 1e-2
    uvw = rng.standard_normal((nrow, 3)) * 300.0
    chan_freq = np.linspace(0.856e9, 2 * 0.856e9, nchan)
    stokes = rng.normal(size=(nsrc, 4))
    stokes[:, 0] = np.sqrt((stokes[:, 1:] ** 2).sum(axis=-1))
    spi = rng.random((nsrc, nspi, 4))
    ref_freq = rng.uniform(0.5 * 0.856e9, 4 * 0.856e9, nsrc)
    ds = dict(time=time, antenna1=ant1, antenna2=ant2, feed1=feed, feed2=feed, radec=radec, phase_dir=phase_dir,
              uvw=uvw, chan_freq=chan_freq, stokes=stokes, spi=spi, ref_freq=ref_freq)
    out = dict(ds)
    lin, circ = "[XX,XY,YX,YY]", "[RR,RL,LR,LL]"
    for tag, corrs, conv, base in (("kb_lin_casa_std", lin, "casa", "standard"), ("kb_circ_fourier_log", circ, "fourier", "log"),
                                   ("kb_lin_fourier_log10", lin, "fourier", "log10"), ("kb_rrll_fourier_std", "[RR,LL]", "fourier", "standard")):
        out[tag] = rime("(Kpq, Bpq): [I,Q,U,V] -> %s" % corrs, ds, convention=conv, spi_base=base)
    # parallactic angles per (time, antenna); one feed, receptor angles zero
    pa = rng.uniform(-1.0, 1.0, (ntime, na))
    feed_pa = np.empty((ntime, 1, na, 2, 2))
    feed_pa[:, 0, :, 0, 0] = feed_pa[:, 0, :, 1, 0] = np.sin(pa)
    feed_pa[:, 0, :, 0, 1] = feed_pa[:, 0, :, 1, 1] = np.cos(pa)
    beam_pa = np.stack((np.sin(pa), np.cos(pa)), axis=-1)[:, None]
    out.update(parallactic_angles=pa, feed_parangle=feed_pa, beam_parangle=beam_pa)
    for tag, corrs in (("lkbl_lin", lin), ("lkbl_circ", circ)):
        out[tag] = rime("(Lp, Kpq, Bpq, Lq): [I,Q,U,V] -> %s" % corrs, {**ds, "feed_parangle": feed_pa},
                        convention="casa", spi_base="standard")
    lw = mh = nud = 10
    beam = rc(rng, (lw, mh, nud, 4))
    ext = np.array([[-0.05, 0.05], [-0.05, 0.05]])
    bfm = np.sort(rng.uniform(chan_freq[0], chan_freq[-1], nud))
    out.update(beam=beam, beam_lm_extents=ext, beam_freq_map=bfm)
    eds = {**ds, "beam": beam, "beam_lm_extents": ext, "beam_freq_map": bfm, "beam_parangle": beam_pa}
    out["ekbe_lin"] = rime("(Ep, Kpq, Bpq, Eq): [I,Q,U,V] -> %s" % lin, eds, convention="casa", spi_base="standard")
    out["lekbel_lin"] = rime("(Lp, Ep, Kpq, Bpq, Eq, Lq): [I,Q,U,V] -> %s" % lin, {**eds, "feed_parangle": feed_pa},
                             convention="fourier", spi_base="standard")
    out["elkble_lin"] = rime("(Ep, Lp, Kpq, Bpq, Lq, Eq): [I,Q,U,V] -> %s" % lin, {**eds, "feed_parangle": feed_pa},
                             convention="fourier", spi_base="standard")
    save("fused_spec", **out)


if __name__ == "__main__":
    only = sys.argv[1:]
    if only:  # e.g. `python oracle/gen_golden.py gen_dft_padded`: (re)write the named files only
        for name in only:
            globals()[name]()
        sys.exit(0)
    gen_phase()
    gen_dft()
    gen_dft_padded()
    gen_predict()
    gen_beam()
    gen_fused()
    gen_wsclean()
    gen_brightness()
    gen_feeds()
    gen_corrupt_vis()
    gen_fused_spec()
