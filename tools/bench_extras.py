"""
Secondary timings reported under "extra" in bench.py's JSON line: every other entry of the
hot path on device-resident synthetic inputs, CUDA-event timed
(1 warm-up + best of 2).  Kept short so the default bench finishes in a few minutes.
"""
import os
import sys

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import synth  # noqa: E402
from codex_africanus_b200 import dft, rime  # noqa: E402
from codex_africanus_b200 import distributed as D  # noqa: E402


def _timed(fn, reps=2):
    fn()
    torch.cuda.synchronize()
    best = 1e30
    for _ in range(reps):
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        out = fn()
        e1.record()
        torch.cuda.synchronize()
        best = min(best, e0.elapsed_time(e1) * 1e-3)
        del out
    return best


def run(dev, fp64_peak):
    rng = np.random.default_rng(3)
    res = {}

    def T(a):
        return torch.from_numpy(np.ascontiguousarray(a)).to(dev)

    def compute_entry(name, seconds, terms, flop_per_term, peak=None, note=None):
        e = {"Gterms_per_s": terms / seconds / 1e9, "ms": 1e3 * seconds, "terms": terms,
             "flop_per_term": flop_per_term}
        if peak:
            e["frac_of_fp64_fma_peak"] = flop_per_term * terms / seconds / peak
        if note:
            e["note"] = note
        res[name] = e

    def bw_entry(name, seconds, nbytes, terms=None, note=None):
        e = {"GB_per_s": nbytes / seconds / 1e9, "ms": 1e3 * seconds, "algorithmic_bytes": nbytes}
        if terms:
            e["Gterms_per_s"] = terms / seconds / 1e9
        if note:
            e["note"] = note
        res[name] = e

    # ---- configs[1] (100-timestep slice): adjoint and FP32 variants
    na, ntime, nchan, nsrc = 64, 100, 256, 10000
    uvw, tidx, a1, a2 = synth.uvw_tracks(na, ntime, rng, ntime_total=1000)
    freq = synth.frequencies(nchan)
    lm = synth.sky_lm(nsrc, rng)
    image = synth.stokes_image(nsrc, nchan, 1, rng, freq)
    d_uvw, d_lm, d_freq, d_img = T(uvw), T(lm), T(freq), T(image)
    terms = float(nsrc) * uvw.shape[0] * nchan
    vis = dft.im_to_vis(d_img, d_uvw, d_lm, d_freq)
    flags = (torch.rand(vis.shape, device=dev) < 0.05)
    t = _timed(lambda: dft.vis_to_im(vis, d_uvw, d_lm, d_freq, flags))
    compute_entry("vis_to_im_f64_cfg2_100steps", t, terms, 11, fp64_peak, "5% flags")
    # FP32 FMA peak measured like the FP64 one (dependent-free FFMA chains on every SM)
    import ctypes
    from codex_africanus_b200 import _lib as _l32
    pk32 = ctypes.c_double()
    _l32.check(_l32.lib().afr_measure_fma_peak(0, 20000, ctypes.byref(pk32), None))
    fp32_peak = pk32.value
    t = _timed(lambda: dft.im_to_vis(d_img, d_uvw, d_lm, d_freq, dtype=np.complex64))
    compute_entry("im_to_vis_c64_cfg2_100steps", t, terms, 11,
                  note="FP32 rotation/accumulate, FP64 phase + anchors")
    res["im_to_vis_c64_cfg2_100steps"].update(
        frac_of_fp32_fma_peak=11 * terms / t / fp32_peak, fp32_fma_peak_TFLOPs=fp32_peak / 1e12)
    t = _timed(lambda: dft.vis_to_im(vis, d_uvw, d_lm, d_freq, flags, dtype=np.float32))
    compute_entry("vis_to_im_f32_cfg2_100steps", t, terms, 11)
    res["vis_to_im_f32_cfg2_100steps"].update(frac_of_fp32_fma_peak=11 * terms / t / fp32_peak)
    image4 = synth.stokes_image(nsrc, nchan, 4, rng, freq)
    d_img4 = T(image4)
    t = _timed(lambda: dft.im_to_vis(d_img4, d_uvw[: uvw.shape[0] // 4], d_lm, d_freq))
    compute_entry("im_to_vis_c128_ncorr4_cfg2_25steps", t, terms / 4, 23, fp64_peak)
    del vis, flags, d_img4, d_img

    # ---- configs[4] slice: vis_to_im onto a 1024 x 1024 image (4" cells), 64 chan, 2 timesteps
    npix = 1024
    cell = 4.0 / 3600.0 * np.pi / 180.0
    gl = (np.arange(npix) - npix // 2) * cell
    lm5 = np.stack(np.meshgrid(gl, gl, indexing="ij"), axis=-1).reshape(-1, 2)
    rows5 = 2 * (uvw.shape[0] // ntime)
    freq5 = synth.frequencies(64)
    vis5 = torch.randn((rows5, 64, 1), dtype=torch.complex128, device=dev)
    flags5 = (torch.rand(vis5.shape, device=dev) < 0.05)
    d_lm5, d_f5 = T(lm5), T(freq5)
    t = _timed(lambda: dft.vis_to_im(vis5, d_uvw[:rows5], d_lm5, d_f5, flags5))
    compute_entry("vis_to_im_f64_cfg5_slice_1024x1024", t, float(npix * npix) * rows5 * 64, 11, fp64_peak,
                  "1,048,576 pixels x 4032 rows (2 of 1550 timesteps) x 64 chan, 5% flags; the "
                  "multi-GPU run sums the per-rank (npix,chan) partials with one NCCL all_reduce")
    del vis5, flags5, d_lm5

    # ---- configs[0]: fused point-source predict, 64 ant x 100 times, 64 chan, 100 src, 2x2
    nchan1, nsrc1 = 64, 100
    freq1 = synth.frequencies(nchan1)
    lm1 = synth.sky_lm(nsrc1, rng)
    bright = synth.brightness_2x2(nsrc1, nchan1, rng, freq1)
    d_b, d_f1, d_lm1 = T(bright), T(freq1), T(lm1)
    d_t, d_a1, d_a2 = T(tidx), T(a1), T(a2)
    terms1 = float(nsrc1) * uvw.shape[0] * nchan1
    t = _timed(lambda: rime.fused_predict_vis(d_lm1, d_uvw, d_f1, d_b, d_t, d_a1, d_a2))
    compute_entry("fused_point_predict_c128_cfg1", t, terms1, 39, fp64_peak)
    die = synth.gains(ntime, na, nchan1, rng)
    d_die = T(die)
    t = _timed(lambda: rime.fused_predict_vis(d_lm1, d_uvw, d_f1, d_b, d_t, d_a1, d_a2,
                                              die1_jones=d_die, die2_jones=d_die))
    compute_entry("fused_point_predict_die_c128_cfg1", t, terms1, 39, fp64_peak)

    # ---- configs[3] block: SKA-Mid 197 antennas (19306 rows = one timestep), 4096 chan, 500 src
    na4, nchan4, nsrc4 = 197, 4096, 500
    uvw4, tidx4, a14, a24 = synth.uvw_tracks(na4, 1, rng, ntime_total=1000, max_radius=150e3)
    freq4 = synth.frequencies(nchan4)
    lm4 = synth.sky_lm(nsrc4, rng)
    b4 = synth.brightness_2x2(nsrc4, nchan4, rng, freq4)
    d_uvw4, d_f4, d_lm4, d_b4 = T(uvw4), T(freq4), T(lm4), T(b4)
    d_t4, d_a14, d_a24 = T(tidx4), T(a14), T(a24)
    terms4 = float(nsrc4) * uvw4.shape[0] * nchan4
    t = _timed(lambda: rime.fused_predict_vis(d_lm4, d_uvw4, d_f4, d_b4, d_t4, d_a14, d_a24))
    compute_entry("fused_point_predict_c128_cfg4_block", t, terms4, 39, fp64_peak,
                  "one row block (1 timestep, 19306 rows) x 4096 chan x 500 sources of configs[3]")
    del d_b4
    # the same block with the brightness generated on the device from (stokes, spi, ref_freq)
    # (SURVEY.md 8f-2), default chunking (one chunk here) and forced 128-source chunks (4 brightness
    # kernels + 4 predict passes, each re-reading and re-writing the 5 GB accumulator)
    stokes4 = np.stack([np.abs(rng.standard_normal(nsrc4))] + [0.1 * rng.standard_normal(nsrc4)] * 3, axis=1)
    spi4 = np.full((nsrc4, 1, 4), -0.7)
    d_st4, d_spi4, d_rf4 = T(stokes4), T(spi4), T(np.full(nsrc4, 1.284e9))
    t = _timed(lambda: rime.fused_predict_vis_stokes(d_lm4, d_uvw4, d_f4, d_st4, d_spi4, d_rf4, d_t4, d_a14,
                                                     d_a24))
    compute_entry("fused_stokes_predict_c128_cfg4_block", t, terms4, 39, fp64_peak,
                  "same block, brightness from (stokes, spi, ref_freq) on the device, default chunking")
    t = _timed(lambda: rime.fused_predict_vis_stokes(d_lm4, d_uvw4, d_f4, d_st4, d_spi4, d_rf4, d_t4, d_a14,
                                                     d_a24, source_chunk=128))
    compute_entry("fused_stokes_predict_c128_cfg4_block_128src_chunks", t, terms4, 39, fp64_peak,
                  "same, forced 128-source chunks: 4 passes over the 5 GB accumulator")
    # row-block streaming driver, host buffers in and out: 3 timesteps of the SKA-Mid layout x 1024
    # channels, one timestep (19306 rows, 1.27 GB of visibilities) per block, D2H of block i under
    # the compute of block i+1
    from codex_africanus_b200.rime.stream import stream_predict_vis_stokes
    uvw4s, tidx4s, a14s, a24s = synth.uvw_tracks(na4, 3, rng, ntime_total=1000, max_radius=150e3)
    freq4s = synth.frequencies(1024)

    nss = 2000  # enough sources that a block's compute (~65 ms) exceeds its D2H copy (~25 ms)
    lm4s = synth.sky_lm(nss, rng)
    stokes4s, spi4s = np.resize(stokes4, (nss, 4)), np.resize(spi4, (nss, 1, 4))

    def stream_all():
        tot = 0.0
        for _, blk in stream_predict_vis_stokes(lm4s, uvw4s, freq4s, stokes4s, spi4s, np.full(nss, 1.284e9),
                                                tidx4s, a14s, a24s, rows_per_block=uvw4.shape[0]):
            tot += float(blk[0, 0, 0, 0].real)
        return tot

    t = _timed(stream_all, reps=1)
    compute_entry("stream_stokes_predict_c128_ska_3blocks_e2e", t, float(nss) * uvw4s.shape[0] * 1024, 39,
                  fp64_peak, "2000 sources; numpy in -> numpy blocks out (3 x 1.27 GB), H2D + D2H inside the "
                  "timed region")
    from codex_africanus_b200 import model
    nsb = 20000
    d_stb, d_spib, d_rfb = (T(np.resize(a, (nsb,) + a.shape[1:])) for a in (stokes4, spi4, np.full(nsrc4, 1.284e9)))
    t = _timed(lambda: model.stokes_brightness(d_stb, d_spib, d_rfb, d_f4))
    bw_entry("stokes_brightness_c128", t, nsb * nchan4 * 64.0,
             note="20000 sources x 4096 chan -> (s,f,2,2) c128, std base, 1 spectral index; bytes = output")

    # ---- pure store reference for the phase_delay number: torch fill of the same bytes
    fill = torch.empty(2 << 30, dtype=torch.uint8, device=dev)
    t = _timed(lambda: fill.fill_(1))
    res["device_fill_reference"] = {"GB_per_s": fill.numel() / t / 1e9, "ms": 1e3 * t,
                                    "note": "torch fill_ of 2 GiB: write-only HBM rate on this GPU"}
    del fill

    # ---- un-fused building blocks on a 10-timestep slice (memory-bound by construction)
    rows10 = 10 * (uvw.shape[0] // ntime)
    K = rime.phase_delay(d_lm1, d_uvw[:rows10], d_f1)
    t = _timed(lambda: rime.phase_delay(d_lm1, d_uvw[:rows10], d_f1))
    bw_entry("phase_delay_c128_store", t, K.numel() * 16, terms=float(K.numel()),
             note="16 B/term store-bound")
    coh = torch.einsum("srf,sfij->srfij", K, d_b).contiguous()
    del K
    t = _timed(lambda: rime.predict_vis(d_t[:rows10], d_a1[:rows10], d_a2[:rows10], None, coh,
                                        None, d_die[:10], None, d_die[:10]))
    bw_entry("predict_vis_c128_2x2_materialised_coh", t, coh.numel() * 16, terms=float(coh.numel() // 4),
             note="64 B/term load-bound")
    del coh

    # ---- wsclean_predict: 2000 components, 20 % Gaussian, 100 timesteps x 256 chan
    nsw = 2000
    st = np.where(rng.random(nsw) < 0.8, "POINT", "GAUSSIAN")
    arcsec = np.pi / 180.0 / 3600.0
    gshape = np.stack([rng.uniform(10, 90, nsw) * arcsec, rng.uniform(3, 10, nsw) * arcsec,
                       rng.uniform(0, np.pi, nsw)], axis=1)
    lmw = synth.sky_lm(nsw, rng)
    fluxw = np.abs(rng.standard_normal(nsw)) + 0.1
    coefw = rng.standard_normal((nsw, 2)) * 0.2
    lpw = rng.random(nsw) < 0.5
    rfw = np.full(nsw, 1.284e9)
    freqw = synth.frequencies(256)
    rows_w = min(uvw.shape[0], 100 * (uvw.shape[0] // ntime))
    wargs = (d_uvw[:rows_w], T(lmw), st, T(fluxw), T(coefw), lpw, T(rfw), T(gshape), T(freqw))
    t = _timed(lambda: rime.wsclean_predict(*wargs))
    terms_w = float(nsw) * rows_w * 256
    res["wsclean_predict_c128_2000src_20pct_gauss"] = {
        "Gterms_per_s": terms_w / t / 1e9, "ms": 1e3 * t, "terms": terms_w,
        "note": "POINT sources on the phasor-stream kernel, GAUSSIAN ones one sincos + exp per term"}
    st_p = np.full(nsw, "POINT")
    t = _timed(lambda: rime.wsclean_predict(wargs[0], wargs[1], st_p, *wargs[3:]))
    res["wsclean_predict_c128_2000src_points_only"] = {
        "Gterms_per_s": terms_w / t / 1e9, "ms": 1e3 * t, "terms": terms_w, "flop_per_term": 11,
        "frac_of_fp64_fma_peak": 11 * terms_w / t / fp64_peak}

    # ---- configs[2] slice: beam_cube_dde -> fused DIE+DDE predict, 4096 chan, 1 timestep
    nchan3, nsrc3 = 4096, 96
    freq3 = synth.frequencies(nchan3)
    lm3 = synth.sky_lm(nsrc3, rng)
    beam, ext, bfreq = synth.beam_cube(257, 64, rng)
    d_beam = T(beam)
    pa = rng.uniform(-0.3, 0.3, (1, na))
    perr = np.zeros((1, na, nchan3, 2))
    ascale = np.ones((na, nchan3, 2))
    d_pa, d_pe, d_as, d_f3, d_lm3 = T(pa), T(perr), T(ascale), T(freq3), T(lm3)
    dde = rime.beam_cube_dde(d_beam, ext, bfreq, d_lm3, d_pa, d_pe, d_as, d_f3)
    t = _timed(lambda: rime.beam_cube_dde(d_beam, ext, bfreq, d_lm3, d_pa, d_pe, d_as, d_f3))
    bw_entry("beam_cube_dde_c128_257x257x64", t, dde.numel() * 16,
             note="bytes = output only; 8 corner gathers/elem from a 270 MB cube")
    rows3 = uvw.shape[0] // ntime
    bright3 = synth.brightness_2x2(nsrc3, nchan3, rng, freq3)
    die3 = synth.gains(1, na, nchan3, rng)
    d_b3, d_die3 = T(bright3), T(die3)
    tz = torch.zeros(rows3, dtype=torch.int32, device=dev)
    terms3 = float(nsrc3) * rows3 * nchan3
    t = _timed(lambda: rime.fused_predict_vis(d_lm3, d_uvw[:rows3], d_f3, d_b3, tz, d_a1[:rows3],
                                              d_a2[:rows3], dde, dde, d_die3, None, d_die3))
    e = {"Gterms_per_s": terms3 / t / 1e9, "ms": 1e3 * t, "terms": terms3, "flop_per_term": 95,
         "frac_of_fp64_fma_peak": 95 * terms3 / t / fp64_peak,
         "dde_GB_per_s_if_read_once": dde.numel() * 16 / t / 1e9,
         "dde_GB_per_s_gathered": 128 * terms3 / t / 1e9,
         "note": "configs[2] slice: 64 ant, 1 timestep, 4096 chan, %d src; DDE %.2f GB" % (
             nsrc3, dde.numel() * 16 / 1e9)}
    from codex_africanus_b200 import _lib as _l
    e["kernel"] = {2: "fused_dde_ws_kernel, antenna-phasor mode (2048 rows x 1 chan per CTA)",
                   3: "fused_dde_ws_kernel, per-row phasor mode", 4: "fused_dde_tiled_kernel",
                   5: "fused_dde_kernel (gather)",
                   6: "fused_dde_mma_kernel, antenna-phasor mode as a DMMA GEMM per (time, chan)"}.get(_l.lib().afr_last_fused_path(), "?")
    res["fused_dde_predict_c128_cfg3_slice"] = e
    # ---- configs[2] proper: 1000 sources x 4 timesteps x 4096 chan, DDEs from the beam cube in
    # source chunks (the full DDE array would be 67 GB; device memory holds one 4 GiB chunk)
    nt3, ns3 = 4, 1000
    rows34 = nt3 * rows3
    lm34 = synth.sky_lm(ns3, rng)
    b34 = T(synth.brightness_2x2(ns3, nchan3, rng, freq3))
    die34 = T(synth.gains(nt3, na, nchan3, rng))
    pa4 = T(rng.uniform(-0.3, 0.3, (nt3, na)))
    pe4 = torch.zeros((nt3, na, nchan3, 2), dtype=torch.float64, device=dev)
    t34 = d_t[:rows34].contiguous()
    torch.cuda.reset_peak_memory_stats(dev)
    tt = _timed(lambda: rime.fused_predict_vis_beam(T(lm34), d_uvw[:rows34], d_f3, b34, t34, d_a1[:rows34],
                                                    d_a2[:rows34], d_beam, ext, bfreq, pa4, pe4, d_as,
                                                    die34, None, die34), reps=1)
    terms34 = float(ns3) * rows34 * nchan3
    res["fused_beam_predict_c128_cfg3_4steps_1000src"] = {
        "Gterms_per_s": terms34 / tt / 1e9, "ms": 1e3 * tt, "terms": terms34, "flop_per_term": 95,
        "frac_of_fp64_fma_peak": 95 * terms34 / tt / fp64_peak,
        "torch_peak_device_GB": torch.cuda.max_memory_allocated(dev) / 1e9,
        "full_dde_array_GB": ns3 * nt3 * na * nchan3 * 64 / 1e9,
        "note": "beam_cube_dde sampling per source chunk + fused DIE/DDE predict, timed together"}
    del b34, die34

    # the same slice with the antenna decomposition disabled (what arbitrary uvw get)
    os.environ["AFR_DDE_ANT"] = "0"
    try:
        t = _timed(lambda: rime.fused_predict_vis(d_lm3, d_uvw[:rows3], d_f3, d_b3, tz, d_a1[:rows3],
                                                  d_a2[:rows3], dde, dde, d_die3, None, d_die3))
    finally:
        del os.environ["AFR_DDE_ANT"]
    res["fused_dde_predict_c128_cfg3_slice_row_phasors"] = {
        "Gterms_per_s": terms3 / t / 1e9, "ms": 1e3 * t, "terms": terms3, "flop_per_term": 95,
        "frac_of_fp64_fma_peak": 95 * terms3 / t / fp64_peak}
    return res
