"""
First-class benchmark entries for the BASELINE configs other than the headline (bench.py reports
them under "configs"): every entry is timed like the headline -- >= 3 warm-up and >= 5 timed steps
per default, a 256 MiB L2 flush before each step, CUDA events on the launching stream, max over
ranks -- and carries its own `roofline`, `e2e` (public API, host buffers in and out, copies inside
the timed region), `cpu_baseline` (the oracle's C port on a bounded sample of the same workload)
and `parity` (the timed call's output against the oracle on a row subsample, project gate
rtol = 1e-10 / atol = 1e-10 max|ref|).

  configs[0]  fused point-source predict, 64 antennas x 100 times, 64 chan, 100 sources, 2x2
  configs[2]  full RIME: beam_cube_dde (257 x 257 x 64 cube) -> DDE -> predict with DIE gains,
              64 antennas, 4096 chan, 1000 sources, 4 timesteps
  configs[3]  SKA-Mid predict through the row-block streaming driver
              (distributed.sharded_stream_predict_vis_stokes): 197 antennas, 4096 chan, 10^4 sources,
              2 whole-timestep blocks per rank
  configs[4]  (N >= 2) vis_to_im onto 1024 x 1024 pixels from 3,124,800 rows x 64 chan, rows
              sharded over the ranks (strong scaling), one NCCL all_reduce of the partial images

The oracle is used as the checker and as the CPU baseline only.
"""
import os
import statistics
import sys
import time

import numpy as np
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
sys.path.insert(0, ROOT)

import synth  # noqa: E402
from codex_africanus_b200 import _lib, dft, rime  # noqa: E402
from codex_africanus_b200 import distributed as D  # noqa: E402

UNIT = "Gterms/s"


def _dist():
    import torch.distributed as dist

    return dist if dist.is_available() and dist.is_initialized() else None


def _max_over_ranks(x, dev):
    dist = _dist()
    if dist is None:
        return x
    t = torch.tensor([x], dtype=torch.float64, device=dev)
    dist.all_reduce(t, op=dist.ReduceOp.MAX)
    return float(t.item())


def _barrier():
    dist = _dist()
    if dist is not None:
        dist.barrier()


def timed_steps(fn, steps, warmup, flush, dev):
    """Mean seconds per step over `steps` (device time, max over ranks) and the last result."""
    out = None
    for _ in range(warmup):
        out = None
        flush.zero_()
        out = fn()
    torch.cuda.synchronize()
    _barrier()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)) for _ in range(steps)]
    for k in range(steps):
        out = None
        flush.zero_()
        ev[k][0].record()
        out = fn()
        ev[k][1].record()
    torch.cuda.synchronize()
    sec = statistics.mean(a.elapsed_time(b) for a, b in ev) * 1e-3
    _barrier()
    return _max_over_ranks(sec, dev), out


def timed_host(fn, steps, dev):
    """End-to-end seconds per step (wall clock around host-buffer calls, max over ranks)."""
    out = fn()  # warm-up: page-locks the staging buffers
    out = None
    torch.cuda.synchronize()
    _barrier()
    t0 = time.perf_counter()
    for _ in range(steps):
        out = None
        out = fn()
    torch.cuda.synchronize()
    sec = (time.perf_counter() - t0) / steps
    _barrier()
    return _max_over_ranks(sec, dev), out


def parity(got, ref, what):
    scale = float(np.max(np.abs(ref))) if ref.size else 0.0
    err = float(np.max(np.abs(got - ref)) / scale) if scale > 0 else 0.0
    ok = bool(np.allclose(got, ref, rtol=1e-10, atol=1e-10 * scale))
    return {"ok": ok, "max_abs_err_over_max_ref": err, "gate": "allclose(rtol=1e-10, atol=1e-10*max|ref|)",
            "checked": what}


def cpu_sample(fn, terms, seconds_target, what):
    """fn() = the oracle port on a bounded sample of `terms` terms; repeated to ~seconds_target."""
    import oracle

    oracle.build()
    oracle.set_threads(os.cpu_count() or 1)
    fn()
    n, t0 = 0, time.perf_counter()
    while True:
        fn()
        n += 1
        dt = time.perf_counter() - t0
        if dt >= seconds_target or n >= 50:
            break
    cores = min(os.cpu_count() or 1, oracle.max_threads() or (os.cpu_count() or 1))
    return {"value": terms * n / dt / 1e9, "unit": UNIT, "cores": cores, "kind": "port",
            "sample": "%s, %d repeat(s), %.1f s" % (what, n, dt)}


def roofline(flop_per_term, terms, kernel_s, fp64_peak, kernel, executed_flop_per_term=None, note=None):
    ach = flop_per_term * terms / kernel_s / 1e12
    r = {"bound": "fp64", "kernel": kernel, "achieved": ach, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
         "frac": ach / (fp64_peak / 1e12), "algorithmic_flop_per_term": flop_per_term,
         "terms_per_step": terms, "kernel_ms": 1e3 * kernel_s,
         "peak_source": "afr_measure_fma_peak in this run (DFMA chains on every SM)", "traffic": None}
    if executed_flop_per_term:
        r["executed_flop_per_term"] = executed_flop_per_term
        r["frac_executed"] = executed_flop_per_term * terms / kernel_s / fp64_peak
    if note:
        r["note"] = note
    return r


def T(a, dev):
    return torch.from_numpy(np.ascontiguousarray(a)).to(dev)


# ------------------------------------------------------------------------------------------ configs[0]
def config0(dev, fp64_peak, steps, warmup, flush, cpu_seconds):
    import oracle

    rng = np.random.default_rng(1)
    na, ntime, nchan, nsrc = 64, 100, 64, 100
    uvw, tidx, a1, a2 = synth.uvw_tracks(na, ntime, rng)
    freq = synth.frequencies(nchan)
    lm = synth.sky_lm(nsrc, rng)
    bright = synth.brightness_2x2(nsrc, nchan, rng, freq)
    terms = float(nsrc) * uvw.shape[0] * nchan
    d = [T(x, dev) for x in (lm, uvw, freq, bright, tidx, a1, a2)]
    sec, out = timed_steps(lambda: rime.fused_predict_vis(*d), steps, warmup, flush, dev)
    path = _lib.describe_dft_path()
    rows = np.linspace(0, uvw.shape[0] - 1, 48).astype(np.int64)
    par = parity(out[T(rows, dev)].cpu().numpy(),
                 oracle.fused_predict(lm, uvw[rows], freq, bright, tidx[rows], a1[rows], a2[rows]),
                 "48 rows spread over the track, all channels and sources")
    out = None
    e2e_s, res = timed_host(lambda: rime.fused_predict_vis(lm, uvw, freq, bright, tidx, a1, a2), steps, dev)
    h2d = sum(x.nbytes for x in (lm, uvw, freq, bright, tidx, a1, a2))
    d2h = res.nbytes
    res = None
    crow = np.linspace(0, uvw.shape[0] - 1, 4032).astype(np.int64)
    cpu = cpu_sample(lambda: oracle.fused_predict(lm, uvw[crow], freq, bright, tidx[crow], a1[crow], a2[crow]),
                     float(nsrc) * crow.size * nchan, cpu_seconds, "4032 of %d rows" % uvw.shape[0])
    return {
        "workload": "configs[0] fused point-source predict: %d antennas x %d times (%d rows) x %d chan x %d "
                    "sources, 2x2 corr, complex128" % (na, ntime, uvw.shape[0], nchan, nsrc),
        "value": terms / sec / 1e9, "unit": UNIT, "ms_per_step": 1e3 * sec, "steps": steps, "warmup": warmup,
        "roofline": roofline(39, terms, sec, fp64_peak, "phasor_stream kernel, complex W, ncorr=4: " + path),
        "e2e": {"value": terms / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s,
                "api": "rime.fused_predict_vis(numpy) -> numpy"},
        "cpu_baseline": cpu, "parity": par,
    }


# ------------------------------------------------------------------------------------------ configs[2]
def config2(dev, fp64_peak, steps, warmup, flush, cpu_seconds):
    import oracle

    rng = np.random.default_rng(3)
    na, ntime, nchan, nsrc = 64, 4, 4096, int(os.environ.get("BENCH_CFG2_NSRC", 1000))
    uvw, tidx, a1, a2 = synth.uvw_tracks(na, ntime, rng, ntime_total=1000)
    freq = synth.frequencies(nchan)
    lm = synth.sky_lm(nsrc, rng)
    bright = synth.brightness_2x2(nsrc, nchan, rng, freq)
    beam, ext, bfreq = synth.beam_cube(257, 64, rng)
    pa = rng.uniform(-0.3, 0.3, (ntime, na))
    perr = np.zeros((ntime, na, nchan, 2))
    ascale = np.ones((na, nchan, 2))
    die = synth.gains(ntime, na, nchan, rng)
    terms = float(nsrc) * uvw.shape[0] * nchan
    host = (lm, uvw, freq, bright, tidx, a1, a2, beam, ext, bfreq, pa, perr, ascale, die, None, die)
    dv = [None if x is None else (x if x is ext or x is bfreq else T(x, dev)) for x in host]
    dv[15] = dv[13]  # die2 is die1: one upload
    torch.cuda.reset_peak_memory_stats(dev)
    sec, out = timed_steps(lambda: rime.fused_predict_vis_beam(*dv), steps, warmup, flush, dev)
    peak_gb = torch.cuda.max_memory_allocated(dev) / 1e9
    path = {6: "fused_dde_mma_kernel (antenna phasors, source sum as a DMMA GEMM)",
            2: "fused_dde_ws_kernel, antenna mode", 3: "fused_dde_ws_kernel, per-row mode"}.get(
        _lib.lib().afr_last_fused_path(), "path %d" % _lib.lib().afr_last_fused_path())
    # parity: 24 rows of timestep 0 x every 64th channel x ALL sources; beam sampling, DDE and
    # predict are independent per channel, so the oracle runs the same chain on the channel subset
    nbl = uvw.shape[0] // ntime
    rows = np.linspace(0, nbl - 1, 24).astype(np.int64)
    fsel = np.arange(0, nchan, 64)
    o_dde = oracle.beam_cube_dde(beam, ext, bfreq, lm, pa[:1], perr[:1, :, fsel], ascale[:, fsel], freq[fsel])
    ref = oracle.fused_predict(lm, uvw[rows], freq[fsel], bright[:, fsel], tidx[rows] - tidx[0], a1[rows],
                               a2[rows], o_dde, o_dde, die[:1, :, fsel], None, die[:1, :, fsel])
    got = out[T(rows, dev)][:, T(fsel, dev)].cpu().numpy()
    par = parity(got, ref, "24 rows of timestep 0 x 64 of %d channels x all %d sources (beam -> DDE -> DIE)"
                 % (nchan, nsrc))
    del o_dde
    out = None
    # beam sampling alone (the gather-heavy stage): GB/s of DDE output
    d_one = rime.beam_cube_dde(dv[7], ext, bfreq, dv[0][:64], dv[10][:1], dv[11][:1], dv[12], dv[2])
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    d_one = rime.beam_cube_dde(dv[7], ext, bfreq, dv[0][:64], dv[10][:1], dv[11][:1], dv[12], dv[2])
    e1.record()
    torch.cuda.synchronize()
    beam_gbs = d_one.numel() * 16 / (e0.elapsed_time(e1) * 1e-3) / 1e9
    del d_one
    hsteps = max(1, min(steps, 3))
    e2e_s, res = timed_host(lambda: rime.fused_predict_vis_beam(*host), hsteps, dev)
    h2d = sum(x.nbytes for x in host[:14] if x is not None)
    d2h = res.nbytes
    res = None
    # CPU: the reference's un-fused chain on 8 sources x 1 timestep x 32 channels
    cs, cf = 8, np.arange(0, nchan, 128)
    r1 = np.arange(nbl)

    def cpu_chain():
        dd = oracle.beam_cube_dde(beam, ext, bfreq, lm[:cs], pa[:1], perr[:1, :, cf], ascale[:, cf], freq[cf])
        return oracle.fused_predict(lm[:cs], uvw[r1], freq[cf], bright[:cs][:, cf], tidx[r1] - tidx[0], a1[r1],
                                    a2[r1], dd, dd, die[:1, :, cf], None, die[:1, :, cf])

    cpu = cpu_sample(cpu_chain, float(cs) * nbl * cf.size, cpu_seconds,
                     "%d sources x 1 timestep (%d rows) x %d channels, beam sampling included" % (cs, nbl, cf.size))
    dde_bytes = float(nsrc) * ntime * na * nchan * 64
    return {
        "workload": "configs[2] full RIME: beam_cube_dde (257x257x64 cube, 2x2) -> DDE -> predict with DIE "
                    "gains, %d antennas x %d times (%d rows) x %d chan x %d sources, complex128; the DDE "
                    "array (%.0f GB) is sampled per source chunk and never exists whole" % (
                        na, ntime, uvw.shape[0], nchan, nsrc, dde_bytes / 1e9),
        "value": terms / sec / 1e9, "unit": UNIT, "ms_per_step": 1e3 * sec, "steps": steps, "warmup": warmup,
        "roofline": roofline(95, terms, sec, fp64_peak, path, executed_flop_per_term=69,
                             note="beam sampling + DDE predict + DIE application timed together; 95 flop/term is "
                                  "SURVEY 8d's figure for per-row phasors, the antenna-mode GEMM executes 64 "
                                  "(32 FMA) per term + 8 % on the diagonal tiles"),
        "hbm": {"dde_GBps_written_and_read_once": 2 * dde_bytes / sec / 1e9,
                "beam_cube_dde_alone_GBps_of_output": beam_gbs,
                "torch_peak_device_GB": peak_gb},
        "e2e": {"value": terms / e2e_s / 1e9, "unit": UNIT, "h2d_bytes_per_step": int(h2d),
                "d2h_bytes_per_step": int(d2h), "ms_per_step": 1e3 * e2e_s, "steps": hsteps,
                "api": "rime.fused_predict_vis_beam(numpy) -> numpy"},
        "cpu_baseline": cpu, "parity": par,
    }


# ------------------------------------------------------------------------------------------ configs[3]
def config3(dev, fp64_peak, steps, warmup, flush, cpu_seconds, rank, world):
    import oracle

    rng = np.random.default_rng(4)
    na, nchan = 197, 4096
    nsrc = int(os.environ.get("BENCH_CFG3_NSRC", 10000))
    nblk = 2  # whole-timestep row blocks per rank and step
    uvw, tidx, a1, a2 = synth.uvw_tracks(na, nblk * world, rng, t0=137, ntime_total=1000, max_radius=150e3)
    freq = synth.frequencies(nchan)
    lm = synth.sky_lm(nsrc, rng)
    stokes = np.stack([np.abs(rng.standard_normal(nsrc))] + [0.1 * rng.standard_normal(nsrc) for _ in range(3)], axis=1)
    spi = np.full((nsrc, 1, 4), -0.7)
    rf = np.full(nsrc, 1.284e9)
    nbl = uvw.shape[0] // (nblk * world)
    r0, r1 = D.row_shards(tidx, world)[rank]
    terms_rank = float(nsrc) * (r1 - r0) * nchan
    terms = terms_rank * world
    host = (lm, uvw, freq, stokes, spi, rf, tidx, a1, a2)
    dvc = [T(x, dev) for x in host]

    def stream(args):
        last = None
        for (b0, b1), blk in D.sharded_stream_predict_vis_stokes(*args, rows_per_block=nbl):
            last = (b0, b1, blk)
        return last

    sec, last = timed_steps(lambda: stream(dvc), steps, warmup, flush, dev)
    path = _lib.describe_dft_path()
    par = None
    if rank == 0:
        b0, b1, blk = last
        rows = np.concatenate([np.argsort(np.linalg.norm(uvw[b0:b1], axis=1))[-4:], [0, nbl // 2]]) + b0
        bright = oracle.convert(oracle.spectral_model(stokes, spi, rf, freq), ["I", "Q", "U", "V"],
                                [["XX", "XY"], ["YX", "YY"]])
        ref = oracle.fused_predict(lm, uvw[rows], freq, bright, tidx[rows], a1[rows], a2[rows])
        got = blk[T(rows - b0, dev)].cpu().numpy()
        par = parity(got, ref, "6 rows of the last block (4 longest baselines, up to 150 km) x all channels "
                                "x all %d sources" % nsrc)
        del bright
    last = None
    e2e_s, last = timed_host(lambda: stream(host), max(1, min(steps, 3)), dev)
    d2h = (r1 - r0) * nchan * 64
    last = None
    cpu = None
    if rank == 0:
        cs, cr = min(nsrc, 400), np.linspace(0, nbl - 1, 512).astype(np.int64)
        bsm = oracle.convert(oracle.spectral_model(stokes[:cs], spi[:cs], rf[:cs], freq), ["I", "Q", "U", "V"],
                             [["XX", "XY"], ["YX", "YY"]])
        cpu = cpu_sample(lambda: oracle.fused_predict(lm[:cs], uvw[cr], freq, bsm, tidx[cr], a1[cr], a2[cr]),
                         float(cs) * cr.size * nchan, cpu_seconds, "%d sources x 512 rows x %d channels" % (cs, nchan))
    return {
        "workload": "configs[3] SKA-Mid predict, streamed: %d antennas (%d baselines, 150 km) x %d whole-timestep "
                    "row blocks per rank x %d chan x %d sources, 2x2 corr, brightness from (stokes, spi, ref_freq) "
                    "on the device, %d rank(s), no collective" % (na, nbl, nblk, nchan, nsrc, world),
        "value": terms / sec / 1e9, "unit": UNIT, "ms_per_step": 1e3 * sec, "steps": steps, "warmup": warmup,
        "n_gpus": world, "scaling": "weak",
        "roofline": roofline(39, terms_rank, sec, fp64_peak, "phasor_stream kernel, complex W, ncorr=4: " + path,
                             note="per rank; device-resident inputs, blocks stay on the device"),
        "e2e": {"value": terms / e2e_s / 1e9, "unit": UNIT,
                "h2d_bytes_per_step": int(sum(x.nbytes for x in host)), "d2h_bytes_per_step": int(d2h),
                "ms_per_step": 1e3 * e2e_s,
                "api": "distributed.sharded_stream_predict_vis_stokes(numpy) -> numpy row blocks (5.06 GB each); "
                       "every rank streams its blocks to the same host concurrently"},
        "cpu_baseline": cpu, "parity": par,
    }


# ------------------------------------------------------------------------------------------ configs[4]
def config4(dev, fp64_peak, rank, world, cpu_seconds):
    """Strong scaling: the 3,124,800 rows of configs[4] split over the ranks; one timed pass (tens of
    seconds of device time at 2 GPUs) after a warm-up pass on a 16-timestep slice."""
    import oracle

    dist = _dist()
    rng = np.random.default_rng(5)
    na, ntime, nchan, npix = 64, int(os.environ.get("BENCH_CFG4_NTIME", 1550)), 64, 1024
    uvw, tidx, a1, a2 = synth.uvw_tracks(na, ntime, rng)
    cell = np.deg2rad(4.0 / 3600.0)
    x = (np.arange(npix) - npix // 2) * cell
    ll, mm = np.meshgrid(x, x, indexing="ij")
    lm = np.stack([ll.ravel(), mm.ravel()], axis=1)
    freq = synth.frequencies(nchan)
    r0, r1 = D.row_shards(tidx, world)[rank]
    g = torch.Generator(device=dev).manual_seed(100 + rank)
    vis = torch.randn((r1 - r0, nchan, 1), dtype=torch.complex128, device=dev, generator=g)
    flags = torch.rand((r1 - r0, nchan, 1), device=dev, generator=g) < 0.05
    d_uvw, d_lm, d_freq = T(uvw[r0:r1], dev), T(lm, dev), T(freq, dev)
    nw = 16 * (uvw.shape[0] // ntime)
    part = dft.vis_to_im(vis[:nw], d_uvw[:nw], d_lm, d_freq, flags[:nw])  # warm-up
    if dist is not None:
        dist.all_reduce(part, op=dist.ReduceOp.SUM)
    torch.cuda.synchronize()
    del part
    _barrier()
    e0, e1, e2 = (torch.cuda.Event(enable_timing=True) for _ in range(3))
    e0.record()
    img = dft.vis_to_im(vis, d_uvw, d_lm, d_freq, flags)
    e1.record()
    if dist is not None:
        dist.all_reduce(img, op=dist.ReduceOp.SUM)
    e2.record()
    torch.cuda.synchronize()
    total_s = _max_over_ranks(e0.elapsed_time(e2) * 1e-3, dev)
    kern_s = _max_over_ranks(e0.elapsed_time(e1) * 1e-3, dev)
    ar_wait_s = _max_over_ranks(e1.elapsed_time(e2) * 1e-3, dev)  # includes waiting for the slowest rank
    path = _lib.describe_dft_path()
    # the collective alone: the same all_reduce on a copy, ranks aligned by a barrier first
    ar_s = 0.0
    if dist is not None:
        tmp = img.clone()
        dist.barrier()
        torch.cuda.synchronize()
        e3, e4 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e3.record()
        dist.all_reduce(tmp, op=dist.ReduceOp.SUM)
        e4.record()
        torch.cuda.synchronize()
        ar_s = _max_over_ranks(e3.elapsed_time(e4) * 1e-3, dev)
        del tmp
    # parity of the ALL-REDUCED image: 3 pixels x all channels; every rank runs the oracle on its
    # own rows (its shard of the visibilities goes to the host), the partial sums are added
    pix = np.array([0, npix * (npix // 2) + npix // 2 + 7, npix * npix - 1])
    ref = oracle.vis_to_im(vis.cpu().numpy(), uvw[r0:r1], lm[pix], freq, flags.cpu().numpy())
    ref_t = torch.from_numpy(ref).to(dev)
    if dist is not None:
        dist.all_reduce(ref_t, op=dist.ReduceOp.SUM)
    par = parity(img[T(pix, dev)].cpu().numpy(), ref_t.cpu().numpy(),
                 "3 pixels (corners + centre) x all %d channels of the all-reduced image, oracle partials summed "
                 "over the %d rank(s)" % (nchan, world))
    terms = float(lm.shape[0]) * uvw.shape[0] * nchan
    return {
        "workload": "configs[4] vis_to_im: %dx%d pixels (4 arcsec) from %d rows x %d chan = %.3g visibilities, "
                    "ncorr=1, float64, 5%% flags; rows sharded over %d rank(s), one NCCL all_reduce(SUM) of the "
                    "%d MiB partial images" % (npix, npix, uvw.shape[0], nchan, float(uvw.shape[0]) * nchan, world,
                                                 lm.shape[0] * nchan * 8 >> 20),
        "value": terms / total_s / 1e9, "unit": UNIT, "ms_per_step": 1e3 * total_s, "steps": 1, "warmup": 1,
        "n_gpus": world, "scaling": "strong", "allreduce_ms": 1e3 * ar_s,
        "allreduce_in_pass_ms": 1e3 * ar_wait_s,  # all_reduce as timed inside the pass: + the wait for the slowest rank
        "allreduce_bus_GBps": (2.0 * (world - 1) / world) * lm.shape[0] * nchan * 8 / ar_s / 1e9 if world > 1 else None,
        "roofline": roofline(11, terms / world, kern_s, fp64_peak, "phasor_stream kernel, adjoint, ncorr=1: " + path,
                             note="per rank, kernel only (the all_reduce is timed separately)"),
        "parity": par,
    }


def run(dev, fp64_peak, steps, warmup, rank=0, world=1):
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)
    cpu_seconds = float(os.environ.get("BENCH_CFG_CPU_SECONDS", 4))
    res = {}

    def guard(name, fn):
        try:
            res[name] = fn()
        except Exception as exc:  # an entry must never sink the headline line
            res[name] = {"error": repr(exc)}
        torch.cuda.empty_cache()

    if world == 1:
        guard("configs[0]", lambda: config0(dev, fp64_peak, steps, warmup, flush, cpu_seconds))
        guard("configs[2]", lambda: config2(dev, fp64_peak, steps, warmup, flush, cpu_seconds))
    guard("configs[3]", lambda: config3(dev, fp64_peak, steps, warmup, flush, cpu_seconds, rank, world))
    if world > 1:
        guard("configs[4]", lambda: config4(dev, fp64_peak, rank, world, cpu_seconds))
    return res
