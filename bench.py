#!/usr/bin/env python
"""
bench.py -- headline benchmark of the B200 RIME/DFT hot path.

    python bench.py --gpus N --steps K --warmup W            (this implementation)
    python bench.py --impl reference --gpus N --steps K ...  (CPU reference arm)
    python -m torch.distributed.run --nproc-per-node N ... bench.py --gpus N ...

Metric (BASELINE.json): G source-visibility terms/s, term = one (source, row, chan) triple.

A "step" is one pass of ``im_to_vis`` (complex128, ncorr = 1) over BASELINE.json configs[1]:
64 antennas x 1000 times (2,016,000 rows) x 256 channels x 10,000 sources = 5.16e12 terms,
synthetic MeerKAT-shaped data (tools/synth.py).  With N > 1 every rank owns its own
1000-timestep row block of an N x 1000-timestep observation (rows shard with no collective
on the data path) -> weak scaling; `value` is the whole-job aggregate.

One JSON line is printed by rank 0 (see DESIGN.md "Measurement" for every field).
"""
import argparse
import ctypes
import json
import os
import statistics
import subprocess
import sys
import tempfile
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)
sys.path.insert(0, os.path.join(ROOT, "tools"))

import synth  # noqa: E402

METRIC = "G source-visibility terms/s"
UNIT = "Gterms/s"
FLOP_PER_TERM = 11  # im_to_vis, real image, ncorr=1: 7 + 4*ncorr


def env_int(name, default):
    return int(os.environ.get(name, default))


def workload(rank, world):
    """configs[1] for this rank (BENCH_* env vars shrink it for smoke runs only)."""
    na = env_int("BENCH_NA", 64)
    ntime = env_int("BENCH_NTIME", 1000)
    nchan = env_int("BENCH_NCHAN", 256)
    nsrc = env_int("BENCH_NSRC", 10000)
    rng = np.random.default_rng(2)
    uvw, tidx, a1, a2 = synth.uvw_tracks(na, ntime, rng, t0=rank * ntime, ntime_total=world * ntime)
    lm = synth.sky_lm(nsrc, rng)
    freq = synth.frequencies(nchan)
    image = synth.stokes_image(nsrc, nchan, 1, rng, freq)
    desc = ("configs[1] im_to_vis direct DFT: %d antennas x %d times (%d rows) x %d chan x %d "
            "sources, ncorr=1, complex128, per GPU" % (na, ntime, uvw.shape[0], nchan, nsrc))
    return dict(na=na, ntime=ntime, nchan=nchan, nsrc=nsrc, uvw=uvw, tidx=tidx, lm=lm, freq=freq,
                image=image, desc=desc, terms=float(nsrc) * uvw.shape[0] * nchan)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    FIELDS = ("index,clocks.sm,clocks.max.sm,power.draw,clocks_event_reasons.active,"
              "clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
              "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")

    def __init__(self, gpu_index):
        self.gpu_index = gpu_index
        self.proc = None
        self.path = None

    def start(self):
        fd, self.path = tempfile.mkstemp(prefix="clocks_", suffix=".csv")
        os.close(fd)
        try:
            self.proc = subprocess.Popen(
                ["nvidia-smi", "--query-gpu=" + self.FIELDS, "--format=csv,noheader,nounits",
                 "-lms", "200", "-i", str(self.gpu_index)],
                stdout=open(self.path, "w"), stderr=subprocess.DEVNULL)
        except OSError:
            self.proc = None

    def stop(self):
        out = {"sm_mhz": None, "sm_max_mhz": None, "reasons": [], "samples": 0}
        if self.proc is None:
            return out
        self.proc.terminate()
        try:
            self.proc.wait(timeout=5)
        except subprocess.TimeoutExpired:
            self.proc.kill()
        sm, smax, power, reasons = [], [], [], set()
        names = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]
        for line in open(self.path):
            parts = [p.strip() for p in line.split(",")]
            if len(parts) < 9:
                continue
            try:
                sm.append(float(parts[1]))
                smax.append(float(parts[2]))
                power.append(float(parts[3]))
            except ValueError:
                continue
            for name, val in zip(names, parts[5:9]):
                if val.lower().startswith("active"):
                    reasons.add(name)
        os.unlink(self.path)
        if sm:
            out.update(sm_mhz=statistics.median(sm), sm_max_mhz=max(smax), samples=len(sm),
                       power_w_max=max(power), reasons=sorted(reasons))
        return out


def physical_gpu_index(local_rank):
    vis = os.environ.get("CUDA_VISIBLE_DEVICES")
    if vis:
        ids = [v for v in vis.split(",") if v.strip()]
        if local_rank < len(ids) and ids[local_rank].strip().isdigit():
            return int(ids[local_rank])
    return local_rank


# ----------------------------------------------------------------------------- CPU arm
def cpu_im_to_vis_rate(wl, seconds_target, threads=None):
    """The reference algorithm's CPU port (oracle/afr_oracle.c, OpenMP over rows) timed on a
    bounded row sample of the same workload.  Returns (Gterms/s, cores, sample, seconds)."""
    import oracle

    oracle.build()
    threads = threads or os.cpu_count() or 1
    oracle.set_threads(threads)
    threads = min(threads, oracle.max_threads()) if oracle.max_threads() > 0 else threads
    nsrc, nchan = wl["nsrc"], wl["nchan"]
    nbl = wl["uvw"].shape[0] // wl["ntime"]

    def run(nrows):
        rows = np.linspace(0, wl["uvw"].shape[0] - 1, nrows).astype(np.int64)  # spread over the track
        uvw = np.ascontiguousarray(wl["uvw"][rows])
        t0 = time.perf_counter()
        oracle.im_to_vis(wl["image"], uvw, wl["lm"], wl["freq"])
        return time.perf_counter() - t0

    probe_rows = max(threads, 8)
    run(probe_rows)  # page in / thread start-up
    tp = run(probe_rows)
    rows = int(min(max(probe_rows, probe_rows * seconds_target / max(tp, 1e-6)), 8 * nbl))
    rows = max(threads, (rows // threads) * threads)
    t = run(rows)
    rate = nsrc * float(rows) * nchan / t / 1e9
    sample = "%d of %d rows (spread over the track) x %d chan x %d sources, %.1f s" % (
        rows, wl["uvw"].shape[0], nchan, nsrc, t)
    return rate, threads, sample, t, rows


def numba_im_to_vis_rate(wl, seconds_target, threads=None):
    """The reference's OWN numba kernel (africanus.dft.im_to_vis, dft/kernels.py:23-69, nogil) from
    baseline/_ref (tools/install_reference.py), threaded over row blocks with a ThreadPoolExecutor --
    what dask's threaded scheduler does with it.  None when the install or numba is absent."""
    ref_dir = os.path.join(ROOT, "baseline", "_ref")
    if not os.path.isdir(os.path.join(ref_dir, "africanus")):
        return None
    os.environ.setdefault("NUMBA_CACHE_DIR", os.path.join(tempfile.gettempdir(), "numba_cache_afr"))
    if ref_dir not in sys.path:
        sys.path.insert(0, ref_dir)
    try:
        from concurrent.futures import ThreadPoolExecutor

        from africanus.dft import im_to_vis as ref_im_to_vis
    except Exception:
        return None
    threads = threads or os.cpu_count() or 1
    image, lm, freq = wl["image"], wl["lm"], wl["freq"]
    nsrc, nchan = wl["nsrc"], wl["nchan"]
    ref_im_to_vis(image[:2], wl["uvw"][:2], lm[:2], freq)  # JIT

    def run(nrows, nthreads):
        rows = np.linspace(0, wl["uvw"].shape[0] - 1, nrows).astype(np.int64)
        uvw = np.ascontiguousarray(wl["uvw"][rows])
        blocks = np.array_split(np.arange(nrows), nthreads)
        t0 = time.perf_counter()
        with ThreadPoolExecutor(nthreads) as ex:
            list(ex.map(lambda b: ref_im_to_vis(image, uvw[b[0]:b[-1] + 1], lm, freq), [b for b in blocks if b.size]))
        return time.perf_counter() - t0

    t1 = run(2, 1)  # two rows on one thread: the per-core rate
    rows = int(max(threads, min(64 * threads, threads * seconds_target / max(t1 / 2, 1e-6))))
    rows = (rows // threads) * threads
    t = run(rows, threads)
    rate = nsrc * float(rows) * nchan / t / 1e9
    return {"value": rate, "unit": UNIT, "cores": threads, "kind": "reference",
            "one_core_Gterms_per_s": nsrc * 2.0 * nchan / t1 / 1e9,
            "sample": "africanus.dft.im_to_vis (numba, nogil) on %d threads: %d of %d rows x %d chan x %d "
                      "sources, %.1f s" % (threads, rows, wl["uvw"].shape[0], nchan, nsrc, t),
            "seconds": t}


def run_reference_arm(args):
    rank = env_int("RANK", 0)
    if rank != 0:
        return 0  # rank 0 alone runs the CPU arm
    wl = workload(0, 1)
    per_step = float(os.environ.get("BENCH_REF_STEP_SECONDS", 6.0))
    # the reference's own numba kernel when it is installed (baseline/_ref), else the C port
    use_numba = numba_im_to_vis_rate(wl, 0.5) is not None
    rates, secs, cores, sample, kind = [], [], 1, "", "port"
    for i in range(args.warmup + args.steps):
        if use_numba:
            r = numba_im_to_vis_rate(wl, per_step)
            rate, cores, sample, t, kind = r["value"], r["cores"], r["sample"], r["seconds"], "reference"
        else:
            rate, cores, sample, t, _ = cpu_im_to_vis_rate(wl, per_step)
        if i >= args.warmup:
            rates.append(rate)
            secs.append(t)
    value = statistics.mean(rates)
    port_rate, port_cores, port_sample, _, _ = cpu_im_to_vis_rate(wl, per_step)
    line = {
        "impl": "reference", "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": args.gpus,
        "steps": args.steps, "warmup": args.warmup, "ms_per_step": 1e3 * statistics.mean(secs),
        "higher_is_better": True, "scaling": "weak", "vs_baseline": None, "dtype": "f64",
        "data": "synthetic",
        "config": {"workload": wl["desc"],
                   "reference_arm": ("the reference's own numba africanus.dft.im_to_vis (baseline/_ref), threaded "
                                     "over row blocks on all host threads" if use_numba else
                                     "CPU port of africanus.dft.im_to_vis (oracle/afr_oracle.c, OpenMP over "
                                     "rows, all host threads)") + "; each step is a bounded row sample of the "
                                    "workload: " + sample},
        "cpu_baseline": {"value": value, "unit": UNIT, "cores": cores, "kind": kind, "sample": sample},
        "cpu_baseline_port": {"value": port_rate, "unit": UNIT, "cores": port_cores, "kind": "port",
                              "sample": port_sample},
        "e2e": {"value": value, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    emit(line)
    return 0


# ----------------------------------------------------------------------------- GPU arm
_JSON_OUT = None


def _claim_stdout():
    """The contract is ONE JSON line on stdout.  Libraries write there too (NCCL prints its
    version banner to stdout when NCCL_DEBUG is set on the box), so keep a private handle on the
    real stdout for the JSON line and point file descriptor 1 at stderr for everything else."""
    global _JSON_OUT
    if _JSON_OUT is None:
        sys.stdout.flush()
        _JSON_OUT = os.fdopen(os.dup(1), "w")
        os.dup2(2, 1)
    return _JSON_OUT


def emit(line):
    out = _claim_stdout()
    out.write(json.dumps(line) + "\n")
    out.flush()


def main():
    _claim_stdout()
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=3)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="b200", choices=["b200", "reference"])
    ap.add_argument("--no-extras", action="store_true", help="skip the secondary kernel timings")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-e2e", action="store_true")
    ap.add_argument("--no-configs", action="store_true", help="skip the configs[0]/[2]/[3]/[4] entries")
    args = ap.parse_args()
    if args.impl == "reference":
        return run_reference_arm(args)

    import torch
    import torch.distributed as dist

    from codex_africanus_b200 import _lib, dft

    world = env_int("WORLD_SIZE", 1)
    rank = env_int("RANK", 0)
    local_rank = env_int("LOCAL_RANK", 0)
    if not torch.cuda.is_available():
        raise SystemExit("bench.py: no CUDA device (the product path has no CPU fallback)")
    torch.cuda.set_device(local_rank)
    dev = torch.device("cuda", local_rank)
    if world > 1:
        os.environ.setdefault("MASTER_ADDR", "127.0.0.1")
        dist.init_process_group("nccl", device_id=dev)
    steps, warmup = args.steps, max(args.warmup, 0)

    def barrier():
        if world > 1:
            dist.barrier()

    def max_over_ranks(x):
        if world == 1:
            return x
        t = torch.tensor([x], dtype=torch.float64, device=dev)
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
        return float(t.item())

    lib = _lib.lib()
    _lib.check(lib.afr_set_device(local_rank))
    wl = workload(rank, world)
    terms = wl["terms"]

    # ---- measured FP64 FMA pipe peak (roofline denominator), before the timed region
    peak = ctypes.c_double()
    _lib.check(lib.afr_measure_fma_peak(1, 20000, ctypes.byref(peak), None))
    fp64_peak = peak.value

    # ---- device-resident inputs
    d_image = torch.from_numpy(wl["image"]).to(dev)
    d_uvw = torch.from_numpy(wl["uvw"]).to(dev)
    d_lm = torch.from_numpy(wl["lm"]).to(dev)
    d_freq = torch.from_numpy(wl["freq"]).to(dev)
    flush = torch.empty(256 << 20, dtype=torch.uint8, device=dev)  # > 126 MB L2

    def step():
        return dft.im_to_vis(d_image, d_uvw, d_lm, d_freq)

    out = None
    for _ in range(warmup):
        out = None
        flush.zero_()
        out = step()
    torch.cuda.synchronize()

    sampler = ClockSampler(physical_gpu_index(local_rank))
    sampler.start()
    launches0 = lib.afr_kernel_launches()
    ev = [(torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True))
          for _ in range(steps)]
    barrier()
    torch.cuda.synchronize()
    t_start = torch.cuda.Event(enable_timing=True)
    t_end = torch.cuda.Event(enable_timing=True)
    t_start.record()
    for k in range(steps):
        out = None
        flush.zero_()  # evict the previous step's data from L2
        ev[k][0].record()
        out = step()
        ev[k][1].record()
    t_end.record()
    torch.cuda.synchronize()
    barrier()
    launches = int(lib.afr_kernel_launches() - launches0)
    clocks = sampler.stop()
    total_s = max_over_ranks(t_start.elapsed_time(t_end) * 1e-3)
    kernel_s = statistics.mean(a.elapsed_time(b) for a, b in ev) * 1e-3
    value = world * terms * steps / total_s / 1e9
    checksum = float(torch.view_as_real(out[:1024]).abs().sum().item())
    dft_path = _lib.describe_dft_path()
    # parity of the timed call's output on a row subsample: 16 rows spread over the
    # track, all channels and sources, against the oracle (the checker; never the thing measured)
    parity = None
    if rank == 0:
        import oracle

        oracle.build()
        rows = np.linspace(0, wl["uvw"].shape[0] - 1, 16).astype(np.int64)
        ref = oracle.im_to_vis(wl["image"], wl["uvw"][rows], wl["lm"], wl["freq"])
        got = out[torch.from_numpy(rows).to(dev)].cpu().numpy()
        scale = float(np.max(np.abs(ref)))
        parity = {"ok": bool(np.allclose(got, ref, rtol=1e-10, atol=1e-10 * scale)),
                  "max_abs_err_over_max_ref": float(np.max(np.abs(got - ref)) / scale),
                  "gate": "allclose(rtol=1e-10, atol=1e-10*max|ref|)",
                  "checked": "16 rows spread over the track x all %d channels x all %d sources vs the oracle"
                             % (wl["nchan"], wl["nsrc"])}
    out = None

    # ---- end to end through the public API with HOST buffers (pinned), H2D + D2H timed
    e2e = None
    if not args.no_e2e:
        pin = {k: torch.from_numpy(wl[k]).pin_memory() for k in ("image", "uvw", "lm", "freq")}
        h = {k: v.numpy() for k, v in pin.items()}
        h2d = sum(v.nbytes for v in h.values())
        res = dft.im_to_vis(h["image"], h["uvw"], h["lm"], h["freq"])  # warm-up (pins buffers)
        d2h = res.nbytes
        res = None
        n_e2e = max(steps, env_int("BENCH_E2E_STEPS", 5))
        barrier()
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(n_e2e):
            res = None
            res = dft.im_to_vis(h["image"], h["uvw"], h["lm"], h["freq"])
        torch.cuda.synchronize()
        e2e_s = max_over_ranks(time.perf_counter() - t0)
        barrier()
        e2e = {"value": world * terms * n_e2e / e2e_s / 1e9, "unit": UNIT,
               "h2d_bytes_per_step": int(h2d), "d2h_bytes_per_step": int(d2h),
               "steps": n_e2e, "ms_per_step": 1e3 * e2e_s / n_e2e,
               "api": "codex_africanus_b200.dft.im_to_vis(numpy) -> numpy; row blocks streamed "
                      "back over a copy stream while later blocks compute"}
        res = None
        del pin, h

    line = {
        "metric": METRIC, "value": value, "unit": UNIT, "n_gpus": world, "steps": steps,
        "warmup": warmup, "ms_per_step": 1e3 * total_s / steps, "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": {"workload": wl["desc"], "parallelism": "rows sharded by timestep, %d rank(s), "
                   "no data-path collective" % world,
                   "l2": "256 MiB flush between steps; per-step output (%.2f GB) exceeds L2"
                         % (terms / wl["nsrc"] * 16 / 1e9),
                   "checksum": checksum, "kernel_path": dft_path},
        "clocks": clocks, "gpu_launches": launches,
    }
    if parity:
        line["parity"] = parity
    if e2e:
        line["e2e"] = e2e

    if rank == 0:
        achieved = FLOP_PER_TERM * terms / kernel_s / 1e12
        # DRAM bytes per launch of this kernel from an `ncu --set full` capture (profiles/): a
        # capture cannot run inside the timed region, so the figure is accepted only while its stamp
        # -- the sha256 of the kernel's source file it was taken with -- matches the source that
        # built the library measured here; otherwise traffic is null
        traffic, traffic_note = None, "no capture"
        tpath = os.path.join(ROOT, "profiles", "traffic.json")
        if os.path.exists(tpath):
            try:
                import hashlib

                tj = json.load(open(tpath))
                src = os.path.join(ROOT, "codex_africanus_b200", "csrc", "afr_dft.cu")
                sha = hashlib.sha256(open(src, "rb").read()).hexdigest()
                ent = tj.get("phasor_stream_im_to_vis_cfg2")
                if ent and ent.get("source_sha256") == sha:
                    traffic = ent.get("dram_bytes_per_launch")
                    traffic_note = "ncu capture %s, source stamp matches" % ent.get("capture")
                else:
                    traffic_note = "capture is stale (kernel source changed since): not reported"
            except Exception as exc:
                traffic_note = "unreadable: %r" % (exc,)
        peaks = {}
        ppath = os.path.join(ROOT, "MEASURED_PEAKS.json")
        if os.path.exists(ppath):
            try:
                peaks = json.load(open(ppath))
            except Exception:
                peaks = {}
        io_bytes = (wl["image"].nbytes + wl["uvw"].nbytes + terms / wl["nsrc"] * 16)
        line["roofline"] = {
            "bound": "fp64", "kernel": "phasor_stream_ws_kernel<ncorr=1, real W, forward, double, CH=16, 16+4 warps>",
            "achieved": achieved, "peak": fp64_peak / 1e12, "unit": "TFLOP/s",
            "frac": achieved / (fp64_peak / 1e12),
            "algorithmic_flop_per_term": FLOP_PER_TERM, "terms_per_launch": terms,
            "kernel_ms": 1e3 * kernel_s,
            "peak_source": "measured in this run: afr_measure_fma_peak (dependent-free DFMA "
                           "chains on all SMs); MEASURED_PEAKS.json has no FP64 entry",
            "traffic": traffic, "traffic_source": traffic_note,
            "hbm": {"algorithmic_GBps": io_bytes / kernel_s / 1e9,
                    "peak_GBps": peaks.get("hbm_gbs", 6650.0),
                    "peak_source": "MEASURED_PEAKS.json" if "hbm_gbs" in peaks else "fallback",
                    "note": "bytes/term << 1: the kernel is FP64-pipe bound, not HBM bound"},
        }

    # ---- CPU baseline: the reference algorithm's port on this box's host cores
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        rate, cores, sample, _, _ = cpu_im_to_vis_rate(wl, float(os.environ.get("BENCH_CPU_SECONDS", 12)))
        line["cpu_baseline"] = {"value": rate, "unit": UNIT, "cores": cores, "kind": "port",
                                "sample": sample}
        try:
            nb = numba_im_to_vis_rate(wl, float(os.environ.get("BENCH_NUMBA_SECONDS", 8)))
        except Exception as exc:
            nb = {"error": repr(exc)}
        if nb:
            nb.pop("seconds", None)
            line["cpu_baseline_reference"] = nb  # the reference's own numba kernel, same box, same run

    # ---- the other BASELINE configs as first-class entries (tools/bench_configs.py)
    if not args.no_configs:
        try:
            import bench_configs

            cfgs = bench_configs.run(dev, fp64_peak, env_int("BENCH_CFG_STEPS", 5), env_int("BENCH_CFG_WARMUP", 3),
                                     rank, world)
            if rank == 0:
                line["configs"] = cfgs
        except Exception as exc:
            if rank == 0:
                line["configs"] = {"error": repr(exc)}

    # ---- secondary kernels (one GPU): every other row of SURVEY.md section 8
    if world == 1 and not args.no_extras:
        try:
            import bench_extras

            line["extra"] = bench_extras.run(dev, fp64_peak)
        except Exception as exc:  # extras must never sink the headline number
            line["extra"] = {"error": repr(exc)}

    if rank == 0:
        emit(line)
    if world > 1:
        dist.barrier()
        dist.destroy_process_group()
    return 0


if __name__ == "__main__":
    sys.exit(main())
