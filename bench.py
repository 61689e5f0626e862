#!/usr/bin/env python
"""Benchmark of the B200 spectral frontend (BASELINE.json metric: audio-seconds / second).

    python bench.py --gpus N --steps K --warmup W            # this repo's CUDA path
    python bench.py --impl reference --gpus N --steps K ...  # CPU oracle port on the host cores

Workload (config.workload): BASELINE.json configs[2] sharded by track -- per GPU a
batch of 3-minute 44.1 kHz stereo tracks, n_fft 2048, hop 512, 128 mels (the
per-track shape of configs[1]); 128 tracks per GPU, i.e. 1024 tracks on 8 GPUs.
One step = one pass of the whole frontend (ta_frontend_run: STFT magnitude, mel,
LTAS/centroid/roll-off, stereo band energies, onset flux, autocorrelation,
K-weighted gated loudness, RMS frames, moments) over that batch.

  value   : tracks resident in HBM (8 GB per GPU >> 126 MB L2, so every step
            streams from HBM), device time by CUDA events, max over ranks.
  e2e     : the same work through engine.HostPipeline -- pinned host PCM copied
            H2D every step and every output copied D2H inside the timed region.
  roofline: the fused STFT kernel (dominant), algorithmic bytes / its event time.
  cpu_baseline: the oracle port (oracle/, numpy+scipy) timed on the host cores.
"""

from __future__ import annotations

import argparse
import json
import multiprocessing as mp
import os
import subprocess
import sys
import threading
import time

import numpy as np

ROOT = os.path.dirname(os.path.abspath(__file__))
sys.path.insert(0, ROOT)

SR, N_FFT, HOP, N_MELS = 44_100, 2048, 512, 128
METRIC = "audio_seconds_per_second"
UNIT = "audio-s/s (xRT)"


# ----------------------------------------------------------------------------- CPU side
def _gen_track(args):
    seed, seconds = args
    from track_analyser_b200 import synth

    return synth.synth_track(seed, seconds, SR, 2)


def _oracle_frontend(x):
    """The reference's frontend arithmetic for one stereo track, every STFT computed once
    (the reference itself recomputes the mono STFT >= 11 times)."""
    from oracle import frontend as ofe
    from oracle import librosa_np as olr
    from oracle import pyloudnorm_np as opl

    mono = np.mean(x, axis=0)
    D = olr.stft(mono, n_fft=N_FFT, hop_length=HOP)
    mag = np.abs(D)
    mel = np.einsum("ft,mf->mt", mag**2, olr.filters_mel(SR, N_FFT, n_mels=N_MELS), optimize=True)
    env = olr.onset_strength(S=olr.power_to_db(mel), sr=SR, hop_length=HOP)
    out = [np.mean(mag, axis=1), olr.autocorrelate(env),
           olr.onset_strength(S=np.asarray(mel, dtype=float), sr=SR, hop_length=HOP)]
    freq = olr.fft_frequencies(SR, N_FFT)[:, None]
    out.append(np.sum(freq * olr.normalize(mag, norm=1, axis=-2), axis=-2))
    total = np.cumsum(mag, axis=-2)
    out.append(np.nanmin(np.where(total < 0.85 * total[-1], np.nan, 1) * freq, axis=-2))
    out.append(ofe.frequency_dependent_width(x, SR))
    out.append(ofe.mid_side_rms(x))
    out.append(ofe.mono_compatibility_correlation(x))
    out.append(opl.integrated_loudness(mono, SR))
    out.append(ofe.windowed_loudness(mono, SR, 0.4))
    out.append(ofe.windowed_loudness(mono, SR, 3.0))
    out.append(ofe.rms_dbfs(mono))
    # chroma_stft on the shared power spectrogram (tuning estimate + filterbank + projection) and the tempogram
    tuning = olr.estimate_tuning(mag**2, SR)
    raw = np.einsum("cf,ft->ct", olr.filters_chroma(SR, N_FFT, tuning=tuning), mag**2, optimize=True)
    out.append(olr.normalize(raw, norm=np.inf, axis=-2))
    out.append(olr.tempogram(onset_envelope=env, sr=SR, hop_length=HOP))
    return len(out)


def _oracle_timed(args):
    seed, seconds = args
    x = _gen_track((seed, seconds))
    t0 = time.perf_counter()
    _oracle_frontend(x)
    return time.perf_counter() - t0


def cpu_workers(world: int = 1, cap: int = 64) -> int:
    """Host processes for the CPU legs: every core this rank may use (at most `cap`: each worker holds a
    float64 STFT of its track, ~0.5 GB for 60 s)."""
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:
        n = os.cpu_count() or 1
    return max(1, min(cap, n // max(1, world)))


def run_cpu_sample(workers: int, seconds: float, rounds: int = 1):
    """`workers` processes, one `seconds`-long track each per round, all at once.

    Each worker times only its oracle call (track synthesis excluded); a round's wall time
    is the slowest worker's.  Returns (audio_seconds, wall_seconds)."""
    ctx = mp.get_context("fork")
    wall = 0.0
    with ctx.Pool(workers) as pool:
        for r in range(rounds):
            dts = pool.map(_oracle_timed, [(13_370 + r * workers + i, seconds) for i in range(workers)], chunksize=1)
            wall += max(dts)
    return workers * rounds * seconds, max(wall, 1e-9)


# ----------------------------------------------------------------------------- clocks
class ClockSampler:
    """SM clock and throttle reasons sampled DURING the timed region: NVML every 10 ms (nvidia-smi every 200 ms as
    a fallback)."""

    Q = ("clocks.sm,clocks.max.sm,clocks_event_reasons.hw_slowdown,clocks_event_reasons.hw_thermal_slowdown,"
         "clocks_event_reasons.sw_thermal_slowdown,clocks_event_reasons.sw_power_cap")
    NAMES = ["hw_slowdown", "hw_thermal_slowdown", "sw_thermal_slowdown", "sw_power_cap"]

    def __init__(self, index: int):
        self.index, self.samples, self._stop = index, [], threading.Event()
        self.th = threading.Thread(target=self._run, daemon=True)
        self.source = "nvidia-smi"

    def _run_nvml(self) -> bool:
        try:
            import pynvml as nv

            nv.nvmlInit()
            h = nv.nvmlDeviceGetHandleByIndex(self.index)
            mx = nv.nvmlDeviceGetMaxClockInfo(h, nv.NVML_CLOCK_SM)
            get_reasons = getattr(nv, "nvmlDeviceGetCurrentClocksEventReasons", None) or nv.nvmlDeviceGetCurrentClocksThrottleReasons
            bits = {
                "hw_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwSlowdown", 0x8),
                "hw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonHwThermalSlowdown", 0x40),
                "sw_thermal_slowdown": getattr(nv, "nvmlClocksThrottleReasonSwThermalSlowdown", 0x20),
                "sw_power_cap": getattr(nv, "nvmlClocksThrottleReasonSwPowerCap", 0x4),
            }
            nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
        except Exception:
            return False
        self.source = "nvml"
        while not self._stop.is_set():
            try:
                sm = nv.nvmlDeviceGetClockInfo(h, nv.NVML_CLOCK_SM)
                r = int(get_reasons(h))
                self.samples.append([str(sm), str(mx)] + ["Active" if r & bits[n] else "Not Active" for n in self.NAMES])
            except Exception:
                pass
            self._stop.wait(0.01)
        return True

    def _run(self):
        if self._run_nvml():
            return
        while not self._stop.is_set():
            try:
                o = subprocess.run(["nvidia-smi", f"--query-gpu={self.Q}", "--format=csv,noheader,nounits", "-i",
                                    str(self.index)], capture_output=True, text=True, timeout=5).stdout.strip()
                if o:
                    self.samples.append([c.strip() for c in o.split(",")])
            except Exception:
                pass
            self._stop.wait(0.2)

    def __enter__(self):
        self.th.start()
        return self

    def __exit__(self, *a):
        self._stop.set()
        self.th.join(timeout=6)

    def summary(self):
        sm = [float(s[0]) for s in self.samples if s and s[0].replace(".", "").isdigit()]
        mx = [float(s[1]) for s in self.samples if len(s) > 1 and s[1].replace(".", "").isdigit()]
        reasons = sorted({self.NAMES[i] for s in self.samples for i in range(4) if len(s) > 2 + i and s[2 + i] == "Active"})
        return {"sm_mhz": float(np.median(sm)) if sm else None, "sm_max_mhz": max(mx) if mx else None,
                "reasons": reasons, "samples": len(sm), "source": self.source}


# ----------------------------------------------------------------------------- reference arm
def reference_arm(args, rank, world):
    if rank != 0:
        return
    workers = cpu_workers(1)
    seconds = args.ref_seconds
    vals = []
    for step in range(args.warmup + args.steps):
        audio, wall = run_cpu_sample(workers, seconds)
        if step >= args.warmup:
            vals.append((audio, wall))
    audio = sum(a for a, _ in vals)
    wall = sum(w for _, w in vals)
    v = audio / wall
    sample = f"{workers} tracks x {seconds:.0f} s per step, one process per track (oracle port, STFTs shared)"
    cfg = workload_config(args, world)
    # the metric is length-normalised (audio seconds per second), so the CPU arm times a bounded sample of the same
    # per-track workload instead of the full 128 x 180 s batch: said here, not only in cpu_baseline.sample
    cfg["timed_sample"] = sample
    cfg["workload"] += f"; this arm times a bounded sample of it per step: {workers} tracks x {seconds:.0f} s"
    line = {
        "impl": "reference", "metric": METRIC, "value": v, "unit": UNIT, "n_gpus": args.gpus, "steps": args.steps,
        "warmup": args.warmup, "ms_per_step": 1e3 * wall / max(1, args.steps), "higher_is_better": True,
        "scaling": "weak", "vs_baseline": None, "dtype": "f64", "data": "synthetic",
        "config": cfg,
        "cpu_baseline": {"value": v, "unit": UNIT, "cores": workers, "kind": "port", "sample": sample},
        "e2e": {"value": v, "unit": UNIT, "h2d_bytes_per_step": 0, "d2h_bytes_per_step": 0},
        "gpu_launches": 0,
    }
    print(json.dumps(line), flush=True)


def workload_config(args, world):
    return {
        "workload": ("BASELINE configs[2] sharded by track: %d tracks/GPU x %d GPU(s) of %.0f s 44.1 kHz stereo, "
                     "n_fft 2048 hop 512 128 mels (per-track shape of configs[1]); full frontend, all outputs"
                     % (args.tracks_per_gpu, world, args.seconds)),
        "tracks_per_gpu": args.tracks_per_gpu, "seconds_per_track": args.seconds, "sample_rate": SR,
        "n_fft": N_FFT, "hop": HOP, "n_mels": N_MELS,
        "l2_policy": "inputs larger than L2 (%.1f GB PCM per GPU per step)" % (args.tracks_per_gpu * args.seconds * SR * 8 / 1e9),
        "parallelism": f"tracks sharded over {world} GPU(s), no collective",
    }


def bind_to_gpu_numa(index: int):
    """Pin this process to the CPUs local to GPU `index` (its PCIe root's NUMA node) before any pinned host buffer is
    allocated, so that every rank's staging memory sits next to its own GPU.  Best effort: returns a note for the
    JSON line."""
    try:
        import pynvml as nv

        nv.nvmlInit()
        bus = nv.nvmlDeviceGetPciInfo(nv.nvmlDeviceGetHandleByIndex(index)).busId
        bus = bus.decode() if isinstance(bus, bytes) else bus
        dom, rest = bus.split(":", 1)
        path = f"/sys/bus/pci/devices/{int(dom, 16):04x}:{rest.lower()}"
        cpus = set()
        for part in open(path + "/local_cpulist").read().strip().split(","):
            if not part:
                continue
            lo, _, hi = part.partition("-")
            cpus.update(range(int(lo), int(hi or lo) + 1))
        allowed = os.sched_getaffinity(0)
        cpus &= allowed
        node = open(path + "/numa_node").read().strip()
        if cpus and cpus != allowed:
            os.sched_setaffinity(0, cpus)
            return f"numa node {node}, {len(cpus)} cpus"
        return f"numa node {node}, affinity unchanged"
    except Exception as exc:  # no sysfs entry, no NVML, restricted container ...
        return f"unbound ({type(exc).__name__})"


def pcie_probe(dev, nbytes=1 << 30, reps=3):
    """GB/s of pinned cudaMemcpyAsync over this GPU's link: host->device alone, device->host alone, both at once."""
    import torch

    h_up = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    h_dn = torch.empty(nbytes, dtype=torch.uint8, pin_memory=True)
    d_up = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    d_dn = torch.empty(nbytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)

    def timed(up, down):
        torch.cuda.synchronize()
        t0 = time.perf_counter()
        for _ in range(reps):
            if up:
                with torch.cuda.stream(s1):
                    d_up.copy_(h_up, non_blocking=True)
            if down:
                with torch.cuda.stream(s2):
                    h_dn.copy_(d_dn, non_blocking=True)
        torch.cuda.synchronize()
        return reps * nbytes / (time.perf_counter() - t0) / 1e9

    timed(True, True)  # first touch of the pinned pages
    out = {"h2d_gbs": timed(True, False), "d2h_gbs": timed(False, True)}
    both = timed(True, True)   # bytes per direction per second while both directions run
    out["bidir_each_gbs"] = both
    del h_up, h_dn, d_up, d_dn
    torch.cuda.empty_cache()
    return out


def pcie_sustained(dev, up_bytes, down_bytes, seconds, barrier):
    """What this rank's link moves while EVERY rank keeps copying: pieces of ``down_bytes`` device->host and ``up_bytes``
    host->device (the end-to-end leg's ratio), two of each in flight, until ``seconds`` have passed on every rank; only
    pieces that completed inside the window count, so no rank is measured on a host the others have already left.
    (A fixed amount of work per rank overstates a shared host: ranks that finish early free the others' path.)
    Returns (GB/s down, GB/s up)."""
    import torch

    h_up = torch.empty(up_bytes, dtype=torch.uint8, pin_memory=True)
    h_dn = [torch.empty(down_bytes, dtype=torch.uint8, pin_memory=True) for _ in range(2)]
    d_up = [torch.empty(up_bytes, dtype=torch.uint8, device=dev) for _ in range(2)]
    d_dn = torch.empty(down_bytes, dtype=torch.uint8, device=dev)
    s1, s2 = torch.cuda.Stream(dev), torch.cuda.Stream(dev)
    for b in range(2):   # first touch
        with torch.cuda.stream(s1):
            d_up[b].copy_(h_up, non_blocking=True)
        with torch.cuda.stream(s2):
            h_dn[b].copy_(d_dn, non_blocking=True)
    barrier()
    ev_dn, ev_up, done_dn, done_up, i = [], [], 0, 0, 0
    t0 = time.perf_counter()
    while True:
        with torch.cuda.stream(s1):
            d_up[i % 2].copy_(h_up, non_blocking=True)
            e = torch.cuda.Event()
            e.record(s1)
            ev_up.append(e)
        with torch.cuda.stream(s2):
            h_dn[i % 2].copy_(d_dn, non_blocking=True)
            e = torch.cuda.Event()
            e.record(s2)
            ev_dn.append(e)
        i += 1
        if i >= 2:
            ev_dn[i - 2].synchronize()   # keep two pieces in flight per direction
            if time.perf_counter() - t0 >= seconds:
                done_dn = i - 1
                done_up = sum(1 for e in ev_up if e.query())
                break
    elapsed = time.perf_counter() - t0
    # keep the link busy until every rank has closed its window, then drain
    for _ in range(2):
        with torch.cuda.stream(s1):
            d_up[0].copy_(h_up, non_blocking=True)
        with torch.cuda.stream(s2):
            h_dn[0].copy_(d_dn, non_blocking=True)
    torch.cuda.synchronize()
    barrier()
    del h_up, h_dn, d_up, d_dn
    torch.cuda.empty_cache()
    return done_dn * down_bytes / elapsed / 1e9, done_up * up_bytes / elapsed / 1e9


def verify_one_track(plan, x, res):
    """One track of the bench batch against the oracle at the parity tolerance (outside every timed region)."""
    from oracle import frontend as ofe
    from oracle import librosa_np as olr
    from oracle import pyloudnorm_np as opl

    mono = np.mean(x, axis=0)
    mag = np.abs(olr.stft(mono, n_fft=N_FFT, hop_length=HOP))
    # rtol 1e-4 / atol 1e-6 per bin; the few bins 1e4 below their frame's peak that miss it must sit inside the float32
    # transform's own rounding floor (4 eps32 of the frame maximum: the oracle's float32 FFT is no closer to a float64 one)
    err = np.abs(res["magnitude"] - mag)
    ok = (err <= 1e-6 + 1e-4 * mag) | (err <= 4.0 * np.finfo(np.float32).eps * np.max(mag, axis=0, keepdims=True))
    checks = {"magnitude": bool(ok.all()) and bool(np.mean(err <= 1e-6 + 1e-4 * mag) >= 0.99999)}
    mel = np.einsum("ft,mf->mt", mag**2, olr.filters_mel(SR, N_FFT, n_mels=N_MELS), optimize=True)
    env = olr.onset_strength(S=olr.power_to_db(mel), sr=SR, hop_length=HOP)
    close = lambda a, b: bool(np.allclose(a, b, rtol=1e-4, atol=1e-6))  # noqa: E731
    checks["mel"] = close(res["mel"], mel)
    checks["onset_env"] = close(res["onset_env"], env)
    checks["autocorr"] = close(res["autocorr"], olr.autocorrelate(env))
    checks["ltas"] = close(res["ltas"], np.mean(mag, axis=1))
    checks["centroid"] = close(res["centroid"], olr.spectral_centroid(mono, SR, N_FFT, HOP)[0])
    freqs = np.fft.rfftfreq(N_FFT, 1.0 / SR)
    checks["rolloff_bins_equal"] = bool(np.mean(freqs[res["rolloff_bin"]] == olr.spectral_rolloff(mono, SR, N_FFT, HOP)[0]) >= 0.9995)
    chroma, tuning = olr.chroma_stft(mono, SR, return_tuning=True)
    checks["tuning"] = bool(abs(res["tuning"] - tuning) < 1e-12)
    checks["chroma"] = close(res["chroma"], chroma)
    checks["tempogram"] = close(res["tempogram"], olr.tempogram(onset_envelope=env, sr=SR, hop_length=HOP))
    checks["lufs"] = bool(abs(res["lufs"] - opl.integrated_loudness(mono, SR)) < 0.01)
    checks["mid_side_rms"] = close(np.sqrt(np.asarray([res["moments"][5], res["moments"][6]]) / res["moments"][7]), ofe.mid_side_rms(x))
    return checks


# ----------------------------------------------------------------------------- our arm
def ours(args, rank, world, local_rank):
    numa_note = bind_to_gpu_numa(local_rank) if world > 1 else "single rank"
    workers = cpu_workers(world)
    # 1. host pool of distinct synthetic tracks (and, on rank 0 at N=1, the CPU baseline) BEFORE CUDA init: fork-safe
    pool_n = min(args.pool, args.tracks_per_gpu)
    # this rank's tracks of the global batch, by the product's own partition rule (contiguous blocks for equal lengths)
    from track_analyser_b200 import sharding

    mine = sharding.partition([int(args.seconds * SR)] * (args.tracks_per_gpu * world), world)[rank]
    assert len(mine) == args.tracks_per_gpu
    ctx = mp.get_context("fork")
    with ctx.Pool(min(workers, pool_n)) as pool:
        host_np = pool.map(_gen_track, [(13_370 + mine[i], args.seconds) for i in range(pool_n)])
    cpu_baseline = None
    if rank == 0 and world == 1 and not args.no_cpu_baseline:
        audio, wall = run_cpu_sample(workers, args.ref_seconds)
        cpu_baseline = {"value": audio / wall, "unit": UNIT, "cores": workers, "kind": "port",
                        "sample": f"{workers} tracks x {args.ref_seconds:.0f} s, one process per track (oracle port, STFTs shared)"}

    import torch
    import torch.distributed as dist

    torch.cuda.set_device(local_rank)
    if world > 1:
        dist.init_process_group("nccl", device_id=torch.device(f"cuda:{local_rank}"))
    from track_analyser_b200 import engine

    dev = torch.device(f"cuda:{local_rank}")
    plan = engine.Plan(SR, N_FFT, HOP, N_MELS, device=local_rank)
    n = host_np[0].shape[1]
    stride = (2 * n + 3) & ~3
    host_pool = []
    for x in host_np:
        t = torch.empty(2 * n, dtype=torch.float32, pin_memory=True)
        t.numpy()[:] = x.reshape(-1)
        host_pool.append(t)
    del host_np

    # 2. resident batch: tracks_per_gpu tracks in HBM (pool repeated)
    nt = args.tracks_per_gpu
    pcm = torch.empty(nt * stride, dtype=torch.float32, device=dev)
    for i in range(nt):
        pcm[i * stride: i * stride + 2 * n].copy_(host_pool[i % pool_n], non_blocking=True)
    batch = engine.DeviceBatch(plan, pcm, np.arange(nt, dtype=np.int64) * stride, np.full(nt, n, dtype=np.int64), 2)
    bufs = engine.FrontendBuffers(batch, engine.FRONTEND_OUTPUTS)
    torch.cuda.synchronize()

    def barrier():
        if world > 1:
            dist.barrier()
        torch.cuda.synchronize()

    audio_per_step = nt * args.seconds
    # ---- kernel leg -------------------------------------------------------------------
    for _ in range(args.warmup):
        engine.run_device(plan, batch, bufs)
    barrier()
    launches0 = engine.launch_count()
    stage = np.zeros(len(engine.STAGE_NAMES))
    with ClockSampler(local_rank) as clk:
        e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
        e0.record()
        for _ in range(args.steps):
            engine.run_device(plan, batch, bufs)   # the production call: time-domain pass overlapped on a second stream
        e1.record()
        barrier()
        ms_kernel = e0.elapsed_time(e1) / args.steps
        launches = engine.launch_count() - launches0
        # per-stage device times (sequential, event-bracketed inside the library) for the roofline figure
        prof_steps = max(1, min(args.steps, 5))
        for _ in range(prof_steps):
            stage += np.asarray(engine.run_device_profiled(plan, batch, bufs))
        barrier()
    stage /= prof_steps
    t = torch.tensor([ms_kernel], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    ms_kernel_max = float(t.item())
    lufs = bufs.t["lufs"].cpu().numpy()
    assert np.all(np.isfinite(lufs)), "frontend produced non-finite loudness"
    verified = None
    if rank == 0 and not args.no_verify:
        # one random track of the timed batch against the oracle, outside the timing (its first `verify_seconds`, so that the
        # float64 oracle stays within seconds: the kernels are re-run on exactly that excerpt)
        pick = int(np.random.default_rng(args.steps).integers(0, nt))
        n_v = min(n, int(args.verify_seconds * SR))
        x_v = host_pool[pick % pool_n].numpy().reshape(2, n)[:, :n_v].copy()
        full = engine.download_track(batch, bufs, pick, ("magnitude", "mel"))
        part = engine.analyse_batch(plan, [x_v], engine.FRONTEND_OUTPUTS)[0]
        T_v = 1 + n_v // HOP
        inner = T_v - 8   # frames whose windows end inside the excerpt see the same samples in the full-length run
        same = bool(np.array_equal(full["magnitude"][:, :inner], part["magnitude"][:, :inner]) and
                    np.array_equal(full["mel"][:, :inner], part["mel"][:, :inner]))
        checks = verify_one_track(plan, x_v, part)
        checks["batch_track_equals_excerpt_run"] = same
        verified = {"ok": bool(all(checks.values())), "track": pick, "seconds": n_v / SR, "checks": checks}

    # ---- end-to-end leg ------------------------------------------------------------------
    chunk = min(args.chunk_tracks, nt)
    del bufs, batch, pcm
    torch.cuda.empty_cache()
    pcie = None
    if not args.no_pcie_probe:
        # the link's own limits: this rank alone (ranks take turns), then every rank at once (what the host can feed)
        alone = None
        for r in range(world):
            barrier()
            if r == rank:
                alone = pcie_probe(dev)
        barrier()
        together = pcie_probe(dev) if world > 1 else alone
        barrier()
        pcie = {"alone": alone, "all_ranks_together": together}
    pipe = engine.HostPipeline(plan, n, 2, chunk, engine.FRONTEND_OUTPUTS)
    nt_e2e, link_rates = nt, None
    if world > 1 and pcie is not None:
        # Every rank runs the end-to-end leg's copy pattern for one whole step at once (no kernels): the rate each GPU's
        # link sustains is what the host gives that GPU, and on this kind of box the GPUs do not share it evenly.  The
        # end-to-end shards are sized by these rates (sharding.partition(weights=...)), so that all ranks finish
        # together; the device-resident `value` above keeps equal shards.
        piece = 8   # an eighth of a chunk per copy: ~100 MB down, ~64 MB up
        down_gbs, up_gbs = pcie_sustained(dev, chunk * 2 * n * 4 // piece, pipe.d2h_bytes_per_chunk // piece, 1.0, barrier)
        mine_r = torch.tensor([down_gbs, up_gbs], dtype=torch.float64, device=dev)
        all_r = [torch.zeros_like(mine_r) for _ in range(world)]
        dist.all_gather(all_r, mine_r)
        link_rates = [[float(v) for v in t.tolist()] for t in all_r]
        down = [r[0] for r in link_rates]
        pcie["sustained_step_pattern"] = {"down_gbs_per_rank": down, "up_gbs_per_rank": [r[1] for r in link_rates],
                                          "sum_down_gbs": float(sum(down)), "min_down_gbs": float(min(down))}
        # whole chunks per rank: a partial last chunk would still copy full-size output buffers back
        unit = chunk if (nt * world) % chunk == 0 else 1
        shard = sharding.partition([n * unit] * (nt * world // unit), world, weights=down)[rank]
        nt_e2e = len(shard) * unit
    tracks = [host_pool[i % pool_n] for i in range(nt_e2e)]
    sink = []

    def consume(ci, first, cnt, host_out):
        sink.append(float(host_out["lufs"][:cnt].sum()))  # host reads the step's result

    d2h_chunk_full = pipe.d2h_bytes_per_chunk
    e2e_steps = max(1, min(args.steps, args.e2e_steps))
    for _ in range(min(args.warmup, 1) or 1):
        pipe.run(tracks, consume)
    barrier()
    t0 = time.perf_counter()
    for _ in range(e2e_steps):
        pipe.run(tracks, consume)
    barrier()
    s_e2e = (time.perf_counter() - t0) / e2e_steps
    t = torch.tensor([s_e2e], dtype=torch.float64, device=dev)
    if world > 1:
        dist.all_reduce(t, op=dist.ReduceOp.MAX)
    s_e2e_max = float(t.item())
    n_chunks = (nt + chunk - 1) // chunk
    # bytes this rank really moved per step (a partial last chunk still copies whole output buffers), summed over ranks
    moved = torch.tensor([nt_e2e * 2 * n * 4, ((nt_e2e + chunk - 1) // chunk) * d2h_chunk_full, nt_e2e], dtype=torch.float64, device=dev)
    per_rank_tracks = [nt_e2e]
    if world > 1:
        gathered = [torch.zeros_like(moved) for _ in range(world)]
        dist.all_gather(gathered, moved)
        per_rank_tracks = [int(g[2].item()) for g in gathered]
        moved = torch.stack(gathered).sum(dim=0)
    e2e_up_total, e2e_down_total = float(moved[0].item()), float(moved[1].item())
    assert sum(per_rank_tracks) == world * nt, per_rank_tracks

    # ---- informational: the byte-reduced pipeline that analyse_track's host stages need -----------------------------
    # (HPSS curves and true peak computed on the device, magnitude and tempogram never downloaded; NOT the contract's
    # e2e figure, which copies every frontend output back)
    e2e_analysis = None
    d2h_full = n_chunks * d2h_chunk_full
    if not args.no_analysis_leg:
        del pipe
        torch.cuda.empty_cache()
        e2e_analysis = {"note": "outputs consumed by analyse_track's host stages (HPSS curves + true peak on the device; "
                                "magnitude and tempogram stay in HBM); pcm16: host tracks are interleaved int16 as a 16-bit "
                                "WAV stores them, decoded on the device"}
        for label, pcm16 in (("f32_input", False), ("pcm16_input", True)):
            src = tracks
            if pcm16:
                pool16 = [torch.from_numpy(np.ascontiguousarray(
                    np.clip(np.round(t.numpy().reshape(2, n) * 32767.0), -32768, 32767).astype(np.int16).T).reshape(-1)).pin_memory()
                    for t in host_pool]
                src = [pool16[i % pool_n] for i in range(nt_e2e)]
            pipe_a = engine.HostPipeline(plan, n, 2, chunk, engine.ANALYSIS_OUTPUTS, pcm16=pcm16)
            pipe_a.run(src, consume)
            barrier()
            t0 = time.perf_counter()
            for _ in range(e2e_steps):
                pipe_a.run(src, consume)
            barrier()
            ta = torch.tensor([(time.perf_counter() - t0) / e2e_steps], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ta, op=dist.ReduceOp.MAX)
            e2e_analysis[label] = {"value": world * audio_per_step / float(ta.item()), "unit": UNIT, "s_per_step": float(ta.item()),
                                   "h2d_bytes_per_step": nt * pipe_a.h2d_bytes_per_track,
                                   "d2h_bytes_per_step": n_chunks * pipe_a.d2h_bytes_per_chunk}
            del pipe_a
            torch.cuda.empty_cache()

    # ---- informational: the batch API down to the reference's dataclasses (pipeline.analyse_tracks) ---------------------
    e2e_track = None
    if not args.no_analysis_leg and args.analysis_tracks > 0:
        from track_analyser_b200 import pipeline
        from track_analyser_b200.utils import AudioInput

        pool_audio = []
        for t_ in host_pool:
            st = t_.numpy().reshape(2, n)
            pool_audio.append(AudioInput(samples=np.mean(st, axis=0), sample_rate=SR, stereo_samples=st))
        na = min(args.analysis_tracks, nt)
        srcs = [pool_audio[i % pool_n] for i in range(na)]
        # warm-up: forks the workers, builds the plans, and runs enough chunks for the device PCM buffers that go round
        # between uploader and kernel thread (five) to exist before the timed call
        pipeline.analyse_tracks(srcs[: min(na, 6 * args.chunk_tracks)], chunk_tracks=args.chunk_tracks)
        runs_s = []
        for _ in range(2):   # host-side scheduling of ~15 processes on 16 cores varies from run to run: both times are reported
            barrier()
            t0 = time.perf_counter()
            out = pipeline.analyse_tracks(srcs, chunk_tracks=args.chunk_tracks)
            ta = torch.tensor([time.perf_counter() - t0], dtype=torch.float64, device=dev)
            if world > 1:
                dist.all_reduce(ta, op=dist.ReduceOp.MAX)
            runs_s.append(float(ta.item()))
            assert len(out) == na and all(90.0 <= o.beat.bpm <= 135.0 for o in out)
        s_mean = float(np.mean(runs_s))
        e2e_track = {"value": world * na * args.seconds / s_mean, "unit": UNIT, "tracks_per_gpu": na,
                     "s_total": s_mean, "s_runs": runs_s, "host_workers": pipeline._worker_count(None),
                     "note": "pipeline.analyse_tracks: host AudioInput in -> TrackAnalysisResult dataclasses out (beats, structure, "
                             "loudness, harmony incl. chroma_cqt, features, stereo); host stages in worker processes"}

    if rank == 0:
        B = N_FFT // 2 + 1
        T = 1 + n // HOP
        k1_bytes = nt * (4 * 2 * n + 4 * B * T + 4 * N_MELS * T + 12 * T)  # PCM read + magnitude + mel + centroid/roll-off
        peaks = {}
        try:
            peaks = json.load(open(os.path.join(ROOT, "MEASURED_PEAKS.json")))
        except Exception:
            pass
        peak = float(peaks.get("hbm_gbs", 6650.0))
        achieved = k1_bytes / (stage[0] * 1e-3) / 1e9
        traffic = None  # DRAM bytes of one K1 launch from the committed ncu --set full capture of this same command
        try:
            tj = json.load(open(os.path.join(ROOT, "profiles", "r2_k1_traffic.json")))
            if args.tracks_per_gpu == 128 and args.seconds == 180.0:
                traffic = tj["traffic_bytes"]
        except Exception:
            pass
        # what else bounds K1: pipe utilisations of the same launch from the committed ncu --set full capture
        # (profiles/r2_k1_ncu_summary.txt); the HBM fraction above is the contract's figure, these say why it is what it is
        other_pipes = None
        try:
            import re

            txt = open(os.path.join(ROOT, "profiles", "r2_k1_ncu_summary.txt")).read()
            if args.tracks_per_gpu == 128 and args.seconds == 180.0:
                pick = lambda key: float(re.search(key + r"=([0-9.eE+-]+)", txt).group(1))  # noqa: E731
                other_pipes = {"fp32_fma_pipe_pct": pick("fma%"), "issue_slots_pct": pick("issue%"), "lsu_pipe_pct": pick("lsu%"),
                               "dram_pct": pick("dram%"), "occupancy_pct": pick("occ%"),
                               "shared_memory_bytes_per_16_frame_tile": 1.4e6, "shared_memory_floor_cycles_per_tile": 11000,
                               "cycles_per_tile": 27700, "source": "profiles/r2_k1_ncu_summary.txt, profiles/r2_notes.md"}
        except Exception:
            pass
        # algorithmic bytes of the other stages, per launch of this batch
        stage_bytes = {
            "stft_mel_features": k1_bytes,
            "onset_flux": nt * (4 * N_MELS * T + 12 * T),
            "autocorrelation": nt * (4 * T + 8 * T),
            "tempogram": nt * (4 * T + 4 * 384 * T),
            "chroma_stft": nt * (4 * B * T + 48 * T),
            "time_domain_loudness": nt * (4 * 2 * n),
        }
        stage_gbs = {k: stage_bytes[k] / (max(v, 1e-9) * 1e-3) / 1e9 for k, v in zip(engine.STAGE_NAMES, stage)}
        line = {
            "metric": METRIC, "value": world * audio_per_step / (ms_kernel_max * 1e-3), "unit": UNIT, "n_gpus": world,
            "steps": args.steps, "warmup": args.warmup, "ms_per_step": ms_kernel_max, "higher_is_better": True,
            "scaling": "weak", "vs_baseline": None, "dtype": "f32", "data": "synthetic",
            "config": workload_config(args, world),
            "stage_ms": {k: float(v) for k, v in zip(engine.STAGE_NAMES, stage)},
            "stage_ms_note": "stages timed one after another; ms_per_step runs the time-domain pass on a second stream",
            "stage_hbm_frac": {k: float(v / peak) for k, v in stage_gbs.items()},
            "roofline": {"bound": "hbm", "kernel": "stft_fused_kernel<2048,16,stereo,4>", "achieved": achieved,
                         "peak": peak, "unit": "GB/s", "frac": achieved / peak, "traffic": traffic,
                         "peak_source": "measured (MEASURED_PEAKS.json)" if peaks else "fallback",
                         "algorithmic_bytes_per_launch": k1_bytes, "other_pipes": other_pipes},
            "e2e": {"value": world * audio_per_step / s_e2e_max, "unit": UNIT,
                    "h2d_bytes_per_step": int(e2e_up_total / world), "d2h_bytes_per_step": int(e2e_down_total / world),
                    "steps": e2e_steps, "s_per_step": s_e2e_max, "chunk_tracks": chunk,
                    "h2d_gbs_per_gpu": e2e_up_total / world / s_e2e_max / 1e9,
                    "d2h_gbs_per_gpu": e2e_down_total / world / s_e2e_max / 1e9,
                    "tracks_per_rank": per_rank_tracks,
                    "sharding": ("equal shards" if link_rates is None else
                                 "shards sized by the host-link rate every rank sustains while all ranks copy "
                                 "(sharding.partition(weights=...)); bytes are the per-GPU average")},
            "gpu_launches": int(launches),
            "verified": (verified or {}).get("ok"),
            "verification": verified,
            "clocks": clk.summary(),
            "host_binding": numa_note,
        }
        if pcie:
            # the busier direction of the end-to-end leg against what the link moves in that direction while both run
            # (all ranks together: the host's ceiling at N > 1)
            # N = 1: against the short both-directions probe.  N > 1: against what the links sustain, summed over ranks, while
            # every rank runs the leg's own copy pattern for a whole step (the host's ceiling)
            sus = pcie.get("sustained_step_pattern")
            lim = sus["sum_down_gbs"] if sus else pcie["all_ranks_together"]["bidir_each_gbs"]
            busy = world * max(line["e2e"]["h2d_gbs_per_gpu"], line["e2e"]["d2h_gbs_per_gpu"])
            line["e2e"]["pcie"] = pcie
            line["e2e"]["pcie_ceiling_gbs"] = lim
            line["e2e"]["pcie_frac"] = busy / lim if lim else None
            # per-GPU efficiency of the end-to-end leg against one GPU that has the host to itself
            share_now = (sus["sum_down_gbs"] / world) if sus else pcie["all_ranks_together"]["bidir_each_gbs"]
            line["e2e"]["per_gpu_link_share"] = (share_now / pcie["alone"]["bidir_each_gbs"]
                                                 if pcie["alone"] and pcie["alone"]["bidir_each_gbs"] else None)
        if e2e_analysis:
            line["e2e_analysis_outputs"] = e2e_analysis
        if e2e_track:
            line["e2e_analyse_track"] = e2e_track
        if cpu_baseline:
            line["cpu_baseline"] = cpu_baseline
        print(json.dumps(line), flush=True)
    if world > 1:
        dist.destroy_process_group()


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--gpus", type=int, default=1)
    ap.add_argument("--steps", type=int, default=20)
    ap.add_argument("--warmup", type=int, default=3)
    ap.add_argument("--impl", default="ours", choices=["ours", "reference"])
    ap.add_argument("--tracks-per-gpu", type=int, default=128)
    ap.add_argument("--pool", type=int, default=16, help="distinct synthetic tracks per GPU (cycled to fill the batch)")
    ap.add_argument("--seconds", type=float, default=180.0)
    ap.add_argument("--chunk-tracks", type=int, default=8)
    ap.add_argument("--e2e-steps", type=int, default=3)
    ap.add_argument("--ref-seconds", type=float, default=60.0, help="track length of the bounded CPU sample (one track per worker)")
    ap.add_argument("--no-cpu-baseline", action="store_true")
    ap.add_argument("--no-analysis-leg", action="store_true", help="skip the informational byte-reduced end-to-end leg")
    ap.add_argument("--analysis-tracks", type=int, default=128, help="tracks per GPU of the pipeline.analyse_tracks leg (0: skip)")
    ap.add_argument("--no-verify", action="store_true", help="skip the oracle check of one track of the batch after the timed region")
    ap.add_argument("--verify-seconds", type=float, default=30.0)
    ap.add_argument("--no-pcie-probe", action="store_true")
    args = ap.parse_args()
    if args.warmup < 3 and args.impl == "ours":
        args.warmup = 3
    rank = int(os.environ.get("RANK", "0"))
    world = int(os.environ.get("WORLD_SIZE", "1"))
    local_rank = int(os.environ.get("LOCAL_RANK", "0"))
    if args.impl == "reference":
        reference_arm(args, rank, world)
    else:
        ours(args, rank, world, local_rank)


if __name__ == "__main__":
    main()
