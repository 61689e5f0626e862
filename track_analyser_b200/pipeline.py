"""Orchestration (mirror of the reference's ``pipeline.py``).

``TrackAnalysisResult`` keeps the reference's field names and order (pipeline.py:17-29) and
``analyse_track`` its signature and stage order (pipeline.py:32-120).  One ``frontend_session``
spans the call, so the >= 11 identical STFT requests of the reference become one
fused GPU run per distinct (buffer, n_fft, hop).  ``structure`` is the reference's ``StructureAnalysis`` (HPSS curves
from csrc/hpss.cu, host logic restated in analysis/structure.py); ``harmonic`` is the reference's ``HarmonyAnalysis``
(chroma_stft and chroma_cqt from csrc/chroma.cu and csrc/cqt.cu).  ``analyse_tracks`` is the batch form of the same call:
kernels on chunks of tracks, the per-track host stages in a pool of worker processes.
"""

from __future__ import annotations

import os
from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Optional

import numpy as np

from . import features, harmony, runtime, stereo
from .analysis import beats, loudness, structure
from .tempo import beat_grid, estimate_bpm
from .utils import DEFAULT_SEED, AudioInput, coerce_audio


@dataclass
class TrackAnalysisResult:
    audio: AudioInput
    beat: beats.BeatAnalysis
    downbeat: Optional[beats.DownbeatAnalysis]
    structure: object
    loudness: loudness.LoudnessAnalysis
    harmonic: object
    features: features.FeatureAnalysis
    stereo: stereo.StereoAnalysis
    stems: Optional[object] = None


def _run_stages(audio: AudioInput, seed: int, tick) -> dict:
    """The stage sequence of pipeline.py:58-118 on one track; every frontend request goes through ``runtime.frontend``."""
    grid = beat_grid(audio.samples, audio.sample_rate)
    bpm = estimate_bpm(audio.samples, audio.sample_rate)
    beat_result = beats.build_beat_analysis(bpm, grid["time"].to_numpy(), audio.sample_rate, grid=grid)
    downbeat_result = beats.analyse_downbeats(audio, beat_result, seed=seed)
    tick("beats")
    structure_result = structure.analyse_structure(audio, beat_result, seed=seed)
    tick("structure")
    loudness_result = loudness.analyse_loudness(audio, seed=seed)
    tick("loudness")
    harmonic_result = harmony.analyse_harmony(audio, beat_result, downbeat_result, seed=seed)
    tick("harmonic")
    feature_result = features.analyse_features(audio)
    tick("features")
    stereo_result = stereo.analyse_stereo(audio)
    tick("stereo")
    return dict(beat=beat_result, downbeat=downbeat_result, structure=structure_result, loudness=loudness_result,
                harmonic=harmonic_result, features=feature_result, stereo=stereo_result)


def analyse_track(source, *, output_dir: Optional[str | Path] = None, use_stems: bool = False,
                  seed: int = DEFAULT_SEED, progress_callback: Optional[Callable[[str], None]] = None) -> TrackAnalysisResult:
    audio = source if isinstance(source, AudioInput) else coerce_audio(source)

    def tick(stage: str) -> None:
        if progress_callback:
            progress_callback(stage)

    if use_stems:
        raise NotImplementedError("stem separation (demucs) is an optional model outside the frontend's scope")
    if output_dir is not None:
        raise NotImplementedError("rendering/export is outside the frontend's scope; pass the result to the reference's renderer")
    tick("audio")
    with runtime.frontend_session():
        if audio.stereo_samples is not None:
            # mono == mid exactly (utils.py:116), so one fused run on the stereo buffer serves the mono stages too
            runtime.alias_mono_to_stereo(audio.samples, audio.stereo_samples)
        stages = _run_stages(audio, seed, tick)
    return TrackAnalysisResult(audio=audio, stems=None, **stages)


# ---------------------------------------------------------------------------------------------------------------------
# batch entry point: the kernels run on whole chunks of tracks, the per-track host stages in a pool of worker processes
# ---------------------------------------------------------------------------------------------------------------------
_PLAN_A = (2048, 512, 128)   # the default plan of beats / structure / loudness / harmony / features / stereo
_PLAN_B = (4096, 1024, 0)    # harmony._spectral_balance (harmony.py:253-267)
_ring = None
_pool = None
_pool_size = 0


def _shutdown():
    """Stop the worker processes and drop the shared-memory ring (also registered with atexit)."""
    global _pool, _ring
    if _pool is not None:
        _pool.shutdown(wait=True, cancel_futures=True)
        _pool = None
    if _ring is not None:
        _ring.close()
        _ring = None


import atexit  # noqa: E402

atexit.register(_shutdown)


def _worker_count(workers: Optional[int]) -> int:
    if workers is not None:
        return max(0, int(workers))
    try:
        n = len(os.sched_getaffinity(0))
    except AttributeError:  # pragma: no cover
        n = os.cpu_count() or 1
    return max(1, n - 1)


def _get_pool(n: int):
    """Forked once and kept: the children never touch CUDA (they only see ``runtime.precomputed_session``)."""
    global _pool, _pool_size
    if _pool is None or _pool_size != n:
        if _pool is not None:
            _pool.shutdown(wait=True, cancel_futures=True)
        import concurrent.futures as cf
        import multiprocessing as mp

        # everything the stages import lazily is loaded here once, so that the forked children inherit it
        import pandas  # noqa: F401
        import scipy.fft  # noqa: F401
        import scipy.ndimage  # noqa: F401
        import scipy.signal  # noqa: F401

        _pool = cf.ProcessPoolExecutor(max_workers=n, mp_context=mp.get_context("fork"), initializer=_worker_init)
        _pool_size = n
        # the executor forks its workers on demand: make it fork all of them now, outside anybody's timed region
        list(_pool.map(_worker_warm, [0.05] * (4 * n), chunksize=1))
    return _pool


def _placeholder_audio(meta: dict) -> AudioInput:
    """An AudioInput with the right shapes and untouched zero pages: inside a precomputed session the stages only hand the
    sample arrays to ``runtime.frontend``, which answers from the precomputed results without reading them."""
    samples = np.zeros(meta["mono_shape"], dtype=np.float32)
    stereo_ = None if meta["stereo_shape"] is None else np.zeros(meta["stereo_shape"], dtype=np.float32)
    return AudioInput(samples, meta["sample_rate"], meta["path"], stereo_)


def _load_results(path: str, base: int, specs: dict, index: int) -> dict:
    """{plan key: TrackResult} of track ``index`` of a chunk from the chunk's shared-memory file: the parent wrote every
    output array of the batch there once; the per-track pieces are views cut out of the mapping (``engine._cut``)."""
    from . import engine

    buf = np.memmap(path, dtype=np.uint8, mode="r")
    out = {}
    for key, (plan_g, batch_g, arrays) in specs.items():
        r = engine.TrackResult(n_samples=int(batch_g.n_samples[index]), n_frames=int(batch_g.n_frames[index]),
                               channels=batch_g.channels)
        for name, (off, shape, dtype) in arrays.items():
            h = np.ndarray(shape, dtype=dtype, buffer=buf, offset=base + off)
            r.data[name] = engine._cut(plan_g, batch_g, index, name, h)
        out[key] = r
    return out


def _worker_init():
    """The host stages are many small numpy / scipy calls: one BLAS / OpenMP thread per worker, or a dozen workers' thread
    pools fight over the same cores."""
    try:
        import threadpoolctl

        threadpoolctl.threadpool_limits(1)
    except Exception:  # pragma: no cover - best effort
        pass
    try:
        import torch

        torch.set_num_threads(1)
    except Exception:  # pragma: no cover
        pass


def _worker_warm(seconds: float) -> int:
    import time

    time.sleep(seconds)
    return os.getpid()


def _stage_worker(task):
    """Worker process: the host stages of one track on precomputed frontend results."""
    import time

    t0 = time.perf_counter()
    path, base, specs, index, meta, seed = task
    results = _load_results(path, base, specs, index)
    audio = _placeholder_audio(meta)
    with runtime.precomputed_session(results):
        stages = _run_stages(audio, seed, lambda stage: None)
    if os.environ.get("TA_TRACE_WORKERS"):
        import sys

        print(f"[ta worker {os.getpid()}] {(time.perf_counter() - t0) * 1e3:.1f} ms", file=sys.stderr)
    return stages


def _mono_fingerprint(mono) -> tuple:
    """(sum of the float32 bit patterns, sum of bit pattern * ((i & 0xffff) + 1)) modulo 2^32: the low halves of what
    ``ta_mono_mix_fingerprint`` forms on the device from the stereo pair.  All sums stay in uint32 (they wrap, which is the
    modulus): numpy then adds at memory speed instead of casting every sample to 64 bits."""
    b = np.ascontiguousarray(mono, dtype=np.float32).view(np.uint32)
    n = b.shape[0]
    m = n // 65536
    col = np.zeros(65536, dtype=np.uint32)
    with np.errstate(over="ignore"):
        if m:
            col += np.add.reduce(b[: m * 65536].reshape(m, 65536), axis=0, dtype=np.uint32)
        col[: n - m * 65536] += b[m * 65536:]
        s1 = int(np.add.reduce(col, dtype=np.uint32))
        s2 = int(np.add.reduce(col * np.arange(1, 65537, dtype=np.uint32), dtype=np.uint32))
    return s1, s2


def _batch_buffer(audio: AudioInput):
    """(buffer for the batched run, channels) from the shapes and dtypes alone: the planar float32 stereo pair, or the mono
    samples of a track without one; None for layouts the batched path does not take."""
    st = audio.stereo_samples
    mono = np.asarray(audio.samples)
    if mono.dtype != np.float32 or mono.ndim != 1:
        return None
    if st is None:
        return mono, 1
    st = np.asarray(st)
    if st.dtype != np.float32 or st.ndim != 2 or st.shape[0] != 2 or st.shape[1] != mono.shape[0] or not st.flags.c_contiguous:
        return None
    return st, 2


class _ShmRing:
    """A few slots of one /dev/shm file, registered with the CUDA driver as pinned memory: the parent copies a chunk's output
    arrays device -> slot at link speed, the workers map the same file and read the slot.  Slots are reused, so their pages
    are touched (and pinned) once, not per chunk."""

    def __init__(self, slots: int = 4):
        import tempfile

        shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
        fd, self.path = tempfile.mkstemp(prefix="ta_b200_", dir=shm_dir)
        os.close(fd)
        self.slots, self.slot_bytes, self.mm, self._registered = slots, 0, None, None

    def _unregister(self):
        if self._registered is not None:
            import torch

            torch.cuda.cudart().cudaHostUnregister(self._registered)
            self._registered = None

    def ensure(self, slot_bytes: int):
        """Grow the file so that every slot holds ``slot_bytes`` (only when no slot is in use)."""
        if slot_bytes <= self.slot_bytes:
            return
        import torch

        self._unregister()
        self.slot_bytes = (max(slot_bytes, 1 << 20) * 5 // 4 + 4095) & ~4095
        self.mm = np.memmap(self.path, dtype=np.uint8, mode="w+", shape=(self.slots * self.slot_bytes,))
        self.mm[:: 4096] = 0   # touch every page once
        ptr = self.mm.ctypes.data
        if int(torch.cuda.cudart().cudaHostRegister(ptr, self.mm.nbytes, 0)) == 0:
            self._registered = ptr

    def close(self):
        self._unregister()
        self.mm = None
        try:
            os.unlink(self.path)
        except OSError:
            pass


def analyse_tracks(sources, *, seed: int = DEFAULT_SEED, workers: Optional[int] = None, chunk_tracks: int = 8,
                   device: Optional[int] = None) -> list:
    """``[analyse_track(s, seed=seed) for s in sources]``, batched: the frontend kernels run on chunks of ``chunk_tracks``
    tracks (two fused runs per chunk: the 2048/512 plan with every output the host stages consume, the 4096/1024 plan of the
    spectral balance on the same resident PCM) while a pool of ``workers`` processes (default: all cores but one;
    0 = in this process) runs the beat / structure / loudness / harmony / feature / stereo host logic of earlier chunks on
    the downloaded arrays.  Three things overlap: a thread verifies and uploads the next chunk, this thread runs the
    kernels and copies their outputs into a pinned shared-memory ring, the workers run the host stages.  Mirrors
    pipeline.py:32-120 per track; results equal ``analyse_track``'s."""
    import concurrent.futures as cf
    import queue
    import sys
    import threading
    import time

    import ctypes as C

    import torch

    from . import _native as nat
    from . import engine

    global _ring
    trace = bool(os.environ.get("TA_TRACE_WORKERS"))
    audios = [s if isinstance(s, AudioInput) else coerce_audio(s) for s in sources]
    results: list = [None] * len(audios)
    if not audios:
        return results
    n_workers = _worker_count(workers)
    pool = _get_pool(n_workers) if n_workers > 0 else None   # forked before this call starts any thread of its own
    dev_index = torch.cuda.current_device() if device is None else int(device)
    main_stream = torch.cuda.current_stream(dev_index)

    # group by (sample rate, channel layout) from the shapes alone; whether a stereo pair may serve the mono view too
    # (mono == mean(stereo), sample for sample) is verified by the uploader thread chunk by chunk
    groups: dict = {}
    for i, a in enumerate(audios):
        st = a.stereo_samples
        ch = 2 if (st is not None and np.ndim(st) == 2 and np.shape(st)[0] == 2) else (1 if st is None else 0)
        groups.setdefault((int(a.sample_rate), ch), []).append(i)
    chunks = []
    for (sr, ch), idxs in groups.items():
        if ch == 0:   # layouts the batched path does not take ((1, N) / (N, 2) stereo_samples ...)
            for i in idxs:
                results[i] = analyse_track(audios[i], seed=seed)
            continue
        chunks += [(sr, ch, idxs[c0: c0 + chunk_tracks]) for c0 in range(0, len(idxs), chunk_tracks)]

    uploaded: "queue.Queue" = queue.Queue(maxsize=2)
    # device PCM buffers go round between the two threads (at most five exist: two being filled, two queued, one in use)
    # instead of being allocated per chunk: a buffer comes back once the kernels that read it have finished
    pcm_free: "queue.Queue" = queue.Queue()
    pcm_made = [0]

    def pcm_buffer(n_floats: int):
        while True:
            try:
                buf = pcm_free.get(block=pcm_made[0] >= 5)
            except queue.Empty:
                buf = None
            if buf is not None and buf.numel() >= n_floats:
                return buf
            if buf is not None:
                pcm_made[0] -= 1   # too small for this chunk: let it go
                continue
            pcm_made[0] += 1
            return torch.empty(n_floats, dtype=torch.float32, device=f"cuda:{dev_index}")

    def uploader():
        # Two halves per chunk, one chunk apart: `issue` enqueues the PCM copies and the device half of the fingerprints on
        # this thread's stream and returns; `complete` waits for them, compares and hands the chunk over.  The next chunk's
        # copies are already enqueued while this one's are being waited for, so the link never idles between chunks.
        try:
            stream = torch.cuda.Stream(dev_index)
            # pinned landing places of the device fingerprints, one per chunk in flight (allocated once: a pinned allocation
            # per chunk costs milliseconds and synchronises with the device)
            fp_host = [torch.empty(2 * max(1, chunk_tracks), dtype=torch.int64, pin_memory=True) for _ in range(3)]
            issued = [0]
            with cf.ThreadPoolExecutor(max_workers=8) as tp, torch.cuda.stream(stream):
                def issue(sr, ch, idxs):
                    st = dict(sr=sr, ch=ch, t0=time.perf_counter(), cand=[], bad=[], batch=None, buf=None, host_fp=[], fps=None,
                              done=None)
                    for i in idxs:
                        c = _batch_buffer(audios[i])
                        (st["cand"] if (c is not None and c[1] == ch and c[0].shape[-1] > 0) else st["bad"]).append((i, c))
                    cand = st["cand"]
                    if cand:
                        plan_a = runtime.get_plan(sr, *_PLAN_A, device=device)
                        # (the host half of the fingerprints runs in the pool while the PCM travels)
                        st["host_fp"] = [tp.submit(_mono_fingerprint, audios[i].samples) for i, _ in cand] if ch == 2 else []
                        need = sum((c[0].size + 3) & ~3 for _, c in cand)
                        st["buf"] = pcm_buffer(max(need, 4))
                        batch = st["batch"] = engine.upload(plan_a, [c[0] for _, c in cand], out=st["buf"], wait=False)
                        if ch == 2:
                            # one stereo run may serve the mono stages only if mono == mean(stereo) sample for sample
                            # (utils.py:116): two 64-bit sums over the bit patterns, formed on the device from the stereo PCM
                            # that is there anyway and on the host from the mono samples (numpy releases the GIL)
                            fps = torch.zeros(2 * len(cand), dtype=torch.int64, device=batch.pcm.device)
                            st_ptr = C.c_void_p(torch.cuda.current_stream(dev_index).cuda_stream)
                            for j in range(len(cand)):
                                nat.check(plan_a.lib.ta_mono_mix_fingerprint(
                                    C.c_void_p(batch.pcm.data_ptr() + 4 * int(batch.offsets[j])), int(batch.n_samples[j]),
                                    C.c_void_p(fps.data_ptr() + 16 * j), st_ptr))
                            st["fps"] = fp_host[issued[0] % 3][: 2 * len(cand)]
                            issued[0] += 1
                            st["fps"].copy_(fps, non_blocking=True)
                        st["done"] = torch.cuda.Event()
                        st["done"].record()
                    return st

                def complete(st):
                    cand, bad, batch, ch = st["cand"], st["bad"], st["batch"], st["ch"]
                    good = []
                    if cand:
                        st["done"].synchronize()
                        ok = [True] * len(cand)
                        if ch == 2:
                            dev_fp = st["fps"].numpy().view(np.uint64).reshape(-1, 2)
                            ok = [tuple(int(v) & 0xffffffff for v in dev_fp[j]) == st["host_fp"][j].result() for j in range(len(cand))]
                        if not all(ok):   # rare: re-upload only the consistent tracks, the others take the single-track path
                            bad += [ic for ic, o in zip(cand, ok) if not o]
                            cand = [ic for ic, o in zip(cand, ok) if o]
                            plan_a = runtime.get_plan(st["sr"], *_PLAN_A, device=device)
                            batch = engine.upload(plan_a, [c[0] for _, c in cand], out=st["buf"]) if cand else None
                            if batch is None:
                                pcm_free.put(st["buf"])
                        good = [(i, c[0]) for i, c in cand]
                    t1 = time.perf_counter()
                    uploaded.put((st["sr"], ch, good, [i for i, _ in bad], batch))
                    if trace:
                        print("[ta uploader] chunk ready %.1f ms after its copies were enqueued, queue %.1f ms"
                              % (1e3 * (t1 - st["t0"]), 1e3 * (time.perf_counter() - t1)), file=sys.stderr)

                prev = None
                for sr, ch, idxs in chunks:
                    cur = issue(sr, ch, idxs)
                    if prev is not None:
                        complete(prev)
                    prev = cur
                if prev is not None:
                    complete(prev)
            uploaded.put(None)
        except BaseException as exc:  # noqa: BLE001 - handed to the consuming thread
            uploaded.put(exc)

    import gc

    gc_was_on = gc.isenabled() and not os.environ.get("TA_KEEP_GC")
    if gc_was_on:
        gc.disable()   # a generation-2 pass over the unpickled results stops every thread of this process for ~0.1 s
    th = threading.Thread(target=uploader, daemon=True)
    th.start()
    if _ring is None:
        _ring = _ShmRing()
    ring, pending = _ring, []   # pending: (futures, indices, slot)
    buffers: dict = {}
    free_slots = list(range(ring.slots))

    def collect(entry):
        futs, idxs, slot = entry
        for i, f in zip(idxs, futs):
            stages = f.result() if hasattr(f, "result") else f
            results[i] = TrackAnalysisResult(audio=audios[i], stems=None, **stages)
        free_slots.append(slot)

    try:
        while True:
            _t = [time.perf_counter()]
            item = uploaded.get()
            if item is None:
                break
            if isinstance(item, BaseException):
                raise item
            sr, channels, good, bad, batch = item
            for i in bad:
                results[i] = analyse_track(audios[i], seed=seed)   # e.g. mono samples that are not the mean of the stereo pair
            if not good:
                continue
            _t.append(time.perf_counter())
            plan_a = runtime.get_plan(sr, *_PLAN_A, device=device)
            plan_b = runtime.get_plan(sr, *_PLAN_B, device=device)
            tracks = [b for _, b in good]
            outs_a = engine.available_outputs(plan_a, engine.ANALYSIS_OUTPUTS)
            if not all(t.shape[-1] >= plan_a.meter_block * sr for t in tracks):
                outs_a = tuple(o for o in outs_a if o not in ("kw_blocks", "lufs"))
            runs = [((*_PLAN_A, channels), plan_a, batch, outs_a), ((*_PLAN_B, channels), plan_b, batch.rebind(plan_b), ("ltas",))]
            if channels == 1:
                # a mono track's stereo stage analyses the duplicated channel pair (stereo.py:42-59): its own small run
                runs.append(((*_PLAN_A, 2), plan_a, engine.upload(plan_a, [np.vstack([t, t]) for t in tracks]),
                             ("moments", "band_energy")))
            launched = []
            for key, plan, bt, outs in runs:
                # output buffers are reused by the next chunk of the same geometry (this thread waits for the copies below
                # before it takes another chunk), which saves ~30 allocations per chunk
                bkey = (key, bt.n_samples.tobytes(), outs)
                bufs = buffers.get(bkey)
                if bufs is None:
                    while len(buffers) >= 3:   # the runs of one chunk; a chunk of another geometry replaces them
                        buffers.pop(next(iter(buffers)))
                    bufs = buffers[bkey] = engine.FrontendBuffers(bt, outs)
                engine.run_device(plan, bt, bufs)
                launched.append((key, plan, bt, bufs))
            # every requested output array of the batch goes to one slot of the shared-memory ring with one copy each; the
            # workers cut their track's views out of the mapping instead of receiving pickled arrays
            specs, cur = {}, 0
            for key, plan, bt, bufs in launched:
                arrays = {}
                for name in bufs.requested:
                    t = bufs.t[name]
                    arrays[name] = (cur, tuple(t.shape), np.dtype(str(t.dtype).replace("torch.", "")).str)
                    cur += (t.numel() * t.element_size() + 63) & ~63
                specs[key] = (plan.geometry(), bt.geometry(with_cqt="chroma_cqt" in bufs.requested), arrays)
            if cur > ring.slot_bytes:
                while pending:   # nobody may be reading the file while it is re-created
                    collect(pending.pop(0))
                ring.ensure(cur)
            while not free_slots:
                collect(pending.pop(0))
            slot = free_slots.pop(0)
            base = slot * ring.slot_bytes
            for key, plan, bt, bufs in launched:
                for name, (off, shape, dtype) in specs[key][2].items():
                    dst = np.ndarray(shape, dtype=dtype, buffer=ring.mm, offset=base + off)
                    torch.from_numpy(dst).copy_(bufs.t[name], non_blocking=True)
            main_stream.synchronize()
            pcm_free.put(batch.pcm._base if batch.pcm._base is not None else batch.pcm)   # the whole buffer, not the slice in use
            del launched, batch
            _t.append(time.perf_counter())
            tasks = []
            for j, (i, _) in enumerate(good):
                a = audios[i]
                meta = dict(sample_rate=a.sample_rate, path=a.path, mono_shape=np.asarray(a.samples).shape,
                            stereo_shape=None if a.stereo_samples is None else np.asarray(a.stereo_samples).shape)
                tasks.append((ring.path, base, specs, j, meta, seed))
            futs = [pool.submit(_stage_worker, t) for t in tasks] if pool is not None else [_stage_worker(t) for t in tasks]
            pending.append((futs, [i for i, _ in good], slot))
            _t.append(time.perf_counter())
            if trace:
                print("[ta analyse_tracks] chunk of %d: wait for upload %.1f, kernels + copy to shm %.1f, submit %.1f ms"
                      % ((len(good),) + tuple(1e3 * (b - a) for a, b in zip(_t, _t[1:]))), file=sys.stderr)
        for entry in pending:
            collect(entry)
        th.join()
    finally:
        if gc_was_on:
            gc.enable()
    return results
