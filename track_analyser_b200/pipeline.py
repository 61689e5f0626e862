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
_pool = None
_pool_size = 0


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

        _pool = cf.ProcessPoolExecutor(max_workers=n, mp_context=mp.get_context("fork"))
        _pool_size = n
    return _pool


def _placeholder_audio(meta: dict) -> AudioInput:
    """An AudioInput with the right shapes and untouched zero pages: inside a precomputed session the stages only hand the
    sample arrays to ``runtime.frontend``, which answers from the precomputed results without reading them."""
    samples = np.zeros(meta["mono_shape"], dtype=np.float32)
    stereo_ = None if meta["stereo_shape"] is None else np.zeros(meta["stereo_shape"], dtype=np.float32)
    return AudioInput(samples, meta["sample_rate"], meta["path"], stereo_)


def _load_results(path: str, layout: dict) -> dict:
    """{plan key: TrackResult} from the chunk's shared-memory file: zero-copy views of what the parent wrote."""
    from .engine import TrackResult

    buf = np.memmap(path, dtype=np.uint8, mode="r")
    out = {}
    for key, entry in layout.items():
        r = TrackResult(n_samples=entry["n_samples"], n_frames=entry["n_frames"], channels=entry["channels"])
        for name, (off, shape, dtype, scalar) in entry["arrays"].items():
            a = np.ndarray(shape, dtype=dtype, buffer=buf, offset=off)
            r.data[name] = a.reshape(()).item() if scalar else a
        out[key] = r
    return out


def _stage_worker(task):
    """Worker process: the host stages of one track on precomputed frontend results."""
    path, layout, meta, seed = task
    results = _load_results(path, layout)
    audio = _placeholder_audio(meta)
    with runtime.precomputed_session(results):
        return _run_stages(audio, seed, lambda stage: None)


def _stereo_buffer(audio: AudioInput):
    """(buffer for the batched run, channels, aliased) -- the planar stereo pair when the mono samples are exactly its
    mean (one run then serves the mono stages too), else None: such a track goes through ``analyse_track``."""
    st = audio.stereo_samples
    mono = np.asarray(audio.samples)
    if mono.dtype != np.float32 or mono.ndim != 1:
        return None
    if st is None:
        return mono, 1
    st = np.asarray(st)
    if st.dtype != np.float32 or st.ndim != 2 or st.shape[0] != 2 or st.shape[1] != mono.shape[0] or not st.flags.c_contiguous:
        return None
    with runtime.frontend_session():
        if not runtime.alias_mono_to_stereo(mono, st):
            return None
    return st, 2


def analyse_tracks(sources, *, seed: int = DEFAULT_SEED, workers: Optional[int] = None, chunk_tracks: int = 8,
                   device: Optional[int] = None) -> list:
    """``[analyse_track(s, seed=seed) for s in sources]``, batched: the frontend kernels run on chunks of ``chunk_tracks``
    tracks (two fused runs per chunk: the 2048/512 plan with every output the host stages consume, the 4096/1024 plan of the
    spectral balance on the same resident PCM) while a pool of ``workers`` processes (default: all cores but one;
    0 = in this process) runs the beat / structure / loudness / harmony / feature / stereo host logic of earlier chunks on
    the downloaded arrays.  Mirrors pipeline.py:32-120 per track; results equal ``analyse_track``'s."""
    import tempfile

    from . import engine

    import concurrent.futures as cf

    audios = [s if isinstance(s, AudioInput) else coerce_audio(s) for s in sources]
    results: list = [None] * len(audios)
    groups: dict = {}
    # mono == mean(stereo) is verified sample for sample before one stereo run may serve both views: numpy releases the
    # GIL in these passes, so a few threads do it for all tracks at once
    with cf.ThreadPoolExecutor(max_workers=min(8, max(1, len(audios)))) as tp:
        checked = list(tp.map(_stereo_buffer, audios))
    for i, a in enumerate(audios):
        buf = checked[i]
        if buf is None or buf[0].shape[-1] == 0:
            results[i] = analyse_track(a, seed=seed)   # layouts the batched path does not take
            continue
        groups.setdefault((int(a.sample_rate), buf[1]), []).append((i, buf[0]))
    n_workers = _worker_count(workers)
    pool = _get_pool(n_workers) if n_workers > 0 else None
    shm_dir = "/dev/shm" if os.path.isdir("/dev/shm") else tempfile.gettempdir()
    pending = []   # (futures, indices, path)

    def collect(entry):
        futs, idxs, path = entry
        for i, f in zip(idxs, futs):
            stages = f.result() if hasattr(f, "result") else f
            results[i] = TrackAnalysisResult(audio=audios[i], stems=None, **stages)
        try:
            os.unlink(path)
        except OSError:
            pass

    for (sr, channels), items in groups.items():
        plan_a = runtime.get_plan(sr, *_PLAN_A, device=device)
        plan_b = runtime.get_plan(sr, *_PLAN_B, device=device)
        outs_a = engine.available_outputs(plan_a, engine.ANALYSIS_OUTPUTS)
        for c0 in range(0, len(items), chunk_tracks):
            chunk = items[c0: c0 + chunk_tracks]
            tracks = [b for _, b in chunk]
            batch = engine.upload(plan_a, tracks)
            long_enough = all(t.shape[-1] >= plan_a.meter_block * sr for t in tracks)
            oa = outs_a if long_enough else tuple(o for o in outs_a if o not in ("kw_blocks", "lufs"))
            res_a = engine.analyse_batch(plan_a, tracks, oa, resident=batch)
            res_b = engine.analyse_batch(plan_b, tracks, ("ltas",), resident=batch)
            # a mono track's stereo stage analyses the duplicated channel pair (stereo.py:42-59): its own small run
            res_d = (engine.analyse_batch(plan_a, [np.vstack([t, t]) for t in tracks], ("moments", "band_energy"))
                     if channels == 1 else [None] * len(tracks))
            # one shared-memory file per chunk: the workers map it instead of receiving pickled arrays
            fd, path = tempfile.mkstemp(prefix="ta_b200_", dir=shm_dir)
            os.close(fd)
            layouts, cur, blobs = [], 0, []
            for ra, rb, rd in zip(res_a, res_b, res_d):
                layout = {}
                for key, r in (((*_PLAN_A, channels), ra), ((*_PLAN_B, channels), rb)) + ((((*_PLAN_A, 2), rd),) if rd is not None else ()):
                    arrays = {}
                    for name, v in r.data.items():
                        a = np.ascontiguousarray(v)
                        arrays[name] = (cur, a.shape, a.dtype.str, not isinstance(v, np.ndarray))
                        blobs.append((cur, a))
                        cur += (a.nbytes + 63) & ~63
                    layout[key] = dict(n_samples=r.n_samples, n_frames=r.n_frames, channels=r.channels, arrays=arrays)
                layouts.append(layout)
            mm = np.memmap(path, dtype=np.uint8, mode="w+", shape=(max(cur, 64),))
            for off, a in blobs:
                mm[off: off + a.nbytes] = a.reshape(-1).view(np.uint8)
            mm.flush()
            del mm
            tasks = []
            for (i, _), layout in zip(chunk, layouts):
                a = audios[i]
                meta = dict(sample_rate=a.sample_rate, path=a.path, mono_shape=np.asarray(a.samples).shape,
                            stereo_shape=None if a.stereo_samples is None else np.asarray(a.stereo_samples).shape)
                tasks.append((path, layout, meta, seed))
            futs = [pool.submit(_stage_worker, t) for t in tasks] if pool is not None else [_stage_worker(t) for t in tasks]
            pending.append((futs, [i for i, _ in chunk], path))
            while len(pending) > 2:   # keep the device at most two chunks ahead of the host stages
                collect(pending.pop(0))
    for entry in pending:
        collect(entry)
    return results
