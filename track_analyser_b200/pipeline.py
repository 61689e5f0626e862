"""Orchestration (mirror of the reference's ``pipeline.py``).

``TrackAnalysisResult`` keeps the reference's field names and order (pipeline.py:17-29) and
``analyse_track`` its signature and stage order (pipeline.py:32-120).  One ``frontend_session``
spans the call, so the >= 11 identical STFT requests of the reference become one
fused GPU run per distinct (buffer, n_fft, hop).  ``structure`` is the reference's ``StructureAnalysis`` (HPSS curves from csrc/hpss.cu, host logic
restated in analysis/structure.py); ``harmonic`` is the reference's ``HarmonyAnalysis`` with the STFT chroma standing in for
chroma_cqt (see harmony.py's docstring: the constant-Q transform is the one section-8f row not on the device).
"""

from __future__ import annotations

from dataclasses import dataclass
from pathlib import Path
from typing import Callable, Optional

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


def analyse_track(source, *, output_dir: Optional[str | Path] = None, use_stems: bool = False,
                  seed: int = DEFAULT_SEED, progress_callback: Optional[Callable[[str], None]] = None) -> TrackAnalysisResult:
    audio = source if isinstance(source, AudioInput) else coerce_audio(source)

    def tick(stage: str) -> None:
        if progress_callback:
            progress_callback(stage)

    tick("audio")
    with runtime.frontend_session():
        if audio.stereo_samples is not None:
            # mono == mid exactly (utils.py:116), so one fused run on the stereo buffer serves the mono stages too
            runtime.alias_mono_to_stereo(audio.samples, audio.stereo_samples)
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
    if use_stems:
        raise NotImplementedError("stem separation (demucs) is an optional model outside the frontend's scope")
    if output_dir is not None:
        raise NotImplementedError("rendering/export is outside the frontend's scope; pass the result to the reference's renderer")
    return TrackAnalysisResult(audio=audio, beat=beat_result, downbeat=downbeat_result, structure=structure_result,
                               loudness=loudness_result, harmonic=harmonic_result, features=feature_result,
                               stereo=stereo_result, stems=None)
