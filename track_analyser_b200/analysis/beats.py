"""Beat containers and grid assembly (mirror of the reference's ``analysis/beats.py``).

Pure host logic on the outputs of ``tempo`` (SURVEY.md section 2.1 marks it out of the hot path);
kept so that ``analyse_track`` returns the reference's dataclasses.  The optional madmom RNN
downbeat tracker of the reference (beats.py:117-141) is not available offline, so the reference's
own heuristic fallback (beats.py:144-155) is what runs, exactly as it does there without madmom.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List, Optional, Tuple

import numpy as np

from .. import hostlogic
from ..utils import AudioInput, seed_everything


@dataclass(slots=True)
class BeatAnalysis:
    bpm: float
    beat_times: List[float]
    beat_frames: List[int]
    confidence: float
    grid: Optional[object] = None  # pandas.DataFrame


@dataclass(slots=True)
class DownbeatAnalysis:
    downbeat_times: List[float]
    beat_positions: List[int]
    source: str


def _compute_confidence(beat_times: np.ndarray) -> float:
    if len(beat_times) < 2:
        return 0.0
    gaps = np.diff(beat_times)
    if np.allclose(gaps, gaps[0]):
        return 1.0
    return float(np.clip(1.0 - np.std(gaps) / (np.mean(gaps) + 1e-9), 0.0, 1.0))


def build_beat_analysis(bpm: float, beat_times: np.ndarray, sr: int, *, hop_length: int = 512, grid=None) -> BeatAnalysis:
    beat_times = np.asarray(beat_times, dtype=float)
    frames = hostlogic.time_to_frames(beat_times, sr, hop_length)
    return BeatAnalysis(bpm=float(bpm), beat_times=beat_times.astype(float).tolist(),
                        beat_frames=frames.astype(int).tolist(), confidence=_compute_confidence(beat_times),
                        grid=grid.copy() if grid is not None else None)


def _fallback_downbeats(beat_result: BeatAnalysis) -> DownbeatAnalysis:
    positions, downbeats = [], []
    for idx, t in enumerate(beat_result.beat_times):
        if idx % 4 == 0:
            downbeats.append(float(t))
            positions.append(1)
        else:
            positions.append((idx % 4) + 1)
    return DownbeatAnalysis(downbeat_times=downbeats, beat_positions=positions, source="heuristic")


def analyse_downbeats(audio: AudioInput | str, beat_result: BeatAnalysis, *, hop_length: int = 512,
                      seed: int) -> Optional[DownbeatAnalysis]:
    if not isinstance(audio, AudioInput):
        raise TypeError("analyse_downbeats expects an AudioInput instance")
    seed_everything(seed)
    return _fallback_downbeats(beat_result)


def analyse_beats(audio: AudioInput | str, *, hop_length: int = 512, seed: int) -> Tuple[BeatAnalysis, Optional[DownbeatAnalysis]]:
    from ..tempo import beat_grid, estimate_bpm
    from .. import runtime

    seed_everything(seed)
    if not isinstance(audio, AudioInput):
        raise TypeError("analyse_beats expects an AudioInput instance")
    with runtime.frontend_session():
        grid = beat_grid(audio.samples, audio.sample_rate, hop_length=hop_length)
        bpm = estimate_bpm(audio.samples, audio.sample_rate, hop_length=hop_length)
    beat = build_beat_analysis(bpm, grid["time"].to_numpy(), audio.sample_rate, hop_length=hop_length, grid=grid)
    return beat, analyse_downbeats(audio, beat, hop_length=hop_length, seed=seed)
