"""Harmony analysis (mirror of the reference's ``harmony.py``), frontend half on the GPU.

On the section-8a path: ``_spectral_balance`` (harmony.py:253-267; a 4096/1024 STFT reduced to
three band ratios) and the ``chroma_stft`` projection (harmony.py:108,149).  ``chroma_cqt``
(harmony.py:107,148) is a multi-rate constant-Q transform outside that path;
until it has a kernel, the key/chord logic below runs on the STFT chroma only and says so in
``HarmonyAnalysis.chroma_source``.  Key scoring, chord hints, change points and MIDI suggestions
are small host-side decisions on (12, T) chroma, restated from harmony.py:192-465.
"""

from __future__ import annotations

from dataclasses import dataclass, field
from typing import List, Optional, Sequence, Tuple

import numpy as np

from . import runtime
from .analysis.beats import BeatAnalysis, DownbeatAnalysis
from .utils import AudioInput, deterministic_rng, seed_everything

MAJOR_PROFILE = np.array([6.35, 2.23, 3.48, 2.33, 4.38, 4.09, 2.52, 5.19, 2.39, 3.66, 2.29, 2.88])
MINOR_PROFILE = np.array([6.33, 2.68, 3.52, 5.38, 2.6, 3.53, 2.54, 4.75, 3.98, 2.69, 3.34, 3.17])
PITCH_CLASS_NAMES = ["C", "C#", "D", "Eb", "E", "F", "F#", "G", "Ab", "A", "Bb", "B"]


@dataclass(slots=True)
class SpectralBalance:
    low_band: float
    mid_band: float
    high_band: float


@dataclass(slots=True)
class StereoImage:
    correlation: float
    balance: float


@dataclass(slots=True)
class KeyEstimate:
    key: str
    confidence: float


@dataclass(slots=True)
class KeyEstimation:
    best: KeyEstimate
    second_best: KeyEstimate


@dataclass(slots=True)
class ChordHint:
    time: float
    chord: str
    confidence: float


@dataclass(slots=True)
class ChordChangePoint:
    time: float
    strength: float


@dataclass(slots=True)
class HarmonyFrontend:
    """GPU outputs of the harmony stage for one track."""

    spectral_balance: SpectralBalance
    chroma_stft: np.ndarray
    tuning: float


def chroma_stft(y: np.ndarray, sr: int) -> Tuple[np.ndarray, float]:
    """librosa.feature.chroma_stft(y=y, sr=sr) -> ((12, T) float32, tuning)."""
    res = runtime.frontend(np.asarray(y, dtype=np.float32), sr, outputs=("chroma", "tuning"))
    return res["chroma"], res["tuning"]


def _spectral_balance(audio: AudioInput) -> SpectralBalance:
    sr = audio.sample_rate
    res = runtime.frontend(np.asarray(audio.samples, dtype=np.float32), sr, n_fft=4096, hop=1024, n_mels=0,
                           outputs=("ltas",))
    sums = res["ltas"].astype(np.float64) * res.n_frames  # per-bin time sums of |X|
    total = float(np.sum(sums))
    if total <= 0:
        return SpectralBalance(0.0, 0.0, 0.0)
    freqs = np.fft.rfftfreq(n=4096, d=1.0 / sr)
    lo, mid, hi = freqs < 200, (freqs >= 200) & (freqs < 2000), freqs >= 2000
    return SpectralBalance(float(sums[lo].sum() / total), float(sums[mid].sum() / total), float(sums[hi].sum() / total))


def _stereo_image(audio: AudioInput) -> StereoImage:
    samples = audio.stereo_samples if audio.stereo_samples is not None else audio.samples
    samples = np.asarray(samples, dtype=np.float32)
    if samples.ndim == 1 or samples.shape[0] < 2:
        return StereoImage(correlation=1.0, balance=0.0)
    left, right = samples[0], samples[1]
    corr = float(np.corrcoef(left, right)[0, 1]) if left.size and right.size else 0.0
    return StereoImage(correlation=corr, balance=float(np.mean(np.abs(left)) - np.mean(np.abs(right))))


def _key_names() -> List[str]:
    return [f"{p} major" for p in PITCH_CLASS_NAMES] + [f"{p} minor" for p in PITCH_CLASS_NAMES]


def _score_keys(chroma_matrices: Sequence[np.ndarray]) -> Tuple[np.ndarray, List[str]]:
    if not chroma_matrices:
        return np.array([]), []
    major = MAJOR_PROFILE / np.linalg.norm(MAJOR_PROFILE)
    minor = MINOR_PROFILE / np.linalg.norm(MINOR_PROFILE)
    total = np.zeros(24, dtype=float)
    for chroma in chroma_matrices:
        if chroma.size == 0:
            continue
        mean = np.mean(chroma, axis=1)
        nrm = np.linalg.norm(mean)
        if nrm <= 0:
            continue
        mean = mean / nrm
        total[:12] += [float(np.dot(mean, np.roll(major, s))) for s in range(12)]
        total[12:] += [float(np.dot(mean, np.roll(minor, s))) for s in range(12)]
    return total, _key_names()


def _rank_keys(scores: np.ndarray, keys: List[str]) -> KeyEstimation:
    if not scores.size:
        fallback = KeyEstimate(key="C major", confidence=0.0)
        return KeyEstimation(best=fallback, second_best=fallback)
    pos = np.maximum(scores, 0.0)
    conf = pos / (float(np.sum(pos)) or 1.0)
    first = int(np.argmax(conf))
    best = KeyEstimate(key=keys[first], confidence=float(conf[first]))
    conf[first] = -np.inf
    second = int(np.argmax(conf))
    return KeyEstimation(best=best, second_best=KeyEstimate(key=keys[second], confidence=float(max(conf[second], 0.0))))


def key_estimate(y: np.ndarray, sr: int) -> KeyEstimation:
    """Best and second-best key from the STFT chroma (the reference also adds chroma_cqt scores)."""
    chroma, _ = chroma_stft(y, sr)
    return _rank_keys(*_score_keys([chroma]))


def key_index(estimate: KeyEstimation) -> int:
    """Integer index (0..23) of the best key: the 'key index' integer output of the north star."""
    return _key_names().index(estimate.best.key)


def harmony_frontend(audio: AudioInput) -> HarmonyFrontend:
    if not isinstance(audio, AudioInput):
        raise TypeError("analyse_harmony expects an AudioInput instance")
    chroma, tuning = chroma_stft(audio.samples, audio.sample_rate)
    return HarmonyFrontend(spectral_balance=_spectral_balance(audio), chroma_stft=chroma, tuning=tuning)
