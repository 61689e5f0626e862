"""Structural segmentation, frontend half (mirror of ``analysis/structure.py:48-59,190-196``).

``structure_frontend`` returns what the reference computes with librosa.stft,
librosa.feature.melspectrogram, librosa.power_to_db and librosa.onset.onset_strength before its
HPSS / MFCC / peak-picking host logic (structure.py:52,199-342 -- SURVEY 8f rank 1, not yet on
the device): the magnitude spectrogram, the mel power spectrogram, log-mel and the spectral
flux computed on LINEAR mel power in float64 (reference quirk).
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import List

import numpy as np

from .. import runtime
from ..utils import AudioInput, seed_everything


@dataclass(slots=True)
class StructuralSegment:
    label: str
    category: str
    start: float
    end: float
    confidence: float
    percussive_energy: float
    harmonic_energy: float
    percussive_ratio: float


@dataclass(slots=True)
class StructureAnalysis:
    segments: List[StructuralSegment]
    novelty_curve: List[float]


@dataclass(slots=True)
class StructureFrontend:
    magnitude: np.ndarray      # (1 + n_fft/2, T) float32  -- structure.py:48-51
    mel: np.ndarray            # (128, T) float32 power    -- structure.py:53-59
    log_mel: np.ndarray        # (128, T) float64          -- structure.py:194
    spectral_flux: np.ndarray  # (T,) float64              -- structure.py:195-196


def power_to_db(S: np.ndarray, amin: float = 1e-10, top_db: float = 80.0) -> np.ndarray:
    """librosa.power_to_db(ref=1.0) on the host for the (128, T) log-mel the MFCC stage consumes."""
    S = np.asarray(S)
    out = 10.0 * np.log10(np.maximum(amin, S))
    return np.maximum(out, out.max() - top_db) if out.size else out


def structure_frontend(audio: AudioInput, *, frame_length: int = 2048, hop_length: int = 512) -> StructureFrontend:
    if not isinstance(audio, AudioInput):
        raise TypeError("analyse_structure expects an AudioInput instance")
    res = runtime.frontend(np.asarray(audio.samples, dtype=np.float32), audio.sample_rate, n_fft=frame_length,
                           hop=hop_length, outputs=("magnitude", "mel", "flux_linear"))
    mel64 = np.asarray(res["mel"], dtype=float)
    return StructureFrontend(magnitude=res["magnitude"], mel=res["mel"], log_mel=power_to_db(mel64 + 1e-9),
                             spectral_flux=np.asarray(res["flux_linear"], dtype=float))
