"""Spectral summary features (mirror of the reference's ``features.py``).

Signatures and dataclasses follow /root/reference/src/track_analyser/features.py
(:18-63 containers, :66-149 functions); the arithmetic of librosa.stft /
spectral_centroid / spectral_rolloff runs in the fused STFT kernel
(csrc/stft_fused.cu) and only small per-frame series come back to the host.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import runtime
from .utils import AudioInput


@dataclass(slots=True)
class LongTermAverageSpectrum:
    frequencies: np.ndarray
    magnitude: np.ndarray

    def as_dict(self) -> dict[str, Sequence[float]]:
        return {"frequencies": self.frequencies.tolist(), "magnitude": self.magnitude.tolist()}


@dataclass(slots=True)
class FeatureSeries:
    values: np.ndarray

    @property
    def mean(self) -> float:
        return float(np.mean(self.values)) if self.values.size else 0.0

    @property
    def median(self) -> float:
        return float(np.median(self.values)) if self.values.size else 0.0

    @property
    def as_list(self) -> list[float]:
        return self.values.tolist()


@dataclass(slots=True)
class FeatureAnalysis:
    ltas: LongTermAverageSpectrum
    spectral_centroid: FeatureSeries
    spectral_rolloff: FeatureSeries


def _mono(samples: np.ndarray) -> np.ndarray:
    x = np.asarray(samples, dtype=np.float32)
    return np.mean(x, axis=0) if x.ndim > 1 else x


def compute_ltas(samples: np.ndarray, sample_rate: int, *, n_fft: int = 2_048, hop_length: int = 512,
                 window: str = "hann") -> LongTermAverageSpectrum:
    res = runtime.frontend(_mono(samples), sample_rate, n_fft=n_fft, hop=hop_length, outputs=("ltas",), window=window)
    return LongTermAverageSpectrum(frequencies=np.fft.rfftfreq(n=n_fft, d=1.0 / sample_rate), magnitude=res["ltas"])


def spectral_centroid_series(samples: np.ndarray, sample_rate: int, *, n_fft: int = 2_048,
                             hop_length: int = 512) -> FeatureSeries:
    res = runtime.frontend(_mono(samples), sample_rate, n_fft=n_fft, hop=hop_length, outputs=("centroid",))
    return FeatureSeries(values=np.asarray(res["centroid"], dtype=np.float64))


def spectral_rolloff_series(samples: np.ndarray, sample_rate: int, *, roll_percent: float = 0.85,
                            n_fft: int = 2_048, hop_length: int = 512) -> FeatureSeries:
    if not 0.0 < roll_percent < 1.0:
        raise ValueError("roll_percent must lie in the range (0, 1)")
    res = runtime.frontend(_mono(samples), sample_rate, n_fft=n_fft, hop=hop_length, roll_percent=roll_percent,
                           outputs=("rolloff_bin",))
    freqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sample_rate)
    return FeatureSeries(values=freqs[res["rolloff_bin"]])


def analyse_features(audio: AudioInput, *, n_fft: int = 2_048, hop_length: int = 512,
                     roll_percent: float = 0.85) -> FeatureAnalysis:
    with runtime.frontend_session():
        return FeatureAnalysis(
            ltas=compute_ltas(audio.samples, audio.sample_rate, n_fft=n_fft, hop_length=hop_length),
            spectral_centroid=spectral_centroid_series(audio.samples, audio.sample_rate, n_fft=n_fft,
                                                       hop_length=hop_length),
            spectral_rolloff=spectral_rolloff_series(audio.samples, audio.sample_rate, roll_percent=roll_percent,
                                                     n_fft=n_fft, hop_length=hop_length),
        )
