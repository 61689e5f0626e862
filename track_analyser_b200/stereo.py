"""Stereo image metrics (mirror of the reference's ``stereo.py``).

Signatures/containers follow /root/reference/src/track_analyser/stereo.py
(:20-39, :42-153).  The passes over the PCM (moments) and over the mid/side
spectra (per-bin energy sums) run on the GPU; the host only combines a handful
of sums per track.
"""

from __future__ import annotations

from dataclasses import dataclass
from typing import Sequence

import numpy as np

from . import runtime
from .utils import AudioInput

_EPS = 1e-12


@dataclass(slots=True)
class StereoWidthBands:
    low: float
    mid: float
    high: float

    def as_dict(self) -> dict[str, float]:
        return {"low": self.low, "mid": self.mid, "high": self.high}


@dataclass(slots=True)
class StereoAnalysis:
    mid_rms: float
    side_rms: float
    correlation: float
    width: StereoWidthBands


def _ensure_stereo_array(audio: AudioInput) -> np.ndarray:
    if audio.stereo_samples is None:
        mono = np.asarray(audio.samples, dtype=np.float32)
        return np.vstack([mono, mono]) if mono.ndim == 1 else mono[:2]
    st = np.asarray(audio.stereo_samples, dtype=np.float32)
    if st.ndim == 1:
        return np.vstack([st, st])
    if st.shape[0] == 2:
        return st
    if st.shape[1] == 2:
        return np.transpose(st)
    if st.shape[0] < 2:
        return np.vstack([st[0], st[0]])
    return st[:2]


# ---- host combination of GPU sums ---------------------------------------------------
# moments = [sum L, sum R, sum L^2, sum R^2, sum LR, sum mid^2, sum side^2, n]

def mid_side_from_moments(m) -> tuple[float, float]:
    n = float(m[7])
    if n == 0:
        return 0.0, 0.0
    return float(np.sqrt(m[5] / n)), float(np.sqrt(m[6] / n))


def correlation_from_moments(m) -> float:
    n = float(m[7])
    if n == 0:
        return 1.0
    var_l = max(m[2] - m[0] * m[0] / n, 0.0)
    var_r = max(m[3] - m[1] * m[1] / n, 0.0)
    denom = float(np.sqrt(var_l) * np.sqrt(var_r))
    if denom <= _EPS:
        return 1.0
    return float(np.clip((m[4] - m[0] * m[1] / n) / denom, -1.0, 1.0))


def width_from_band_energy(band_energy, freqs, n_frames, bands, sample_rate, moments=None) -> dict[str, float]:
    """Band-wise sqrt(mean |S|^2 / mean |M|^2) from the per-bin time sums of the STFT kernel.

    If the time-domain side signal is identically zero (``moments[6] == 0``: L == R sample for
    sample) its STFT is identically zero too, and the reference returns exactly 0.0
    (tests/test_stereo.py:15-27); the packed fp32 FFT would instead leave rounding noise of the
    mid channel in the side spectrum, so that case is answered exactly here."""
    nyq = sample_rate / 2.0
    if bands is None:
        bands = (("low", 0.0, min(200.0, nyq)), ("mid", 200.0, min(2_000.0, nyq)), ("high", 2_000.0, nyq))
    out = {"low": 0.0, "mid": 0.0, "high": 0.0}
    if moments is not None and float(moments[6]) == 0.0:
        return out
    for name, lo, hi in bands:
        sel = (freqs >= lo) & (freqs <= hi)
        if not np.any(sel):
            out[name] = 0.0
            continue
        cnt = float(np.count_nonzero(sel)) * float(n_frames)
        mid_e = float(np.sum(band_energy[0][sel]) / cnt)
        side_e = float(np.sum(band_energy[1][sel]) / cnt)
        out[name] = 0.0 if mid_e <= _EPS else float(np.sqrt(side_e / mid_e))
    return out


def _as_pair(stereo: np.ndarray) -> np.ndarray:
    a = np.asarray(stereo, dtype=np.float32)
    if a.ndim == 2 and a.shape[0] == 2 and a.flags.c_contiguous:
        return a  # already the planar pair: keep the buffer so a frontend session recognises it
    left, right = a  # same unpacking (and the same ValueError for anything but two rows) as stereo.py:62
    return np.ascontiguousarray(np.stack([left, right]))


def mid_side_rms(stereo: np.ndarray) -> tuple[float, float]:
    st = _as_pair(stereo)
    if st.shape[1] == 0:
        return 0.0, 0.0
    return mid_side_from_moments(runtime.frontend(st, 44_100, outputs=("moments",))["moments"])


def mono_compatibility_correlation(stereo: np.ndarray) -> float:
    st = _as_pair(stereo)
    if st.shape[1] == 0:
        return 1.0
    return correlation_from_moments(runtime.frontend(st, 44_100, outputs=("moments",))["moments"])


def frequency_dependent_width(stereo: np.ndarray, sample_rate: int, *,
                              bands: Sequence[tuple[str, float, float]] | None = None, n_fft: int = 2_048,
                              hop_length: int = 512) -> StereoWidthBands:
    st = _as_pair(stereo)
    res = runtime.frontend(st, sample_rate, n_fft=n_fft, hop=hop_length, outputs=("band_energy", "moments"))
    freqs = np.fft.rfftfreq(n=n_fft, d=1.0 / sample_rate)
    w = width_from_band_energy(res["band_energy"], freqs, res.n_frames, bands, sample_rate, res["moments"])
    return StereoWidthBands(low=w.get("low", 0.0), mid=w.get("mid", 0.0), high=w.get("high", 0.0))


def analyse_stereo(audio: AudioInput, *, n_fft: int = 2_048, hop_length: int = 512,
                   bands: Sequence[tuple[str, float, float]] | None = None) -> StereoAnalysis:
    st = np.ascontiguousarray(_ensure_stereo_array(audio))
    with runtime.frontend_session():
        res = runtime.frontend(st, audio.sample_rate, n_fft=n_fft, hop=hop_length,
                               outputs=("moments", "band_energy"))
        mid_v, side_v = mid_side_from_moments(res["moments"]) if st.shape[1] else (0.0, 0.0)
        corr = correlation_from_moments(res["moments"]) if st.shape[1] else 1.0
        width = frequency_dependent_width(st, audio.sample_rate, bands=bands, n_fft=n_fft, hop_length=hop_length)
    return StereoAnalysis(mid_rms=mid_v, side_rms=side_v, correlation=corr, width=width)
