"""CPU restatement of the reference's frontend call sites.

TEST INFRASTRUCTURE (see ``oracle/__init__.py``).  Each function names the
reference file:line it follows under /root/reference/src/track_analyser and
returns plain numpy values, so the parity tests can compare the CUDA path
tensor by tensor.
"""

from __future__ import annotations

import numpy as np

from . import librosa_np as lr
from . import pyloudnorm_np as pl

EPS = 1e-12


def _as_mono(samples):
    x = np.asarray(samples, dtype=np.float32)
    return np.mean(x, axis=0) if x.ndim > 1 else x


# ---- features.py -----------------------------------------------------------


def compute_ltas(samples, sample_rate, n_fft=2048, hop_length=512):
    """features.py:66-82 -> (frequencies f64, magnitude f32)."""
    D = lr.stft(_as_mono(samples), n_fft=n_fft, hop_length=hop_length)
    return lr.fft_frequencies(sample_rate, n_fft), np.mean(np.abs(D), axis=1)


def spectral_centroid_series(samples, sample_rate, n_fft=2048, hop_length=512):
    """features.py:85-100."""
    return lr.spectral_centroid(_as_mono(samples), sample_rate, n_fft, hop_length)[0]


def spectral_rolloff_series(samples, sample_rate, roll_percent=0.85, n_fft=2048, hop_length=512):
    """features.py:103-123."""
    return lr.spectral_rolloff(_as_mono(samples), sample_rate, n_fft, hop_length, roll_percent)[0]


# ---- stereo.py -------------------------------------------------------------


def ensure_stereo(samples, stereo_samples):
    """stereo.py:42-59 (_ensure_stereo_array) on the two AudioInput arrays."""
    if stereo_samples is None:
        mono = np.asarray(samples, dtype=np.float32)
        return np.vstack([mono, mono]) if mono.ndim == 1 else mono[:2]
    st = np.asarray(stereo_samples, dtype=np.float32)
    if st.ndim == 1:
        return np.vstack([st, st])
    if st.shape[0] == 2:
        return st
    if st.shape[1] == 2:
        return st.T
    if st.shape[0] < 2:
        return np.vstack([st[0], st[0]])
    return st[:2]


def mid_side_rms(stereo):
    """stereo.py:62-70."""
    left, right = np.asarray(stereo, dtype=np.float32)
    if left.size == 0:
        return 0.0, 0.0
    mid = 0.5 * (left + right)
    side = 0.5 * (left - right)
    return float(np.sqrt(np.mean(np.square(mid)))), float(np.sqrt(np.mean(np.square(side))))


def mono_compatibility_correlation(stereo):
    """stereo.py:73-83."""
    left, right = np.asarray(stereo, dtype=np.float32)
    if left.size == 0:
        return 1.0
    lc = left - np.mean(left)
    rc = right - np.mean(right)
    denom = float(np.linalg.norm(lc) * np.linalg.norm(rc))
    if denom <= EPS:
        return 1.0
    return float(np.clip(float(np.dot(lc, rc) / denom), -1.0, 1.0))


def default_bands(sample_rate):
    nyq = sample_rate / 2.0
    return (("low", 0.0, min(200.0, nyq)), ("mid", 200.0, min(2000.0, nyq)), ("high", 2000.0, nyq))


def frequency_dependent_width(stereo, sample_rate, bands=None, n_fft=2048, hop_length=512):
    """stereo.py:86-128 -> dict(low, mid, high)."""
    left, right = np.asarray(stereo, dtype=np.float32)
    DL = lr.stft(left, n_fft=n_fft, hop_length=hop_length)
    DR = lr.stft(right, n_fft=n_fft, hop_length=hop_length)
    mid_e = np.abs(0.5 * (DL + DR)) ** 2
    side_e = np.abs(0.5 * (DL - DR)) ** 2
    freqs = lr.fft_frequencies(sample_rate, n_fft)
    out = {"low": 0.0, "mid": 0.0, "high": 0.0}
    for name, lo, hi in (bands or default_bands(sample_rate)):
        sel = (freqs >= lo) & (freqs <= hi)
        if not sel.any():
            out[name] = 0.0
            continue
        m = float(np.mean(mid_e[sel]))
        s = float(np.mean(side_e[sel]))
        out[name] = 0.0 if m <= EPS else float(np.sqrt(s / m))
    return out


# ---- analysis/loudness.py --------------------------------------------------


def windowed_loudness(samples, sample_rate, meter_block_size):
    """analysis/loudness.py:30-42 (unweighted RMS dB, global 80 dB floor)."""
    frame_length = max(1024, int(round(sample_rate * meter_block_size)))
    if frame_length % 2:
        frame_length += 1
    hop = max(1, frame_length // 2)
    r = lr.rms(samples, frame_length=frame_length, hop_length=hop)[0]
    return lr.amplitude_to_db(r + 1e-9, ref=1.0)


def measure_loudness(samples, sample_rate, meter_block_size=0.4):
    """analysis/loudness.py:45-78 with pyloudnorm 0.1.1 (no Meter.loudness_range)."""
    x = np.asarray(samples, dtype=np.float32)
    if x.ndim != 1:
        raise ValueError("measure_loudness expects mono audio samples")
    short_term = windowed_loudness(x, sample_rate, 3.0)
    momentary = windowed_loudness(x, sample_rate, meter_block_size)
    integrated = pl.integrated_loudness(x, sample_rate, meter_block_size)
    lra = float(np.percentile(momentary, 95) - np.percentile(momentary, 5))
    return (integrated, np.asarray(short_term, dtype=float).tolist(),
            np.asarray(momentary, dtype=float).tolist(), lra)


def true_peak_dbtp(samples, sample_rate, oversample=8):
    """analysis/loudness.py:81-97 (section 8f rank 2)."""
    import scipy.signal

    x = np.asarray(samples, dtype=np.float32)
    up = x if oversample == 1 else scipy.signal.resample_poly(x, oversample, 1)
    return float(20.0 * np.log10(float(np.max(np.abs(up))) + 1e-12))


def rms_dbfs(samples):
    """analysis/loudness.py:118-119."""
    x = np.asarray(samples).astype(np.float32)
    return float(20.0 * np.log10(float(np.sqrt(np.mean(x**2))) + 1e-12))


# ---- tempo.py ---------------------------------------------------------------


def onset_envelope(y, sr, hop_length=512):
    """tempo.py:16-24."""
    env = lr.onset_strength(y=np.asarray(y), sr=sr, hop_length=hop_length, aggregate=np.mean)
    return np.zeros(1, dtype=float) if env.size == 0 else env


def onset_autocorrelation(env):
    """tempo.py:38."""
    return lr.autocorrelate(env)


# ---- analysis/structure.py ---------------------------------------------------


def structure_frontend(samples, sample_rate, frame_length=2048, hop_length=512):
    """analysis/structure.py:48-59 and :190-196 -> (magnitude, mel, log_mel, flux)."""
    x = np.asarray(samples)
    magnitude = np.abs(lr.stft(x, n_fft=frame_length, hop_length=hop_length))
    mel = lr.melspectrogram(x, sample_rate, n_fft=frame_length, hop_length=hop_length, power=2.0)
    mel64 = np.asarray(mel, dtype=float)
    log_mel = lr.power_to_db(mel64 + 1e-9)
    flux = np.asarray(lr.onset_strength(S=mel64, sr=sample_rate, hop_length=hop_length), dtype=float)
    return magnitude, mel, log_mel, flux


# ---- harmony.py ---------------------------------------------------------------


def spectral_balance(samples, sample_rate):
    """harmony.py:253-267 -> (low, mid, high)."""
    spec = np.abs(lr.stft(np.asarray(samples), n_fft=4096, hop_length=1024))
    freqs = lr.fft_frequencies(sample_rate, 4096)
    total = np.sum(spec)
    if total <= 0:
        return 0.0, 0.0, 0.0
    lo = freqs < 200
    mid = (freqs >= 200) & (freqs < 2000)
    hi = freqs >= 2000
    return (float(np.sum(spec[lo]) / total), float(np.sum(spec[mid]) / total),
            float(np.sum(spec[hi]) / total))


def chroma_stft(samples, sample_rate, return_tuning=False):
    """harmony.py:108,149."""
    return lr.chroma_stft(np.asarray(samples), sample_rate, return_tuning=return_tuning)


# ---- report.py ------------------------------------------------------------------


def tempogram(samples, sample_rate, hop_length=512):
    """report.py:254-262 (samples promoted to float64 by the reference first)."""
    x = np.asarray(samples, dtype=float)
    if x.ndim > 1:
        x = np.mean(x, axis=0)
    return np.asarray(lr.tempogram(y=x, sr=sample_rate, hop_length=hop_length), dtype=float)
