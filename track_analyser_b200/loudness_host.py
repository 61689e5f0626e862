"""Host-side finishing of the loudness outputs (a few hundred values per track).

The pass over the PCM -- K-weighting scan, gating blocks, gating, frame mean
squares -- runs in csrc/timedomain.cu.  What remains is the reference's dB
conversion of the frame RMS values (analysis/loudness.py:39-42 ->
librosa.amplitude_to_db with its global 80 dB floor), done in float32 like numpy.
"""

from __future__ import annotations

import numpy as np


def frames_to_db(mean_square: np.ndarray) -> np.ndarray:
    rms = np.sqrt(np.asarray(mean_square, dtype=np.float64)).astype(np.float32)
    mag = np.abs(rms + np.float32(1e-9))
    power = np.square(mag)
    amin = 1e-5**2
    db = 10.0 * np.log10(np.maximum(np.float32(amin), power))
    db = (db - np.float32(10.0 * np.log10(max(amin, 1.0)))).astype(np.float32)
    return np.maximum(db, db.max() - np.float32(80.0)) if db.size else db


def rms_dbfs_from_moments(moments: np.ndarray, stereo_run: bool) -> float:
    """20 log10(sqrt(mean(mono^2)) + 1e-12) (analysis/loudness.py:118-119) from the time-domain moments: sum of mono^2 is
    moments[2] of a mono run and moments[5] (mid^2) when the stereo run served the request; moments[7] is the count."""
    sq = moments[5] if stereo_run else moments[2]
    rms_val = float(np.sqrt(sq / moments[7])) if moments[7] else 0.0
    return float(20.0 * np.log10(rms_val + 1e-12))
