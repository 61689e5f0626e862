"""The reference's own test assertions for the hot path, run against the CPU oracle.

Each test cites the reference test it restates (/root/reference/tests/...).  These
property-level checks are all the reference pins for this path.
"""

import numpy as np
import pytest

from oracle import frontend as fe
from track_analyser_b200 import hostlogic
from track_analyser_b200 import tempo as ptempo

from . import signals


def test_ltas_peak_at_440():  # test_features.py:15-24
    f, m = fe.compute_ltas(signals.sine(440.0), 22_050)
    assert float(f[np.argmax(m)]) == pytest.approx(440.0, abs=5.0)


def test_centroid_of_1k_sine():  # test_features.py:27-34
    c = fe.spectral_centroid_series(signals.sine(1000.0), 22_050)
    assert float(np.mean(c)) == pytest.approx(1000.0, abs=20.0)


def test_rolloff_of_white_noise():  # test_features.py:37-44
    noise = np.random.default_rng(1337).normal(size=22_050).astype(np.float32)
    assert np.all(fe.spectral_rolloff_series(noise, 22_050) > 5_000.0)


def test_mono_has_no_side():  # test_stereo.py:15-27
    mono = signals.sine(440.0)
    st = fe.ensure_stereo(mono, None)
    assert fe.mid_side_rms(st)[1] == pytest.approx(0.0, abs=1e-6)
    assert fe.mono_compatibility_correlation(st) == pytest.approx(1.0, abs=1e-6)
    w = fe.frequency_dependent_width(st, 22_050)
    assert all(v == pytest.approx(0.0, abs=1e-6) for v in w.values())


def test_mid_side_imbalanced():  # test_stereo.py:30-39
    left = signals.sine(440.0)
    m, s = fe.mid_side_rms(np.vstack([left, 0.5 * left]))
    assert m > s > 0.0


def test_width_phase_difference():  # test_stereo.py:42-54
    st = np.vstack([signals.sine(440.0), signals.sine(440.0, phase=np.pi / 2)])
    w = fe.frequency_dependent_width(st, 22_050)
    assert min(w.values()) >= 0.0 and max(w.values()) > 0.0


def test_constant_channels_correlation():  # test_stereo.py:57-64
    st = np.vstack([np.ones(10, np.float32), np.ones(10, np.float32)])
    assert fe.mono_compatibility_correlation(st) == pytest.approx(1.0)


def test_minus18_lufs():  # test_loudness.py:33-43
    integrated, short_term, momentary, _ = fe.measure_loudness(signals.minus18_sine(48_000), 48_000)
    assert integrated == pytest.approx(-18.0, abs=0.3)
    assert integrated == pytest.approx(-18.0347, abs=1e-3)  # SURVEY A.9 cross-check value
    assert short_term and momentary


def test_true_peak():  # test_loudness.py:46-55
    x = signals.minus18_sine(44_100)
    expected = 20.0 * np.log10(float(np.max(np.abs(x))))
    assert fe.true_peak_dbtp(x, 44_100) == pytest.approx(expected, abs=0.2)


def test_click_track_tempo_and_grid():  # test_tempo.py:39-53 (host logic from the product, envelope from the oracle)
    y, sr, expected = signals.noisy_click_track()  # 64 bars like the reference
    env = fe.onset_envelope(y, sr)
    ac = fe.onset_autocorrelation(env)
    bpm = ptempo._bpm_from_autocorr(env, ac, sr, 90.0, 135.0, 512)
    assert abs(bpm - 120.0) <= 0.1
    fit = ptempo._fit_onset_regression(env, sr, 512, 60.0 / bpm)
    assert fit is not None
    times = max(fit[0], 0.0) + np.arange(expected.size) * (60.0 / bpm)
    assert float(np.max(np.abs(times - expected))) <= 0.005
    frames = hostlogic.onset_detect(env, sr, 512, backtrack=True)
    assert frames.size >= 250
