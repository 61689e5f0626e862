"""Host-side pieces of the install() shims (track_analyser_b200/compat.py) against the oracle's restatements of the same
librosa functions: everything here runs without a GPU.  The device-backed shims are covered by tests/test_gpu_install.py."""

import numpy as np
import pytest

from oracle import librosa_np as olr
from track_analyser_b200 import compat, hostlogic


def test_db_conversions_match_the_oracle():
    rng = np.random.default_rng(3)
    S = rng.random((40, 50)) ** 4 * 100.0
    np.testing.assert_array_equal(compat.power_to_db(S), olr.power_to_db(S))
    np.testing.assert_array_equal(compat.power_to_db(S, ref=np.max), olr.power_to_db(S, ref=float(np.max(S))))   # callable ref
    np.testing.assert_array_equal(compat.power_to_db(S, top_db=None), olr.power_to_db(S, top_db=None))
    # amplitude_to_db(S, ref) == power_to_db(S**2, ref**2) with amin**2 (analysis/loudness.py:42)
    A = np.sqrt(S)
    np.testing.assert_allclose(compat.amplitude_to_db(A, ref=np.max), olr.power_to_db(A**2, ref=float(np.max(A)) ** 2, amin=1e-10),
                               rtol=0, atol=1e-12)


def test_normalize_is_librosa_util_normalize():
    x = np.array([0.5, -2.0, 1.0])
    np.testing.assert_array_equal(compat.normalize(x), x / 2.0)
    np.testing.assert_array_equal(compat.normalize(np.zeros(4)), np.zeros(4))         # below tiny: left alone
    np.testing.assert_allclose(compat.normalize(x, norm=1), x / 3.5)
    np.testing.assert_allclose(compat.normalize(x, norm=2), x / np.sqrt(5.25))
    m = np.array([[1.0, -4.0], [2.0, 2.0]])
    np.testing.assert_array_equal(compat.normalize(m, axis=0), m / np.array([[2.0, 4.0]]))
    with pytest.raises(ValueError):
        compat.normalize(x, norm=3)


def test_pitch_and_time_helpers():
    assert compat.midi_to_hz(69) == 440.0
    np.testing.assert_allclose(compat.midi_to_hz([60, 81]), [261.6255653005986, 880.0])
    np.testing.assert_allclose(compat.hz_to_midi(compat.midi_to_hz(np.arange(20, 100))), np.arange(20, 100), atol=1e-9)
    np.testing.assert_array_equal(compat.fft_frequencies(sr=44_100, n_fft=2048), np.fft.rfftfreq(2048, 1 / 44_100))
    np.testing.assert_array_equal(compat.frames_to_time([0, 1, 10], sr=22_050, hop_length=512), np.array([0, 512, 5120]) / 22_050)
    np.testing.assert_array_equal(compat.time_to_frames([0.0, 1.0], sr=22_050, hop_length=512), [0, 43])
    tf = compat.tempo_frequencies(5, hop_length=512, sr=22_050)
    assert np.isinf(tf[0])
    np.testing.assert_allclose(tf[1:], 60.0 * 22_050 / (512 * np.arange(1, 5)))


def test_onset_strength_and_autocorrelation_fallbacks_match_the_oracle():
    rng = np.random.default_rng(5)
    mel = rng.random((128, 300)) ** 3
    np.testing.assert_allclose(compat._onset_strength_from_S(mel), olr.onset_strength(S=mel, sr=22_050, hop_length=512), rtol=1e-12, atol=0)
    env = rng.random(1000).astype(np.float32)
    np.testing.assert_allclose(compat._autocorrelate_host(env), olr.autocorrelate(env), rtol=1e-10, atol=1e-12)
    direct = np.array([np.dot(env[: 1000 - k].astype(np.float64), env[k:].astype(np.float64)) for k in range(6)])
    np.testing.assert_allclose(compat._autocorrelate_host(env)[:6], direct, rtol=1e-10)


def test_peak_pick_shim_is_the_host_logic_one():
    x = np.array([0, 1, 0, 3, 0, 2, 0, 5, 0], dtype=float)
    np.testing.assert_array_equal(compat.peak_pick(x, pre_max=1, post_max=1, pre_avg=1, post_avg=1, delta=0.1, wait=0),
                                  hostlogic.peak_pick(x, 1, 1, 1, 1, 0.1, 0))
