"""Independent cross-checks of the oracle against implementations available offline
(torchaudio, scipy, direct O(n^2) sums) and against the committed golden fixtures."""

import os

import numpy as np
import pytest
import scipy.signal

from oracle import frontend as fe
from oracle import librosa_np as lr
from oracle import pyloudnorm_np as pl

from . import signals

GOLDEN = os.path.join(os.path.dirname(__file__), "golden")


def test_stft_matches_direct_dft():
    rng = np.random.default_rng(7)
    y = rng.normal(size=3000).astype(np.float32)
    D = lr.stft(y, n_fft=256, hop_length=64)
    assert D.shape == (129, 1 + 3000 // 64) and D.dtype == np.complex64
    w = scipy.signal.get_window("hann", 256, fftbins=True)
    ypad = np.pad(y.astype(np.float64), 128)
    n = np.arange(256)
    for t in (0, 5, 46):
        fr = ypad[t * 64: t * 64 + 256] * w
        X = np.array([np.sum(fr * np.exp(-2j * np.pi * k * n / 256)) for k in range(129)])
        np.testing.assert_allclose(D[:, t], X.astype(np.complex64), rtol=1e-5, atol=1e-5)


def test_mel_filterbank_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    for sr, n_fft, n_mels in ((44_100, 2048, 128), (48_000, 2048, 128), (44_100, 4096, 256)):
        ours = lr.filters_mel(sr, n_fft, n_mels=n_mels)
        theirs = ta.functional.melscale_fbanks(n_fft // 2 + 1, 0.0, sr / 2, n_mels, sr, norm="slaney",
                                               mel_scale="slaney").T.numpy()
        assert np.abs(ours - theirs).max() < 5e-7
    assert int((lr.filters_mel(44_100, 2048) != 0).sum()) == 2014  # SURVEY A.2 probe


def test_loudness_matches_torchaudio():
    ta = pytest.importorskip("torchaudio")
    import torch

    x = signals.minus18_sine(48_000, seconds=3.0)
    ours = pl.integrated_loudness(x, 48_000)
    theirs = ta.functional.loudness(torch.from_numpy(x)[None], 48_000).item()
    assert abs(ours - theirs) < 0.01


def test_k_weighting_coefficients_probe_values():  # SURVEY A.9 probe
    (b1, a1), (b2, a2) = pl.k_weighting_coefficients(44_100)
    np.testing.assert_allclose(b1, [1.53090959, -2.65116903, 1.16916686], atol=1e-8)
    np.testing.assert_allclose(a1, [1.0, -1.66375011, 0.71265753], atol=1e-8)
    np.testing.assert_allclose(b2, [0.99460781, -1.98921562, 0.99460781], atol=1e-8)
    np.testing.assert_allclose(a2, [1.0, -1.98920104, 0.9892302], atol=1e-8)


def test_block_bounds_are_multiples_of_quarter_block():
    for sr, n in ((44_100, 7_938_000), (48_000, 172_800_000 // 16), (22_050, 22_050 * 32)):
        lo, hi = pl.block_bounds(n, sr)
        step = int(round(0.1 * sr))
        assert np.all(lo % step == 0) and np.all(hi % step == 0)
    assert len(pl.block_bounds(7_938_000, 44_100)[0]) == 1797
    assert len(pl.block_bounds(48_000, 48_000)[0]) == 7


def test_autocorrelate_matches_direct_sum():
    rng = np.random.default_rng(3)
    x = rng.random(175).astype(np.float32)
    ac = lr.autocorrelate(x)
    direct = np.array([np.dot(x[: x.size - k].astype(float), x[k:].astype(float)) for k in range(x.size)])
    np.testing.assert_allclose(ac, direct, rtol=1e-10, atol=1e-10)
    assert ac.dtype == np.float64


def test_onset_envelope_shape_and_padding():
    y, sr, _ = signals.noisy_click_track(bars=2)
    env = fe.onset_envelope(y, sr)
    assert env.dtype == np.float32 and env.shape == (1 + y.size // 512,)
    assert np.all(env[:3] == 0)


def test_golden_fixtures_match_oracle():
    """tests/golden/*.npz were produced by tests/golden/make_golden.py from this oracle; a drift
    in the oracle (or in numpy/scipy) shows up here."""
    path = os.path.join(GOLDEN, "tiny_click.npz")
    g = np.load(path)
    x = g["samples"]
    np.testing.assert_array_equal(x, signals.tiny_click())
    sr = 44_100
    env = fe.onset_envelope(x, sr)
    np.testing.assert_allclose(env, g["onset_env"], rtol=1e-6, atol=1e-7)
    np.testing.assert_allclose(fe.onset_autocorrelation(env), g["autocorr"], rtol=1e-9, atol=1e-9)
    f, m = fe.compute_ltas(x, sr)
    np.testing.assert_allclose(m, g["ltas"], rtol=1e-6, atol=1e-8)
    np.testing.assert_allclose(fe.spectral_centroid_series(x, sr), g["centroid"], rtol=1e-9)
    np.testing.assert_array_equal(fe.spectral_rolloff_series(x, sr), g["rolloff"])
    assert fe.measure_loudness(x, sr)[0] == pytest.approx(float(g["lufs"]), abs=1e-9)
