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
    # rows added with the widened scope
    chroma, tuning = fe.chroma_stft(x, sr, return_tuning=True)
    np.testing.assert_allclose(chroma, g["chroma"], rtol=1e-6, atol=1e-7)
    assert tuning == pytest.approx(float(g["tuning"]), abs=1e-12)
    np.testing.assert_allclose(lr.tempogram(onset_envelope=env, sr=sr, hop_length=512)[:, ::4], g["tempogram_cols"],
                               rtol=1e-5, atol=1e-6)
    mag, mel, _, _ = fe.structure_frontend(x, sr)
    harm, perc = lr.hpss(mag)
    np.testing.assert_allclose(np.sum(harm, axis=0, dtype=np.float64), g["hpss_harmonic"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(np.sum(perc, axis=0, dtype=np.float64), g["hpss_percussive"], rtol=1e-6, atol=1e-9)
    np.testing.assert_allclose(lr.mfcc(lr.power_to_db(np.asarray(mel, dtype=float) + 1e-9)), g["mfcc"], rtol=1e-9, atol=1e-9)
    assert fe.true_peak_dbtp(x, sr) == pytest.approx(float(g["true_peak_db"]), abs=1e-9)


def test_chroma_mel_db_and_spectrogram_match_transformers_port():
    """transformers.audio_utils carries its own port of librosa's filters.chroma / filters.mel / power_to_db / stft
    framing: an implementation independent of oracle/librosa_np.py with the same specification."""
    au = pytest.importorskip("transformers.audio_utils")
    sr, n_fft = 44_100, 2048
    for tuning in (0.0, 0.23, -0.41):  # the tuning estimate shifts the filterbank (harmony.py:108 -> estimate_tuning)
        theirs = au.chroma_filter_bank(num_frequency_bins=n_fft, num_chroma=12, sampling_rate=sr, tuning=tuning)
        np.testing.assert_allclose(lr.filters_chroma(sr, n_fft, tuning=tuning), theirs, rtol=0, atol=1e-7)
    mel = au.mel_filter_bank(num_frequency_bins=n_fft // 2 + 1, num_mel_filters=128, min_frequency=0.0,
                             max_frequency=sr / 2, sampling_rate=sr, norm="slaney", mel_scale="slaney")
    np.testing.assert_allclose(lr.filters_mel(sr, n_fft, 128), mel.T, rtol=0, atol=1e-8)
    x = np.abs(np.random.default_rng(0).standard_normal((40, 30))) ** 2
    np.testing.assert_array_equal(lr.power_to_db(x), au.power_to_db(x, reference=1.0, min_value=1e-10, db_range=80.0))
    y = np.random.default_rng(1).standard_normal(20_000).astype(np.float32)
    win = au.window_function(n_fft, "hann")
    np.testing.assert_allclose(lr.get_window("hann", n_fft), win, rtol=0, atol=1e-15)
    # librosa 0.10 centres with zero padding (pad_mode="constant")
    spec = au.spectrogram(y, win, frame_length=n_fft, hop_length=512, fft_length=n_fft, power=2.0, center=True,
                          pad_mode="constant")
    ours = lr.spectrogram(y, n_fft, 512, 2.0)
    assert spec.shape == ours.shape and np.max(np.abs(spec - ours)) <= 1e-6 * ours.max()
    melspec = au.spectrogram(y, win, frame_length=n_fft, hop_length=512, fft_length=n_fft, power=2.0, center=True,
                             pad_mode="constant", mel_filters=mel)
    ours = lr.melspectrogram(y, sr, n_fft=n_fft, hop_length=512, n_mels=128)
    assert np.max(np.abs(melspec - ours)) <= 1e-6 * ours.max()


def test_mfcc_and_resampler_match_torchaudio():
    """MFCC: torchaudio's orthonormal DCT-II matrix.  Resampler: torchaudio's sinc_interp_kaiser with the
    kaiser_best design (64 zero crossings, roll-off 0.9476, beta 14.7697 -- torchaudio's own default beta is
    resampy's) is a different realisation of the same filter (exact kernel per phase instead of a 512-per-crossing
    table with linear interpolation), so the two agree to ~1e-4, not bit for bit."""
    ta = pytest.importorskip("torchaudio")
    import torch

    from oracle import resampy_np

    rng = np.random.default_rng(0)
    S = rng.standard_normal((128, 50))
    D = ta.functional.create_dct(13, 128, norm="ortho").numpy().T
    np.testing.assert_allclose(lr.mfcc(S), D @ S, rtol=0, atol=2e-5)  # torchaudio's matrix is float32
    for sr0, sr1 in ((22_050, 44_100), (48_000, 44_100), (44_100, 48_000)):
        n = sr0 // 2
        t = np.arange(n) / sr0
        x = (0.4 * np.sin(2 * np.pi * 997 * t) + 0.2 * np.sin(2 * np.pi * 5000 * t) + 0.01 * rng.standard_normal(n)).astype(np.float32)
        ours = resampy_np.resample(x, sr0, sr1)
        theirs = ta.functional.resample(torch.from_numpy(x), sr0, sr1, lowpass_filter_width=64,
                                        rolloff=resampy_np.KAISER_BEST["rolloff"], resampling_method="sinc_interp_kaiser",
                                        beta=resampy_np.KAISER_BEST["beta"]).numpy()
        assert len(ours) == len(theirs)
        assert np.max(np.abs(ours - theirs)) < 5e-4


def test_tempogram_matches_direct_windowed_autocorrelation():
    """librosa.feature.tempogram restated (oracle) against the definition: pad by win/2 with a linear ramp, frame,
    Hann, direct O(win^2) autocorrelation per frame, divide by the frame's largest magnitude."""
    rng = np.random.default_rng(2)
    env = np.abs(rng.standard_normal(90)).astype(np.float32)
    win = 16
    got = lr.tempogram(onset_envelope=env, sr=22_050, hop_length=512, win_length=win)
    pad = np.pad(env, (win // 2, win // 2), mode="linear_ramp", end_values=[0, 0]).astype(np.float64)
    w = scipy.signal.get_window("hann", win, fftbins=True)
    want = np.zeros((win, len(env)))
    for t in range(len(env)):
        z = pad[t:t + win] * w
        ac = np.array([np.dot(z[: win - l], z[l:]) for l in range(win)])
        want[:, t] = ac / max(np.max(np.abs(ac)), np.finfo(np.float64).tiny)
    np.testing.assert_allclose(got, want, rtol=1e-6, atol=1e-6)
