"""The boundary beyond the default configuration: every power-of-two n_fft from 256 to 4096, any hop,
other analysis windows, other true-peak oversampling factors, and the host-buffer C-ABI entry point.  Run with -m gpu."""

import os

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

from oracle import frontend as ofe  # noqa: E402
from oracle import librosa_np as olr  # noqa: E402
from track_analyser_b200 import engine, features, synth  # noqa: E402
from track_analyser_b200.analysis import loudness  # noqa: E402

RTOL, ATOL = 1e-4, 1e-6


def _magnitude_close(got, ref, n_fft):
    frame_norm = np.sqrt(np.sum(ref.astype(np.float64) ** 2, axis=0, keepdims=True) * 2 / n_fft)
    assert got.shape == ref.shape
    assert np.all(np.abs(got.astype(np.float64) - ref) <= 1e-6 + 1e-4 * ref + 1.5e-6 * frame_norm)


@pytest.mark.parametrize("n_fft,hop,mels,channels", [(256, 64, 40, 1), (256, 100, 40, 2), (512, 128, 64, 2), (512, 200, 64, 1),
                                                     (1024, 300, 64, 2), (2048, 441, 128, 2), (2048, 441, 128, 1),
                                                     (4096, 1000, 128, 1), (2048, 2, 128, 1)])
def test_every_power_of_two_n_fft_and_any_hop(n_fft, hop, mels, channels):
    sr = 22_050
    seconds = 0.3 if hop == 2 else 2.1
    x = synth.synth_track(11 + n_fft + hop, seconds, sr, channels)
    mono = np.mean(x, axis=0) if channels == 2 else x
    plan = engine.Plan(sr, n_fft, hop, mels, device=0)
    outs = ("magnitude", "mel", "ltas", "centroid", "rolloff_bin", "onset_env", "band_energy", "chroma", "tuning")
    r = engine.analyse_batch(plan, [x], outs)[0]
    mag = np.abs(olr.stft(mono, n_fft=n_fft, hop_length=hop))
    _magnitude_close(r["magnitude"], mag, n_fft)
    mel = np.einsum("ft,mf->mt", mag**2, olr.filters_mel(sr, n_fft, n_mels=mels), optimize=True)
    np.testing.assert_allclose(r["mel"], mel, rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(r["ltas"], np.mean(mag, axis=1), rtol=RTOL, atol=ATOL)
    np.testing.assert_allclose(r["centroid"], olr.spectral_centroid(mono, sr, n_fft, hop)[0], rtol=RTOL, atol=ATOL)
    freqs = np.fft.rfftfreq(n_fft, 1.0 / sr)
    assert np.mean(freqs[r["rolloff_bin"]] == olr.spectral_rolloff(mono, sr, n_fft, hop)[0]) >= 0.999
    # librosa.onset.onset_strength pads lag + 2048 // (2 hop) frames whatever the plan's n_fft is
    env = olr.onset_strength(S=olr.power_to_db(mel), sr=sr, hop_length=hop)
    np.testing.assert_allclose(r["onset_env"], env[: r["onset_env"].shape[0]], rtol=RTOL, atol=ATOL)
    ref_chroma, tuning = olr.chroma_stft(mono, sr, n_fft=n_fft, hop_length=hop, return_tuning=True)
    assert r["tuning"] == pytest.approx(tuning, abs=1e-12)
    np.testing.assert_allclose(r["chroma"], ref_chroma, rtol=RTOL, atol=ATOL)
    if channels == 2:
        from track_analyser_b200 import stereo as pstereo

        w = pstereo.width_from_band_energy(r["band_energy"], freqs, r.n_frames, None, sr)
        ref = ofe.frequency_dependent_width(x, sr, n_fft=n_fft, hop_length=hop)
        for k in ("low", "mid", "high"):
            assert w[k] == pytest.approx(ref[k], rel=RTOL, abs=ATOL)


@pytest.mark.parametrize("window", ["hamming", "blackman", "boxcar"])
def test_compute_ltas_with_other_windows(window):
    """features.compute_ltas(window=...) (features.py:66-82 passes it to librosa.stft)."""
    sr = 22_050
    x = synth.synth_track(5, 1.5, sr, 1)
    got = features.compute_ltas(x, sr, n_fft=1024, hop_length=256, window=window)
    ref = np.mean(np.abs(olr.stft(x, n_fft=1024, hop_length=256, window=window)), axis=1)
    np.testing.assert_allclose(got.magnitude, ref, rtol=RTOL, atol=ATOL)
    np.testing.assert_array_equal(got.frequencies, np.fft.rfftfreq(1024, 1.0 / sr))


@pytest.mark.parametrize("oversample", [1, 2, 3, 4, 16, 32])
def test_true_peak_other_oversampling_factors(oversample):
    """loudness.true_peak_dbtp(oversample=...) (analysis/loudness.py:81-97): scipy.signal.resample_poly(x, oversample, 1)."""
    sr = 44_100
    rng = np.random.default_rng(oversample)
    t = np.arange(int(0.7 * sr)) / sr
    x = (0.5 * np.sin(2 * np.pi * 10_000.5 * t + 0.3) + 0.2 * rng.standard_normal(t.size)).astype(np.float32)
    assert loudness.true_peak_dbtp(x, sr, oversample=oversample) == pytest.approx(ofe.true_peak_dbtp(x, sr, oversample), abs=1e-4)
    with pytest.raises(ValueError):
        loudness.true_peak_dbtp(x, sr, oversample=0)


def test_host_buffer_entry_point_equals_the_device_path():
    """ta_frontend_run_host: numpy in, numpy out, no torch tensor; same numbers as the torch-owned path."""
    sr = 44_100
    tracks = [synth.synth_track(60 + i, 1.0 + 0.83 * i, sr, 2) for i in range(3)]
    plan = engine.Plan(sr, 2048, 512, 128, device=0)
    outs = ("magnitude", "mel", "onset_env", "autocorr", "flux_linear", "ltas", "centroid", "rolloff_bin", "band_energy", "moments",
            "kw_blocks", "lufs", "rms_momentary", "rms_short", "chroma", "tuning", "tempogram", "true_peak", "hpss_harmonic",
            "hpss_percussive", "mfcc", "chroma_cqt", "cqt_tuning")
    a = engine.analyse_batch(plan, tracks, outs)
    b = engine.analyse_host(plan, tracks, outs)
    for ra, rb in zip(a, b):
        for k in outs:
            if k in ("ltas", "band_energy", "moments", "kw_blocks", "lufs", "rms_momentary", "rms_short"):
                # sums gathered with float64 atomics: the order of the additions varies from run to run
                np.testing.assert_allclose(rb[k], ra[k], rtol=1e-6 if k in ("ltas", "band_energy") else 1e-11, atol=1e-9, err_msg=k)
            else:
                np.testing.assert_array_equal(rb[k], ra[k], err_msg=k)
    # a subset that needs intermediates the caller did not ask for (mel for the envelope, magnitude for the chroma)
    c = engine.analyse_host(plan, tracks[:1], ("onset_env", "chroma", "lufs"))[0]
    np.testing.assert_array_equal(c["onset_env"], a[0]["onset_env"])
    np.testing.assert_array_equal(c["chroma"], a[0]["chroma"])
    assert c["lufs"] == pytest.approx(a[0]["lufs"], abs=1e-9)


def test_the_ctypes_stub_printed_in_integration_md_runs_as_printed():
    """the numpy-only stub over ta_frontend_run_host, executed verbatim (library path aside)."""
    import re

    root = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
    text = open(os.path.join(root, "INTEGRATION.md")).read()
    block = re.search(r"```python\n(# src/track_analyser/_b200\.py.*?)```", text, re.S).group(1)
    block = block.replace('C.CDLL("libta_b200.so")', "C.CDLL(%r)" % os.path.join(root, "track_analyser_b200", "libta_b200.so"))
    ns: dict = {}
    exec(compile(block, "INTEGRATION.md", "exec"), ns)
    sr = 22_050
    y = synth.synth_track(77, 4.0, sr, 1).reshape(-1)
    mag, env = ns["stft_magnitude_and_onsets"](y, sr)
    ref = np.abs(olr.stft(y, n_fft=2048, hop_length=512))
    _magnitude_close(mag, ref, 2048)
    mel = np.einsum("ft,mf->mt", ref**2, olr.filters_mel(sr, 2048, n_mels=128), optimize=True)
    np.testing.assert_allclose(env, olr.onset_strength(S=olr.power_to_db(mel), sr=sr, hop_length=512), rtol=RTOL, atol=ATOL)
