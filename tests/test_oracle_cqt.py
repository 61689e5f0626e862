"""The constant-Q oracle (oracle/cqt_np.py) against independent evidence (CPU only).

The restatement of librosa's recursive CQT cannot be pinned against librosa itself (not installable here), so it is
checked against (i) a brute-force evaluation from the definition -- direct float64 correlation of the FULL-RATE signal with
every constant-Q atom, no decimation / FFT / sparsification; agreement is bounded by the 1 % basis sparsification --
(ii) the stated decimator's own specification, and (iii) the reference's harmony test property (C major for the
C-F-G-C triad progression, /root/reference/tests/test_harmony.py:37-64)."""

import numpy as np
import scipy.signal

from oracle import cqt_np as cq

from . import signals


def test_decimator_meets_the_published_soxr_hq_band_edges():
    h = cq.decimator_taps()
    assert len(h) == 381 and np.allclose(h, h[::-1]) and abs(h.sum() - 1.0) < 1e-12
    w, H = scipy.signal.freqz(h, worN=1 << 15)
    f = w / np.pi * 2.0                                  # in units of the output Nyquist
    mag = np.abs(H)
    assert np.max(np.abs(mag[f <= cq.SOXR_HQ_PASSBAND_END] - 1.0)) < 1e-6      # pass band flat to 1e-6
    assert 20 * np.log10(np.max(mag[f >= cq.SOXR_HQ_STOPBAND_BEGIN])) < -125.0  # 20-bit rejection


def test_decimate2_length_scale_and_alignment():
    sr = 8000
    t = np.arange(4001) / sr
    y = np.sin(2 * np.pi * 200.0 * t).astype(np.float32)
    out = cq.decimate2(y)
    assert out.shape == (2001,) and out.dtype == np.float32
    ref = np.sqrt(2.0) * np.sin(2 * np.pi * 200.0 * t[::2])   # zero-phase: sample m sits at input sample 2m; scale=True
    np.testing.assert_allclose(out[300:-300], ref[300:-300], atol=2e-6)


def test_cqt_plan_matches_librosa_recursion_for_44k1_and_22k05():
    p = cq.cqt_plan(44_100)
    assert p["early"] == 1 and [o["hop"] for o in p["octaves"]] == [256, 128, 64, 32, 16, 8, 4]
    assert [o["stage"] for o in p["octaves"]] == [1, 2, 3, 4, 5, 6, 7]
    p = cq.cqt_plan(22_050)
    assert p["early"] == 0 and [o["hop"] for o in p["octaves"]] == [512, 256, 128, 64, 32, 16, 8]
    fb, n_fft, lengths = cq.vqt_filter_fft(p["octaves"][0]["sr"], p["octaves"][0]["freqs"], p["alpha"])
    assert n_fft == 1024 and fb.shape == (36, 513)
    nnz = (fb != 0).sum(axis=1)
    assert nnz.min() >= 8 and nnz.max() <= 32            # ~1 % of the L1 mass dropped: a dozen bins around the centre


def test_cqt_matches_brute_force_definition():
    sr = 44_100
    rng = np.random.default_rng(3)
    n = int(sr * 2.5)
    t = np.arange(n) / sr
    y = sum(a * np.sin(2 * np.pi * f * t) for a, f in ((0.3, 261.63), (0.2, 329.63), (0.2, 392.0), (0.1, 1046.5), (0.1, 65.41)))
    y = (y + 0.005 * rng.standard_normal(n)).astype(np.float32)
    chroma, C, tuning = cq.chroma_cqt(y, sr, return_parts=True)
    frames = np.arange(30, C.shape[1] - 30, 23)
    B = cq.brute_force_cqt(y, sr, tuning=tuning, frames=frames)
    A = C[:, frames]
    assert np.max(np.abs(A - B)) <= 0.012 * np.max(B)    # bounded by the 1 % sparsification of the spectral basis
    big = B > 0.1 * B.max()
    assert np.median(np.abs(A - B)[big] / B[big]) < 3e-3
    # and the chroma built from either agrees on the pitch-class ranking (C, E, G on top)
    fold = cq.cq_to_chroma(252)
    top_a = set(np.argsort(np.mean(fold @ A, axis=1))[-3:])
    top_b = set(np.argsort(np.mean(fold @ B, axis=1))[-3:])
    assert top_a == top_b == {0, 4, 7}
    assert chroma.shape[0] == 12 and np.all(chroma <= 1.0 + 1e-6) and np.all(chroma >= 0.0)


def test_cq_to_chroma_layout():
    m = cq.cq_to_chroma(252)
    assert m.shape == (12, 252) and np.all(m.sum(axis=0) == 1.0)
    # pitch class c collects bins 3c-1, 3c, 3c+1 of every octave (36 bins per octave, fmin = C1)
    for c in range(12):
        cols = np.nonzero(m[c, :36])[0]
        assert sorted(cols) == sorted({(3 * c - 1) % 36, 3 * c, 3 * c + 1})


def test_detuned_tone_tuning_estimate():
    sr = 22_050
    t = np.arange(int(sr * 2.0)) / sr
    f = 440.0 * 2.0 ** (0.25 / 12.0)                      # a quarter of a semitone sharp = 0.75 of a 36-per-octave bin
    y = (0.5 * np.sin(2 * np.pi * f * t)).astype(np.float32)
    tun = cq.estimate_tuning_y(y, sr, 36)
    assert abs(((tun - 0.75 + 0.5) % 1.0) - 0.5) <= 0.05  # residual modulo one bin (parabolic peak interpolation is good to a few cents)


def test_reference_harmony_property_c_major():
    x, sr = signals.triad_progression()
    chroma = cq.chroma_cqt(x, sr)
    mean = chroma.mean(axis=1)
    assert int(np.argmax(mean)) in (0, 7)                # C or G dominate the C-F-G-C progression
    assert chroma.shape[1] in (1 + len(x) // 512, 2 + len(x) // 512)
