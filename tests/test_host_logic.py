"""Host-side glue of the product: peak picking, stereo sums, dB conversion, sharding, coercion."""

import numpy as np
import pytest

from oracle import frontend as fe
from oracle import librosa_np as lr
from track_analyser_b200 import hostlogic, loudness_host, sharding, stereo, utils

from . import signals


def brute_peak_pick(x, pre_max, post_max, pre_avg, post_avg, delta, wait):
    peaks, last = [], -10**9
    for n in range(len(x)):
        w = x[max(0, n - pre_max): n + post_max]
        a = x[max(0, n - pre_avg): n + post_avg]
        if x[n] == w.max() and x[n] >= a.astype(np.float64).mean() + delta and n > last + wait:
            peaks.append(n)
            last = n
    return np.array(peaks, dtype=int)


def test_peak_pick_matches_definition():
    rng = np.random.default_rng(11)
    for _ in range(5):
        x = rng.random(400).astype(np.float32)
        got = hostlogic.peak_pick(x, 2, 1, 8, 9, 0.07, 2)
        np.testing.assert_array_equal(got, brute_peak_pick(x, 2, 1, 8, 9, 0.07, 2))


def test_onset_backtrack_moves_to_previous_minimum():
    e = np.array([3, 1, 2, 5, 4, 0.5, 0.7, 9], dtype=np.float32)
    np.testing.assert_array_equal(hostlogic.onset_backtrack(np.array([3, 7]), e), [1, 5])
    np.testing.assert_array_equal(hostlogic.onset_backtrack(np.array([0]), e), [0])


def test_time_frame_round_trip():
    t = np.array([0.0, 0.5, 1.2345])
    np.testing.assert_array_equal(hostlogic.time_to_frames(t, 44_100, 512), (t * 44_100).astype(int) // 512)
    assert hostlogic.frames_to_time(10, 44_100, 512) == 10 * 512 / 44_100


def test_stereo_sums_reproduce_oracle():
    st = np.vstack([signals.sine(440.0), 0.5 * signals.sine(523.0, phase=0.3)]).astype(np.float32)
    L, R = st.astype(np.float64)
    mid, side = 0.5 * (st[0] + st[1]), 0.5 * (st[0] - st[1])
    m = [L.sum(), R.sum(), (L * L).sum(), (R * R).sum(), (L * R).sum(), (mid.astype(float) ** 2).sum(),
         (side.astype(float) ** 2).sum(), float(L.size)]
    np.testing.assert_allclose(stereo.mid_side_from_moments(m), fe.mid_side_rms(st), rtol=1e-6)
    assert stereo.correlation_from_moments(m) == pytest.approx(fe.mono_compatibility_correlation(st), abs=1e-6)
    ones = [10.0, 10.0, 10.0, 10.0, 10.0, 10.0, 0.0, 10.0]
    assert stereo.correlation_from_moments(ones) == 1.0  # reference: test_stereo.py:57-64
    # width from per-bin energy sums
    DL, DR = lr.stft(st[0]), lr.stft(st[1])
    be = np.stack([(np.abs(0.5 * (DL + DR)) ** 2).sum(axis=1), (np.abs(0.5 * (DL - DR)) ** 2).sum(axis=1)]).astype(float)
    w = stereo.width_from_band_energy(be, lr.fft_frequencies(22_050, 2048), DL.shape[1], None, 22_050)
    ref = fe.frequency_dependent_width(st, 22_050)
    for k in ("low", "mid", "high"):
        assert w[k] == pytest.approx(ref[k], rel=1e-5, abs=1e-9)


def test_frames_to_db_matches_oracle():
    x = signals.minus18_sine(44_100, seconds=3.0)
    for seconds in (0.4, 3.0):
        frame = max(1024, int(round(44_100 * seconds)))
        frame += frame % 2
        ms = lr.rms(x, frame_length=frame, hop_length=frame // 2)[0].astype(np.float64) ** 2
        np.testing.assert_allclose(loudness_host.frames_to_db(ms), fe.windowed_loudness(x, 44_100, seconds),
                                   rtol=1e-6, atol=1e-5)


def test_partition_covers_every_track_once():
    for lengths, world in (([100] * 10, 4), ([5, 9, 1, 7, 7, 3, 2], 3), ([4], 8), ([], 2)):
        shards = sharding.partition(lengths, world)
        assert len(shards) == world
        assert sorted(i for s in shards for i in s) == list(range(len(lengths)))
    eq = sharding.partition([7] * 1024, 8)
    assert all(len(s) == 128 for s in eq) and eq[1][0] == 128
    ragged = sharding.partition([10, 1, 1, 1, 1, 1, 1, 1, 1, 1, 1], 2)
    assert sum(l for i, l in enumerate([10] + [1] * 10) if i in ragged[0]) == 10


def test_coerce_audio_shapes():
    st = np.vstack([signals.sine(440.0), signals.sine(220.0)])
    a = utils.coerce_audio(st)
    assert a.sample_rate == utils.DEFAULT_SR and a.stereo_samples.shape == st.shape
    np.testing.assert_array_equal(a.samples, np.mean(st, axis=0))
    assert utils.coerce_audio(a).samples is not None and a.duration == st.shape[1] / 44_100
    with pytest.raises(TypeError):
        utils.coerce_audio(123)
    import torch

    if not torch.cuda.is_available():  # another rate means resampling on the device: no CPU fallback, fail loudly
        with pytest.raises((RuntimeError, AssertionError)):
            utils.coerce_audio(utils.AudioInput(st[0], 48_000))


def test_wav_reader_round_trip(tmp_path):
    import struct

    from track_analyser_b200 import io as tio

    x = (np.random.default_rng(0).uniform(-0.5, 0.5, size=(1000, 2)) * 32767).astype("<i2")
    p = tmp_path / "t.wav"
    with open(p, "wb") as fh:
        data = x.tobytes()
        fh.write(b"RIFF" + struct.pack("<I", 36 + len(data)) + b"WAVEfmt " +
                 struct.pack("<IHHIIHH", 16, 1, 2, 44_100, 44_100 * 4, 4, 16) + b"data" + struct.pack("<I", len(data)) + data)
    s, sr, meta = tio.load_audio(str(p), mono=False)
    assert sr == 44_100 and s.shape == (2, 1000) and meta["channels"] == 2
    np.testing.assert_array_equal(s, (x.astype(np.float32) / 32768.0).T)
    m, _, _ = tio.load_audio(str(p))  # the reference's default is mono=True: 1-d channel mean (io.py:59,129-138)
    np.testing.assert_array_equal(m, np.mean(s, axis=0))
    # WAVE_FORMAT_EXTENSIBLE carrying IEEE float: the SubFormat GUID decides, not the width
    f = np.random.default_rng(1).uniform(-0.5, 0.5, size=(100, 2)).astype("<f4")
    guid = struct.pack("<H", 3) + bytes.fromhex("000000001000800000aa00389b71")
    fmt = struct.pack("<HHIIHH", 0xFFFE, 2, 48_000, 48_000 * 8, 8, 32) + struct.pack("<HHI", 22, 32, 3) + guid
    with open(p, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", 20 + len(fmt) + f.nbytes) + b"WAVEfmt " + struct.pack("<I", len(fmt)) + fmt +
                 b"data" + struct.pack("<I", f.nbytes) + f.tobytes())
    s, sr, _ = tio.load_audio(str(p), mono=False)
    assert sr == 48_000
    np.testing.assert_array_equal(s, f.T)
    # a mono file keeps the (1, N) shape under mono=False; a 24-bit data chunk with a ragged tail is cut to whole frames
    raw = bytes(range(1, 3 * 7 + 2 + 1))
    with open(p, "wb") as fh:
        fh.write(b"RIFF" + struct.pack("<I", 36 + len(raw) + 1) + b"WAVEfmt " + struct.pack("<IHHIIHH", 16, 1, 1, 8_000, 24_000, 3, 24) +
                 b"data" + struct.pack("<I", len(raw)) + raw + b"\x00")
    s, sr, meta = tio.load_audio(str(p), mono=False)
    assert s.shape == (1, 7) and meta["channels"] == 1
    assert s[0, 0] == np.float32((1 | (2 << 8) | (3 << 16)) / 8388608.0)


def test_structure_host_logic_on_oracle_curves_like_reference_test():
    """analysis/structure.py host logic (novelty mix, peak picking, refinement, beat snapping, labelling) fed with the
    oracle's arrays: the reference's own structure test (tests/test_structure.py) must hold without a GPU."""
    from track_analyser_b200.analysis import structure
    from track_analyser_b200.analysis.beats import BeatAnalysis

    samples, sr, beat_times = signals.drums_muted_track()
    beat = BeatAnalysis(bpm=120.0, beat_times=beat_times.astype(float).tolist(),
                        beat_frames=(beat_times * sr / 512).astype(int).tolist(), confidence=1.0)
    mag, mel, log_mel, flux = fe.structure_frontend(samples, sr)
    harm, perc = lr.hpss(mag)
    front = structure.StructureFrontend(magnitude=mag, mel=mel, log_mel=log_mel, spectral_flux=flux,
                                        harmonic_curve=np.sum(harm, axis=0), percussive_curve=np.sum(perc, axis=0))
    res = structure.segments_from_curves(front, beat, sample_rate=sr, hop_length=512, duration=len(samples) / sr)
    starts = [s.start for s in res.segments[1:]]
    assert any(abs(b - 12.0) <= 0.5 for b in starts)
    assert [s.label for s in res.segments] == [chr(ord("A") + i) for i in range(len(res.segments))]
    assert res.segments[0].category == "intro" and res.segments[-1].category == "outro"
    assert all(s.end > s.start for s in res.segments) and res.segments[0].start == 0.0
    assert all(set(beat_times).__contains__(s.start) for s in res.segments)  # boundaries are snapped to beats
    for s in res.segments:
        assert 0.0 <= s.percussive_ratio <= 1.0 and 0.0 <= s.confidence <= 1.0
    # the segment that starts where the drums drop out is less percussive than the one before it
    after = next(s for s in res.segments if abs(s.start - 12.0) <= 0.5)
    assert after.percussive_ratio < res.segments[0].percussive_ratio
    assert len(res.novelty_curve) == mel.shape[1]
    empty = structure.StructureFrontend(magnitude=mag[:, :0], mel=mel[:, :0], log_mel=log_mel[:, :0], spectral_flux=flux[:0])
    with pytest.raises(ValueError):
        structure.segments_from_curves(empty, beat, sample_rate=sr, hop_length=512, duration=0.0)


def test_structure_spacing_helpers():
    from track_analyser_b200.analysis import structure

    nov = np.array([0.0, 0.9, 0.1, 0.8, 0.2, 1.0, 0.0])
    np.testing.assert_array_equal(structure._space_frames(np.array([1, 3, 5]), nov, 3), [1, 5])
    np.testing.assert_array_equal(structure._space_frames(np.array([3, 1]), nov, 1), [1, 3])
    mask = structure._space_times([0.0, 1.0, 2.0, 20.0, 21.0, 40.0], [0, 1, 3, 4, 5, 6], nov, 8.0)
    assert mask.tolist() == [True, False, False, False, True, True]  # 21.0 replaces 20.0: higher novelty, too close
    assert structure._refine(np.array([3]), np.array([0.0, 0.0, 0.0, 0.1, 0.9, 0.0]), 1).tolist() == [4]
    assert structure._classify([0.1, 0.7, 0.5, 0.2, 0.4, 0.3], [1, 7, 5, 2, 4, 3], [9, 3, 5, 8, 6, 7]) == [
        "intro", "drop", "groove", "breakdown", "bridge", "outro"]


def test_harmony_host_logic_on_oracle_chroma():
    """Key ranking, chord hints, change points and seeded MIDI suggestions (harmony.py:190-465) on the oracle's
    chroma_stft of the reference's own C-F-G-C test signal (tests/test_harmony.py)."""
    from track_analyser_b200 import harmony
    from track_analyser_b200.analysis import beats

    x, sr = signals.triad_progression()
    chroma = lr.chroma_stft(x, sr)
    keys = harmony._rank_keys(*harmony._score_keys([chroma, chroma]))
    assert keys.best.key == "C major" and keys.second_best.key in {"G major", "F major"}
    assert harmony.key_index(keys) == 0
    beat = beats.build_beat_analysis(bpm=60.0, beat_times=np.arange(4) * 1.0, sr=sr)
    hints = harmony._estimate_chords(chroma, beat, np.random.default_rng(123))
    assert [h.chord[:-3] for h in hints[1:3]] == ["F", "G"] or len(hints) >= 3
    changes = harmony._detect_chord_changes(chroma, beat, hints)
    times = np.array([c.time for c in changes])
    assert sum(bool(np.any(np.abs(times - b) <= 0.25)) for b in (1.0, 2.0, 3.0)) >= 2
    assert harmony._scale_for_key("A minor") == [9, 11, 0, 2, 4, 5, 7] and harmony._scale_for_key("C major")[:3] == [0, 2, 4]
    midi = harmony._generate_midi(chroma, beat, keys.best, np.random.default_rng(5), name="bass", octave=-1, start_offset=0.0)
    assert list(midi.notes.columns) == ["start", "duration", "pitch", "velocity", "channel"] and len(midi.notes) == 4
    assert midi.notes["pitch"].between(48, 59).all() and midi.notes["velocity"].between(20, 127).all()
    assert len(harmony._chord_templates()) == 60
    assert harmony._score_keys([])[0].size == 0 and harmony._rank_keys(np.array([]), []).best.key == "C major"


def test_vectorised_chord_estimate_equals_per_template_loop():
    """harmony._estimate_chords scores all templates with one matrix product; it must reproduce the reference's
    per-template np.dot loop (harmony.py:305-312) exactly, including the seeded tie-breaker draws."""
    from types import SimpleNamespace
    from track_analyser_b200 import harmony

    rng0 = np.random.default_rng(3)
    chroma = rng0.random((12, 3000)).astype(np.float32)
    frames = sorted(set(rng0.integers(0, 3000, size=300).tolist()))
    br = SimpleNamespace(beat_frames=frames, beat_times=[f * 512 / 44100 for f in frames])
    names, mats = zip(*harmony._chord_templates().items())
    rng = np.random.default_rng(7)
    want = []
    for idx, profile in harmony._beat_profiles(chroma, br):
        scores = np.array([float(np.dot(t, profile)) for t in mats])
        best = int(np.argmax(scores + rng.normal(0.0, 1e-6, size=scores.shape)))
        want.append((float(br.beat_times[idx]), names[best], float(scores[best] / float(np.max(scores + 1e-9)))))
    got = [(h.time, h.chord, h.confidence) for h in harmony._estimate_chords(chroma, br, np.random.default_rng(7))]
    assert got == want and len(got) > 250


def test_resampy_restatement_properties():
    """oracle/resampy_np.py is unpinned (resampy is not installed); these are the properties band-limited
    interpolation must have: output length int(n * ratio), unit DC gain and float32-accurate reconstruction of an
    in-band tone when up-sampling, stop-band rejection when down-sampling."""
    from oracle import resampy_np as R

    win, num_table = R.kaiser_best()
    assert win.shape == (64 * 512 + 1,) and num_table == 512 and win[0] == R.KAISER_BEST["rolloff"]
    sr0, sr1 = 22_050, 44_100
    t = np.arange(sr0 // 2) / sr0
    x = (0.5 * np.sin(2 * np.pi * 1000.0 * t)).astype(np.float32)
    y = R.resample(x, sr0, sr1)
    assert y.dtype == np.float32 and y.shape == (int(len(x) * 2.0),)
    tt = np.arange(len(y)) / sr1
    assert np.max(np.abs(y - 0.5 * np.sin(2 * np.pi * 1000.0 * tt))[3000:-3000]) < 1e-6
    dc = R.resample(np.ones(3000, np.float32), sr0, sr1)
    assert np.max(np.abs(dc[1000:-1000] - 1.0)) < 2e-6
    assert R.resample(np.ones(48_000, np.float32), 48_000, 44_100).shape == (44_100,)
    tone = np.sin(2 * np.pi * 23_000.0 * np.arange(24_000) / 48_000).astype(np.float32)  # above the new Nyquist
    assert np.max(np.abs(R.resample(tone, 48_000, 44_100)[2000:-2000])) < 2e-3
    st = np.stack([x, -x])
    np.testing.assert_array_equal(R.resample(st, sr0, sr1)[1], R.resample(-x, sr0, sr1))
    with pytest.raises(ValueError):
        R.resample(np.zeros(1, np.float32), 48_000, 8_000)
