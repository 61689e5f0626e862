"""chroma_cqt on the device (csrc/cqt.cu) against the CPU oracle and a brute-force CQT.  Run with -m gpu.

Reference call sites: harmony.py:107 (key_estimate) and :148 (analyse_harmony) -> librosa.feature.chroma_cqt(y=y, sr=sr).
PARITY UNPINNED against librosa itself (see oracle/cqt_np.py): the oracle restates librosa's recursive CQT with a STATED
decimator (libsoxr cannot be restated here), the kernels use the same one, and the restatement is cross-checked against
the brute-force definition in tests/test_oracle_cqt.py and below (key index / chord labels equal across all three)."""

import ctypes as C

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

from oracle import cqt_np as ocq  # noqa: E402
from track_analyser_b200 import _native as nat  # noqa: E402
from track_analyser_b200 import engine, harmony, synth  # noqa: E402
from track_analyser_b200.analysis.beats import BeatAnalysis  # noqa: E402

from . import signals  # noqa: E402

RTOL, ATOL = 1e-4, 1e-6
_plans = {}


def plan_for(sr):
    if sr not in _plans:
        _plans[sr] = engine.Plan(sr, 2048, 512, 128, device=0)
    return _plans[sr]


def assert_chroma_cqt(got, ref):
    """rtol 1e-4 / atol 1e-6 on >= 99.5 % of the values and atol 4e-6 on all of them: a chroma value is a sum of 21 float32
    constant-Q magnitudes divided by the frame's maximum, and BOTH sides accumulate their projections in float32 (the oracle
    as librosa does, in complex64), so values two orders below the frame's maximum carry ~1e-6 of evaluation-order noise."""
    got, ref = np.asarray(got), np.asarray(ref)
    assert got.shape == ref.shape
    if ref.size:
        assert np.mean(np.abs(got - ref) <= ATOL + RTOL * np.abs(ref)) >= 0.995
    np.testing.assert_allclose(got, ref, rtol=RTOL, atol=4e-6)


def detuned_tone(sr=44_100, seconds=3.0, cents=31.0):
    t = np.arange(int(sr * seconds)) / sr
    f0 = 220.0 * 2.0 ** (cents / 1200.0)
    y = sum(a * np.sin(2 * np.pi * f0 * h * t) for h, a in ((1, 0.4), (2, 0.2), (3, 0.1), (5, 0.05)))
    return y.astype(np.float32)


def cases():
    x, sr = signals.triad_progression()
    yield "triad_22k05", x, sr
    yield "detuned_44k1", detuned_tone(), 44_100
    yield "synth_stereo_44k1", synth.synth_track(synth.DEFAULT_SEED, 7.0, 44_100, 2), 44_100
    yield "synth_mono_48k", synth.synth_track(77, 5.3, 48_000, 1), 48_000


@pytest.mark.parametrize("name,x,sr", list(cases()), ids=[c[0] for c in cases()])
def test_chroma_cqt_matches_oracle(name, x, sr):
    res = engine.analyse_batch(plan_for(sr), [x], ("chroma_cqt", "cqt_tuning", "cqt_mag"))[0]
    mono = np.mean(x, axis=0) if x.ndim == 2 else x
    chroma, Cq, tuning = ocq.chroma_cqt(mono, sr, return_parts=True)
    assert res["cqt_tuning"] == pytest.approx(tuning, abs=1e-12)          # histogram arg-max bin: exact
    assert res["cqt_mag"].shape == Cq.shape and res["chroma_cqt"].shape == chroma.shape
    # |CQT|: fp32 transform noise scales with the loudest component of the frame, like the STFT magnitude
    scale = float(np.max(Cq))
    np.testing.assert_allclose(res["cqt_mag"], Cq, rtol=RTOL, atol=2e-6 * max(scale, 1.0))
    assert_chroma_cqt(res["chroma_cqt"], chroma)


def test_key_and_chords_equal_across_gpu_oracle_and_brute_force():
    """The integer outputs chroma_cqt feeds (key index, chord labels) agree whether the chroma comes from the kernels,
    from the multi-rate oracle, or from the brute-force definition (full-rate correlation with every atom)."""
    x, sr = signals.triad_progression()
    plan = plan_for(sr)
    res = engine.analyse_batch(plan, [x], ("chroma_cqt", "cqt_tuning", "chroma", "tuning"))[0]
    o_chroma, o_C, o_tuning = ocq.chroma_cqt(x, sr, return_parts=True)
    frames = np.arange(o_C.shape[1])
    B = ocq.brute_force_cqt(x, sr, tuning=o_tuning, frames=frames)
    raw = ocq.cq_to_chroma(252) @ B.astype(np.float32)
    b_chroma = raw / np.maximum(raw.max(axis=0, keepdims=True), np.finfo(np.float32).tiny)
    stft_chroma = res["chroma"]
    keys = [harmony.key_index(harmony._rank_keys(*harmony._score_keys([c, stft_chroma])))
            for c in (res["chroma_cqt"], o_chroma, b_chroma)]
    assert keys[0] == keys[1] == keys[2]
    assert harmony._key_names()[keys[0]] == "C major"                      # reference tests/test_harmony.py:37-64
    beat_frames = list(range(8, o_C.shape[1] - 8, 11))
    beats = BeatAnalysis(bpm=120.0, beat_times=[f * 512 / sr for f in beat_frames], beat_frames=beat_frames, confidence=1.0)
    labels = []
    for c in (res["chroma_cqt"], o_chroma, b_chroma):
        hints = harmony._estimate_chords(np.asarray(c), beats, np.random.default_rng(0))
        labels.append([h.chord for h in hints])
    assert labels[0] == labels[1]                      # kernels vs multi-rate oracle: every label
    agree = np.mean([a == b for a, b in zip(labels[1], labels[2])])
    assert agree >= 0.9                                # vs the unsparsified brute-force definition: transitions may differ


def test_key_estimate_api_uses_the_constant_q_chroma():
    x, sr = signals.triad_progression()
    est = harmony.key_estimate(x, sr)
    o_cqt = ocq.chroma_cqt(x, sr)
    from oracle import frontend as ofe

    o_stft = ofe.chroma_stft(x, sr)
    want = harmony._rank_keys(*harmony._score_keys([o_cqt, o_stft]))
    assert est.best.key == want.best.key == "C major" and est.second_best.key == want.second_best.key
    assert est.best.confidence == pytest.approx(want.best.confidence, rel=1e-4)
    got = harmony._chroma_cqt(x, sr)
    assert_chroma_cqt(got, o_cqt)


def test_ragged_batch_and_c_abi_entry_point():
    sr = 44_100
    tracks = [synth.synth_track(300 + i, 2.0 + 1.7 * i, sr, 2) for i in range(3)]
    plan = plan_for(sr)
    res = engine.analyse_batch(plan, tracks, ("chroma_cqt", "cqt_tuning", "magnitude", "frame_max"))
    for r, x in zip(res, tracks):
        ref, _, tun = ocq.chroma_cqt(np.mean(x, axis=0), sr, return_parts=True)
        assert r["cqt_tuning"] == pytest.approx(tun, abs=1e-12)
        assert_chroma_cqt(r["chroma_cqt"], ref)
    # the stand-alone stage through the C ABI on the magnitude the fused run produced
    batch = engine.upload(plan, tracks)
    bufs = engine.FrontendBuffers(batch, ("magnitude", "frame_max"))
    engine.run_device(plan, batch, bufs, stage="stft")
    frames, pitch, off = batch.cqt_layout()
    out = torch.empty(12 * int(off[-1]), dtype=torch.float32, device="cuda:0")
    tun = torch.empty(len(tracks), dtype=torch.float64, device="cuda:0")
    need = plan.lib.ta_cqt_scratch_bytes(plan._h, C.byref(batch.c_batch))
    assert need > 0
    scratch = torch.empty(need, dtype=torch.uint8, device="cuda:0")
    ws = engine.workspace(plan, batch)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)
    nat.check(plan.lib.ta_chroma_cqt(plan._h, C.byref(batch.c_batch), C.c_void_p(bufs.t["magnitude"].data_ptr()),
                                     C.c_void_p(bufs.t["frame_max"].data_ptr()), C.c_void_p(out.data_ptr()), None,
                                     C.c_void_p(tun.data_ptr()), C.c_void_p(scratch.data_ptr()), need,
                                     C.c_void_p(ws.data_ptr()), ws.numel(), stream))
    torch.cuda.synchronize()
    h = out.cpu().numpy()
    for i, r in enumerate(res):
        got = h[12 * int(off[i]): 12 * int(off[i] + pitch[i])].reshape(12, int(pitch[i]))[:, : int(frames[i])]
        np.testing.assert_array_equal(got, r["chroma_cqt"])
    # too small a scratch is refused, not overrun
    rc = plan.lib.ta_chroma_cqt(plan._h, C.byref(batch.c_batch), C.c_void_p(bufs.t["magnitude"].data_ptr()),
                                C.c_void_p(bufs.t["frame_max"].data_ptr()), C.c_void_p(out.data_ptr()), None,
                                C.c_void_p(tun.data_ptr()), C.c_void_p(scratch.data_ptr()), 1024,
                                C.c_void_p(ws.data_ptr()), ws.numel(), stream)
    assert rc == nat.TA_ERR_INVALID


@pytest.mark.parametrize("n", [0, 5, 700, 2047, 4097])
def test_short_and_silent_inputs(n):
    sr = 22_050
    rng = np.random.default_rng(n)
    x = (0.1 * rng.standard_normal(n)).astype(np.float32)
    res = engine.analyse_batch(plan_for(sr), [x], ("chroma_cqt", "cqt_tuning"))[0]
    ref, _, tun = ocq.chroma_cqt(x, sr, return_parts=True)
    assert res["chroma_cqt"].shape == ref.shape
    assert res["cqt_tuning"] == pytest.approx(tun, abs=1e-12)
    assert_chroma_cqt(res["chroma_cqt"], ref)
    z = np.zeros(3000, dtype=np.float32)
    res = engine.analyse_batch(plan_for(sr), [z], ("chroma_cqt", "cqt_tuning"))[0]
    assert np.all(res["chroma_cqt"] == 0.0) and res["cqt_tuning"] == 0.0


def test_unsupported_configurations_fail_loudly():
    # 8 kHz: the top constant-Q filter exceeds the Nyquist frequency (librosa raises ParameterError)
    p8 = engine.Plan(8_000, 2048, 512, 128, device=0)
    assert not p8.cqt_ok
    with pytest.raises(nat.NativeError):
        engine.analyse_batch(p8, [np.zeros(8000, np.float32)], ("chroma_cqt",))
    # a plan that is not librosa's chroma_cqt default (hop 512, tuning from a 2048-point STFT)
    p = engine.Plan(44_100, 4096, 1024, 0, device=0)
    assert not p.cqt_ok
    with pytest.raises(nat.NativeError):
        engine.analyse_batch(p, [np.zeros(44_100, np.float32)], ("chroma_cqt",))
    with pytest.raises(ValueError):
        harmony.key_estimate(np.zeros(8000, np.float32), 8_000)


def test_full_size_properties_of_the_round_two_outputs():
    """3-minute 44.1 kHz stereo track (BASELINE configs[1]): properties that need no oracle.  A power-of-two gain is exact in
    binary floating point, so the inf-normalised constant-Q chroma, its tuning estimate and the roll-off bins must not move
    by a single bit; a track's results must not depend on its neighbours in the batch."""
    sr = 44_100
    x = synth.synth_track(synth.DEFAULT_SEED, 180.0, sr, 2)
    plan = plan_for(sr)
    outs = ("chroma_cqt", "cqt_tuning", "rolloff_bin", "self_similarity", "chroma", "tuning")
    r1 = engine.analyse_batch(plan, [x], outs)[0]
    assert r1["chroma_cqt"].shape == (12, 15_504) and np.all(np.isfinite(r1["chroma_cqt"]))
    assert float(np.max(r1["chroma_cqt"])) == 1.0 and float(np.min(r1["chroma_cqt"])) >= 0.0
    r2 = engine.analyse_batch(plan, [0.5 * x], outs)[0]
    for k in ("chroma_cqt", "cqt_tuning", "rolloff_bin", "chroma", "tuning"):
        np.testing.assert_array_equal(np.asarray(r2[k]), np.asarray(r1[k]), err_msg=k)
    other = synth.synth_track(91, 47.0, sr, 2)
    rb = engine.analyse_batch(plan, [other, x, other[:, :30_000]], outs)[1]
    for k in outs:
        np.testing.assert_array_equal(np.asarray(rb[k]), np.asarray(r1[k]), err_msg=k)
