"""track_analyser_b200.install(): the UNMODIFIED reference package (pip-installed without its dependencies into
baseline/_ref by tools/install_reference.sh; git-ignored) imported and run on the B200 frontend through shim modules for
librosa / pyloudnorm / resampy / audioread (compat.py).  Skipped when baseline/_ref is absent.  Run with -m gpu.

What this shows (north star: "the beat, structure, harmonic ... code runs unchanged on its outputs"): the reference's own
host logic, fed by the kernels, returns what the mirrored package returns, and the reference's own test-suite for the
frontend modules passes on the kernels."""

import os
import subprocess
import sys

import numpy as np
import pytest

torch = pytest.importorskip("torch")
ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))
REF = os.path.join(ROOT, "baseline", "_ref")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device"),
              pytest.mark.skipif(not os.path.isdir(os.path.join(REF, "track_analyser")),
                                 reason="reference not installed (tools/install_reference.sh)")]

_SCRIPT = r"""
import sys, json
import numpy as np
sys.path.insert(0, {root!r}); sys.path.insert(0, {ref!r})
import track_analyser_b200
ref = track_analyser_b200.install()
assert "librosa" in ref.__ta_b200_installed__, ref.__ta_b200_installed__
import librosa
assert getattr(librosa, "__ta_b200_shim__", False)
from track_analyser.utils import AudioInput as RefAudio
from track_analyser_b200 import pipeline as ours, synth
from track_analyser_b200.utils import AudioInput
x = synth.synth_track(4242, 40.0, 44_100, 2)
mono = np.mean(x, axis=0)
r = ref.analyse_track(RefAudio(samples=mono, sample_rate=44_100, stereo_samples=x))
o = ours.analyse_track(AudioInput(samples=mono, sample_rate=44_100, stereo_samples=x))
out = dict(
    bpm=(r.beat.bpm, o.beat.bpm), beats=(list(r.beat.beat_frames), list(o.beat.beat_frames)),
    downbeats=(len(r.downbeat.downbeat_times), len(o.downbeat.downbeat_times)),
    seg=([(s.label, round(s.start, 6), round(s.end, 6)) for s in r.structure.segments],
         [(s.label, round(s.start, 6), round(s.end, 6)) for s in o.structure.segments]),
    novelty=float(np.max(np.abs(np.asarray(r.structure.novelty_curve) - np.asarray(o.structure.novelty_curve)))),
    lufs=(r.loudness.integrated_lufs, o.loudness.integrated_lufs), tp=(r.loudness.true_peak_dbfs, o.loudness.true_peak_dbfs),
    rms=(r.loudness.rms_dbfs, o.loudness.rms_dbfs), lra=(r.loudness.loudness_range, o.loudness.loudness_range),
    key=(r.harmonic.primary_key.key, o.harmonic.primary_key.key), key2=(r.harmonic.secondary_key.key, o.harmonic.secondary_key.key),
    chords=([h.chord for h in r.harmonic.chord_hints], [h.chord for h in o.harmonic.chord_hints]),
    balance=((r.harmonic.spectral_balance.low_band, r.harmonic.spectral_balance.mid_band, r.harmonic.spectral_balance.high_band),
             (o.harmonic.spectral_balance.low_band, o.harmonic.spectral_balance.mid_band, o.harmonic.spectral_balance.high_band)),
    ltas=float(np.max(np.abs(np.asarray(r.features.ltas.magnitude) - np.asarray(o.features.ltas.magnitude)))),
    centroid=float(np.max(np.abs(np.asarray(r.features.spectral_centroid.values) - np.asarray(o.features.spectral_centroid.values)))),
    rolloff=bool(np.array_equal(np.asarray(r.features.spectral_rolloff.values), np.asarray(o.features.spectral_rolloff.values))),
    stereo=((r.stereo.mid_rms, r.stereo.side_rms, r.stereo.correlation, r.stereo.width.low, r.stereo.width.mid, r.stereo.width.high),
            (o.stereo.mid_rms, o.stereo.side_rms, o.stereo.correlation, o.stereo.width.low, o.stereo.width.mid, o.stereo.width.high)),
)
print("RESULT" + json.dumps(out))
"""


def _run(code: str, timeout=900):
    return subprocess.run([sys.executable, "-c", code], capture_output=True, text=True, timeout=timeout, cwd=ROOT)


def test_reference_analyse_track_runs_on_the_kernels_and_agrees_with_the_mirror():
    import json

    p = _run(_SCRIPT.format(root=ROOT, ref=REF))
    assert p.returncode == 0, p.stderr[-4000:]
    out = json.loads([line for line in p.stdout.splitlines() if line.startswith("RESULT")][0][6:])
    assert out["bpm"][0] == pytest.approx(out["bpm"][1], rel=1e-9) and out["beats"][0] == out["beats"][1]
    assert out["downbeats"][0] == out["downbeats"][1]
    assert out["seg"][0] == out["seg"][1] and out["novelty"] < 1e-9
    for k in ("lufs", "rms", "lra"):
        assert out[k][0] == pytest.approx(out[k][1], abs=1e-6), k
    # the reference takes max|.| of the float32 signal the resampy shim returns; the mirror's kernel keeps the peak in float64
    assert out["tp"][0] == pytest.approx(out["tp"][1], abs=1e-4)
    assert out["key"][0] == out["key"][1] and out["key2"][0] == out["key2"][1] and out["chords"][0] == out["chords"][1]
    assert out["balance"][0] == pytest.approx(out["balance"][1], rel=1e-6)
    assert out["ltas"] < 1e-6 and out["centroid"] < 1e-6 and out["rolloff"]
    # np.corrcoef on float32 mid / side (stereo.py:60-66) against the kernel's float64 moments: float32 rounding apart
    assert out["stereo"][0] == pytest.approx(out["stereo"][1], rel=1e-5, abs=1e-9)


def test_reference_test_suite_passes_on_the_kernels():
    tests = os.path.join(REF, "_reference_tests")
    if not os.path.isdir(tests):
        pytest.skip("reference tests not copied (tools/install_reference.sh)")
    files = [os.path.join(tests, f) for f in ("test_features.py", "test_stereo.py", "test_loudness.py", "test_tempo.py",
                                              "test_structure.py", "test_harmony.py")]
    code = ("import sys; sys.path.insert(0, %r); sys.path.insert(0, %r); import track_analyser_b200; track_analyser_b200.install(); "
            "import pytest; sys.exit(pytest.main(['-q', '-x', '-p', 'no:cacheprovider', '--noconftest'] + %r))" % (ROOT, REF, files))
    p = _run(code, timeout=1500)
    assert p.returncode == 0, (p.stdout[-3000:] + p.stderr[-3000:])
