"""pipeline.analyse_tracks (the batch form of analyse_track, reference pipeline.py:32-120): kernels on chunks of tracks, host
stages in worker processes on precomputed results.  Must return what analyse_track returns per track.  Run with -m gpu."""

import dataclasses

import numpy as np
import pytest

torch = pytest.importorskip("torch")
pytestmark = [pytest.mark.gpu, pytest.mark.skipif(not torch.cuda.is_available(), reason="needs a CUDA device")]

from track_analyser_b200 import pipeline, runtime, synth  # noqa: E402
from track_analyser_b200.utils import AudioInput  # noqa: E402


def assert_same(a, b, path="result"):
    """Deep comparison: integers / strings / lists exact, floats to 1e-6 relative (sums gathered with float64 atomics differ
    in their last bits from run to run)."""
    import pandas as pd

    if dataclasses.is_dataclass(a):
        assert type(a) is type(b), path
        for f in dataclasses.fields(a):
            assert_same(getattr(a, f.name), getattr(b, f.name), f"{path}.{f.name}")
    elif isinstance(a, pd.DataFrame):
        assert list(a.columns) == list(b.columns) and len(a) == len(b), path
        for c in a.columns:
            assert_same(a[c].to_numpy(), b[c].to_numpy(), f"{path}[{c}]")
    elif isinstance(a, np.ndarray):
        assert a.shape == b.shape and a.dtype == b.dtype, path
        if a.dtype.kind in "fc":
            np.testing.assert_allclose(a, b, rtol=1e-6, atol=1e-9, err_msg=path)
        else:
            np.testing.assert_array_equal(a, b, err_msg=path)
    elif isinstance(a, (list, tuple)):
        assert len(a) == len(b), path
        for i, (x, y) in enumerate(zip(a, b)):
            assert_same(x, y, f"{path}[{i}]")
    elif isinstance(a, dict):
        assert a.keys() == b.keys(), path
        for k in a:
            assert_same(a[k], b[k], f"{path}[{k!r}]")
    elif isinstance(a, float):
        assert b == pytest.approx(a, rel=1e-6, abs=1e-9) or (np.isnan(a) and np.isnan(b)), path
    else:
        assert a == b, path


def _audio(seed, seconds, sr, channels):
    x = synth.synth_track(seed, seconds, sr, channels)
    if channels == 2:
        return AudioInput(samples=np.mean(x, axis=0), sample_rate=sr, stereo_samples=x)
    return AudioInput(samples=x, sample_rate=sr)


@pytest.mark.parametrize("workers", [0, 3])
def test_analyse_tracks_equals_analyse_track(workers):
    sr = 44_100
    audios = [_audio(900, 9.0, sr, 2), _audio(901, 6.5, sr, 2), _audio(902, 7.25, sr, 1), _audio(903, 8.0, 22_050, 2)]
    got = pipeline.analyse_tracks(audios, workers=workers, chunk_tracks=2)
    assert len(got) == len(audios)
    for a, g in zip(audios, got):
        want = pipeline.analyse_track(a)
        assert isinstance(g, pipeline.TrackAnalysisResult) and g.audio is a
        for stage in ("beat", "downbeat", "structure", "loudness", "harmonic", "features", "stereo"):
            assert_same(getattr(want, stage), getattr(g, stage), stage)


def test_precomputed_session_refuses_what_was_not_computed():
    with runtime.precomputed_session({}):
        assert runtime.is_precomputed()
        with pytest.raises(RuntimeError):
            runtime.frontend(np.zeros(4096, np.float32), 44_100, outputs=("ltas",))
    assert not runtime.is_precomputed()


def test_inconsistent_mono_view_takes_the_single_track_path():
    """An AudioInput whose mono samples are not the mean of its stereo pair cannot share one stereo run."""
    sr = 44_100
    x = synth.synth_track(77, 5.0, sr, 2)
    a = AudioInput(samples=(0.5 * x[0]).astype(np.float32), sample_rate=sr, stereo_samples=x)
    got = pipeline.analyse_tracks([a], workers=0)[0]
    want = pipeline.analyse_track(a)
    assert_same(want.loudness, got.loudness, "loudness")
    assert_same(want.stereo, got.stereo, "stereo")
