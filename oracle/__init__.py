"""CPU oracle for the track-analyser spectral frontend.  TEST INFRASTRUCTURE ONLY.

This package is a numpy/scipy restatement of the arithmetic the reference
(cillianjoy/track-analyser) delegates to librosa 0.10.2.post1, pyloudnorm 0.1.1,
scipy 1.11.4 and numpy 1.26.4 on its hot path.

Who may import it: ``tests/``, ``__graft_entry__.smoke()`` and ``bench.py``'s
``cpu_baseline`` / ``--impl reference`` legs, and only as the *checker* or the
*timed CPU baseline*.  The product package ``track_analyser_b200`` never imports
it and has no CPU fallback.

PARITY UNPINNED at the north-star tolerance (rtol 1e-4 / atol 1e-6 / 0.01 LU):
none of librosa, pyloudnorm, soundfile is installed in this image or on the GPU
box, the reference cannot be imported (``ModuleNotFoundError: librosa`` raised at
``analysis/beats.py:20``) and the reference ships no golden vectors for this
path.  What *is* pinned:
  * every property-level assertion of the reference's own tests for the path
    (tests/test_loudness.py:33-43, test_features.py:15-44, test_stereo.py:15-64,
    test_tempo.py:39-53) -- re-run against this oracle in
    ``tests/test_oracle_reference_properties.py``;
  * independent cross-checks available offline: ``torchaudio`` mel filterbank and
    BS.1770 loudness, ``scipy.signal.get_window``, ``scipy.signal.lfilter``,
    direct O(n^2) DFT / autocorrelation sums (``tests/test_oracle_crosschecks.py``).

Float64 FFTs are used on purpose: numpy 1.26 (the reference's pin) always
transforms in double and librosa then rounds to complex64.
"""

from . import librosa_np, pyloudnorm_np, frontend  # noqa: F401
