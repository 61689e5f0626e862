"""Generates tests/golden/*.npz from the CPU oracle (the reference itself cannot be imported:
librosa/pyloudnorm are not installed -- see oracle/__init__.py).  Run from the repo root:
    python tests/golden/make_golden.py
"""

import os
import sys

import numpy as np

ROOT = os.path.dirname(os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, ROOT)

from oracle import frontend as fe  # noqa: E402
from oracle import librosa_np as lr  # noqa: E402
from oracle import pyloudnorm_np as pl  # noqa: E402
from tests import signals  # noqa: E402
from track_analyser_b200 import synth  # noqa: E402

HERE = os.path.dirname(os.path.abspath(__file__))


def bundle(x, sr, stereo=None):
    mono = np.asarray(x, dtype=np.float32)
    env = fe.onset_envelope(mono, sr)
    mag, mel, _, flux = fe.structure_frontend(mono, sr)
    out = dict(
        samples=mono, onset_env=env, autocorr=fe.onset_autocorrelation(env), ltas=fe.compute_ltas(mono, sr)[1],
        centroid=fe.spectral_centroid_series(mono, sr), rolloff=fe.spectral_rolloff_series(mono, sr),
        mel=mel.astype(np.float32), flux_linear=flux, lufs=np.float64(pl.integrated_loudness(mono, sr)),
        kw_blocks=pl.block_energies(mono, sr), momentary_db=fe.windowed_loudness(mono, sr, 0.4),
        short_db=fe.windowed_loudness(mono, sr, 3.0),
        magnitude_rows=mag[::64].astype(np.float32),  # every 64th bin keeps the fixture small
    )
    # rows added after the first fixtures: chroma_stft + tuning, tempogram (every 4th frame), HPSS column sums,
    # MFCC of the structure stage, true peak
    chroma, tuning = fe.chroma_stft(mono, sr, return_tuning=True)
    harm, perc = lr.hpss(mag)
    out.update(
        chroma=chroma.astype(np.float32), tuning=np.float64(tuning),
        tempogram_cols=lr.tempogram(onset_envelope=env, sr=sr, hop_length=512)[:, ::4].astype(np.float32),
        hpss_harmonic=np.sum(harm, axis=0, dtype=np.float64), hpss_percussive=np.sum(perc, axis=0, dtype=np.float64),
        mfcc=lr.mfcc(lr.power_to_db(np.asarray(mel, dtype=float) + 1e-9)),
        true_peak_db=np.float64(fe.true_peak_dbtp(mono, sr)),
    )
    if stereo is not None:
        w = fe.frequency_dependent_width(stereo, sr)
        out.update(stereo=stereo, mid_side_rms=np.array(fe.mid_side_rms(stereo)),
                   correlation=np.float64(fe.mono_compatibility_correlation(stereo)),
                   width=np.array([w["low"], w["mid"], w["high"]]))
    return out


def main():
    np.savez_compressed(os.path.join(HERE, "tiny_click.npz"), **bundle(signals.tiny_click(), 44_100))
    st = synth.synth_track(synth.DEFAULT_SEED, 4.0, 44_100, 2)
    np.savez_compressed(os.path.join(HERE, "synth_stereo_4s.npz"), **bundle(np.mean(st, axis=0), 44_100, st))
    print("wrote", sorted(f for f in os.listdir(HERE) if f.endswith(".npz")))


if __name__ == "__main__":
    main()
