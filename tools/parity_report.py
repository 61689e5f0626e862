"""Per-tensor parity statistics of the CUDA frontend against the CPU oracle.

Run on a GPU box:  python tools/parity_report.py [--seconds 20] [--tracks 2]
Writes gpurun_out/parity_report.json.  (Checker use of oracle/ -- this is a test tool.)
"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

from oracle import frontend as ofe  # noqa: E402
from oracle import librosa_np as olr  # noqa: E402
from oracle import pyloudnorm_np as opl  # noqa: E402
from track_analyser_b200 import engine, synth  # noqa: E402


def stats(name, got, ref, rtol=1e-4, atol=1e-6):
    got = np.asarray(got, dtype=np.float64)
    ref = np.asarray(ref, dtype=np.float64)
    if got.shape != ref.shape:
        return {"name": name, "shape_mismatch": [list(got.shape), list(ref.shape)]}
    err = np.abs(got - ref)
    ok = err <= atol + rtol * np.abs(ref)
    denom = np.maximum(np.abs(ref), 1e-30)
    return {
        "name": name, "shape": list(ref.shape), "max_abs": float(err.max(initial=0)),
        "max_rel": float((err / denom).max(initial=0)), "ref_absmax": float(np.abs(ref).max(initial=0)),
        "pass_rate": float(ok.mean()) if ok.size else 1.0, "allclose": bool(ok.all()),
    }


def oracle_track(x, sr, n_fft, hop, n_mels):
    st = np.asarray(x, dtype=np.float32)
    mono = np.mean(st, axis=0) if st.ndim == 2 else st
    D = olr.stft(mono, n_fft=n_fft, hop_length=hop)
    mag = np.abs(D)
    S = mag**2
    mel = np.einsum("ft,mf->mt", S, olr.filters_mel(sr, n_fft, n_mels=n_mels), optimize=True)
    env = olr.onset_strength(S=olr.power_to_db(mel), sr=sr, hop_length=hop)
    o = {
        "magnitude": mag, "mel": mel, "onset_env": env, "autocorr": olr.autocorrelate(env),
        "flux_linear": olr.onset_strength(S=np.asarray(mel, dtype=float), sr=sr, hop_length=hop),
        "ltas": np.mean(mag, axis=1),
        "centroid": olr.spectral_centroid(mono, sr, n_fft, hop)[0],
        "rolloff": olr.spectral_rolloff(mono, sr, n_fft, hop)[0],
        "kw_blocks": opl.block_energies(mono, sr), "lufs": opl.integrated_loudness(mono, sr),
        "momentary_db": ofe.windowed_loudness(mono, sr, 0.4), "short_db": ofe.windowed_loudness(mono, sr, 3.0),
        "rms_dbfs": ofe.rms_dbfs(mono),
    }
    if st.ndim == 2:
        o["mid_side_rms"] = np.array(ofe.mid_side_rms(st))
        o["correlation"] = ofe.mono_compatibility_correlation(st)
        w = ofe.frequency_dependent_width(st, sr, n_fft=n_fft, hop_length=hop)
        o["width"] = np.array([w["low"], w["mid"], w["high"]])
    return o


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--seconds", type=float, default=20.0)
    ap.add_argument("--tracks", type=int, default=2)
    ap.add_argument("--sr", type=int, default=44100)
    ap.add_argument("--n-fft", type=int, default=2048)
    ap.add_argument("--hop", type=int, default=512)
    ap.add_argument("--mels", type=int, default=128)
    ap.add_argument("--mono", action="store_true")
    args = ap.parse_args()

    from track_analyser_b200 import stereo as pstereo, loudness_host  # product host glue

    tracks = [synth.synth_track(synth.DEFAULT_SEED + i, args.seconds * (1 + 0.13 * i), args.sr,
                                1 if args.mono else 2) for i in range(args.tracks)]
    plan = engine.Plan(args.sr, args.n_fft, args.hop, args.mels)
    t0 = time.time()
    res = engine.analyse_batch(plan, tracks, engine.ALL_OUTPUTS)
    t_gpu = time.time() - t0
    report = {"gpu_seconds_first_call": t_gpu, "tracks": []}
    freqs = np.fft.rfftfreq(args.n_fft, 1.0 / args.sr)
    for i, (x, r) in enumerate(zip(tracks, res)):
        t0 = time.time()
        o = oracle_track(x, args.sr, args.n_fft, args.hop, args.mels)
        rows = [stats(k, r[k], o[k]) for k in ("magnitude", "mel", "onset_env", "autocorr", "flux_linear", "ltas",
                                                "centroid", "kw_blocks")]
        rows.append(stats("rolloff", freqs[r["rolloff_bin"]], o["rolloff"]))
        rows.append({"name": "rolloff_bins_equal", "pass_rate": float(np.mean(freqs[r["rolloff_bin"]] == o["rolloff"]))})
        rows.append({"name": "lufs", "got": r["lufs"], "ref": o["lufs"], "abs_err": abs(r["lufs"] - o["lufs"])})
        mono = np.mean(x, axis=0) if x.ndim == 2 else x
        ref_chroma, ref_tuning = olr.chroma_stft(mono, args.sr, n_fft=args.n_fft, hop_length=args.hop, return_tuning=True)
        rows.append(stats("chroma", r["chroma"], ref_chroma))
        rows.append({"name": "tuning", "got": r["tuning"], "ref": ref_tuning, "equal": bool(abs(r["tuning"] - ref_tuning) < 1e-12)})
        rows.append(stats("tempogram", r["tempogram"], olr.tempogram(onset_envelope=o["onset_env"], sr=args.sr, hop_length=args.hop)))
        harm, perc = olr.hpss(o["magnitude"])
        rows.append(stats("hpss_harmonic", r["hpss_harmonic"], np.sum(harm, axis=0, dtype=np.float64)))
        rows.append(stats("hpss_percussive", r["hpss_percussive"], np.sum(perc, axis=0, dtype=np.float64)))
        rows.append(stats("mfcc", r["mfcc"], olr.mfcc(olr.power_to_db(np.asarray(o["mel"], dtype=float) + 1e-9)), atol=5e-3))
        tp_ref = ofe.true_peak_dbtp(mono, args.sr)
        tp_got = 20.0 * np.log10(r["true_peak"] + 1e-12)
        rows.append({"name": "true_peak_db", "got": tp_got, "ref": tp_ref, "abs_err": abs(tp_got - tp_ref)})
        from track_analyser_b200 import hostlogic
        on_g = hostlogic.onset_detect(r["onset_env"], args.sr, args.hop, backtrack=True)
        on_o = hostlogic.onset_detect(o["onset_env"], args.sr, args.hop, backtrack=True)
        rows.append({"name": "onset_frames", "count": int(on_o.size), "bit_exact": bool(np.array_equal(on_g, on_o))})
        rows.append(stats("momentary_db", loudness_host.frames_to_db(r["rms_momentary"]), o["momentary_db"]))
        rows.append(stats("short_db", loudness_host.frames_to_db(r["rms_short"]), o["short_db"]))
        mo = r["moments"]
        if x.ndim == 2:
            rows.append(stats("mid_side_rms", np.array(pstereo.mid_side_from_moments(mo)), o["mid_side_rms"]))
            rows.append(stats("correlation", pstereo.correlation_from_moments(mo), o["correlation"]))
            wd = pstereo.width_from_band_energy(r["band_energy"], freqs, r.n_frames, None, args.sr)
            rows.append(stats("width", np.array([wd["low"], wd["mid"], wd["high"]]), o["width"]))
        report["tracks"].append({"index": i, "n_samples": r.n_samples, "oracle_seconds": time.time() - t0, "rows": rows})
        for row in rows:
            print(i, json.dumps(row))
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/parity_report.json", "w") as fh:
        json.dump(report, fh, indent=1)
    # compact table for profiles/
    with open("gpurun_out/parity_report.md", "w") as fh:
        fh.write("| track | output | max abs err | max rel err | |ref| max | pass rate at rtol 1e-4 / atol 1e-6 |\n|---|---|---|---|---|---|\n")
        for t in report["tracks"]:
            for row in t["rows"]:
                if "max_abs" in row:
                    fh.write(f"| {t['index']} | {row['name']} | {row['max_abs']:.3e} | {row['max_rel']:.3e} | {row['ref_absmax']:.3e} | {row['pass_rate']:.6f} |\n")
                else:
                    fh.write(f"| {t['index']} | {row['name']} | " + ", ".join(f"{k}={v}" for k, v in row.items() if k != "name") + " | | | |\n")


if __name__ == "__main__":
    main()
