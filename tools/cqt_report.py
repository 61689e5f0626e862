"""chroma_cqt on the device against the oracle: error statistics and kernel time.  Run on a GPU box:
    python tools/cqt_report.py [--tracks 32] [--seconds 180]
Writes gpurun_out/cqt_report.json.  (Checker use of oracle/ -- this is a test tool.)"""

from __future__ import annotations

import argparse
import json
import os
import sys
import time

import numpy as np

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))

import torch  # noqa: E402

from oracle import cqt_np as ocq  # noqa: E402
from track_analyser_b200 import engine, synth  # noqa: E402


def main():
    ap = argparse.ArgumentParser()
    ap.add_argument("--tracks", type=int, default=32)
    ap.add_argument("--seconds", type=float, default=180.0)
    ap.add_argument("--check-seconds", type=float, default=12.0)
    args = ap.parse_args()
    sr = 44_100
    plan = engine.Plan(sr, 2048, 512, 128, device=0)
    report = {"cases": []}
    for seed, ch in ((13370, 2), (7, 1)):
        x = synth.synth_track(seed, args.check_seconds, sr, ch)
        res = engine.analyse_batch(plan, [x], ("chroma_cqt", "cqt_tuning", "cqt_mag"))[0]
        mono = np.mean(x, axis=0) if ch == 2 else x
        t0 = time.time()
        chroma, Cq, tuning = ocq.chroma_cqt(mono, sr, return_parts=True)
        cpu_s = time.time() - t0
        em = np.abs(res["cqt_mag"] - Cq)
        ec = np.abs(res["chroma_cqt"] - chroma)
        report["cases"].append({
            "seed": seed, "channels": ch, "tuning_gpu": res["cqt_tuning"], "tuning_oracle": tuning, "oracle_seconds": cpu_s,
            "cqt_absmax": float(Cq.max()), "cqt_max_abs_err": float(em.max()),
            "cqt_pass_rate_1e-4_1e-6": float(np.mean(em <= 1e-6 + 1e-4 * Cq)),
            "cqt_pass_rate_1e-4_2e-6scaled": float(np.mean(em <= 2e-6 * max(1.0, float(Cq.max())) + 1e-4 * Cq)),
            "chroma_max_abs_err": float(ec.max()), "chroma_pass_rate_1e-4_1e-6": float(np.mean(ec <= 1e-6 + 1e-4 * chroma)),
            "chroma_pass_rate_1e-4_2e-6": float(np.mean(ec <= 2e-6 + 1e-4 * chroma)),
        })
    # timing: batch resident, chroma_cqt stage alone through the C ABI (magnitude from the fused STFT stage)
    import ctypes as C

    from track_analyser_b200 import _native as nat

    tracks = [synth.synth_track(100 + i, args.seconds, sr, 2) for i in range(min(args.tracks, 4))]
    tracks = (tracks * ((args.tracks + len(tracks) - 1) // len(tracks)))[: args.tracks]
    batch = engine.upload(plan, tracks)
    bufs = engine.FrontendBuffers(batch, ("magnitude", "frame_max", "chroma_cqt", "cqt_tuning"))
    engine.run_device(plan, batch, bufs, stage="stft")
    ws = engine.workspace(plan, batch)
    stream = C.c_void_p(torch.cuda.current_stream().cuda_stream)

    def run():
        nat.check(plan.lib.ta_chroma_cqt(plan._h, C.byref(batch.c_batch), C.c_void_p(bufs.t["magnitude"].data_ptr()),
                                         C.c_void_p(bufs.t["frame_max"].data_ptr()), C.c_void_p(bufs.t["chroma_cqt"].data_ptr()),
                                         None, C.c_void_p(bufs.t["cqt_tuning"].data_ptr()),
                                         C.c_void_p(bufs.cqt_scratch.data_ptr()), bufs.cqt_scratch.numel(),
                                         C.c_void_p(ws.data_ptr()), ws.numel(), stream))

    for _ in range(2):
        run()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        run()
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1) / 5
    report["timing"] = {"tracks": args.tracks, "seconds": args.seconds, "ms_per_batch": ms, "ms_per_track": ms / args.tracks,
                        "scratch_MB": bufs.cqt_scratch.numel() / 1e6}
    os.makedirs("gpurun_out", exist_ok=True)
    with open("gpurun_out/cqt_report.json", "w") as f:
        json.dump(report, f, indent=1)
    print(json.dumps(report, indent=1))


if __name__ == "__main__":
    main()
