"""Device times of the BASELINE.json configs that are not the bench line (single tracks and the 4096 sweep).

    python tools/configs_bench.py            # prints one JSON line per config
Inputs are resident in HBM; times are CUDA-event averages over `reps` runs after two warm-ups.
"""
import json
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_analyser_b200 import engine, synth  # noqa: E402


def timed(plan, batch, outs, reps=5):
    bufs = engine.FrontendBuffers(batch, outs)
    for _ in range(2):
        engine.run_device(plan, batch, bufs)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(reps):
        engine.run_device(plan, batch, bufs)
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / reps


def resident(plan, x, nt):
    n = x.shape[-1]
    ch = 1 if x.ndim == 1 else x.shape[0]
    stride = (ch * n + 3) & ~3
    pcm = torch.zeros(nt * stride, dtype=torch.float32, device="cuda")
    src = torch.from_numpy(np.ascontiguousarray(x).reshape(-1)).cuda()
    for i in range(nt):
        pcm[i * stride: i * stride + ch * n].copy_(src)
    return engine.DeviceBatch(plan, pcm, np.arange(nt, dtype=np.int64) * stride, np.full(nt, n, dtype=np.int64), ch)


def main():
    out = []
    # configs[0]: tiny click, 44.1 kHz mono, N = 89 523
    g = np.load(os.path.join(os.path.dirname(__file__), "..", "tests", "golden", "tiny_click.npz"))
    plan = engine.Plan(44_100, 2048, 512, 128, device=0)
    ms = timed(plan, resident(plan, g["samples"], 1), engine.FRONTEND_OUTPUTS, 20)
    out.append({"config": 0, "what": "tiny_click_120 (2.03 s mono), full frontend", "ms": ms, "xrt": 2.03 / (ms * 1e-3)})
    # configs[1]: one 3-minute 44.1 kHz stereo track
    x = synth.synth_track(synth.DEFAULT_SEED, 180.0, 44_100, 2)
    ms = timed(plan, resident(plan, x, 1), engine.FRONTEND_OUTPUTS, 10)
    out.append({"config": 1, "what": "one 3-min 44.1 kHz stereo track, full frontend", "ms": ms, "xrt": 180.0 / (ms * 1e-3)})
    ms = timed(plan, resident(plan, x, 1), engine.available_outputs(plan, engine.CORE_OUTPUTS + ('tempogram',)), 10)
    out.append({"config": 1, "what": "same + true peak + HPSS curves", "ms": ms, "xrt": 180.0 / (ms * 1e-3)})
    # configs[3]: one 60-minute 48 kHz stereo track
    plan48 = engine.Plan(48_000, 2048, 512, 128, device=0)
    base = synth.synth_track(7, 60.0, 48_000, 2)
    long = np.concatenate([base] * 60, axis=1)
    ms = timed(plan48, resident(plan48, long, 1), engine.FRONTEND_OUTPUTS, 3)
    out.append({"config": 3, "what": "one 60-min 48 kHz stereo track, full frontend", "ms": ms, "xrt": 3600.0 / (ms * 1e-3)})
    del long
    torch.cuda.empty_cache()
    # configs[4]: 256 tracks, n_fft 4096, hop 256, 256 mels + chroma (60 s each keeps the magnitude at 22 GB)
    plan5 = engine.Plan(44_100, 4096, 256, 256, device=0)
    x5 = synth.synth_track(5, 60.0, 44_100, 2)
    outs5 = tuple(o for o in engine.FRONTEND_OUTPUTS if o != "tempogram")
    ms = timed(plan5, resident(plan5, x5, 256), outs5, 3)
    out.append({"config": 4, "what": "256 x 60 s stereo, n_fft 4096 hop 256 256 mels + chroma (no tempogram)", "ms": ms,
                "xrt": 256 * 60.0 / (ms * 1e-3)})
    for o in out:
        print(json.dumps(o))


if __name__ == "__main__":
    main()
