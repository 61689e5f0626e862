"""A/B of the chroma filterbank contraction: CUDA cores (TMA-fed, exact fp32) against tcgen05.mma kind::tf32 (TA_PROJECT=umma).
Usage: python tools/umma_ab.py <n_tracks> <seconds> <out.npy>   (run once per TA_PROJECT setting; compare the saved arrays)."""
import os, sys
import numpy as np, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from track_analyser_b200 import engine, synth

nt, seconds, out = int(sys.argv[1]), float(sys.argv[2]), sys.argv[3]
SR = 44_100
x = synth.synth_track(7, seconds, SR, 2)
n = x.shape[1]
plan = engine.Plan(SR, 2048, 512, 128, device=0)
pcm = torch.from_numpy(x.reshape(-1)).cuda().repeat(nt)
batch = engine.DeviceBatch(plan, pcm, np.arange(nt, dtype=np.int64) * 2 * n, np.full(nt, n, dtype=np.int64), 2)
bufs = engine.FrontendBuffers(batch, ("chroma",))
for _ in range(2):
    engine.run_device_profiled(plan, batch, bufs)
ms = float(np.mean([engine.run_device_profiled(plan, batch, bufs)[4] for _ in range(5)]))   # stage 4: chroma_stft
torch.cuda.synchronize()
res = engine.download(batch, bufs)
c = np.asarray(res[0]["chroma"])
np.save(out, c)
print(f"TA_PROJECT={os.environ.get('TA_PROJECT', 'tma')}: chroma stage (peaks + tuning + filterbank + projection) {ms:.3f} ms for {nt} x {seconds:.0f} s; "
      f"chroma[0] shape {c.shape}, max {c.max():.6f}, finite {bool(np.all(np.isfinite(c)))}")
