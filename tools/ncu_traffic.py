"""Extract per-launch DRAM traffic of one kernel from an ncu report into a small JSON that bench.py reads.

    python tools/ncu_traffic.py gpurun_out/prof.ncu-rep stft_fused profiles/r1_k1_traffic.json "<command that was profiled>"
"""
import csv
import json
import subprocess
import sys

rep, pattern, out, cmd = sys.argv[1:5]
raw = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hdr, units = rows[0], rows[1]
col = {h: i for i, h in enumerate(hdr)}
scale = {"byte": 1.0, "Kbyte": 1e3, "Mbyte": 1e6, "Gbyte": 1e9}
best = None
for r in rows[2:]:
    if pattern not in r[col["Kernel Name"]]:
        continue
    rd = float(r[col["dram__bytes_read.sum"]]) * scale[units[col["dram__bytes_read.sum"]]]
    wr = float(r[col["dram__bytes_write.sum"]]) * scale[units[col["dram__bytes_write.sum"]]]
    dur = float(r[col["gpu__time_duration.sum"]])
    best = {"kernel": r[col["Kernel Name"]], "dram_read_bytes": rd, "dram_write_bytes": wr, "traffic_bytes": rd + wr,
            "duration": dur, "duration_unit": units[col["gpu__time_duration.sum"]], "command": cmd, "report": rep}
json.dump(best, open(out, "w"), indent=1)
print(best)
