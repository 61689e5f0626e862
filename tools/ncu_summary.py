"""Summarise an ncu report (raw page) per kernel: time, DRAM traffic, pipe use, stalls.

    python tools/ncu_summary.py gpurun_out/prof.ncu-rep [--md]
"""
import csv
import subprocess
import sys

WANT = [
    ("gpu__time_duration.sum", "us", 1e-3),
    ("dram__bytes_read.sum", "dram_rd_MB", None),
    ("dram__bytes_write.sum", "dram_wr_MB", None),
    ("launch__registers_per_thread", "regs", None),
    ("sm__warps_active.avg.pct_of_peak_sustained_active", "occ%", None),
    ("smsp__issue_active.avg.pct_of_peak_sustained_active", "issue%", None),
    ("sm__pipe_fma_cycles_active.avg.pct_of_peak_sustained_active", "fma%", None),
    ("sm__pipe_fp64_cycles_active.avg.pct_of_peak_sustained_active", "fp64%", None),
    ("sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "tensor%", None),
    ("sm__inst_executed_pipe_lsu.avg.pct_of_peak_sustained_active", "lsu%", None),
    ("l1tex__data_pipe_lsu_wavefronts_mem_shared.avg.pct_of_peak_sustained_elapsed", "smem_wf%", None),
    ("gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "dram%", None),
    ("smsp__inst_executed.sum", "warp_inst", None),
    ("l1tex__data_bank_conflicts_pipe_lsu_mem_shared.sum", "bank_conf", None),
]
STALLS = "smsp__average_warps_issue_stalled_{}_per_issue_active.ratio"
STALL_NAMES = ["long_scoreboard", "short_scoreboard", "wait", "barrier", "math_pipe_throttle", "mio_throttle", "lg_throttle",
               "no_instruction", "not_selected", "dispatch_stall", "branch_resolving", "membar", "selected"]


def main():
    rep = sys.argv[1]
    out = subprocess.run(["ncu", "-i", rep, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    col = {h: i for i, h in enumerate(hdr)}
    for r in rows[2:]:
        name = r[col["Kernel Name"]][:60]
        parts = []
        for key, label, _ in WANT:
            if key in col:
                v = r[col[key]]
                u = units[col[key]]
                try:
                    f = float(v.replace(",", ""))
                    if label == "us":
                        f = f / 1e3 if u in ("ns", "nsecond") else (f if u in ("us", "usecond") else f * 1e3 if u in ("ms", "msecond") else f)
                    if label.endswith("_MB"):
                        f = {"byte": f / 1e6, "Kbyte": f / 1e3, "Mbyte": f, "Gbyte": f * 1e3}.get(u, f)
                    parts.append(f"{label}={f:.4g}")
                except ValueError:
                    parts.append(f"{label}={v}")
        st = []
        for s in STALL_NAMES:
            k = STALLS.format(s)
            if k in col:
                try:
                    st.append((float(r[col[k]]), s))
                except ValueError:
                    pass
        st.sort(reverse=True)
        tot = sum(v for v, _ in st) or 1.0
        parts.append("stalls: " + " ".join(f"{s}={100 * v / tot:.0f}%" for v, s in st[:6]))
        print(name, "|", " ".join(parts))


if __name__ == "__main__":
    main()
