"""Attribute ncu warp-stall samples (SASS source page) to CUDA source lines.

usage: python tools/ncu_lines.py <report.ncu-rep> <kernel-regex> <cubin.sass from `nvdisasm -g -c`> <function substring>
The SASS page of ncu has no line numbers; nvdisasm's `//## File ..., line N` markers are
aligned with it by instruction order.
"""
import csv, re, subprocess, sys, collections

rep, kre, sass, fsub = sys.argv[1:5]
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--kernel-name", f"regex:{kre}"],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = rows[1]
ns, ie = hdr.index("# Samples"), hdr.index("Instructions Executed")
ins = [(int(r[ns] or 0), int(r[ie] or 0), r[1]) for r in rows[2:] if len(r) >= len(hdr) and r[ns].isdigit()]
# parse nvdisasm: find function, collect (line, file) per instruction
lines, cur, infunc = [], None, False
for l in open(sass):
    if l.startswith(".text.") or l.lstrip().startswith(".section\t.text."):
        infunc = fsub in l
    if not infunc:
        continue
    m = re.search(r'//## File "([^"]+)", line (\d+)', l)
    if m:
        cur = (m.group(1).split("/")[-1], int(m.group(2)))
        continue
    if re.match(r"\s+/\*[0-9a-f]{4,}\*/", l):
        lines.append(cur)
print(f"ncu instrs {len(ins)}  nvdisasm instrs {len(lines)}", file=sys.stderr)
n = min(len(ins), len(lines))
agg = collections.defaultdict(lambda: [0, 0])
for (s, e, _), loc in zip(ins[:n], lines[:n]):
    agg[loc][0] += s
    agg[loc][1] += e
ts, te = sum(v[0] for v in agg.values()), sum(v[1] for v in agg.values())
src = {}
for loc, (s, e) in sorted(agg.items(), key=lambda kv: -kv[1][0])[:int(sys.argv[5]) if len(sys.argv) > 5 else 40]:
    text = ""
    if loc:
        for base in ("track_analyser_b200/csrc/",):
            try:
                src.setdefault(loc[0], open(base + loc[0]).read().splitlines())
                text = src[loc[0]][loc[1] - 1].strip()
            except Exception:
                pass
    print(f"{str(loc):28s} {100*s/ts:5.1f}% samples {100*e/te:5.1f}% instr   {text[:100]}")
