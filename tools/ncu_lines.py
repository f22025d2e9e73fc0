#!/usr/bin/env python
"""Per-source-line stall samples / instruction counts of one kernel from an .ncu-rep (needs -lineinfo + --import-source on).
    python tools/ncu_lines.py gpurun_out/prof.ncu-rep mta_kernel [top_n]"""
import csv
import subprocess
import sys

rep, kernel = sys.argv[1], sys.argv[2]
top = int(sys.argv[3]) if len(sys.argv) > 3 else 25
out = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass", "--kernel-name", kernel],
                     capture_output=True, text=True).stdout
rows = list(csv.reader(out.splitlines()))
hdr = None
lines = []
fname = ""
for r in rows:
    if len(r) == 2 and r[0] == "File Path":
        fname = r[1].split("/")[-1]
    if len(r) > 8 and r[0] == "Line No":
        hdr = r
        continue
    if hdr and len(r) == len(hdr) and r[0] not in ("", "Line No"):
        try:
            lines.append((fname, int(r[0]), r[1], int(r[hdr.index("# Samples")]), int(r[hdr.index("Instructions Executed")]),
                          int(r[hdr.index("L1 Wavefronts Shared")] or 0)))
        except ValueError:
            pass
tot = sum(l[3] for l in lines) or 1
tot_i = sum(l[4] for l in lines) or 1
print(f"{kernel}: {tot} samples, {tot_i} warp instructions")
for f, n, src, s, i, w in sorted(lines, key=lambda l: -l[3])[:top]:
    print(f"{100 * s / tot:5.1f}% samples {100 * i / tot_i:5.1f}% instr  smem_wf {w:>10}  {f}:{n}: {src.strip()[:110]}")
