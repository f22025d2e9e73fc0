#!/usr/bin/env python
"""Digest of a JCB_GEMM_TRACE file (clock64 stamps of CTA 0's producer / MMA / epilogue roles per tile):
median steady-state intervals in cycles per GEMM shape.   python tools/gemm_trace_report.py trace.txt"""
import re
import statistics as st
import sys

blocks, cur = {}, None
for line in open(sys.argv[1]):
    if line.startswith('#'):
        m = re.search(r'M=(\d+) N=(\d+) K=(\d+) epi=(\d+)', line)
        cur = tuple(int(m.group(i)) for i in range(1, 5))
        blocks[cur] = []          # keep the last launch of each shape
    elif cur:
        blocks[cur].append([int(x) for x in line.split()])
for key, rows in blocks.items():
    n = len(rows)
    if n < 30:
        continue
    lo, hi = 10, min(70, n - 2)

    def med(f):
        return st.median([f(i) for i in range(lo, hi)])
    kb = key[2] // 64
    print(f"M={key[0]} N={key[1]} K={key[2]} epi={key[3]}  ({n} tiles traced)")
    print(f"  tile period {med(lambda i: rows[i][1] - rows[i - 1][1]):.0f} cycles;"
          f" MMA issue loop {med(lambda i: rows[i][2] - rows[i][1]):.0f} ({med(lambda i: rows[i][2] - rows[i][1]) / (kb * 4):.1f} per UMMA);"
          f" MMA idle between tiles {med(lambda i: rows[i][1] - rows[i - 1][2]):.0f}")
    print(f"  epilogue: accumulator visible {med(lambda i: rows[i][3] - rows[i][2]):.0f} after the commit is issued;"
          f" TMEM held {med(lambda i: rows[i][4] - rows[i][3]):.0f}; busy {med(lambda i: rows[i][5] - rows[i][3]):.0f};"
          f" idle until the next accumulator {med(lambda i: rows[i + 1][3] - rows[i][5]):.0f}")
    print(f"  producer: first load of a tile {med(lambda i: rows[i][1] - rows[i][6]):.0f} cycles ahead of its first MMA;"
          f" MMA start after the stage was released {med(lambda i: rows[i][1] - rows[i - 2][4]):.0f}")
