#!/usr/bin/env python
"""Turn ncu outputs brought back in gpurun_out/ into the small text summaries committed under profiles/.

  python tools/ncu_summary.py launches gpurun_out/launches.csv            > profiles/rNN_launches.md
  python tools/ncu_summary.py full     gpurun_out/prof_gemm.ncu-rep       > profiles/rNN_gemm_full.md
"""
import collections
import csv
import re
import subprocess
import sys

KEEP = [
    "gpu__time_duration.sum", "dram__bytes_read.sum", "dram__bytes_write.sum",
    "gpu__dram_throughput.avg.pct_of_peak_sustained_elapsed", "lts__throughput.avg.pct_of_peak_sustained_elapsed",
    "l1tex__m_xbar2l1tex_read_bytes.sum", "l1tex__m_xbar2l1tex_read_bytes.sum.per_second",
    "l1tex__m_xbar2l1tex_read_bytes.sum.pct_of_peak_sustained_elapsed",
    "sm__pipe_tensor_cycles_active.avg.pct_of_peak_sustained_active", "sm__throughput.avg.pct_of_peak_sustained_elapsed",
    "sm__warps_active.avg.pct_of_peak_sustained_active", "launch__registers_per_thread", "launch__grid_size",
    "launch__block_size", "launch__cluster_size", "sm__cycles_elapsed.avg", "sm__cycles_elapsed.avg.per_second",
    "smsp__cycles_active.avg", "l1tex__data_pipe_lsu_wavefronts_mem_shared.sum",
    "smsp__warp_issue_stalled_long_scoreboard_per_warp_active.pct", "smsp__warp_issue_stalled_barrier_per_warp_active.pct",
    "launch__shared_mem_per_block_dynamic",
]


def short(name):
    m = re.search(r"(\w+_kernel)(<[^>]*>)?", name)
    return (m.group(1) + (m.group(2) or "")) if m else name[:60]


def launches(path):
    rows = [r for r in csv.reader(open(path)) if len(r) > 10]
    hdr = rows[0]
    ki, vi, ui = hdr.index("Kernel Name"), hdr.index("Metric Value"), hdr.index("Metric Unit")
    agg = collections.defaultdict(lambda: [0, 0.0])
    for r in rows[1:]:
        v = float(r[vi].replace(",", ""))
        v = v / 1000 if r[ui] in ("ns", "nsecond") else v * 1000 if r[ui] in ("ms", "msecond") else v
        a = agg[short(r[ki])]
        a[0] += 1
        a[1] += v
    tot = sum(v[1] for v in agg.values())
    print(f"# ncu launch list: {path} ({sum(v[0] for v in agg.values())} launches, {tot / 1000:.2f} ms total; cold-cache, serialised "
          f"-- compare SHARES, not absolutes)\n")
    print("| kernel | launches | total us | share | avg us |\n|---|---:|---:|---:|---:|")
    for k, v in sorted(agg.items(), key=lambda kv: -kv[1][1]):
        print(f"| `{k}` | {v[0]} | {v[1]:.1f} | {100 * v[1] / tot:.1f}% | {v[1] / v[0]:.1f} |")


def full(path):
    out = subprocess.run(["ncu", "-i", path, "--page", "raw", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(out.splitlines()))
    hdr, units = rows[0], rows[1]
    print(f"# ncu --set full: {path}\n")
    for r in rows[2:]:
        d = dict(zip(hdr, r))
        u = dict(zip(hdr, units))
        print(f"## `{short(d['Kernel Name'])}`  grid {d.get('launch__grid_size')} block {d.get('launch__block_size')}\n")
        for k in KEEP:
            if k in d and d[k] != "":
                print(f"- {k} [{u[k]}] = {d[k]}")
        rd, wr = d.get("dram__bytes_read.sum"), d.get("dram__bytes_write.sum")
        if rd and wr:
            scale = {"byte": 1e-9, "Kbyte": 1e-6, "Mbyte": 1e-3, "Gbyte": 1.0, "Tbyte": 1e3}
            gb = float(rd) * scale[u["dram__bytes_read.sum"]] + float(wr) * scale[u["dram__bytes_write.sum"]]
            print(f"- traffic (dram read + write) [Gbyte] = {gb:.3f}")
        print()


if __name__ == "__main__":
    {"launches": launches, "full": full}[sys.argv[1]](sys.argv[2])
