#!/usr/bin/env python
"""Per-source-line instruction counts of one kernel from an ncu report captured with --import-source on.
Usage: python scripts/ncu_lines.py report.ncu-rep [top_n]   (container: ncu reads the report, no GPU needed)"""
import csv
import subprocess
import sys

rep, top = sys.argv[1], int(sys.argv[2]) if len(sys.argv) > 2 else 60
txt = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
rows = list(csv.reader(txt.splitlines()))
fname, hdr, out, tot, tot_s = "", None, [], 0, 0
for r in rows:
    if not r:
        continue
    if r[0] == "File Path":
        fname = r[1].split("/")[-1]
        continue
    if r[0] == "Line No":
        hdr = r
        ii, si = hdr.index("Instructions Executed"), hdr.index("# Samples")
        continue
    if hdr is None or len(r) <= ii or r[2] != "-":       # only the per-CUDA-line summary rows (Address == "-")
        continue
    try:
        n, s = int(float(r[ii])), int(float(r[si]))
    except ValueError:
        continue
    if n or s:
        out.append((n, s, fname, r[0], r[1].strip()[:120]))
        tot += n
        tot_s += s
print(f"total warp instructions {tot}, samples {tot_s}")
for n, s, f, ln, src in sorted(out, reverse=True)[:top]:
    print(f"{n:9d} {100 * n / max(tot, 1):5.1f}%  smp {100 * s / max(tot_s, 1):5.1f}%  {f}:{ln}: {src}")
