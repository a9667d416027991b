#!/usr/bin/env python
"""Warp-stall samples per CUDA source line from an .ncu-rep captured with --import-source on (cuda,sass view).
usage: python tools/ncu_hot_cuda.py file.ncu-rep [N]"""
import csv, subprocess, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv", "--print-source", "cuda,sass"], capture_output=True, text=True).stdout
cur_file = "?"
agg = []
hdr = None
for r in csv.reader(raw.splitlines()):
    if not r:
        continue
    if r[0] == "File Path":
        cur_file = r[1].split("/")[-1]; continue
    if r[0] == "Line No":
        hdr = r; s_all = hdr.index("Warp Stall Sampling (All Samples)"); ex = hdr.index("Instructions Executed"); continue
    if hdr is None or len(r) <= s_all:
        continue
    if r[0] != "" and r[s_all].isdigit():          # a CUDA line row (aggregated over its SASS)
        agg.append((int(r[s_all]), cur_file, r[0], r[1].strip(), r[ex]))
total = sum(a[0] for a in agg)
print("total samples", total)
for n, f, line, src, e in sorted(agg, reverse=True)[:topn]:
    print(f"{n:8d} {100 * n / max(total, 1):5.1f}%  exec={e:>10s}  {f}:{line:>4s}  {src[:100]}")
