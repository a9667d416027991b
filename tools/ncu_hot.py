#!/usr/bin/env python
"""Top SASS instructions by warp-stall samples from an .ncu-rep (source page).
usage: python tools/ncu_hot.py file.ncu-rep [N]"""
import csv, subprocess, sys
path = sys.argv[1]; topn = int(sys.argv[2]) if len(sys.argv) > 2 else 30
raw = subprocess.run(["ncu", "-i", path, "--page", "source", "--csv"], capture_output=True, text=True).stdout
rows = list(csv.reader(raw.splitlines()))
hi = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
hdr = rows[hi]
s_all = hdr.index("Warp Stall Sampling (All Samples)")
src = hdr.index("Source")
ex = hdr.index("Instructions Executed")
body = [r for r in rows[hi + 1:] if len(r) > s_all and r[s_all].isdigit()]
total = sum(int(r[s_all]) for r in body)
print(f"total samples {total}")
idx = {id(r): i for i, r in enumerate(body)}
for r in sorted(body, key=lambda r: -int(r[s_all]))[:topn]:
    print(f"{int(r[s_all]):8d} {100*int(r[s_all])/total:5.1f}%  exec={r[ex]:>9s}  #{idx[id(r)]:5d} {r[src].strip()[:100]}")
