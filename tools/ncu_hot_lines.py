#!/usr/bin/env python
"""Per-source-line stall samples of one kernel in an .ncu-rep (the CSV source page is SASS-only):
joins `ncu --page source --csv` (samples per SASS address) with `nvdisasm -g` line info of the object file.
usage: ncu_hot_lines.py REPORT.ncu-rep OBJECT.o KERNEL_SUBSTRING [top]"""
import csv
import io
import re
import subprocess
import sys
import tempfile
import os
import collections


def main():
    rep, obj, kern = sys.argv[1:4]
    top = int(sys.argv[4]) if len(sys.argv) > 4 else 40
    src = subprocess.run(["ncu", "-i", rep, "--page", "source", "--csv"], capture_output=True, text=True).stdout
    rows = list(csv.reader(io.StringIO(src)))
    hdr = next(i for i, r in enumerate(rows) if r and r[0] == "Address")
    h = rows[hdr]
    ci = {n: i for i, n in enumerate(h)}
    body = [r for r in rows[hdr + 1:] if len(r) >= len(h) - 2 and r[0].startswith("0x")]
    base = int(body[0][0], 16)
    samples = {}
    for r in body:
        samples[int(r[0], 16) - base] = (int(r[ci["# Samples"]] or 0), r[ci["Source"]].strip(), r)
    with tempfile.TemporaryDirectory() as td:
        subprocess.run(["cuobjdump", "-xelf", "all", os.path.abspath(obj)], cwd=td, capture_output=True)
        cubin = [f for f in os.listdir(td) if f.endswith(".cubin")][0]
        dis = subprocess.run(["nvdisasm", "-g", "-c", os.path.join(td, cubin)], capture_output=True, text=True).stdout
    line, infunc = None, False
    per_line = collections.Counter()
    per_line_stall = collections.defaultdict(collections.Counter)
    stall_cols = [n for n in h if n.startswith("stall_") and "Not Issued" not in n]
    for l in dis.splitlines():
        if l.startswith("\t.section") or l.startswith(".text."):
            infunc = kern in l
        if not infunc:
            continue
        m = re.search(r'//## File "([^"]+)", line (\d+)', l)
        if m:
            line = (os.path.basename(m.group(1)), int(m.group(2)))
            continue
        m = re.match(r"\s+/\*([0-9a-f]+)\*/", l)
        if m and line is not None:
            off = int(m.group(1), 16)
            if off in samples:
                n, _, r = samples[off]
                per_line[line] += n
                for sc in stall_cols:
                    v = int(r[ci[sc]] or 0)
                    if v:
                        per_line_stall[line][sc] += v
    total = sum(per_line.values())
    print(f"total samples {total}")
    for (f, ln), n in per_line.most_common(top):
        st = ", ".join(f"{k[6:]} {v}" for k, v in per_line_stall[(f, ln)].most_common(3))
        print(f"{n:7d} {100.0 * n / max(total, 1):5.1f}%  {f}:{ln}  [{st}]")


if __name__ == "__main__":
    main()
