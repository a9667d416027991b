#!/bin/bash
# Launch lists (ncu gpu__time_duration) of one 1024-query batch: full kernel vs filter skipped (VB_GEMM_DEBUG=8).
mkdir -p gpurun_out
for dbg in 0 8; do
  VB_GEMM_DEBUG=$dbg timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 60 --csv --log-file gpurun_out/k2_launches_$dbg.csv \
    python tools/bench_batch.py --steps 1 --no-check > gpurun_out/k2_list_$dbg.log 2>&1
  python - <<PY
import csv
rows = list(csv.reader(open("gpurun_out/k2_launches_$dbg.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
seq = [(r[h.index("Kernel Name")][:50], float(r[h.index("Metric Value")].replace(",", "")) / 1000) for r in rows[hdr + 2:] if len(r) >= len(h)]
print("debug $dbg: last 12 launches")
for n, t in seq[-12:]:
    print(f"{t:10.1f} us  {n}")
PY
done
