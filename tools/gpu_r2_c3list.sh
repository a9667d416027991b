#!/bin/bash
# Launch list (ncu gpu__time_duration) of one C3-shard batch: 12.5M x 768, 1024 queries, k = 100.
mkdir -p gpurun_out
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k regex:'gemm|rescore|pack_queries|sample_select' -c 40 --csv --log-file gpurun_out/c3_launches.csv \
  python tools/bench_scale.py --mode batch --rows 12500000 --steps 1 > gpurun_out/c3_list.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/c3_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
seq = [(r[h.index("Kernel Name")][:60], float(r[h.index("Metric Value")].replace(",", "")) / 1000) for r in rows[hdr + 2:] if len(r) >= len(h)]
for n, t in seq[:14]:
    print(f"{t:10.1f} us  {n}")
PY
