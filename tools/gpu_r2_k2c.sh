#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_batch_gpu.py -m gpu -x -q > gpurun_out/pytest_k2.log 2>&1; tail -3 gpurun_out/pytest_k2.log
for smp in 9472 18944 32768 65536; do
  VB_GEMM_SAMPLE=$smp timeout 300 python tools/bench_batch.py --steps 5 > gpurun_out/k2_s$smp.log 2>&1; echo "sample $smp: $(tail -1 gpurun_out/k2_s$smp.log | cut -c300-420)"
done
KF='regex:gemm|queries|unpack|merge'
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 60 --csv --log-file gpurun_out/r2_k2_launches.csv \
  python tools/bench_batch.py --steps 2 > gpurun_out/k2_ncu_list.log 2>&1
python - <<'PY'
import csv
rows = list(csv.reader(open("gpurun_out/r2_k2_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
for r in rows[hdr + 2:][-8:]:
    if len(r) >= len(h): print(f'{float(r[h.index("Metric Value")].replace(",", "")) / 1000:9.2f} us  {r[h.index("Kernel Name")][:70]}')
PY
