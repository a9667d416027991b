#!/bin/bash
# Round-2 third pass (one GPU): new tests, funnel bench with / without the dense prefix mirror, launch lists
# (ncu) of the quantized pipeline shard and the headline bench.
mkdir -p gpurun_out
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -15 gpurun_out/pytest_gpu.log
for st in "128,384" "64,256"; do
  timeout 300 python tools/bench_funnel.py --stages $st --candidates 100 --iters 50 > gpurun_out/funnel_mirror_$st.log 2>&1; tail -1 gpurun_out/funnel_mirror_$st.log
  VB_NO_PREFIX_MIRROR=1 timeout 300 python tools/bench_funnel.py --stages $st --candidates 100 --iters 50 > gpurun_out/funnel_nomirror_$st.log 2>&1; tail -1 gpurun_out/funnel_nomirror_$st.log
done
# launch list of the quantized pipeline on a 12.5M x 1024 shard (C4 at G=8): where do the ~95 us above the scan go?
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 300 --csv --log-file gpurun_out/r2_c4_launches.csv \
  python tools/bench_scale.py --mode quantized --rows 12500000 --dim 1024 > gpurun_out/c4_ncu.log 2>&1
tail -2 gpurun_out/c4_ncu.log | cut -c1-600
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_c4_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
agg = collections.OrderedDict()
for r in rows[hdr + 2:]:
    if len(r) < len(h): continue
    name = r[h.index("Kernel Name")][:70]
    v = float(r[h.index("Metric Value")].replace(",", ""))
    a = agg.setdefault(name, [0, 0.0]); a[0] += 1; a[1] += v
for n, (c, t) in agg.items():
    print(f"{c:4d} x {t / c / 1000:9.2f} us  {n}")
PY
