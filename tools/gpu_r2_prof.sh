#!/bin/bash
# Round-2 profiling pass (one GPU): launch list of the quantized pipeline shard, full captures of K2 (1024 queries)
# and of funnel stage 1 over the dense prefix mirror, funnel bench for a float metric with / without the mirror.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_muvera.py tests/test_nif_shim.py tests/test_sharded_gpu.py -m gpu -x -q > gpurun_out/pytest_new.log 2>&1; tail -3 gpurun_out/pytest_new.log
for m in inner_product l2; do
  timeout 300 python tools/bench_funnel.py --metric $m --stages 128,384 --candidates 100 --iters 50 > gpurun_out/funnel_${m}_mirror.log 2>&1; tail -1 gpurun_out/funnel_${m}_mirror.log | cut -c1-420
  VB_NO_PREFIX_MIRROR=1 timeout 300 python tools/bench_funnel.py --metric $m --stages 128,384 --candidates 100 --iters 50 > gpurun_out/funnel_${m}_nomirror.log 2>&1; tail -1 gpurun_out/funnel_${m}_nomirror.log | cut -c1-420
done
KF='regex:hamming|topk|sign_pack|unpack|extract|flat_s|arm_ctrl|collect|merge|peer'
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -k "$KF" -c 200 --csv --log-file gpurun_out/r2_c4_launches.csv \
  python tools/bench_scale.py --mode quantized --rows 12500000 --dim 1024 > gpurun_out/c4_ncu.log 2>&1
python - <<'PY'
import csv, collections
rows = list(csv.reader(open("gpurun_out/r2_c4_launches.csv")))
hdr = [i for i, r in enumerate(rows) if r and r[0] == "ID"][0]
h = rows[hdr]
seq = []
for r in rows[hdr + 2:]:
    if len(r) < len(h): continue
    seq.append((r[h.index("Kernel Name")][:60], float(r[h.index("Metric Value")].replace(",", "")) / 1000))
# one query = the repeating tail of the sequence: print the last 14 launches
for n, t in seq[-14:]:
    print(f"{t:9.2f} us  {n}")
PY
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_gemm_topk -s 4 -c 1 -f -o gpurun_out/r2_flat_gemm \
  python tools/bench_batch.py --steps 2 > gpurun_out/k2_ncu.log 2>&1; tail -1 gpurun_out/k2_ncu.log | cut -c1-300
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_stream -s 2 -c 1 -f -o gpurun_out/r2_funnel_stage1 \
  python tools/bench_funnel.py --stages 128,384 --candidates 100 --iters 3 > gpurun_out/funnel_ncu.log 2>&1; tail -1 gpurun_out/funnel_ncu.log | cut -c1-200
ls -la gpurun_out/*.ncu-rep
