#!/bin/bash
# Round-2: full ncu capture of the single-pass K2 kernel (1024 queries x 1M x 768) + pre-pass sample sweep on the C3 shard.
mkdir -p gpurun_out
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_gemm1_topk -s 3 -c 1 -f -o gpurun_out/r2_flat_gemm1 \
  python tools/bench_batch.py --steps 2 > gpurun_out/k2_ncu.log 2>&1; tail -1 gpurun_out/k2_ncu.log | cut -c1-300
for s in 9472 37888 151552; do
  VB_GEMM_SAMPLE=$s timeout 600 python tools/bench_scale.py --mode batch --rows 12500000 --steps 3 > gpurun_out/c3_sample_$s.log 2>&1
  python - <<PY
import json
d = json.loads(open("gpurun_out/c3_sample_$s.log").read().strip().splitlines()[-1])
print("sample $s", {k: round(v["device_ms"], 2) for k, v in d["batch"].items() if isinstance(v, dict)})
PY
done
