#!/bin/bash
# Pre-pass sample sweep on the C3 shard (12.5M x 768, 1024 queries, k = 100) and on C2 (1M rows, k = 10 / 100).
mkdir -p gpurun_out
timeout 300 python -m pytest tests/test_batch_gpu.py -x -q 2>&1 | tail -2
for s in 0 37888 75776 151552 303104; do
  if [ $s = 0 ]; then unset VB_GEMM_SAMPLE; else export VB_GEMM_SAMPLE=$s; fi
  timeout 600 python tools/bench_scale.py --mode batch --rows 12500000 --steps 3 > gpurun_out/c3_sample_$s.log 2>&1
  python - <<PY
import json
d = json.loads(open("gpurun_out/c3_sample_$s.log").read().strip().splitlines()[-1])
print("c3 sample $s", {k: round(v["device_ms"], 2) for k, v in d["batch"].items() if isinstance(v, dict)})
PY
done
for s in 0 18944 37888 75776; do
  if [ $s = 0 ]; then unset VB_GEMM_SAMPLE; else export VB_GEMM_SAMPLE=$s; fi
  for k in 10 100; do
    echo "c2 sample $s k $k: $(timeout 300 python tools/bench_batch.py --k $k --steps 10 2>&1 | tail -1 | cut -c1-300)"
  done
done
