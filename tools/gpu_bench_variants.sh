#!/bin/bash
# usage: tools/gpu_bench_variants.sh "name ENV=.. ENV=.." ...   (each arg = one bench run)
mkdir -p gpurun_out
if [ -z "$SKIP_TESTS" ]; then
  timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
  tail -4 gpurun_out/pytest_gpu.log
fi
for spec in "$@"; do
  set -- $spec; name=$1; shift
  env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_$name.log 2>&1
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/bench_$name.log").read().strip().splitlines()[-1])
    print("$name", "qps=%.1f" % l["value"], "e2e=%.1f" % l["e2e"]["value"], "GB/s=%.0f" % l["roofline"]["achieved"], "frac=%.3f" % l["roofline"]["frac"], "kernel_ms=%.4f" % l["roofline"]["kernel_ms"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_$name.log").read()[-1500:])
PY
done
