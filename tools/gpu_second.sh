#!/bin/bash
set -x
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -5 gpurun_out/pytest_gpu.log
run() { # name, env...
  name=$1; shift
  env "$@" timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_$name.log 2>&1
  python - <<PY
import json
try:
    l = json.loads(open("gpurun_out/bench_$name.log").read().strip().splitlines()[-1])
    print("$name", "qps=%.1f" % l["value"], "e2e=%.1f" % l["e2e"]["value"], "GB/s=%.0f" % l["roofline"]["achieved"], "frac=%.3f" % l["roofline"]["frac"], "kernel_ms=%.4f" % l["roofline"]["kernel_ms"])
except Exception as e:
    print("$name FAILED", e); print(open("gpurun_out/bench_$name.log").read()[-2000:])
PY
}
run stream X=1
run stream_s4 VB_STREAM_STAGES=4
run stream_s6 VB_STREAM_STAGES=6
run stream_rpw2 VB_STREAM_RPW=2
run ldg VB_SCAN_NO_STREAM=1
run ldg_r1 VB_SCAN_NO_STREAM=1 VB_SCAN_R=1
run ldg_r4 VB_SCAN_NO_STREAM=1 VB_SCAN_R=4
