#!/bin/bash
# K2 single-pass kernel: self-test, parity, timings in both modes; merge-tree rank merge check via the bench.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_tc_gpu.py tests/test_batch_gpu.py -m gpu -x -q > gpurun_out/pytest_k2.log 2>&1; tail -5 gpurun_out/pytest_k2.log
timeout 600 python -m pytest tests/test_bench_shapes_gpu.py tests/test_flat_gpu.py tests/test_pipelines_gpu.py -m gpu -x -q > gpurun_out/pytest_k2b.log 2>&1; tail -3 gpurun_out/pytest_k2b.log
for t in 1 3; do
  VB_GEMM_TERMS=$t timeout 300 python tools/bench_batch.py --steps 5 > gpurun_out/k2_terms$t.log 2>&1; tail -1 gpurun_out/k2_terms$t.log | cut -c1-500
done
timeout 900 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_n1_b.log 2>&1
python - <<'PY'
import json
l = json.loads([x for x in open("gpurun_out/bench_n1_b.log") if x.startswith("{")][-1])
c = l["config"]
print("headline", round(l["value"], 1), "e2e", round(l["e2e"]["value"], 1), "frac", round(l["roofline"]["frac"], 3))
print("c2", c["c2_batch_1024"]["ms_per_batch"], c["c2_batch_1024"]["queries_per_sec"])
print("c3", c["c3"].get("local_scan_ms"), c["c3"].get("error"))
print("c4", c["c4"].get("hamming_pass", {}).get("local_scan_ms"), c["c4"].get("pipeline", {}).get("step_ms"), c["c4"].get("error"))
print("c5", c["c5"].get("local_scan_ms"), c["c5"].get("error"))
PY
