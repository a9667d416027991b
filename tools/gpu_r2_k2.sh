#!/bin/bash
# Round-2: K2 single TF32 pass for every eligible k (k' = 256 above k = 32): batch tests, C2/C3 blocks of bench.py.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_batch_gpu.py tests/test_bench_shapes_gpu.py tests/test_sharded_gpu.py -x -q > gpurun_out/pytest_k2.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_k2.log
tail -8 gpurun_out/pytest_k2.log
timeout 900 python bench.py --steps 100 --warmup 10 --configs c2,c3 --no-cpu-baseline > gpurun_out/bench_k2.log 2>&1
python - <<'PY'
import json
for l in open("gpurun_out/bench_k2.log"):
    if l.startswith("{"):
        d = json.loads(l)
        for c in ("c2_batch_1024", "c3"):
            b = d["config"].get(c)
            if b: print(c, {k: b[k] for k in b if k in ("ms_per_batch", "local_scan_ms", "queries_per_sec", "tf32_passes", "parity", "kernel")}, b["roofline"]["frac"])
    else:
        print(l.rstrip()[:300])
PY
