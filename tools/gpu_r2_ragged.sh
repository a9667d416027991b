#!/bin/bash
# Round-2: ragged tensor-core MaxSim — parity tests, then the kernel on corpora of different length profiles.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_maxsim_ragged_gpu.py tests/test_maxsim_gpu.py -x -q > gpurun_out/pytest_ragged.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ragged.log
tail -4 gpurun_out/pytest_ragged.log
for spec in "200000 40,180 32" "1000000 5,40 32" "4000000 1,8 32" "172000 128,128 32" "200000 40,180 64"; do
  set -- $spec
  VB_MAXSIM_NO_TCU=1 timeout 300 python tools/bench_maxsim.py --docs $1 --ragged $2 --tq $3 > gpurun_out/maxsim_ragged_$2_$3.log 2>&1
  python - <<PY
import json
try:
    d = json.loads(open("gpurun_out/maxsim_ragged_$2_$3.log").read().strip().splitlines()[-1])
    print("$spec", d["config"]["kernel"], round(d["ms_per_query"], 3), "ms", round(d["roofline"]["frac"], 3))
except Exception as e:
    print("$spec failed", open("gpurun_out/maxsim_ragged_$2_$3.log").read()[-300:])
PY
done
