#!/bin/bash
# compute-sanitizer racecheck (shared-memory hazards) on the ragged MaxSim kernel and the single-pass K2 kernel, small cases.
mkdir -p gpurun_out
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_maxsim_ragged_gpu.py -x -q \
  -k "long-120 and inner or one_document" > gpurun_out/racecheck_tcr.log 2>&1; echo "tcr rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/racecheck_tcr.log | sort | uniq -c | sort -rn | head -8
timeout 1500 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_batch_gpu.py -x -q \
  -k "ties_resolve and auto" > gpurun_out/racecheck_k2.log 2>&1; echo "k2 rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/racecheck_k2.log | sort | uniq -c | sort -rn | head -8
