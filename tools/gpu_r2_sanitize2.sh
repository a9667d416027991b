#!/bin/bash
# compute-sanitizer on what the end of round 2 added: kernel C (prefix_lane.cu) and the segmented bitonic compaction.
mkdir -p gpurun_out
timeout 100 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_prefix_lane_gpu.py -x -q \
  -k "ring_geometries or overflowing" > gpurun_out/memcheck_lane.log 2>&1; echo "memcheck rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_lane.log | tail -3
timeout 100 compute-sanitizer --tool racecheck --racecheck-report analysis --error-exitcode 9 python -m pytest tests/test_prefix_lane_gpu.py -x -q \
  -k "ring_geometries and 4-0" > gpurun_out/racecheck_lane.log 2>&1; echo "racecheck rc=$?"
grep -E "RACECHECK SUMMARY|hazard|passed|failed" gpurun_out/racecheck_lane.log | sort | uniq -c | sort -rn | head -6
