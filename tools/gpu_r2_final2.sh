#!/bin/bash
# Closing pass after the lane-per-row prefix kernel, the segmented bitonic compaction and the merge-tree threshold:
# whole GPU suite, smoke, default bench line, funnel timings, one full capture of the lane kernel.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 200 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log | cut -c1-200
timeout 900 python bench.py > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -c 300 gpurun_out/bench_n1.log
timeout 200 python tools/bench_funnel.py --stages 128,384 --candidates 100 --iters 100 > gpurun_out/funnel_lane.log 2>&1; tail -1 gpurun_out/funnel_lane.log | cut -c150-420
VB_SCAN_NO_LANE=1 timeout 200 python tools/bench_funnel.py --stages 128,384 --candidates 100 --iters 100 > gpurun_out/funnel_warp.log 2>&1; tail -1 gpurun_out/funnel_warp.log | cut -c150-420
timeout 200 ncu --set full --clock-control none --import-source on -k regex:prefix_lane -s 4 -c 1 -f -o gpurun_out/r2_prefix_lane \
  python tools/bench_funnel.py --stages 128,384 --candidates 100 --iters 3 > gpurun_out/lane_ncu.log 2>&1
ls -la gpurun_out/r2_prefix_lane.ncu-rep
