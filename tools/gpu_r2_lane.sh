#!/bin/bash
# Kernel C: checkpoint cadence sweep (rounds between collector checkpoints), parity first.
mkdir -p gpurun_out
timeout 200 python -m pytest tests/test_prefix_lane_gpu.py -x -q 2>&1 | tail -2
for r in 1 2 4 8; do
  echo "sync rounds $r: $(VB_LANE_SYNC_ROUNDS=$r timeout 100 python tools/bench_funnel.py --stages 128,384 --candidates 100 --iters 200 2>&1 | tail -1 | cut -c150-330)"
done
