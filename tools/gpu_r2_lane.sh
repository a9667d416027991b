#!/bin/bash
# Merge-tree threshold sweep: K1 over k, prefix scan over k (the static is read once per process).
mkdir -p gpurun_out
for m in 16384 4096 1024; do
  export VB_MERGE_TREE_MIN=$m
  echo "tree above $m: $(timeout 300 python tools/bench_k_sweep.py 2>&1 | tail -1 | cut -c1-600)"
  timeout 300 python tools/bench_prefix_dbg.py 2>&1 | grep "debug 0"
done
