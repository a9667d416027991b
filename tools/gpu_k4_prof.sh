#!/bin/bash
# ncu captures of the K4 prefix-scoring kernel (funnel stage 1 over all rows): d=128 of 768, 100 candidates
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:flat_scan_kernel -s 0 -c 4 -f -o gpurun_out/k4_prefix \
  python tools/bench_funnel.py --stages 128,384 --candidates 100 --iters 3 > gpurun_out/k4_prefix.log 2>&1
ls -la gpurun_out/k4_prefix.ncu-rep
