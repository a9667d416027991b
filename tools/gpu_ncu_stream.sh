#!/bin/bash
mkdir -p gpurun_out
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_stream -s 10 -c 1 -f -o gpurun_out/prof_stream2 python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full2.log 2>&1
tail -2 gpurun_out/ncu_full2.log
timeout 300 ncu --set full --clock-control none -k regex:read_tma -s 14 -c 1 -f -o gpurun_out/prof_probe ./build/bw_probe > gpurun_out/ncu_probe.log 2>&1
tail -2 gpurun_out/ncu_probe.log
