#!/bin/bash
# ncu capture of the tensor-core MaxSim kernel (200k docs x 128 x 128, 32-token query)
mkdir -p gpurun_out
timeout 300 ncu --set full --clock-control none --import-source on -k regex:maxsim_tc -s 3 -c 1 -f -o gpurun_out/maxsim_tc \
  python tools/bench_maxsim.py --steps 6 --check 0 > gpurun_out/maxsim_tc.log 2>&1
ls -la gpurun_out/maxsim_tc.ncu-rep
