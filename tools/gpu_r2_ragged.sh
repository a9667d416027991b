#!/bin/bash
# Round-2: ragged tensor-core MaxSim — parity tests, the ragged / uniform / general kernels on comparable corpora,
# one full ncu capture of the ragged kernel.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_maxsim_ragged_gpu.py tests/test_maxsim_gpu.py -x -q > gpurun_out/pytest_ragged.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ragged.log
tail -25 gpurun_out/pytest_ragged.log
timeout 300 python tools/bench_maxsim.py --docs 200000 --ragged 40,180 > gpurun_out/maxsim_ragged_40_180.log 2>&1; tail -1 gpurun_out/maxsim_ragged_40_180.log
VB_MAXSIM_NO_TCU=1 timeout 300 python tools/bench_maxsim.py --docs 172000 --ragged 128,128 > gpurun_out/maxsim_ragged_128.log 2>&1; tail -1 gpurun_out/maxsim_ragged_128.log
timeout 300 python tools/bench_maxsim.py --docs 200000 --ragged 40,180 --tq 64 > gpurun_out/maxsim_ragged_tq64.log 2>&1; tail -1 gpurun_out/maxsim_ragged_tq64.log
timeout 300 python tools/bench_maxsim.py --docs 1000000 --ragged 5,40 > gpurun_out/maxsim_ragged_5_40.log 2>&1; tail -1 gpurun_out/maxsim_ragged_5_40.log
if [ -n "$PROFILE" ]; then
timeout 600 ncu --set full --clock-control none --import-source on -k regex:maxsim_tcr -s 3 -c 1 -f -o gpurun_out/r2_maxsim_tcr \
  python tools/bench_maxsim.py --docs 100000 --ragged 40,180 --steps 2 > gpurun_out/tcr_ncu.log 2>&1; tail -1 gpurun_out/tcr_ncu.log | cut -c1-300
fi
