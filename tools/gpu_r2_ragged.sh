#!/bin/bash
# Round-2: ragged tensor-core MaxSim — parity tests, then the ragged / uniform / general kernels on comparable corpora.
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_maxsim_ragged_gpu.py tests/test_maxsim_gpu.py -x -q > gpurun_out/pytest_ragged.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_ragged.log
tail -25 gpurun_out/pytest_ragged.log
timeout 300 python tools/bench_maxsim.py --docs 200000 --ragged 40,180 > gpurun_out/maxsim_ragged_40_180.log 2>&1; tail -1 gpurun_out/maxsim_ragged_40_180.log
VB_MAXSIM_NO_TCU=1 timeout 300 python tools/bench_maxsim.py --docs 172000 --ragged 128,128 > gpurun_out/maxsim_ragged_128.log 2>&1; tail -1 gpurun_out/maxsim_ragged_128.log
VB_MAXSIM_NO_TCR=1 timeout 300 python tools/bench_maxsim.py --docs 172000 > gpurun_out/maxsim_uniform_128.log 2>&1; tail -1 gpurun_out/maxsim_uniform_128.log
timeout 300 python tools/bench_maxsim.py --docs 200000 --ragged 40,180 --tq 64 > gpurun_out/maxsim_ragged_tq64.log 2>&1; tail -1 gpurun_out/maxsim_ragged_tq64.log
VB_MAXSIM_NO_TCR=1 timeout 300 python tools/bench_maxsim.py --docs 50000 --ragged 40,180 --steps 5 > gpurun_out/maxsim_ragged_general.log 2>&1; tail -1 gpurun_out/maxsim_ragged_general.log
