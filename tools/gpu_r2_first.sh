#!/bin/bash
# Round-2 first pass on one GPU: parity tests, smoke, the full default bench (all config blocks).
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1; nproc >> gpurun_out/gpus.txt; free -g >> gpurun_out/gpus.txt
timeout 1500 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -3 gpurun_out/smoke.log
( time timeout 900 python bench.py --steps 200 --warmup 20 ) > gpurun_out/bench_n1.log 2>&1; echo "rc=$?" >> gpurun_out/bench_n1.log
tail -c 6000 gpurun_out/bench_n1.log
