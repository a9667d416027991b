#!/bin/bash
# First GPU pass: parity tests, smoke, a short bench and scan-variant probes.
set -x
mkdir -p gpurun_out
nvidia-smi -L > gpurun_out/gpus.txt 2>&1
nproc >> gpurun_out/gpus.txt
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -30 gpurun_out/pytest_gpu.log
timeout 300 python __graft_entry__.py --smoke > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log
tail -5 gpurun_out/smoke.log
timeout 600 python bench.py --steps 200 --warmup 20 > gpurun_out/bench_default.log 2>&1; echo "rc=$?" >> gpurun_out/bench_default.log
tail -3 gpurun_out/bench_default.log
for R in 1 4; do
  VB_SCAN_R=$R timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_R$R.log 2>&1
  tail -1 gpurun_out/bench_R$R.log
done
VB_SCAN_CTAS_PER_SM=1 timeout 300 python bench.py --steps 200 --warmup 20 --no-cpu-baseline > gpurun_out/bench_cta1.log 2>&1
tail -1 gpurun_out/bench_cta1.log
