#!/bin/bash
# K2 as CTA pairs (cta_group::2): parity tests under a hard timeout, then C2 / C3 timing with and without pairs.
mkdir -p gpurun_out
VB_GEMM_PAIR=1 timeout 420 python -m pytest tests/test_batch_gpu.py -x -q > gpurun_out/pytest_pair.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_pair.log
tail -6 gpurun_out/pytest_pair.log | cut -c1-300
VB_GEMM_PAIR=1 timeout 200 python tools/bench_batch.py --steps 5 > gpurun_out/pair_c2.log 2>&1; echo "c2 rc=$?"; tail -1 gpurun_out/pair_c2.log | cut -c1-330
timeout 200 python tools/bench_batch.py --steps 5 > gpurun_out/nopair_c2.log 2>&1; tail -1 gpurun_out/nopair_c2.log | cut -c1-330
