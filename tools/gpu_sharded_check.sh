#!/bin/bash
# usage: tools/gpu_sharded_check.sh N  -> emulated sharded tests on one GPU, then the NCCL check on N GPUs
N=${1:-2}
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_sharded_gpu.py -x -q 2>&1 | tail -4
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
  tools/sharded_check.py > gpurun_out/sharded_check_n$N.log 2>&1
echo "rc=$?"; tail -5 gpurun_out/sharded_check_n$N.log | cut -c1-900
