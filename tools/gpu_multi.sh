#!/bin/bash
# usage: tools/gpu_multi.sh N   -> tests + bench at N GPUs (torchrun) + N=1
N=${1:-2}
mkdir -p gpurun_out
timeout 900 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -4 gpurun_out/pytest_gpu.log
timeout 600 python bench.py --gpus 1 --steps 200 --warmup 20 > gpurun_out/bench_n1.log 2>&1; tail -1 gpurun_out/bench_n1.log | cut -c1-400
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 200 --warmup 20 > gpurun_out/bench_n$N.log 2>&1
tail -3 gpurun_out/bench_n$N.log | cut -c1-1200
timeout 600 python bench.py --impl reference --gpus 1 --steps 4 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -1 gpurun_out/bench_ref.log | cut -c1-600
