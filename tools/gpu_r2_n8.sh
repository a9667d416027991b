#!/bin/bash
# usage: tools/gpu_r2_n8.sh N  -> sharded_check + the full bench line at N GPUs (torchrun), end of round 2
N=${1:-8}
mkdir -p gpurun_out
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/sharded_check.py > gpurun_out/sharded_check_n$N.log 2>&1; echo "sharded_check rc=$?"
tail -2 gpurun_out/sharded_check_n$N.log | cut -c1-600
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 400 --warmup 20 ) > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"
grep -E '^\{|real|Error|error' gpurun_out/bench_n$N.log | tail -4 | cut -c1-400
