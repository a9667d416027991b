#!/bin/bash
# usage: tools/gpu_r2_multi.sh N   -> sharded_check + full bench at N GPUs (torchrun), NCCL-exchange headline for comparison
N=${1:-2}
mkdir -p gpurun_out
nvidia-smi topo -m > gpurun_out/topo_n$N.txt 2>&1
timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29533 \
    tools/sharded_check.py > gpurun_out/sharded_check_n$N.log 2>&1; echo "sharded_check rc=$?"
tail -3 gpurun_out/sharded_check_n$N.log | cut -c1-1500
( time timeout 1200 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 \
    bench.py --gpus $N --steps 200 --warmup 20 ) > gpurun_out/bench_n$N.log 2>&1; echo "bench rc=$?"
grep -E '^\{|real|Error|error' gpurun_out/bench_n$N.log | tail -5 | cut -c1-9000
VB_EXCHANGE=nccl timeout 600 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29512 \
    bench.py --gpus $N --steps 200 --warmup 20 --configs '' > gpurun_out/bench_n${N}_nccl.log 2>&1; echo "bench(nccl) rc=$?"
grep -E '^\{' gpurun_out/bench_n${N}_nccl.log | tail -1 | cut -c1-1500
