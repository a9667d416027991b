#!/bin/bash
# usage: tools/gpu_scale.sh N  -> bench at N GPUs under torchrun (line to gpurun_out/bench_nN.json)
N=${1:-8}
mkdir -p gpurun_out
timeout 900 python -m torch.distributed.run --nnodes=1 --nproc-per-node $N --master-addr 127.0.0.1 --master-port 29511 bench.py --gpus $N --steps 1000 --warmup 50 > gpurun_out/bench_n$N.json 2> gpurun_out/bench_n$N.err
tail -1 gpurun_out/bench_n$N.json | python -c "
import sys, json
l = json.loads(sys.stdin.read())
print('N=%d value=%.1f ms/step=%.4f e2e=%.1f frac=%.3f clocks=%s' % (l['n_gpus'], l['value'], l['ms_per_step'], l['e2e']['value'], l['roofline']['frac'], l['clocks']))
"
