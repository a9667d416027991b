#!/bin/bash
# ncu capture of the Hamming scan kernel (k from $1, default 10) plus the launch list of one k=1000 scan.
mkdir -p gpurun_out
k=${1:-10}
timeout 300 ncu --set full --clock-control none --import-source on -k regex:hamming_s -s 3 -c 1 -f -o gpurun_out/ham_k$k \
  python tools/bench_hamming.py --k $k --iters 5 > gpurun_out/ham_k$k.log 2>&1
ls -la gpurun_out/*.ncu-rep
