#!/bin/bash
# Round-2 closing pass (one GPU): the whole GPU suite, smoke, the default bench line, the ncu launch list of the same
# command and full captures of the two kernels this session changed most (ragged MaxSim, single-pass K2).
mkdir -p gpurun_out
timeout 1700 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -6 gpurun_out/pytest_gpu.log
timeout 300 python -c "import __graft_entry__ as g; g.smoke()" > gpurun_out/smoke.log 2>&1; echo "smoke rc=$?" >> gpurun_out/smoke.log; tail -2 gpurun_out/smoke.log
timeout 1200 python bench.py > gpurun_out/bench_n1.log 2>&1; echo "bench rc=$?"; tail -c 600 gpurun_out/bench_n1.log
timeout 600 python bench.py --impl reference --steps 8 --warmup 1 > gpurun_out/bench_ref.log 2>&1; tail -c 300 gpurun_out/bench_ref.log
timeout 900 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/r2_launches.csv \
  python bench.py --steps 100 --warmup 5 --configs '' --no-cpu-baseline > gpurun_out/bench_ncu.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:maxsim_tcr -s 3 -c 1 -f -o gpurun_out/r2_maxsim_tcr \
  python tools/bench_maxsim.py --docs 100000 --ragged 40,180 --steps 2 > gpurun_out/tcr_ncu.log 2>&1
timeout 900 ncu --set full --clock-control none --import-source on -k regex:flat_gemm1_topk -s 3 -c 1 -f -o gpurun_out/r2_flat_gemm1 \
  python tools/bench_batch.py --steps 2 > gpurun_out/k2_ncu.log 2>&1
timeout 300 python tools/bench_maxsim.py --docs 200000 --ragged 40,180 > gpurun_out/maxsim_ragged_40_180.log 2>&1; tail -1 gpurun_out/maxsim_ragged_40_180.log | cut -c1-400
ls -la gpurun_out/*.ncu-rep | tail -3
