#!/bin/bash
# Round profile capture: GPU tests, bench line, ncu launch list of the bench command and
# one full capture of the dominant kernel. Outputs land in gpurun_out/.
mkdir -p gpurun_out
timeout 1200 python -m pytest tests -m gpu -x -q > gpurun_out/pytest_gpu.log 2>&1; echo "pytest rc=$?" >> gpurun_out/pytest_gpu.log
tail -3 gpurun_out/pytest_gpu.log
timeout 600 python bench.py > gpurun_out/bench_default.json 2> gpurun_out/bench_default.err; tail -1 gpurun_out/bench_default.json | cut -c1-300
timeout 600 ncu --metrics gpu__time_duration.sum --clock-control none -c 400 --csv --log-file gpurun_out/launches.csv python bench.py --steps 30 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_launches.log 2>&1
timeout 600 ncu --set full --clock-control none --import-source on -k regex:flat_stream -s 10 -c 1 -f -o gpurun_out/prof_flat_stream python bench.py --steps 20 --warmup 5 --no-cpu-baseline > gpurun_out/ncu_full.log 2>&1
tail -1 gpurun_out/ncu_full.log
