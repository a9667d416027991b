#!/bin/bash
mkdir -p gpurun_out
run() { name=$1; shift; rows=$1; shift
  env "$@" timeout 400 python bench.py --steps 100 --warmup 10 --no-cpu-baseline --rows $rows > gpurun_out/bench_$name.log 2>&1
  tail -1 gpurun_out/bench_$name.log | python -c "
import sys, json
t = sys.stdin.read().strip()
try:
    l = json.loads(t); print('$name rows=$rows GB/s=%.0f kernel_ms=%.4f' % (l['roofline']['achieved'], l['roofline']['kernel_ms']))
except Exception as e:
    print('$name fail', t[-200:])
"
}
run r500k_base 500000 X=1
run r1m_base 1000000 X=1
run r2m_base 2000000 X=1
run r500k_nothing 500000 VB_SCAN_DEBUG=7
run r1m_nothing 1000000 VB_SCAN_DEBUG=7
run r2m_nothing 2000000 VB_SCAN_DEBUG=7
run r1m_w8_nothing 1000000 VB_SCAN_DEBUG=7 VB_STREAM_WARPS=8
run r1m_w8r2_nothing 1000000 VB_SCAN_DEBUG=7 VB_STREAM_WARPS=8 VB_STREAM_RPW=2
