#!/bin/bash
# compute-sanitizer memcheck over the kernels this round added or reshaped (small cases; the sanitizer slows kernels 10-50x).
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_maxsim_ragged_gpu.py -x -q \
  -k "tiny or long-120 or compaction or overflowing or one_document" > gpurun_out/memcheck_tcr.log 2>&1; echo "tcr rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_tcr.log | tail -3
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_batch_gpu.py -x -q \
  -k "cta_pair or overflow_semantics or ties_resolve" > gpurun_out/memcheck_k2.log 2>&1; echo "k2 rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_k2.log | tail -3
timeout 900 compute-sanitizer --tool memcheck --error-exitcode 9 python -m pytest tests/test_sharded_process_gpu.py -x -q \
  -k "multi_vector and 2-inner or quantized_search_equals and 2" > gpurun_out/memcheck_sharded.log 2>&1; echo "sharded rc=$?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/memcheck_sharded.log | tail -3
