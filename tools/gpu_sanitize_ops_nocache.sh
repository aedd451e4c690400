#!/bin/bash
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 600 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 4 python -m pytest tests/test_ops_gpu.py -q -m gpu -p no:cacheprovider > gpurun_out/sanitize_nocache_ops.log 2>&1; echo "memcheck(no cache, ops) exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_nocache_ops.log | tail -n 3
grep -n "=========" gpurun_out/sanitize_nocache_ops.log | head -14 | cut -c1-230
