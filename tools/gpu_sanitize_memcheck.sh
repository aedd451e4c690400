#!/bin/bash
# compute-sanitizer memcheck over a representative subset of the kernel tests (OOB / misaligned accesses).
mkdir -p gpurun_out
timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 20 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider \
  -k "test_linear_plain and (1000 or 300 or 8-1280) or conv_small_cin or conv_stride or layernorm_block or short_kv or sdpa_varlen or sdpa_kv_lens or groupnorm or glu or test_sdpa and (200 or 4250)" \
  > gpurun_out/sanitize.log 2>&1; echo "memcheck exit $?"
grep -E "ERROR SUMMARY|Invalid|passed|failed|misaligned|out of bounds" gpurun_out/sanitize.log | sort | uniq -c | head -20
tail -n 3 gpurun_out/sanitize.log
