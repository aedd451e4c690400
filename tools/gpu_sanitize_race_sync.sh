#!/bin/bash
mkdir -p gpurun_out
for tool in racecheck synccheck; do
timeout 900 compute-sanitizer --tool $tool --error-exitcode 7 --print-limit 10 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider \
  -k "test_linear_plain and 1000 and (0-256 or 1-224) or layernorm_block and 4099 or short_kv and 130 or test_conv_small_cin and 17 or sdpa_kv_lens" \
  > gpurun_out/sanitize_$tool.log 2>&1; echo "$tool exit $?"
grep -E "SUMMARY|hazard|passed|failed|Barrier|error" gpurun_out/sanitize_$tool.log | sort | uniq -c | head -12
done
