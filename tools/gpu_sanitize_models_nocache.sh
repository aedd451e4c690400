#!/bin/bash
# memcheck with PyTorch's caching allocator off: every tensor is its own cudaMalloc, so an out-of-bounds access of a
# few bytes past a tensor is caught (inside a cached block it would be invisible).
mkdir -p gpurun_out
PYTORCH_NO_CUDA_MEMORY_CACHING=1 timeout 1200 compute-sanitizer --tool memcheck --error-exitcode 7 --print-limit 4 python -m pytest tests/test_unet_gpu.py tests/test_sd3_gpu.py tests/test_stdit2_gpu.py tests/test_qwen2vl_gpu.py -x -q -m gpu -p no:cacheprovider -k "tiny_parity or batch_vs_single or special_attn or prefill_parity or decode" > gpurun_out/sanitize_nocache.log 2>&1; echo "memcheck(no cache) exit $?"
grep -E "ERROR SUMMARY|passed|failed" gpurun_out/sanitize_nocache.log | tail -n 3
grep -n "=========" gpurun_out/sanitize_nocache.log | head -16 | cut -c1-230
