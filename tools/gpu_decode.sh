#!/bin/bash
# decode-path check: skinny-M GEMM tests, Qwen2-VL tests, L2-cold skinny rates, the decode probe
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -k "skinny" -p no:cacheprovider -x 2>&1 | tail -4
timeout 900 python -m pytest tests/test_qwen2vl_gpu.py tests/test_clip_llava_gpu.py -q -m "gpu and not slow" -p no:cacheprovider -x 2>&1 | tail -8
timeout 300 python tools/skinny_probe.py 4 2>&1 | tail -5
timeout 600 python tools/decode_probe.py > gpurun_out/r2_decode_probe.log 2>&1; echo "probe exit $?"; tail -22 gpurun_out/r2_decode_probe.log
