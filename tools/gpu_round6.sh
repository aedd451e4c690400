#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py tests/test_stdit2_gpu.py -q -m gpu -x -k "kv_lens or small_attention or stdit" -p no:cacheprovider > gpurun_out/test_stdit.log 2>&1; echo "== stdit exit $?"; tail -n 25 gpurun_out/test_stdit.log
