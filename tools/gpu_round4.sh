#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "sdpa or norm" -p no:cacheprovider > gpurun_out/test_ops3.log 2>&1; echo "== ops exit $?"; tail -n 8 gpurun_out/test_ops3.log
timeout 900 python -m pytest tests/test_qwen2vl_gpu.py tests/test_unet_gpu.py tests/test_sd3_gpu.py -q -m gpu -x -p no:cacheprovider -k "not sd15" > gpurun_out/test_models.log 2>&1; echo "== models exit $?"; tail -n 25 gpurun_out/test_models.log
BNS=0 timeout 900 python tools/gemm_bench.py > gpurun_out/gemm_bench2.log 2>&1; echo "== gemm_bench exit $?"; tail -n 8 gpurun_out/gemm_bench2.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench3.log 2> gpurun_out/bench3.err; echo "== bench exit $?"; tail -n 2 gpurun_out/bench3.log; tail -n 5 gpurun_out/bench3.err
