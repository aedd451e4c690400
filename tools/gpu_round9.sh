#!/bin/bash
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "linear_plain" -p no:cacheprovider > gpurun_out/test_lin.log 2>&1; echo "== linear exit $?"; tail -n 3 gpurun_out/test_lin.log
timeout 900 python tools/gemm_bench.py > gpurun_out/gemm_bench3.log 2>&1; echo "== gemm_bench exit $?"; head -n 16 gpurun_out/gemm_bench3.log
timeout 600 ncu --set full --clock-control none --import-source on -k regex:igemm -s 4 -c 2 -o gpurun_out/prof_out1280 python tools/one_gemm.py 8192 1280 1280 residual > gpurun_out/ncu_out1280.log 2>&1; echo "== ncu out_1280 exit $?"; tail -n 2 gpurun_out/ncu_out1280.log
timeout 900 python tools/bench_models.py > gpurun_out/bench_models.log 2>&1; echo "== bench_models exit $?"; cat gpurun_out/bench_models.log | tail -n 8
