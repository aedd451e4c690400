#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "linear or conv" -p no:cacheprovider > gpurun_out/test_ops2.log 2>&1; echo "== ops exit $?"; tail -n 6 gpurun_out/test_ops2.log
timeout 900 python tools/gemm_bench.py > gpurun_out/gemm_bench.log 2>&1; echo "== gemm_bench exit $?"; cat gpurun_out/gemm_bench.log
timeout 900 python bench.py --steps 10 --warmup 3 --no-cpu-baseline > gpurun_out/bench2.log 2> gpurun_out/bench2.err; echo "== bench exit $?"; tail -n 2 gpurun_out/bench2.log; tail -n 5 gpurun_out/bench2.err
