#!/bin/bash
# One GPU call: re-check the changed kernels, model parity, smoke, then the bench.
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_ops_gpu.py -q -m gpu -x -k "groupnorm" -p no:cacheprovider > gpurun_out/test_gn.log 2>&1; echo "== gn exit $?"; tail -n 5 gpurun_out/test_gn.log
timeout 1500 python -m pytest tests/test_unet_gpu.py -q -m gpu -x -s -p no:cacheprovider > gpurun_out/test_unet.log 2>&1; echo "== unet exit $?"; tail -n 30 gpurun_out/test_unet.log
timeout 600 python __graft_entry__.py smoke > gpurun_out/smoke.log 2>&1; echo "== smoke exit $?"; tail -n 5 gpurun_out/smoke.log
timeout 1200 python bench.py --steps 10 --warmup 3 > gpurun_out/bench.log 2> gpurun_out/bench.err; echo "== bench exit $?"; tail -n 3 gpurun_out/bench.log; tail -n 15 gpurun_out/bench.err
