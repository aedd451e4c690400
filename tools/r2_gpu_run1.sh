#!/bin/bash
# round 2, first GPU pass: new attention kernels / masks / activation edges, then the whole suite, parity at BASELINE shapes
mkdir -p gpurun_out
timeout 600 python -m pytest tests/test_attention_gpu.py -x -q -p no:cacheprovider > gpurun_out/r2_attn_tests.log 2>&1; echo "== attention tests exit $?"; tail -n 15 gpurun_out/r2_attn_tests.log
timeout 300 python -m pytest tests/test_activation_edges_gpu.py -x -q -p no:cacheprovider > gpurun_out/r2_act_tests.log 2>&1; echo "== activation tests exit $?"; tail -n 5 gpurun_out/r2_act_tests.log
timeout 300 python tools/attn_probe.py > gpurun_out/r2_attn_probe.log 2>&1; echo "== attn probe exit $?"; cat gpurun_out/r2_attn_probe.log
timeout 1500 python -m pytest tests/ -q -m "gpu and not slow" -p no:cacheprovider --deselect tests/test_attention_gpu.py --deselect tests/test_activation_edges_gpu.py > gpurun_out/r2_gpu_tests.log 2>&1; echo "== gpu tests exit $?"; tail -n 15 gpurun_out/r2_gpu_tests.log
timeout 1500 python -m pytest tests/test_baseline_parity_gpu.py tests/test_unet_gpu.py -q -m "gpu and slow" -p no:cacheprovider > gpurun_out/r2_slow_tests.log 2>&1; echo "== slow parity tests exit $?"; tail -n 15 gpurun_out/r2_slow_tests.log; cat gpurun_out/parity_baseline.json
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen > gpurun_out/r2_bench1.log 2> gpurun_out/r2_bench1.err; echo "== bench exit $?"; tail -c 2500 gpurun_out/r2_bench1.log; tail -n 5 gpurun_out/r2_bench1.err
