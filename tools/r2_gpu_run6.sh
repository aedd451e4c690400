#!/bin/bash
mkdir -p gpurun_out
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_ops_gpu.py tests/test_activation_edges_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -3
B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_prof.so timeout 200 python tools/attn_prof.py 2>&1 | tee gpurun_out/r2_attn_prof3.log
timeout 300 python tools/attn_probe.py 2>&1 | tee gpurun_out/r2_attn_probe3.log
timeout 600 python bench.py --steps 10 --warmup 3 --no-cpu-baseline --no-qwen --no-extras > gpurun_out/r2_bench3.log 2> gpurun_out/r2_bench3.err; tail -c 1800 gpurun_out/r2_bench3.log; tail -3 gpurun_out/r2_bench3.err
timeout 600 python tools/shape_profile.py > gpurun_out/r2_shape_profile3.log 2>&1; head -40 gpurun_out/r2_shape_profile3.log
