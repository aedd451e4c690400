#!/bin/bash
mkdir -p gpurun_out
B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_prof.so timeout 200 python tools/attn_prof.py 2>&1 | tee gpurun_out/r2_attn_prof2.log
timeout 300 python tools/attn_probe.py 2>&1 | tee gpurun_out/r2_attn_probe2.log
timeout 900 python -m pytest tests/test_attention_gpu.py tests/test_ops_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -3
bash tools/ab.sh main igold 2>&1 | tee gpurun_out/r2_ab_roles.log
