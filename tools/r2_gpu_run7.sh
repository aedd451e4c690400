#!/bin/bash
mkdir -p gpurun_out
B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_prof.so timeout 200 python tools/attn_prof.py 2>&1 | tee gpurun_out/r2_attn_prof4.log
for v in main poly8 poly5 poly4 poly3; do
  if [ "$v" = main ]; then unset B200MIX_LIB; else export B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so; fi
  echo "== variant $v"
  timeout 300 python tools/attn_probe.py 2>&1 | tee gpurun_out/r2_attn_probe4_$v.log
  timeout 600 python -m pytest tests/test_attention_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -2
done
