#!/bin/bash
mkdir -p gpurun_out
timeout 200 python tools/mufu_bench.py > gpurun_out/r2_mufu_bench.log 2>&1; echo "== mufu bench exit $?"; cat gpurun_out/r2_mufu_bench.log
for v in main intpack intpack4; do
  if [ "$v" = main ]; then unset B200MIX_LIB; else export B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so; fi
  echo "== variant $v"
  timeout 300 python tools/attn_probe.py 2>&1 | tee gpurun_out/r2_attn_probe_$v.log
done
export B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_intpack.so
timeout 600 python -m pytest tests/test_attention_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -3
unset B200MIX_LIB
timeout 600 python -m pytest tests/test_scheduler_steps_gpu.py -x -q -p no:cacheprovider 2>&1 | tail -5
