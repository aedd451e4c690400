for v in main pfb4 pfb8 main; do
  if [ "$v" = main ]; then unset B200MIX_LIB; else export B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so; fi
  echo "== $v"; timeout 300 python tools/skinny_probe.py 4 2>&1 | tail -5
done
unset B200MIX_LIB
timeout 900 python -m pytest tests/test_qwen2vl_gpu.py -q -m "gpu and not slow" -p no:cacheprovider -x 2>&1 | tail -8
