timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider -k "sdpa or attention" 2>&1 | tail -n 5
for v in ptmem0 ptmem1 ptmem0 ptmem1; do echo "== $v"; B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so timeout 300 python tools/attn_probe.py 2>&1 | tail -n 3 | cut -c1-60; done
