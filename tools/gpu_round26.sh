for v in attn_old attn_ilp attn_old attn_ilp; do echo "== $v"; B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so BNS=0 timeout 600 python tools/gemm_bench.py 2>&1 | tail -n 2; done
B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_attn_ilp.so timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider -k "sdpa or attention" 2>&1 | tail -n 3
