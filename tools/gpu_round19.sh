bash tools/ab.sh pair0 pair1
for v in pair0 pair1; do echo "== $v"; B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so BNS=0 timeout 600 python tools/gemm_bench.py 2>&1 | head -n 16; done
