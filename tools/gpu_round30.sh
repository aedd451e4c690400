timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider -k "glu or linear" 2>&1 | tail -n 3
timeout 300 python tools/misc_probe.py 2>&1 | tail -n 12
for v in glu0 glu1; do echo "== $v"; B200MIX_LIB=$PWD/paddlemix_b200/csrc/build/variants/libb200mix_$v.so timeout 300 python tools/epi_probe.py 2>&1 | grep -E "geglu"; done
bash tools/ab.sh glu0 glu1
