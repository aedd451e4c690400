timeout 600 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider -k "sdpa or attention" 2>&1 | tail -n 4
timeout 300 python tools/attn_probe.py 2>&1 | tail -n 4
