timeout 300 python -m pytest tests/test_ops_gpu.py -x -q -m gpu -p no:cacheprovider -k "sdpa" 2>&1 | tail -n 12
timeout 300 python - <<'PY'
import sys, torch
sys.path.insert(0, "tools")
from paddlemix_b200 import ops
from paddlemix_b200._lib import lib
from gemm_bench import rnd, timeit
for B, S, H in [(8, 1024, 20), (8, 4096, 10)]:
    q, k, v = rnd(B, S, H, 64), rnd(B, 77, H, 64), rnd(B, 77, H, 64)
    for mode in (0, 1, 0, 1):
        lib.b200mix_debug_no_shortkv(mode)
        ms = timeit(lambda: ops.sdpa(q, k, v))
        print(f"cross-attn B{B} Sq{S} H{H} Sk77  {'general' if mode else 'short-kv'}: {ms*1e3:7.1f} us", flush=True)
lib.b200mix_debug_no_shortkv(0)
PY
