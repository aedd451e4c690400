"""Sensitivity probes for the small-K GEMM shapes: residual / bias / K / N effects (graph-timed)."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops
from paddlemix_b200._lib import lib
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_bench import timeit, rnd  # noqa
ops.init(0)
for name, M, N, K, bias, res in [("out_640 bias+res", 32768, 640, 640, True, True), ("out_640 bias", 32768, 640, 640, True, False),
                                 ("out_640 plain", 32768, 640, 640, False, False), ("N=640 K=1280 res", 32768, 640, 1280, True, True),
                                 ("N=1280 K=640 res", 32768, 1280, 640, True, True), ("N=2560 K=640 plain", 32768, 2560, 640, False, False),
                                 ("out_1280 bias+res", 8192, 1280, 1280, True, True), ("out_1280 plain", 8192, 1280, 1280, False, False),
                                 ("M=16384 N=1280 K=1280 res", 16384, 1280, 1280, True, True)]:
    a, w = rnd(M, K), rnd(N, K)
    b = torch.zeros(N, device="cuda") if bias else None
    r = rnd(M, N) if res else None
    row = f"{name:28s}"
    for bn in (0, 160, 256):
        lib.b200mix_debug_force_bn(bn)
        ms = timeit(lambda: ops.linear(a, w, b, residual=r))
        row += f"  bn={bn or 'auto'}: {2.0 * M * N * K / ms / 1e9:6.0f}"
    print(row, flush=True)
lib.b200mix_debug_force_bn(0)
