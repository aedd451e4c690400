"""Mainloop rate of each tile width (and of the pair / multicast MMA modes) on a GEMM with enough tiles that wave
quantisation is < 2 %: the efficiency table of pick_bn() in gemm.cu comes from this."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from paddlemix_b200 import ops  # noqa: E402
from paddlemix_b200._lib import lib  # noqa: E402
from gemm_bench import rnd, timeit  # noqa: E402

M, N = 16384, 8192
bns = [256, 224, 192, 160, 128, 64, 32]
print(f"{'K':>6s} {'mode':>6s}" + "".join(f"{'bn=%d' % b:>9s}" for b in bns) + "   (TFLOP/s)")
for K in (2048, 640):
    a, w = rnd(M, K), rnd(N, K)
    for pair in (1, 0):
        lib.b200mix_debug_gemm_pair(pair)
        row = f"{K:6d} {'pair' if pair else 'mcast':>6s}"
        for bn in bns:
            lib.b200mix_debug_force_bn(bn)
            ms = timeit(lambda: ops.linear(a, w), iters=6)
            row += f"{2.0 * M * N * K / ms / 1e9:9.0f}"
        print(row, flush=True)
lib.b200mix_debug_force_bn(0)
lib.b200mix_debug_gemm_pair(1)
