"""Does the big-GEMM rate scale with the number of SMs? (It does not beyond ~100 SMs: the mainloop is bound by the
L2 -> SM fill bandwidth, ~6.3 KB/clk chip-wide, not by the tensor pipes.)"""
import ctypes
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from paddlemix_b200 import ops  # noqa: E402
from paddlemix_b200._lib import lib  # noqa: E402
from gemm_bench import rnd, timeit  # noqa: E402

lib.b200mix_debug_max_clusters.argtypes = [ctypes.c_int]
lib.b200mix_debug_max_clusters.restype = None
M, N, K = 16384, 8192, 2048
a, w = rnd(M, K), rnd(N, K)
for bn in (256, 128):
    lib.b200mix_debug_force_bn(bn)
    row = f"bn={bn}: "
    for clusters in (74, 64, 56, 48, 40, 32, 24):
        lib.b200mix_debug_max_clusters(clusters)
        ms = timeit(lambda: ops.linear(a, w), iters=6)
        row += f"  {2 * clusters}SM:{2.0 * M * N * K / ms / 1e9:6.0f}"
    print(row, flush=True)
lib.b200mix_debug_max_clusters(0)
lib.b200mix_debug_force_bn(0)
