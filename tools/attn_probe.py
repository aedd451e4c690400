"""Self-attention rate at the SDXL / SD3 shapes (D = 64): persistent ping-pong kernel (default) vs the one-tile-per-CTA
kernel it replaced (b200mix_debug_attn_pingpong(0)); plus the cross-attention (short-KV) shapes."""
import os
import sys

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from paddlemix_b200 import ops  # noqa: E402
from paddlemix_b200._lib import lib  # noqa: E402
from gemm_bench import rnd, timeit  # noqa: E402

ops.init(0)
for B, S, H in [(8, 4096, 10), (8, 1024, 20), (4, 4250, 24), (2, 4096, 10), (1, 1024, 20)]:
    q, k, v = rnd(B, S, H, 64), rnd(B, S, H, 64), rnd(B, S, H, 64)
    row = f"attn B{B} S{S} H{H} d64:"
    for rep in range(2):
        for pp in (1, 0):
            lib.b200mix_debug_attn_pingpong(pp)
            ms = timeit(lambda: ops.sdpa(q, k, v), iters=10)
            row += f"  {'pingpong' if pp else 'one-tile'} {4.0 * B * H * S * S * 64 / ms / 1e9:6.0f}"
    print(row + "  TFLOP/s", flush=True)
lib.b200mix_debug_attn_pingpong(1)
