"""Times the small non-GEMM kernels of the SDXL forward in isolation (graph-timed)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from paddlemix_b200 import ops  # noqa: E402
from gemm_bench import rnd, timeit  # noqa: E402

ops.init(0)
x = torch.randn(8, 128, 128, 4, device="cuda")
w, b = rnd(320, 3, 3, 4), torch.zeros(320, device="cuda")
ms = timeit(lambda: ops.conv3x3_small_cin(x, w, b))
print(f"conv_in 8x128x128 4->320 (fp32 in): {ms * 1e3:8.1f} us   ({8 * 128 * 128 * 320 * 2 / ms / 1e6:6.0f} GB/s of output)")
xb = x.to(torch.bfloat16)
ms = timeit(lambda: ops.conv3x3_small_cin(xb, w, b))
print(f"conv_in 8x128x128 4->320 (bf16 in): {ms * 1e3:8.1f} us")
for B, HW, C in [(8, 16384, 320), (8, 16384, 640), (8, 4096, 640), (8, 1024, 1280), (8, 1024, 2560), (8, 4096, 1920)]:
    xx, g, bb = rnd(B, HW, C), torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
    ms = timeit(lambda: ops.groupnorm_nhwc(xx, g, bb, silu=True))
    print(f"groupnorm {B}x{HW}x{C}: {ms * 1e3:8.1f} us  {2 * B * HW * C * 2 / ms / 1e6:6.0f} GB/s (1R+1W)")
up = rnd(8, 32, 32, 1280)
ms = timeit(lambda: ops.upsample_nearest2x(up))
print(f"upsample2x 8x32x32x1280: {ms * 1e3:8.1f} us")
