"""One self-attention launch per shape (for ncu captures)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops  # noqa: E402

ops.init(0)
for B, S, H in [(8, 4096, 10), (8, 1024, 20)]:
    q, k, v = (torch.randn(B, S, H, 64, device="cuda").to(torch.bfloat16) for _ in range(3))
    for _ in range(3):
        ops.sdpa(q, k, v)
    torch.cuda.synchronize()
