"""One launch each of the kernels the SDXL forward does not use (for ncu captures): causal GQA attention at d = 128 (Qwen2-VL
prefill shape), STDiT2's temporal small attention, the VAE row softmax, the skinny decode GEMM."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops  # noqa: E402

ops.init(0)
bf = torch.bfloat16
q = torch.randn(4, 768, 28, 128, device="cuda").to(bf)
k = torch.randn(4, 768, 4, 128, device="cuda").to(bf)
v = torch.randn(4, 768, 4, 128, device="cuda").to(bf)
B, T, S, H, d = 1, 16, 1024, 16, 72
qkv = torch.randn(B * T * S, 3 * H * d, device="cuda").to(bf)
x = torch.randn(4096, 16384, device="cuda")
for i in range(3):
    if i == 2:
        torch.cuda.synchronize()
        torch.cuda.cudart().cudaProfilerStart()
    ops.sdpa(q, k, v, causal=True)
    ops.small_attention(qkv, B, T, S, H, d, scale=d ** -0.5)
    ops.softmax_rows(x, scale=0.044)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("done")
