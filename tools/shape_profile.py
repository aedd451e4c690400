"""Per-shape device time of one eager SDXL UNet forward (batch 8, 1024^2): CUDA events around every C-ABI call,
aggregated by shape tag. Eager launches add launch gaps, so compare shares and per-shape TFLOP/s, not the total."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from bench import SDXL  # noqa: E402
from paddlemix_b200 import ops  # noqa: E402
from paddlemix_b200.ppdiffusers.unet_2d_condition import UNet2DConditionModel  # noqa: E402

B = int(os.environ.get("B200MIX_PROFILE_BATCH", "8"))
unet = UNet2DConditionModel(**SDXL).init_synthetic_weights(seed=1, device=0)
g = torch.Generator().manual_seed(2)
x = ops.nchw_to_nhwc(torch.randn(B, 4, 128, 128, generator=g).cuda())
ctx = torch.randn(B, 77, 2048, generator=g).to(torch.bfloat16).cuda()
added = {"text_embeds": torch.randn(B, 1280, generator=g).to(torch.bfloat16).cuda(),
         "time_ids": torch.tensor([[1024., 1024., 0, 0, 1024., 1024.]] * B).cuda()}
t = torch.full((B,), 981.0, device="cuda")
for _ in range(2):
    unet.forward_nhwc(x, t, ctx, added)
torch.cuda.synchronize()
REPS = 3
ops.profile_begin()
for _ in range(REPS):
    unet.forward_nhwc(x, t, ctx, added)
prof = ops.profile_end(by_tag=True)
tot = sum(d["ms"] for d in prof.values()) / REPS
print(f"# total {tot:.2f} ms per forward (sum of per-call event times)")
for tag, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"]):
    ms = d["ms"] / REPS
    rate = d["work"] / d["ms"] / (1e9 if d["unit"] == "flop" else 1e6)
    print(f"{tag:44s} n={d['calls'] // REPS:4d} {ms:8.3f} ms {100 * ms / tot:5.1f}%  {rate:8.1f} {'TFLOP/s' if d['unit'] == 'flop' else 'GB/s'}")
