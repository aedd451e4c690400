"""One eager SDXL UNet forward (batch 8, 1024^2) between cudaProfilerStart/Stop, for ncu launch lists / captures:
  ncu --profile-from-start off --metrics gpu__time_duration.sum --clock-control none --csv --log-file launches.csv python tools/profile_step.py
"""
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
torch.cuda.cudart().cudaProfilerStart()
unet.forward_nhwc(x, t, ctx, added)
torch.cuda.synchronize()
torch.cuda.cudart().cudaProfilerStop()
print("profiled one forward, launches:", ops.launches())
