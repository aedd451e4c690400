"""Device-time of one denoiser forward for the other BASELINE configs (SD3 MMDiT C3 per-GPU share, STDiT2 C5
per-GPU share, SD1.5 C1), random-init weights, CUDA events, eager launches (launch queue runs ahead of the GPU)."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops  # noqa: E402

bf = torch.bfloat16


def timeit(fn, iters=5, warmup=2):
    for _ in range(warmup):
        fn()
    torch.cuda.synchronize()
    n0 = ops.launches()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(iters):
        fn()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters, (ops.launches() - n0) // iters


def sd3():
    from paddlemix_b200.ppdiffusers.transformer_sd3 import SD3Transformer2DModel
    cfg = dict(sample_size=128, patch_size=2, in_channels=16, num_layers=24, attention_head_dim=64, num_attention_heads=24,
               joint_attention_dim=4096, caption_projection_dim=1536, pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=192)
    m = SD3Transformer2DModel(**cfg).init_synthetic_weights(seed=3, device=0)
    B = 4  # configs[2]: batch 32 over 8 GPUs
    x, ctx = torch.randn(B, 16, 128, 128, device="cuda"), torch.randn(B, 154, 4096, device="cuda").to(bf)
    pooled, t = torch.randn(B, 2048, device="cuda").to(bf), torch.full((B,), 500.0, device="cuda")
    ms, n = timeit(lambda: m(x, ctx, pooled, t))
    print(f"SD3-medium MMDiT  B={B} 1024^2 (4096+154 tokens): {ms:8.2f} ms/forward, {B / ms * 1e3:7.1f} sample-forwards/s, "
          f"{8.437 * B / ms * 1e3:6.1f} TFLOP/s, {n} launches")
    del m
    torch.cuda.empty_cache()


def stdit2():
    from paddlemix_b200.opensora import STDiT2
    m = STDiT2(dict(qk_norm=True)).init_synthetic_weights(seed=5, device=0)
    B = 1  # configs[4]: batch 4 over 4 GPUs
    x, y = torch.randn(B, 4, 16, 64, 64, device="cuda"), torch.randn(B, 1, 120, 4096, device="cuda").to(bf)
    kw = dict(num_frames=torch.full((B,), 16.0), height=torch.full((B,), 512.0), width=torch.full((B,), 512.0),
              ar=torch.full((B,), 1.0), fps=torch.full((B,), 24.0))
    t = torch.full((B,), 500.0, device="cuda")
    ms, n = timeit(lambda: m(x, t, y, **kw))
    print(f"STDiT2-XL        B={B} 16x512^2 (16x1024 tokens):  {ms:8.2f} ms/forward, {B / ms * 1e3:7.1f} sample-forwards/s, "
          f"{24.39 * B / ms * 1e3:6.1f} TFLOP/s, {n} launches")
    del m
    torch.cuda.empty_cache()


def sd15():
    from paddlemix_b200.ppdiffusers.unet_2d_condition import UNet2DConditionModel
    m = UNet2DConditionModel(cross_attention_dim=768, attention_head_dim=8).init_synthetic_weights(seed=1, device=0)
    for B in (1, 8):
        x, ctx = torch.randn(B, 4, 64, 64, device="cuda"), torch.randn(B, 77, 768, device="cuda").to(bf)
        ms, n = timeit(lambda: m(x, 981, ctx))
        print(f"SD1.5 UNet       B={B} 512^2:                       {ms:8.2f} ms/forward, {B / ms * 1e3:7.1f} sample-forwards/s, "
              f"{0.803 * B / ms * 1e3:6.1f} TFLOP/s, {n} launches (eager, launch-bound at B=1)")
    del m
    torch.cuda.empty_cache()


if __name__ == "__main__":
    ops.init(0)
    which = sys.argv[1:] or ["sd3", "stdit2", "sd15"]
    for w in which:
        globals()[w]()
