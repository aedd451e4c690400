"""Per-shape device time of the other BASELINE configs: Qwen2-VL-7B prefill (configs[3]) and SD3-medium (configs[2])."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops  # noqa: E402

bf = torch.bfloat16


def report(name, fn, reps=3):
    for _ in range(2):
        fn()
    torch.cuda.synchronize()
    ops.profile_begin()
    for _ in range(reps):
        fn()
    prof = ops.profile_end(by_tag=True)
    tot = sum(d["ms"] for d in prof.values()) / reps
    print(f"# {name}: {tot:.2f} ms (sum of tagged per-call event times)")
    for tag, d in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:16]:
        ms = d["ms"] / reps
        rate = d["work"] / d["ms"] / (1e9 if d["unit"] == "flop" else 1e6)
        print(f"  {tag:46s} n={d['calls'] // reps:4d} {ms:8.3f} ms {100 * ms / tot:5.1f}%  {rate:8.1f} {'TFLOP/s' if d['unit'] == 'flop' else 'GB/s'}")


def qwen():
    from paddlemix_b200.qwen2_vl import Qwen2VLForConditionalGeneration
    model = Qwen2VLForConditionalGeneration({}).init_synthetic_weights(seed=4, device=0)
    c = model.config
    g = torch.Generator().manual_seed(4)
    B, n_img_tok, n_txt = 4, 256, 510
    grid = [[1, 32, 32]] * B
    pv = torch.randn(B * 1024, 1176, generator=g).to(bf).cuda()
    rows = [[c.vision_start_token_id] + [c.image_token_id] * n_img_tok + [c.vision_end_token_id] +
            torch.randint(0, 151643, (n_txt,), generator=g).tolist() for _ in range(B)]
    ids_h = torch.tensor(rows)
    S = ids_h.shape[1]
    pos, _ = model.get_rope_index(ids_h, torch.tensor(grid))
    cos, sin = model._mrope_tables(pos)
    ids_d = ids_h.cuda().reshape(-1)
    idx_d = (ids_h.reshape(-1) == c.image_token_id).nonzero().reshape(-1).cuda()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    fn = lambda: model.prefill_device(ids_d, B, S, cos, sin, pv, grid, idx_d)  # noqa: E731
    fn(), fn()
    e0.record()
    for _ in range(3):
        fn()
    e1.record()
    torch.cuda.synchronize()
    print(f"# Qwen2-VL-7B prefill wall: {e0.elapsed_time(e1) / 3:.2f} ms")
    report("Qwen2-VL-7B prefill 4x768 tokens", fn)


def sd3():
    from paddlemix_b200.ppdiffusers.transformer_sd3 import SD3Transformer2DModel
    cfg = dict(sample_size=128, patch_size=2, in_channels=16, num_layers=24, attention_head_dim=64, num_attention_heads=24,
               joint_attention_dim=4096, caption_projection_dim=1536, pooled_projection_dim=2048, out_channels=16, pos_embed_max_size=192)
    m = SD3Transformer2DModel(**cfg).init_synthetic_weights(seed=3, device=0)
    B = 4
    x, ctx = torch.randn(B, 16, 128, 128, device="cuda"), torch.randn(B, 154, 4096, device="cuda").to(bf)
    pooled, t = torch.randn(B, 2048, device="cuda").to(bf), torch.full((B,), 500.0, device="cuda")
    report("SD3-medium B=4 1024^2", lambda: m(x, ctx, pooled, t))


if __name__ == "__main__":
    qwen()
    torch.cuda.empty_cache()
    sd3()
