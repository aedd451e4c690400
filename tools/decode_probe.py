"""Qwen2-VL-7B decode latency (B=4, 768-token prefill with one 448x448 image each, then single-token steps against the KV
cache), eager launches, random-init weights."""
import os
import sys
import time

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops  # noqa: E402
from paddlemix_b200.qwen2_vl import Qwen2VLForConditionalGeneration  # noqa: E402

model = Qwen2VLForConditionalGeneration({}).init_synthetic_weights(seed=4, device=0)
c = model.config
g = torch.Generator().manual_seed(4)
B, n_img_tok, n_txt = 4, 256, 510
grid = [[1, 32, 32]] * B
pv = torch.randn(B * 1024, 1176, generator=g).to(torch.bfloat16).cuda()
rows = [[c.vision_start_token_id] + [c.image_token_id] * n_img_tok + [c.vision_end_token_id] +
        torch.randint(0, 151643, (n_txt,), generator=g).tolist() for _ in range(B)]
ids = torch.tensor(rows)
model.cache_headroom = 64
out = model(input_ids=ids, pixel_values=pv, image_grid_thw=torch.tensor(grid), use_cache=True)
cache, deltas = out.past_key_values, out.rope_deltas
nxt = out.logits[:, -1].argmax(-1).cpu().unsqueeze(1)
for _ in range(3):
    nxt = model(input_ids=nxt, past_key_values=cache, rope_deltas=deltas, use_cache=True).logits[:, -1].argmax(-1).cpu().unsqueeze(1)
torch.cuda.synchronize()
n0 = ops.launches()
t0 = time.perf_counter()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
N = 16
for _ in range(N):
    nxt = model(input_ids=nxt, past_key_values=cache, rope_deltas=deltas, use_cache=True).logits[:, -1].argmax(-1).cpu().unsqueeze(1)
e1.record()
torch.cuda.synchronize()
wall = (time.perf_counter() - t0) * 1e3 / N
print(f"decode B={B} past~{cache.length}: {e0.elapsed_time(e1) / N:.2f} ms/step device, {wall:.2f} ms/step wall, "
      f"{B / wall * 1e3:.0f} tokens/s, {(ops.launches() - n0) // N} launches/step")
# device-only time of one step (no host sync inside): launch a few steps back to back with fixed inputs
cos, sin = model._mrope_tables((cache.length + deltas.reshape(B, 1).long()).unsqueeze(0).expand(3, -1, -1))
ids_dev = nxt.cuda().reshape(-1)
e0.record()
for _ in range(8):
    model.decode_device(ids_dev, B, cos, sin, cache)
e1.record()
torch.cuda.synchronize()
print(f"decode_device back-to-back: {e0.elapsed_time(e1) / 8:.2f} ms/step (weights 15.2 GB bf16 -> {15.2 / (e0.elapsed_time(e1) / 8):.2f} TB/s if weight-bound)")
# the CUDA-graph step (device-side positions, skinny-M weight-streaming GEMMs, GQA rows folded into one attention tile)
from paddlemix_b200.qwen2_vl import GraphedDecodeStep  # noqa: E402
stepper = GraphedDecodeStep(model, cache, deltas)
tok = nxt.cuda().reshape(-1)
for _ in range(3):
    tok = stepper.step(tok).argmax(-1)
torch.cuda.synchronize()
e0.record()
t0 = time.perf_counter()
N = 32
for _ in range(N):
    tok = stepper.step(tok).argmax(-1)
e1.record()
torch.cuda.synchronize()
ms = e0.elapsed_time(e1) / N
print(f"graphed decode step B={B} past~{cache.length}: {ms:.3f} ms/step device, {(time.perf_counter() - t0) * 1e3 / N:.3f} ms/step wall, "
      f"{B / ms * 1e3:.0f} tokens/s, {stepper.launches_per_step} launches/step, weights 15.2 GB bf16 -> {15.2 / ms:.2f} TB/s of weight reads")
