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
# per-call CUDA-event profile of one eager decode step (which kernels the 8 ms go to)
ops.profile_begin()
model.decode_body_static(stepper.ids, stepper.cos, stepper.sin, cache, stepper.rows - 1, stepper.kv_lens - 1)
prof = ops.profile_end(by_tag=True)
tot = sum(v["ms"] for v in prof.values())
print(f"# eager decode step, per-call events: {tot:.3f} ms in timed ops.* calls")
for k, v in sorted(prof.items(), key=lambda kv: -kv[1]["ms"])[:14]:
    rate = v["work"] / (v["ms"] * 1e-3) / (1e12 if v["unit"] == "flop" else 1e9)
    print(f"  {k:44s} n={v['calls']:4d} {v['ms']:8.3f} ms  {rate:9.1f} {'TFLOP/s' if v['unit'] == 'flop' else 'GB/s'}")
# isolated skinny GEMMs (graph-timed): GB/s of weight streaming
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_bench import timeit  # noqa: E402
for name, N, K, glu in [("qkv", 4608, 3584, 0), ("o_proj", 3584, 3584, 0), ("gate_up swiglu", 37888, 3584, 2), ("down", 3584, 18944, 0),
                        ("lm_head f32", 152064, 3584, 0)]:
    a = (torch.randn(4, K, device="cuda") * 0.05).to(torch.bfloat16)
    w = (torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16)
    ms = timeit(lambda: ops.linear(a, w, glu=glu, out_fp32=(name.startswith("lm_head"))), iters=10)
    print(f"  skinny {name:16s} N={N:6d} K={K:6d}: {ms * 1e3:8.1f} us  {N * K * 2 / ms / 1e6:8.1f} GB/s")
