"""Skinny-M (decode) Linear rates on the Qwen2-VL-7B decode problems, L2-cold: every problem rotates over enough distinct
weight matrices (> 300 MB in total) that no launch finds its weights in the 126 MB L2. GB/s = N*K*2 bytes / launch time.
usage: python tools/skinny_probe.py [M]"""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops  # noqa: E402

M = int(sys.argv[1]) if len(sys.argv) > 1 else 4
torch.manual_seed(0)
for name, N, K, kw in [("qkv bias", 4608, 3584, dict(bias=True)), ("o_proj res", 3584, 3584, dict(res=True)),
                       ("gate_up swiglu", 37888, 3584, dict(glu=2)), ("down res", 3584, 18944, dict(res=True)),
                       ("lm_head f32", 152064, 3584, dict(out_fp32=True))]:
    copies = max(2, int(320e6 // (N * K * 2)) + 1)
    ws = [(torch.randn(N, K, device="cuda") * 0.02).to(torch.bfloat16) for _ in range(copies)]
    a = (torch.randn(M, K, device="cuda") * 0.05).to(torch.bfloat16)
    bias = torch.randn(N, device="cuda") if kw.get("bias") else None
    res = torch.randn(M, N, device="cuda").to(torch.bfloat16) if kw.get("res") else None
    call = lambda w: ops.linear(a, w, bias, residual=res, glu=kw.get("glu", 0), out_fp32=kw.get("out_fp32", False))
    for w in ws:
        call(w)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    reps = max(1, 24 // copies)
    with torch.cuda.graph(g):
        for _ in range(reps):
            for w in ws:
                call(w)
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    for _ in range(5):
        g.replay()
    e1.record()
    torch.cuda.synchronize()
    us = e0.elapsed_time(e1) * 1e3 / (5 * reps * copies)
    print(f"skinny M={M} {name:16s} N={N:6d} K={K:6d}: {us:8.1f} us  {N * K * 2 / us / 1e3:8.1f} GB/s  ({copies} weight copies)", flush=True)
    del ws
