"""Epilogue-bound tile time: a GEMM with K=64 has a one-k-chunk mainloop, so its time per 128x256 tile is the cost of
the epilogue variant (TMEM -> registers -> math -> global)."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from paddlemix_b200 import ops  # noqa: E402
from paddlemix_b200._lib import GLU_GEGLU, lib  # noqa: E402
from gemm_bench import rnd, timeit  # noqa: E402

M, N = 32768, 5120
lib.b200mix_debug_force_bn(256)
tiles_per_cta = (M // 128) * (N // 256) / 148.0
for K in (64, 640):
    a, w = rnd(M, K), rnd(N, K)
    bias = torch.zeros(N, device="cuda")
    res = rnd(M, N)
    temb = torch.zeros(8, N, device="cuda")
    out32 = torch.empty(M, N, device="cuda", dtype=torch.float32)
    cases = [("plain", {}), ("bias", dict(bias=bias)), ("bias+res", dict(bias=bias, residual=res)),
             ("bias+silu", dict(bias=bias, act=1)), ("bias+geglu", dict(bias=bias, glu=GLU_GEGLU)),
             ("bias+radd", dict(bias=bias, row_add=temb, rows_per_group=M // 8)), ("fp32 out", dict(out_fp32=True, out=out32))]
    for name, kw in cases:
        b = kw.pop("bias", None)
        ms = timeit(lambda: ops.linear(a, w, b, **kw), iters=10)
        print(f"K={K:4d} {name:12s} {ms * 1e3:8.1f} us  {ms * 1e6 / tiles_per_cta:7.0f} ns/tile/CTA  {2.0 * M * N * K / ms / 1e9:7.0f} TFLOP/s",
              flush=True)
lib.b200mix_debug_force_bn(0)
