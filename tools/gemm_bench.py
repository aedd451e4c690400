"""Times the tcgen05 implicit-GEMM kernel on the GEMM / conv shapes of the SDXL forward (B=8), per tile width."""
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops  # noqa: E402
from paddlemix_b200._lib import GLU_GEGLU, lib  # noqa: E402

ops.init(0)
bf = torch.bfloat16


def timeit(fn, iters=20):
    """Device time per call: `iters` calls captured in one CUDA graph (no Python / launch overhead in the number)."""
    s = torch.cuda.Stream()
    s.wait_stream(torch.cuda.current_stream())
    with torch.cuda.stream(s):
        for _ in range(3):
            fn()
    torch.cuda.current_stream().wait_stream(s)
    torch.cuda.synchronize()
    g = torch.cuda.CUDAGraph()
    with torch.cuda.graph(g):
        for _ in range(iters):
            fn()
    g.replay()
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    g.replay()
    e1.record()
    torch.cuda.synchronize()
    return e0.elapsed_time(e1) / iters


def rnd(*s):
    return (torch.randn(*s, device="cuda") * 0.05).to(bf)


def main():
    LIN = [("ff1_640 geglu", 32768, 5120, 640, True), ("ff2_640", 32768, 640, 2560, False), ("qkv_640", 32768, 1920, 640, False),
           ("out_640", 32768, 640, 640, False), ("ff1_1280 geglu", 8192, 10240, 1280, True), ("ff2_1280", 8192, 1280, 5120, False),
           ("qkv_1280", 8192, 3840, 1280, False), ("out_1280", 8192, 1280, 1280, False), ("big 8192^3", 8192, 8192, 8192, False)]
    CONV = [("conv 320@128", 8, 128, 128, 320, 320), ("conv 640@64", 8, 64, 64, 640, 640), ("conv 1280@32", 8, 32, 32, 1280, 1280),
            ("conv 2560->1280@32", 8, 32, 32, 2560, 1280), ("conv 960->320@128", 8, 128, 128, 960, 320)]
    bns = [int(b) for b in os.environ.get("BNS", "0,160,192,224,256").split(",")]
    print(f"{'shape':24s}" + "".join(f"{('bn=%d' % b) if b else 'auto':>12s}" for b in bns) + "   (TFLOP/s)")
    for name, M, N, K, glu in LIN:
        a, w, bias = rnd(M, K), rnd(N, K), torch.zeros(N, device="cuda")
        res = None if glu else rnd(M, N)
        row = f"{name:24s}"
        for bn in bns:
            lib.b200mix_debug_force_bn(bn)
            ms = timeit(lambda: ops.linear(a, w, bias, glu=GLU_GEGLU if glu else 0, residual=res))
            row += f"{2.0 * M * N * K / ms / 1e9:12.0f}"
        print(row, flush=True)
    for name, B, H, W, Ci, Co in CONV:
        x, w, bias = rnd(B, H, W, Ci), rnd(Co, 3, 3, Ci), torch.zeros(Co, device="cuda")
        row = f"{name:24s}"
        for bn in bns:
            lib.b200mix_debug_force_bn(bn)
            ms = timeit(lambda: ops.conv3x3(x, w, bias))
            row += f"{2.0 * B * H * W * Co * 9 * Ci / ms / 1e9:12.0f}"
        print(row, flush=True)
    lib.b200mix_debug_force_bn(0)
    for name, M, N in [("layernorm 32768x640", 32768, 640), ("layernorm 8192x1280", 8192, 1280)]:
        x, w, b = rnd(M, N), torch.ones(N, device="cuda"), torch.zeros(N, device="cuda")
        ms = timeit(lambda: ops.layernorm(x, w, b))
        print(f"{name:24s} {2 * M * N * 2 / ms / 1e6:10.0f} GB/s")
    for name, B, HW, C in [("groupnorm 8x16384x320", 8, 16384, 320), ("groupnorm 8x4096x640", 8, 4096, 640), ("groupnorm 8x1024x1280", 8, 1024, 1280)]:
        x, w, b = rnd(B, HW, C), torch.ones(C, device="cuda"), torch.zeros(C, device="cuda")
        ms = timeit(lambda: ops.groupnorm_nhwc(x, w, b, silu=True))
        print(f"{name:24s} {2 * B * HW * C * 2 / ms / 1e6:10.0f} GB/s (algorithmic 1R+1W)")
    for name, B, S, H, D in [("attn 8x4096 h10 d64", 8, 4096, 10, 64), ("attn 8x1024 h20 d64", 8, 1024, 20, 64)]:
        qkv = rnd(B, S, 3 * H * D)
        q, k, v = (qkv[:, :, i * H * D:(i + 1) * H * D].unflatten(-1, (H, D)) for i in range(3))
        ms = timeit(lambda: ops.sdpa(q, k, v))
        print(f"{name:24s} {4.0 * B * H * S * S * D / ms / 1e9:10.0f} TFLOP/s")


if __name__ == "__main__":
    main()
