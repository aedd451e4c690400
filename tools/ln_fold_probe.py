"""Cost of the folded-LayerNorm epilogues on the SDXL projection shapes (graph-timed, one problem at a time):
consumer = plain / +bias / +ln (statistics + column sums + transform), producer = bias+res with / without statistics."""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops
sys.path.insert(0, os.path.dirname(os.path.abspath(__file__)))
from gemm_bench import timeit, rnd  # noqa
ops.init(0)
M = 8192
for name, N, K, glu in [("qkv 3840x1280", 3840, 1280, 0), ("ff1 10240x1280 geglu", 10240, 1280, 1), ("q 1280x1280", 1280, 1280, 0)]:
    h, w = rnd(M, K), rnd(N, K)
    bias, colsum = torch.zeros(N, device="cuda"), torch.zeros(N, device="cuda")
    _, st = ops.linear(rnd(M, 256), rnd(K, 256), residual=h, stats=True)
    fl = 2.0 * M * N * K
    row = f"consumer {name:22s}"
    for tag, fn in [("plain", lambda: ops.linear(h, w, glu=glu)), ("bias", lambda: ops.linear(h, w, bias, glu=glu)),
                    ("ln", lambda: ops.linear(h, w, bias, glu=glu, ln=(st, colsum, 1e-5)))]:
        row += f"  {tag}: {fl / timeit(fn) / 1e9:6.0f}"
    print(row, "TFLOP/s", flush=True)
for name, N, K in [("out 1280x1280", 1280, 1280), ("ff2 1280x5120", 1280, 5120)]:
    a, w, res, bias = rnd(M, K), rnd(N, K), rnd(M, N), torch.zeros(N, device="cuda")
    tab = torch.zeros(M, 2, device="cuda", dtype=torch.int64)
    fl = 2.0 * M * N * K
    t0 = timeit(lambda: ops.linear(a, w, bias, residual=res))
    t1 = timeit(lambda: ops.linear(a, w, bias, residual=res, stats=tab))
    print(f"producer {name:22s}  bias+res: {fl / t0 / 1e9:6.0f}  +stats: {fl / t1 / 1e9:6.0f} TFLOP/s", flush=True)
