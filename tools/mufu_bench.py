"""Issue-rate microbenchmarks behind the attention softmax design: which pipe each instruction of the inner loop uses and
what a mix sustains (thread-instructions per clock per SM; 8 x 256-thread CTAs per SM, long dependent chains x 4)."""
import ctypes
import os
import sys

import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops  # noqa: E402
from paddlemix_b200._lib import lib  # noqa: E402

ops.init(0)
lib.b200mix_debug_mufu_bench.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
out = torch.empty(148 * 8 * 256, device="cuda")
MODES = [(0, "4 ex2.f32", 4), (1, "4 ex2.f16x2", 4), (2, "4 cvt.rn.bf16x2.f32 (F2FP)", 4), (3, "4 ex2 + 2 cvt.bf16x2", 4),
         (4, "4 ex2 + 2 cvt.bf16x2 + 2 FFMA2 + 2 FADD2", 4), (5, "4 ex2 + integer bf16 pack (4 IADD + 2 PRMT)", 4),
         (6, "4 ex2 + integer pack + 2 FFMA2 + 2 FADD2", 4), (7, "4 FFMA2", 4), (8, "4 FMNMX3", 4), (9, "4 FFMA", 4)]
for mode, name, per in MODES:
    for _ in range(2):
        lib.b200mix_debug_mufu_bench(out.data_ptr(), 148 * 8, 256, 20000, mode, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.b200mix_debug_mufu_bench(out.data_ptr(), 148 * 8, 256, 20000, mode, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    iters = 148 * 8 * 256 * 20000
    print(f"mode {mode} [{name}]: {ms:8.3f} ms, {iters * per / ms / 1e6 / 148:8.1f} G 'unit'-ops/s/SM "
          f"= {iters * per / ms / 1e6 / 148 / 1.9:6.2f} per clk per SM @1.9 GHz", flush=True)
