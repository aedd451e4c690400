"""Runs one GEMM shape a few times (for ncu captures): python tools/one_gemm.py M N K [residual] [glu]"""
import os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops
from paddlemix_b200._lib import GLU_GEGLU
ops.init(0)
M, N, K = (int(v) for v in sys.argv[1:4])
res = "residual" in sys.argv
glu = "glu" in sys.argv
bf = torch.bfloat16
a, w = (torch.randn(M, K, device="cuda") * 0.05).to(bf), (torch.randn(N, K, device="cuda") * 0.05).to(bf)
bias = torch.zeros(N, device="cuda")
r = (torch.randn(M, N, device="cuda") * 0.05).to(bf) if res else None
for _ in range(5):
    ops.linear(a, w, bias, residual=r, glu=GLU_GEGLU if glu else 0)
torch.cuda.synchronize()
e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
e0.record()
ops.linear(a, w, bias, residual=r, glu=GLU_GEGLU if glu else 0)
e1.record()
torch.cuda.synchronize()
print(f"{M}x{N}x{K}: {e0.elapsed_time(e1) * 1e3:.1f} us (single launch incl. host gap)")
