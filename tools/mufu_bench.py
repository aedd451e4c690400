import ctypes, os, sys, torch
sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops
from paddlemix_b200._lib import lib
ops.init(0)
lib.b200mix_debug_mufu_bench.argtypes = [ctypes.c_void_p, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_int, ctypes.c_void_p]
out = torch.empty(148 * 8 * 256, device="cuda")
for mode, name, per in ((0, "ex2.approx.ftz.f32", 1), (1, "ex2.approx.f16x2", 2)):
    for _ in range(2):
        lib.b200mix_debug_mufu_bench(out.data_ptr(), 148 * 8, 256, 20000, mode, None)
    torch.cuda.synchronize()
    e0, e1 = torch.cuda.Event(enable_timing=True), torch.cuda.Event(enable_timing=True)
    e0.record()
    lib.b200mix_debug_mufu_bench(out.data_ptr(), 148 * 8, 256, 20000, mode, None)
    e1.record()
    torch.cuda.synchronize()
    ms = e0.elapsed_time(e1)
    n_instr = 148 * 8 * 256 * 20000 * 4
    print(f"{name}: {n_instr / ms / 1e6:.1f} G thread-instr/s, {n_instr * per / ms / 1e6:.1f} G exps/s "
          f"({n_instr / ms / 1e6 / 148 / 1.9:.2f} thread-instr/clk/SM @1.9GHz)")
