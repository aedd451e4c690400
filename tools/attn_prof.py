"""Where the ping-pong attention kernel's roles spend their cycles (build with -DATTN_PROF=1: tools/build_variant.sh prof
"-DATTN_PROF=1", run with B200MIX_LIB=.../libb200mix_prof.so). Prints per-key-block averages over all CTAs."""
import ctypes
import os
import sys

import numpy as np
import torch

sys.path.insert(0, os.path.dirname(os.path.dirname(os.path.abspath(__file__))))
from paddlemix_b200 import ops  # noqa: E402
from paddlemix_b200._lib import lib  # noqa: E402

ops.init(0)
MMA = ["wait q/k_full", "wait s_empty0", "wait s_empty1", "wait p_full0", "wait p_full1", "wait v_full"]
SM = ["wait s_full", "ld S + arrive", "mask/max/grow", "wait pv_done", "exp + st P + arrive", "wait pv_done (item end)",
      "item epilogue", "TOTAL loop"]
for B, S, H in [(8, 4096, 10), (8, 1024, 20)]:
    q, k, v = (torch.randn(B, S, H, 64, device="cuda").to(torch.bfloat16) for _ in range(3))
    for _ in range(2):
        ops.sdpa(q, k, v)
    torch.cuda.synchronize()
    buf = np.zeros(148 * 32, dtype=np.uint64)
    assert lib.b200mix_debug_attn_prof_read(buf.ctypes.data_as(ctypes.c_void_p)) == 0
    a = buf.reshape(148, 32).astype(np.float64)
    items = B * H * ((S + 255) // 256)
    blocks_per_cta = items / 148 * (S / 128)
    print(f"== B{B} S{S} H{H}: {items} items, {blocks_per_cta:.1f} key blocks per CTA; cycles per key block (mean over CTAs)")
    for i, n in enumerate(MMA):
        print(f"   MMA  {n:28s} {a[:, i].mean() / blocks_per_cta:9.1f}")
    for t in range(2):
        for i, n in enumerate(SM):
            print(f"   WG{t}  {n:28s} {a[:, 8 + 8 * t + i].mean() / blocks_per_cta:9.1f}")
