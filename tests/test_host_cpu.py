"""CPU-side checks: the C-ABI library loads and exports every symbol include/b200mix.h declares, refuses to run
without a GPU (no fallback), the host mirror builds the same structure as the oracle, and the multi-rank plumbing
(batch sharding + all_gather of finished latents) works under gloo with world_size 2."""
import ctypes
import os
import re
import subprocess
import sys

import pytest
import torch

ROOT = os.path.dirname(os.path.dirname(os.path.abspath(__file__)))


def test_library_exports_every_declared_symbol():
    from paddlemix_b200 import _lib
    hdr = open(os.path.join(ROOT, "include", "b200mix.h")).read()
    declared = set(re.findall(r"\b(b200mix_[a-z0-9_]+)\s*\(", hdr))
    declared -= {"b200mix_epilogue", "b200mix_status"}
    assert len(declared) >= 20
    for name in sorted(declared):
        assert hasattr(_lib.lib, name), f"{name} declared in include/b200mix.h but not exported"
    assert set(_lib.SIGNATURES) | set(_lib.STRING_GETTERS) == declared


def test_no_cpu_fallback():
    from paddlemix_b200 import _lib
    if torch.cuda.is_available():
        pytest.skip("GPU present")
    rc = _lib.lib.b200mix_init(0)
    assert rc == -3
    assert b"no CPU fallback" in _lib.lib.b200mix_last_error()
    from paddlemix_b200 import ops
    with pytest.raises(_lib.B200MixError):
        ops.linear(torch.zeros(8, 64, dtype=torch.bfloat16), torch.zeros(8, 64, dtype=torch.bfloat16))


def test_unet_structure_matches_oracle():
    from oracle import unet as O
    from paddlemix_b200.ppdiffusers.unet_2d_condition import UNet2DConditionModel
    for name in ("sd15", "sdxl", "tiny_sd", "tiny_xl"):
        cfg = O.UNET_CONFIGS[name]
        keys = ("in_channels", "out_channels", "down_block_types", "up_block_types", "block_out_channels",
                "layers_per_block", "cross_attention_dim", "transformer_layers_per_block", "attention_head_dim",
                "use_linear_projection", "addition_embed_type", "addition_time_embed_dim",
                "projection_class_embeddings_input_dim")
        m = UNet2DConditionModel(**{k: cfg[k] for k in keys})
        shapes = O.unet_param_shapes(cfg)
        assert m.state_dict_shapes() == shapes  # same parameter names and (Paddle-layout) shapes as the oracle
        names = {r.name for r in m.resnets} | {t.name for t in m.transformers}
        for r in m.resnets:
            assert shapes[r.name + ".conv1.weight"] == (r.cout, r.cin, 3, 3)
            assert ((r.name + ".conv_shortcut.weight") in shapes) == (r.shortcut is not None)
        n_attn = sum(2 * len(t.blocks) for t in m.transformers)
        assert len(m.attn_processors) == n_attn
        assert n_attn == sum(1 for k in shapes if k.endswith(".to_q.weight"))
        assert len(names) == len(m.resnets) + len(m.transformers)
    assert len(UNet2DConditionModel(**{k: O.UNET_CONFIGS["sdxl"][k] for k in keys}).attn_processors) == 140


def test_sd3_structure_matches_oracle():
    from oracle import sd3 as O3
    from paddlemix_b200.ppdiffusers.transformer_sd3 import SD3Transformer2DModel
    for name in ("sd3_medium", "tiny"):
        cfg = O3.SD3_CONFIGS[name]
        assert SD3Transformer2DModel(**cfg).state_dict_shapes() == O3.sd3_param_shapes(cfg)
    assert abs(O3.sd3_flops(O3.SD3_CONFIGS["sd3_medium"], 1, 128, 128, 154) / 1e12 - 8.437) < 5e-3


def test_qwen2vl_structure_and_rope_index_match_oracle():
    from oracle import qwen2vl as OQ
    from paddlemix_b200.qwen2_vl import Qwen2VLForConditionalGeneration
    for name in ("qwen2vl_7b", "tiny"):
        cfg = OQ.QWEN2VL_CONFIGS[name]
        assert Qwen2VLForConditionalGeneration(cfg).state_dict_shapes() == OQ.qwen2vl_param_shapes(cfg)
    cfg = OQ.QWEN2VL_CONFIGS["qwen2vl_7b"]
    m = Qwen2VLForConditionalGeneration(cfg)
    g = torch.Generator().manual_seed(0)
    grid = [[1, 32, 32]] * 2  # BASELINE config C4: 448x448 image -> 1024 patches -> 256 merged tokens + 510 text
    rows = [[cfg["vision_start_token_id"]] + [cfg["image_token_id"]] * 256 + [cfg["vision_end_token_id"]] +
            torch.randint(0, 151643, (510,), generator=g).tolist() for _ in range(2)]
    ids = torch.tensor(rows)
    po, do = OQ.get_rope_index(cfg, ids, grid)
    pm, dm = m.get_rope_index(ids, torch.tensor(grid))
    assert ids.shape == (2, 768) and torch.equal(po, pm) and torch.equal(do, dm)
    assert torch.equal(m.rot_pos_emb(grid), OQ.rot_pos_emb(cfg, grid))


def test_stdit2_structure_matches_oracle():
    from oracle import stdit2 as OS2
    from paddlemix_b200.opensora import STDiT2
    for name in ("stdit2_xl", "tiny"):
        cfg = OS2.STDIT2_CONFIGS[name]
        assert STDiT2(cfg).state_dict_shapes() == OS2.stdit2_param_shapes(cfg)
    assert abs(OS2.stdit2_flops(OS2.STDIT2_CONFIGS["stdit2_xl"], 1, 16, 1024, 120) / 1e12 - 24.4) < 0.1


def test_config_errors_mirror_reference():
    from paddlemix_b200.ppdiffusers.unet_2d_condition import UNet2DConditionModel
    with pytest.raises(ValueError, match="same number of `down_block_types`"):
        UNet2DConditionModel(down_block_types=("DownBlock2D",), up_block_types=("UpBlock2D", "UpBlock2D"), block_out_channels=(64,))
    with pytest.raises(ValueError, match="num_attention_heads"):
        UNet2DConditionModel(num_attention_heads=8)
    m = UNet2DConditionModel()
    assert m.config.in_channels == 4 and m.in_channels == 4 and m.dtype == torch.bfloat16
    with pytest.raises(Exception):
        m.config.in_channels = 3
    with pytest.raises(RuntimeError):
        m.forward(torch.zeros(1, 4, 8, 8), 1, torch.zeros(1, 77, 1280))
    # options outside the hot path are refused unless they sit at the reference's own default; True == 1.0 / 0 == False
    # style coincidences must not slip through (each option is compared with ITS default, by type and value)
    for kw in (dict(dual_cross_attention=True), dict(resnet_skip_time_act=True), dict(class_embeddings_concat=True),
               dict(class_embed_type="timestep"), dict(conv_in_kernel=5), dict(conv_in_kernel=3.0),
               dict(time_embedding_type="fourier"), dict(attention_type="gated"), dict(not_a_reference_option=None)):
        with pytest.raises(NotImplementedError):
            UNet2DConditionModel(**kw)
    UNet2DConditionModel(dual_cross_attention=False, conv_in_kernel=3, upcast_attention=True, class_embed_type=None,
                         _class_name="UNet2DConditionModel")


def test_flops_match_survey():
    from oracle import unet as O
    assert abs(O.unet_flops(O.UNET_CONFIGS["sdxl"], 1, 128, 128, 77) / 1e12 - 6.761) < 5e-3
    assert abs(O.unet_flops(O.UNET_CONFIGS["sd15"], 1, 64, 64, 77) / 1e12 - 0.803) < 5e-3


def test_oracle_self_consistency():
    # ModelTesterMixin-style properties on the reference's own tiny config (test_models_unet_2d_condition.py:181-194)
    from oracle import unet as O
    cfg = O.UNET_CONFIGS["ref_tiny"]
    P = O.init_params(O.unet_param_shapes(cfg))
    g = torch.Generator().manual_seed(0)
    x, ctx = torch.randn(4, 4, 32, 32, generator=g), torch.randn(4, 4, 32, generator=g)
    y = O.unet_forward(cfg, P, x, 10, ctx)
    assert y.shape == x.shape and torch.equal(y, O.unet_forward(cfg, P, x, 10, ctx))
    one = O.unet_forward(cfg, P, x[1:2], torch.tensor([10]), ctx[1:2])
    assert torch.allclose(one, y[1:2], atol=1e-4)

    def sdpa_proc(x, c, P, p, heads, mask):  # processor-swap equivalence (test_modeling_common.py:197-256)
        c = x if c is None else c
        q, k, v = O.linear(x, P, p + ".to_q"), O.linear(c, P, p + ".to_k"), O.linear(c, P, p + ".to_v")
        B, S, C = q.shape
        sp = lambda t: t.reshape(B, -1, heads, C // heads).transpose(1, 2)
        o = torch.nn.functional.scaled_dot_product_attention(sp(q), sp(k), sp(v)).transpose(1, 2).reshape(B, S, C)
        return O.linear(o, P, p + ".to_out.0")
    y2 = O.unet_forward(cfg, P, x, 10, ctx, processor=sdpa_proc)
    assert (y - y2).abs().max() < 1e-3


WORKER = r'''
import os, sys, torch, torch.distributed as dist
sys.path.insert(0, sys.argv[1])
from paddlemix_b200.ppdiffusers.pipelines import all_gather_latents, shard_batch
dist.init_process_group("gloo", init_method="tcp://127.0.0.1:%s" % sys.argv[2], rank=int(sys.argv[3]), world_size=2)
rank = dist.get_rank()
full = torch.arange(8 * 4 * 2 * 2, dtype=torch.float32).reshape(8, 4, 2, 2)
lo, hi = shard_batch(8, rank, 2)
mine = full[lo:hi] * 2.0            # "denoise" the local shard
out = all_gather_latents(mine)
assert out.shape == full.shape and torch.equal(out, full * 2.0), rank
assert shard_batch(5, 0, 2) == (0, 3) and shard_batch(5, 1, 2) == (3, 5)
dist.destroy_process_group()
print("ok", rank)
'''


def test_data_parallel_plumbing_gloo(tmp_path):
    script = tmp_path / "worker.py"
    script.write_text(WORKER)
    port = str(29500 + os.getpid() % 2000)
    procs = [subprocess.Popen([sys.executable, str(script), ROOT, port, str(r)], stdout=subprocess.PIPE,
                              stderr=subprocess.STDOUT) for r in range(2)]
    for p in procs:
        out, _ = p.communicate(timeout=120)
        assert p.returncode == 0, out.decode()
